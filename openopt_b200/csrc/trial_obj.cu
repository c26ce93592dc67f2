// Every kernel instantiation of ONE registered objective (-DOODE_OBJ=k) and its launchers (de_launch.h).
#include "de_launch.h"
#include "de_trial_tma.cuh"
#include "de_trial_rows.cuh"

#ifndef OODE_OBJ
#error "compile with -DOODE_OBJ=<objective id>"
#endif
#define OODE_CAT2(a, b) a##b
#define OODE_CAT(a, b) OODE_CAT2(a, b)

namespace oode {

constexpr int OBJ = OODE_OBJ;
constexpr bool HALO = Obj<OBJ>::HALO != 0;

// ---- quad-per-leaf form (narrow rows, hndvi > 1, neighbour-coupled objectives) ----
template <bool UNI> static cudaError_t launch_quad(int mode, int grid, cudaStream_t s, const GenParams &p) {
    switch (mode) {
        case MODE_EVAL: de_trial_kernel<OBJ, MODE_EVAL, true><<<grid, 256, 0, s>>>(p); break;
        case MODE_PHILOX: de_trial_kernel<OBJ, MODE_PHILOX, UNI><<<grid, 256, 0, s>>>(p); break;
        default: de_trial_kernel<OBJ, MODE_REPLAY, UNI><<<grid, 256, 0, s>>>(p); break;
    }
    return cudaGetLastError();
}

// ---- round-1 TMA forms: neighbour-coupled objectives (and A/B builds with -DOODE_KEEP_OLD_TMA) ----
#ifndef OODE_TMA_NC
#define OODE_TMA_NC 20
#endif
#ifndef OODE_TMA2_NC
#define OODE_TMA2_NC 12
#endif
#if defined(OODE_KEEP_OLD_TMA)
constexpr bool OLD_TMA = true;
#else
constexpr bool OLD_TMA = HALO;
#endif

template <bool UNI, bool HALF> static cudaError_t launch_old_tma(int mode, bool cos_checked, int sm_count, cudaStream_t s,
                                                                 const GenParams &p) {
    if constexpr (OLD_TMA) {
        constexpr int NC = HALF ? OODE_TMA2_NC : OODE_TMA_NC;
        constexpr size_t smem = HALF ? Tma2Smem<NC, 2>::total : TmaSmem<NC, 2>::total;
        const int per_cta = HALF ? 2 * NC : NC;
        const int grid = (int)max((int64_t)1, min(((int64_t)p.nrows + per_cta - 1) / per_cta, (int64_t)sm_count));
#define OODE_OLD_LAUNCH(M, U, C)                                                                                \
    do {                                                                                                        \
        if constexpr (HALF) {                                                                                   \
            auto k = de_trial_tma2_kernel<OBJ, M, U, C, NC, 2>;                                                 \
            cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);    \
            if (e != cudaSuccess) return e;                                                                     \
            k<<<grid, NC * 32, smem, s>>>(p);                                                                   \
        } else {                                                                                                \
            auto k = de_trial_tma_kernel<OBJ, M, U, C, NC, 2>;                                                  \
            cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);    \
            if (e != cudaSuccess) return e;                                                                     \
            k<<<grid, NC * 32, smem, s>>>(p);                                                                   \
        }                                                                                                       \
    } while (0)
        switch (mode) {
            case MODE_EVAL: OODE_OLD_LAUNCH(MODE_EVAL, true, true); break;
            case MODE_PHILOX:
                if (cos_checked) OODE_OLD_LAUNCH(MODE_PHILOX, UNI, true);
                else OODE_OLD_LAUNCH(MODE_PHILOX, UNI, false);
                break;
            default: OODE_OLD_LAUNCH(MODE_REPLAY, UNI, true); break;
        }
#undef OODE_OLD_LAUNCH
        return cudaGetLastError();
    } else {
        return cudaErrorInvalidDeviceFunction;      // this build carries no round-1 TMA kernels for the objective
    }
}

// ---- row-walking form (de_trial_rows.cuh) ----
template <int MODE, bool UNI, bool FAST, int WPW> static cudaError_t launch_rows_one(const TrialLaunch &L, const GenParams &p) {
    if constexpr (!HALO) {
        constexpr int MAXW = WPW == 1 ? OODE_ROWS_MAXW1 : OODE_ROWS_MAXW2;
        auto k = de_trial_rows_kernel<OBJ, MODE, UNI, FAST, WPW, MAXW>;
        static bool attr_set = false;
        if (!attr_set) {
            cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, ROWS_SMEM_MAX);
            if (e != cudaSuccess) return e;
            attr_set = true;
        }
        const int warps = max(1, min(L.warps, MAXW));
        const size_t smem = rows_smem_bytes(warps, WPW, p.rs);
        const int per_cta = warps * WPW;
        const int grid = (int)max((int64_t)1, min(((int64_t)p.nrows + per_cta - 1) / per_cta, (int64_t)L.sm_count));
        k<<<grid, warps * 32, smem, L.stream>>>(p);
        return cudaGetLastError();
    } else {
        return cudaErrorInvalidDeviceFunction;
    }
}

#if OODE_OBJ >= 6
// run-time registered objective: one worker per warp, SAFE instantiations only (a third of the compile time)
template <int WPW> static cudaError_t launch_rows_wpw(const TrialLaunch &L, int mode, const GenParams &p) {
    switch (mode) {
        case MODE_EVAL: return launch_rows_one<MODE_EVAL, true, false, 1>(L, p);
        case MODE_PHILOX:
            return L.uni ? launch_rows_one<MODE_PHILOX, true, false, 1>(L, p) : launch_rows_one<MODE_PHILOX, false, false, 1>(L, p);
        default:
            return L.uni ? launch_rows_one<MODE_REPLAY, true, false, 1>(L, p) : launch_rows_one<MODE_REPLAY, false, false, 1>(L, p);
    }
}
#else
template <int WPW> static cudaError_t launch_rows_wpw(const TrialLaunch &L, int mode, const GenParams &p) {
    switch (mode) {
        case MODE_EVAL:
            return L.fast ? launch_rows_one<MODE_EVAL, true, true, WPW>(L, p) : launch_rows_one<MODE_EVAL, true, false, WPW>(L, p);
        case MODE_PHILOX:
            if (L.uni) return L.fast ? launch_rows_one<MODE_PHILOX, true, true, WPW>(L, p) : launch_rows_one<MODE_PHILOX, true, false, WPW>(L, p);
            return L.fast ? launch_rows_one<MODE_PHILOX, false, true, WPW>(L, p) : launch_rows_one<MODE_PHILOX, false, false, WPW>(L, p);
        default:
            return L.uni ? launch_rows_one<MODE_REPLAY, true, false, WPW>(L, p) : launch_rows_one<MODE_REPLAY, false, false, WPW>(L, p);
    }
}
#endif

cudaError_t OODE_CAT(launch_trial_obj_, OODE_OBJ)(const TrialLaunch &L, int mode, const GenParams &p) {
    switch (L.form) {
        case FORM_ROWS:
            if (L.wpw == 1) return launch_rows_wpw<1>(L, mode, p);
            return launch_rows_wpw<2>(L, mode, p);
        case FORM_TMA:
            return L.uni ? launch_old_tma<true, false>(mode, L.cos_checked, L.sm_count, L.stream, p)
                         : launch_old_tma<false, false>(mode, L.cos_checked, L.sm_count, L.stream, p);
        case FORM_TMA2:
            return L.uni ? launch_old_tma<true, true>(mode, L.cos_checked, L.sm_count, L.stream, p)
                         : launch_old_tma<false, true>(mode, L.cos_checked, L.sm_count, L.stream, p);
        default: {
            const int64_t nitems = (int64_t)p.nrows * p.nleaves;
            const int64_t want = (nitems * 4 + 255) / 256;     // one quad of lanes per (row, leaf) item
            const int grid = (int)max((int64_t)1, min(want, (int64_t)L.sm_count * L.ctas_per_sm));
            return L.uni ? launch_quad<true>(mode, grid, L.stream, p) : launch_quad<false>(mode, grid, L.stream, p);
        }
    }
}

cudaError_t OODE_CAT(launch_select_obj_, OODE_OBJ)(bool constrained, cudaStream_t s, const SelParams &p) {
    const int grid = (int)((p.nrows + 255) / 256);
    if (constrained) de_select_kernel<OBJ, true><<<grid, 256, 0, s>>>(p);
    else de_select_kernel<OBJ, false><<<grid, 256, 0, s>>>(p);
    return cudaGetLastError();
}

cudaError_t OODE_CAT(launch_finalize_obj_, OODE_OBJ)(cudaStream_t s, const PartLayout &part, const int *prog, int nprog, int D,
                                                     int invert, int64_t n, double *out) {
    const int grid = (int)((n + 255) / 256);
    de_finalize_kernel<OBJ><<<grid, 256, 0, s>>>(part, prog, nprog, D, invert, n, out);
    return cudaGetLastError();
}

#if OODE_OBJ >= 6
}  // namespace oode
extern "C" int oode_custom_objective_v1(oode::ObjEntry *e) {
    using namespace oode;
    e->abi = OODE_ABI_VERSION;
    e->size_gen = (int)sizeof(GenParams);
    e->size_sel = (int)sizeof(SelParams);
    e->nsum = Obj<OBJ>::NS; e->nprod = Obj<OBJ>::NPR; e->halo = Obj<OBJ>::HALO; e->needs_aux = Obj<OBJ>::AUX ? 1 : 0;
    e->trial = OODE_CAT(launch_trial_obj_, OODE_OBJ);
    e->select = OODE_CAT(launch_select_obj_, OODE_OBJ);
    e->finalize = OODE_CAT(launch_finalize_obj_, OODE_OBJ);
    return 0;
}
namespace oode {
#endif

}  // namespace oode
