// Kernel launchers, one translation unit per registered objective (trial_obj.cu compiled with -DOODE_OBJ=k):
// the objective is a template parameter of every hot kernel, so the six instantiation sets compile in
// parallel and oode.cu (host logic) only sees these plain functions.
#pragma once
#include "de_kernels.cuh"

namespace oode {

enum { FORM_QUAD = 0, FORM_TMA = 1, FORM_TMA2 = 2, FORM_ROWS = 3 };

// most warps per CTA of de_trial_rows_kernel by workers per warp (register budget; -D overrides for A/B builds)
#ifndef OODE_ROWS_MAXW1
#define OODE_ROWS_MAXW1 20
#endif
#ifndef OODE_ROWS_MAXW2
#define OODE_ROWS_MAXW2 12
#endif
constexpr int ROWS_S = 2;                  // shared-memory stages per worker of de_trial_rows_kernel
constexpr int XBUF_DOUBLES = 64;           // a worker's exchange buffer: the leaf sums of one row chunk (peer exchange)
// Shared memory of a CTA: [workers][S] stages of 4 input rows of `rs` doubles, one mbarrier per stage, then one
// exchange buffer per worker.
inline size_t rows_smem_bytes(int warps, int wpw, int rs) {
    return (size_t)warps * wpw * (ROWS_S * (4 * (size_t)rs * 8 + 8) + XBUF_DOUBLES * 8);
}
constexpr int ROWS_SMEM_MAX = 232448;      // 227 KB: the most dynamic shared memory a CTA can have on sm_100
inline int rows_maxw(int wpw) { return wpw == 1 ? OODE_ROWS_MAXW1 : OODE_ROWS_MAXW2; }

struct TrialLaunch {
    int form;              // FORM_*
    int wpw, warps;        // FORM_ROWS: workers per warp (1 or 2) and warps per CTA
    bool uni;              // every column has the same bounds
    bool fast;             // FORM_ROWS: single-reflection repair is provably enough and cos arguments are small
    bool cos_checked;      // FORM_TMA/TMA2: keep the |y| <= 1e6 check of obj_cos
    int sm_count, ctas_per_sm;
    cudaStream_t stream;
};

typedef cudaError_t (*TrialFn)(const TrialLaunch &, int mode, const GenParams &);
typedef cudaError_t (*SelectFn)(bool constrained, cudaStream_t, const SelParams &);
typedef cudaError_t (*FinalizeFn)(cudaStream_t, const PartLayout &, const int *prog, int nprog, int D, int invert, int64_t n,
                                  double *out);

// What a separately compiled objective (trial_obj.cu with -DOODE_OBJ=OODE_OBJ_CUSTOM) hands to oode_register_objective
constexpr int OODE_OBJ_CUSTOM = 6;
struct ObjEntry {
    int abi;                    // OODE_ABI_VERSION of the headers it was compiled against
    int size_gen, size_sel;     // sizeof(GenParams), sizeof(SelParams): the kernel parameter blocks must agree
    int nsum, nprod, halo, needs_aux;
    TrialFn trial;
    SelectFn select;
    FinalizeFn finalize;
};

#define OODE_DECLARE_OBJ(k)                                                           \
    cudaError_t launch_trial_obj_##k(const TrialLaunch &, int, const GenParams &);    \
    cudaError_t launch_select_obj_##k(bool, cudaStream_t, const SelParams &);         \
    cudaError_t launch_finalize_obj_##k(cudaStream_t, const PartLayout &, const int *, int, int, int, int64_t, double *);
OODE_DECLARE_OBJ(0) OODE_DECLARE_OBJ(1) OODE_DECLARE_OBJ(2) OODE_DECLARE_OBJ(3) OODE_DECLARE_OBJ(4) OODE_DECLARE_OBJ(5)
#undef OODE_DECLARE_OBJ

}  // namespace oode
