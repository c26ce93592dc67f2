// K2a, wide-row form: warp-specialised trial kernel fed by TMA bulk copies.
//
// Why: the population rows a trial needs (parent + three random donors) are scattered over HBM,
// and a leaf of the reduction tree is <= 1 KB per row.  Reading such a leaf 64 bytes at a time
// (de_trial_kernel, one quad per leaf) spreads the 16 accesses to one DRAM page over thousands of
// cycles, so almost every access re-opens its page; ncu showed that form stuck near 60 % DRAM
// throughput with warps waiting on loads issued > 2000 cycles earlier.  Here each (row, leaf) work
// item is fetched as four contiguous bulk copies (cp.async.bulk, up to 1040 B each) into shared
// memory, completion signalled on an mbarrier: DRAM sees whole-leaf bursts, no registers are
// spent on prefetching, and the loads of S items per warp are in flight while it computes.
//
//   every warp       walks whole rows and owns S stages.  Per (row, leaf) item: wait full[s]; lane L
//                    makes the genes of columns 2L,2L+1 and 64+2L,64+2L+1 of the leaf (one Philox
//                    call per pair), writes the trial row as 16-byte vectors (512 contiguous bytes
//                    per warp instruction); the objective terms overwrite the consumed inputs in
//                    the stage and 8 lanes per sum slot walk them in numpy's r[j] chain order
//                    (bit-identical pairwise sum); then lane 0 refills the stage with the warp's
//                    item S positions ahead (a new row's pointers were resolved by loads issued
//                    before / in the middle of the compute phase).
//   (A first version with a dedicated producer warp starved the consumers: TMA issue is a
//    uniform-datapath instruction, so 30 lanes x 4 copies were serialised in one warp.)
#pragma once
#include "de_kernels.cuh"

namespace oode {

constexpr int TMA_ROW_DOUBLES = 130;                       // 128-term leaf + halo gene + pad
constexpr int TMA_ROW_BYTES = TMA_ROW_DOUBLES * 8;         // 1040, multiple of 16
constexpr int TMA_STAGE_BYTES = 4 * TMA_ROW_BYTES;

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred P1;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\n\t"
        "selp.b32 %0, 1, 0, P1;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ bool mbar_test_wait(uint32_t bar, uint32_t parity) {   // never suspends
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred P1;\n\t"
        "mbarrier.test_wait.parity.shared::cta.b64 P1, [%1], %2;\n\t"
        "selp.b32 %0, 1, 0, P1;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    while (!mbar_try_wait(bar, parity)) {
    }
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
                 "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}

__device__ __forceinline__ void split_item(const GenParams &p, int64_t nitems, int64_t item, int64_t &lrow, int &leaf) {
    if (p.nleaves == 1) { lrow = item; leaf = 0; }
    else if (nitems < 0x7FFFFFFFll) {
        const unsigned q = (unsigned)item / (unsigned)p.nleaves;
        lrow = q;
        leaf = (int)((unsigned)item - q * (unsigned)p.nleaves);
    } else { lrow = item / p.nleaves; leaf = (int)(item - lrow * p.nleaves); }
}

// Philox round keys held in registers (the asm move keeps the compiler from re-reading them from
// the constant bank at every call).
struct RegKeys {
    uint32_t k[20];
    // zero: a value the compiler cannot see (always 0, read from memory).  With OODE_EXP_PIN_KEYS the keys are
    // XOR-ed with it, which really pins them in vector registers (ptxas propagates through the plain `mov` and
    // re-reads them with LDCU.128 at every use); that needs ~20 more registers, i.e. fewer warps per CTA
    // (-DOODE_TMA_NC=16): an A/B candidate (scripts/ab.py), not the default.
    __device__ __forceinline__ void load(const PhiloxKeys &K, uint32_t zero) {
#pragma unroll
        for (int i = 0; i < 20; ++i) {
#ifdef OODE_EXP_PIN_KEYS
            k[i] = K.rk[i] ^ zero;
#else
            (void)zero;
            asm volatile("mov.u32 %0, %1;" : "=r"(k[i]) : "r"(K.rk[i]));
#endif
        }
    }
};

__device__ __forceinline__ void philox_rk(uint32_t &c0, uint32_t &c1, uint32_t &c2, uint32_t &c3, const RegKeys &K) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint64_t p0 = (uint64_t)PHILOX_M0 * c0;
        const uint64_t p1 = (uint64_t)PHILOX_M1 * c2;
        const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ K.k[2 * r];
        const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ K.k[2 * r + 1];
        c1 = (uint32_t)p1;
        c3 = (uint32_t)p0;
        c0 = n0;
        c2 = n2;
    }
}

// per-kernel invariants kept in registers
struct TrialConsts {
    double Fa, Fb;           // Ft = fma(uf, Fa, Fb): (F, 0), or (0, F) for the 'constant' difference factor
    double Cr, lo, hi;
    uint32_t gen;
    bool cf;                 // 'constant' difference factor: Ft = F
    uint32_t crT;            // crossover threshold on the raw Philox word (keep_parent)
    bool crAll;
};

// genes of one column pair from staged inputs (smem), lc = column inside the leaf (even)
template <int MODE, bool UNI>
__device__ __forceinline__ double2 gene_pair_smem(const GenParams &p, const TrialConsts &tc, const RegKeys &K,
                                                  const double *st, int lc, int gcol_local, uint32_t row_lo,
                                                  uint32_t row_hi, const double *Uf, const double *Uc, bool ok1,
                                                  unsigned &nrep) {
    const double2 xo = *reinterpret_cast<const double2 *>(st + lc);
    if (MODE == MODE_EVAL) return xo;
    const double2 xb = *reinterpret_cast<const double2 *>(st + TMA_ROW_DOUBLES + lc);
    const double2 x1 = *reinterpret_cast<const double2 *>(st + 2 * TMA_ROW_DOUBLES + lc);
    const double2 x2 = *reinterpret_cast<const double2 *>(st + 3 * TMA_ROW_DOUBLES + lc);
    const int col = p.col0 + gcol_local;
    double uf0, uf1;
    bool keep0, keep1;
    if (MODE == MODE_PHILOX) {
        uint32_t w0 = (uint32_t)col >> 1, w1 = row_lo, w2 = row_hi | (TAG_ELEM << 24), w3 = tc.gen;
        philox_rk(w0, w1, w2, w3, K);
        uf0 = u32_to_unit(w0); keep0 = keep_parent(w1, tc.crT, tc.crAll);
        uf1 = u32_to_unit(w2); keep1 = keep_parent(w3, tc.crT, tc.crAll);
    } else {
        keep0 = Uc[col] > tc.Cr;
        keep1 = ok1 ? Uc[col + 1] > tc.Cr : true;
        uf0 = tc.cf ? 0.0 : Uf[col];
        uf1 = (ok1 && !tc.cf) ? Uf[col + 1] : 0.0;
    }
    double lo0, lo1, hi0, hi1;
    if (UNI) { lo0 = lo1 = tc.lo; hi0 = hi1 = tc.hi; }
    else {
        const double2 l = *reinterpret_cast<const double2 *>(p.lb + gcol_local);
        const double2 h = *reinterpret_cast<const double2 *>(p.ub + gcol_local);
        lo0 = l.x; lo1 = l.y; hi0 = h.x; hi1 = h.y;
    }
    unsigned r1 = 0;
    double2 r;
    r.x = finish_gene(xo.x, xb.x, x1.x, x2.x, uf0, keep0, tc.Fa, tc.Fb, lo0, hi0, nrep);
    r.y = finish_gene(xo.y, xb.y, x1.y, x2.y, uf1, keep1, tc.Fa, tc.Fb, lo1, hi1, r1);
    nrep += ok1 ? r1 : 0u;
    return r;
}

// one gene from staged inputs (halo neighbour)
template <int MODE, bool UNI>
__device__ __forceinline__ double gene_one_smem(const GenParams &p, const TrialConsts &tc, const RegKeys &K,
                                                const double *st, int lc, int gcol_local, uint32_t row_lo,
                                                uint32_t row_hi, const double *Uf, const double *Uc) {
    if (MODE == MODE_EVAL) return st[lc];
    const int col = p.col0 + gcol_local;
    double uf;
    bool keep;
    if (MODE == MODE_PHILOX) {
        uint32_t w0 = (uint32_t)col >> 1, w1 = row_lo, w2 = row_hi | (TAG_ELEM << 24), w3 = tc.gen;
        philox_rk(w0, w1, w2, w3, K);
        uf = u32_to_unit((col & 1) ? w2 : w0);
        keep = keep_parent((col & 1) ? w3 : w1, tc.crT, tc.crAll);
    } else { uf = tc.cf ? 0.0 : Uf[col]; keep = Uc[col] > tc.Cr; }
    const double lo = UNI ? tc.lo : p.lb[gcol_local], hi = UNI ? tc.hi : p.ub[gcol_local];
    unsigned dummy = 0;
    return finish_gene(st[lc], st[TMA_ROW_DOUBLES + lc], st[2 * TMA_ROW_DOUBLES + lc], st[3 * TMA_ROW_DOUBLES + lc], uf,
                       keep, tc.Fa, tc.Fb, lo, hi, dummy);
}

template <int NC, int S> struct TmaSmem {
    static constexpr size_t stages = (size_t)NC * S * TMA_STAGE_BYTES;
    static constexpr size_t bars = (size_t)NC * S * sizeof(uint64_t);
    static constexpr size_t total = stages + bars;
};

// Source rows of one trial (slot-resolved base pointers, no leaf offset).
struct RowSrc {
    const double *P, *B, *R1, *R2;
};

// Genes, trial-row store and objective terms of one leaf held in a stage.  FULL: the leaf has
// exactly 128 terms and 128 genes, so no lane predicate is needed anywhere.  The terms overwrite
// the consumed inputs (slot q over input row q of the stage).
template <int OBJ, int MODE, bool UNI, bool FULL, bool CK, typename MidFn>
__device__ __forceinline__ void leaf_body(const GenParams &p, const TrialConsts &tc, const RegKeys &K, double *st,
                                          int off, int len, int glen, double *Wrow, const double *Uf, const double *Uc,
                                          uint32_t row_lo, uint32_t row_hi, int lane, unsigned &nrep, bool mid_flag,
                                          MidFn mid) {
    using O = Obj<OBJ>;
    constexpr int NSL = O::NS + O::NPR;
    // Not FULL: genes are still made and stored pair-wise; an odd count just drags one pad gene
    // along (its inputs were copied with the 16-byte rounding, its slot in the row is the pad
    // column), so a single predicate per half is enough.
    const int glen2 = (glen + 1) & ~1;
    double t[2][2][NSL];
#pragma unroll
    for (int half = 0; half < 2; ++half) {
        const int lc = 64 * half + 2 * lane;                         // column inside the leaf
        if (FULL || lc < glen2) {
            const bool ok1 = FULL || lc + 1 < glen;
            const double2 x = gene_pair_smem<MODE, UNI>(p, tc, K, st, lc, off + lc, row_lo, row_hi, Uf, Uc, ok1, nrep);
            if (MODE != MODE_EVAL) *reinterpret_cast<double2 *>(Wrow + off + lc) = x;
            double xn1 = 0.0;
            if (O::HALO && lc + 1 < len)
                xn1 = gene_one_smem<MODE, UNI>(p, tc, K, st, lc + 2, off + lc + 2, row_lo, row_hi, Uf, Uc);
            double auxa = 0.0, auxb = 0.0;
            if (O::AUX) { const double2 av = *reinterpret_cast<const double2 *>(p.aux + off + lc); auxa = av.x; auxb = av.y; }
            O::template term<CK>(x.x, x.y, auxa, p.col0 + off + lc, t[half][0]);
            O::template term<CK>(x.y, xn1, auxb, p.col0 + off + lc + 1, t[half][1]);
        }
        if (half == 0 && mid_flag) mid();
    }
    __syncwarp();                            // all inputs of the stage have been read
#pragma unroll
    for (int half = 0; half < 2; ++half) {
        const int lc = 64 * half + 2 * lane;
        if (FULL || lc < glen2) {
#pragma unroll
            for (int q = 0; q < NSL; ++q)
                *reinterpret_cast<double2 *>(st + q * TMA_ROW_DOUBLES + lc) = make_double2(t[half][0][q], t[half][1][q]);
        }
    }
    __syncwarp();
}

// numpy's pairwise leaf sum over the terms in the stage: lanes 8q..8q+7 are the r[j] chains.
template <int OBJ, bool FULL>
__device__ __forceinline__ void leaf_sum(const GenParams &p, const double *st, int len, int lane, int64_t out_index) {
    using O = Obj<OBJ>;
    constexpr int NSL = O::NS + O::NPR;
    const int q = lane >> 3, jn = lane & 7;
    if (q < NSL) {
        const double *tq = st + q * TMA_ROW_DOUBLES;
        const unsigned gm = 0xFFu << (lane & 24);
        double res;
        if (FULL) {
            double r = tq[jn];
#pragma unroll
            for (int k = 1; k < 16; ++k) r = slot_op<OBJ>(q, r, tq[8 * k + jn]);
            r = slot_op<OBJ>(q, r, __shfl_xor_sync(gm, r, 1));
            r = slot_op<OBJ>(q, r, __shfl_xor_sync(gm, r, 2));
            r = slot_op<OBJ>(q, r, __shfl_xor_sync(gm, r, 4));
            put_partial<8>(p, out_index + q, r, jn);
        } else if (len >= 96 && (len & 7) == 0) {     // 12..15 full steps, no tail (the 120-term leaves of D = 1000)
            const int nfull = len >> 3;
            double r = tq[jn];
#pragma unroll
            for (int k = 1; k < 12; ++k) r = slot_op<OBJ>(q, r, tq[8 * k + jn]);
#pragma unroll
            for (int k = 12; k < 16; ++k)
                if (k < nfull) r = slot_op<OBJ>(q, r, tq[8 * k + jn]);
            r = slot_op<OBJ>(q, r, __shfl_xor_sync(gm, r, 1));
            r = slot_op<OBJ>(q, r, __shfl_xor_sync(gm, r, 2));
            r = slot_op<OBJ>(q, r, __shfl_xor_sync(gm, r, 4));
            put_partial<8>(p, out_index + q, r, jn);
        } else {
            const int nfull = len >> 3;
            int i = nfull * 8;
            if (nfull > 0) {
                double r = tq[jn];
                for (int k = 1; k < nfull; ++k) r = slot_op<OBJ>(q, r, tq[8 * k + jn]);
                r = slot_op<OBJ>(q, r, __shfl_xor_sync(gm, r, 1));
                r = slot_op<OBJ>(q, r, __shfl_xor_sync(gm, r, 2));
                r = slot_op<OBJ>(q, r, __shfl_xor_sync(gm, r, 4));
                res = r;
            } else {
                res = (q < O::NS) ? 0.0 : 1.0;     // numpy: n < 8 starts from 0. (1. for a product)
            }
            if (jn == 0) {
                for (; i < len; ++i) res = slot_op<OBJ>(q, res, tq[i]);
                put_partial<1>(p, out_index + q, res, 0);
            }
        }
    }
    __syncwarp();
}

// Every warp is a self-serving consumer that walks WHOLE ROWS (row, leaf 0..nleaves-1), so the
// row-level work -- donor indices, row-version selectors, base pointers -- is done once per row.
// It owns S stages; after finishing an item it refills the drained stage with the item S positions
// ahead in its own (row, leaf) sequence (lane 0 issues the four bulk copies).  When that look-ahead
// position enters a new row, the row's pointers are resolved by loads issued before / in the middle
// of the current item's compute phase.
// CK: keep the |y| <= 1e6 check of obj_cos (false when the host knows the bounds make it redundant).
template <int OBJ, int MODE, bool UNI, bool CK, int NC, int S>
__global__ void __launch_bounds__(NC * 32, 1) de_trial_tma_kernel(const GenParams p) {
    using O = Obj<OBJ>;
    constexpr int NSL = O::NS + O::NPR;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double *stage_base = reinterpret_cast<double *>(smem_raw);
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem_raw + TmaSmem<NC, S>::stages);
    const int c = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int i = 0; i < NC * S; ++i) mbar_init(smem_u32(bars + i), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    const int nleaves = p.nleaves;
    const int G = (int)gridDim.x * NC;                          // rows advance by G per warp (32-bit: NP < 2^31)
    const int first = (int)blockIdx.x * NC + c;
    const uint32_t bar0 = smem_u32(bars + c * S);
    const uint32_t dst0 = smem_u32(stage_base) + (uint32_t)(c * S) * TMA_STAGE_BYTES;
    double *my_stages = stage_base + (size_t)(c * S) * (TMA_STAGE_BYTES / 8);
    RegKeys K;
    if (MODE == MODE_PHILOX) K.load(p.keys, (uint32_t)*p.n_repaired_zero);
    // kernel-wide constants, passed through a value the compiler cannot see (always 0) so that
    // they are held in registers instead of being re-read from the constant bank at every use
    TrialConsts tc;
    int ld;
    {
        const long long z = (long long)*p.n_repaired_zero;
        tc.Fa = __longlong_as_double(__double_as_longlong(p.const_F ? 0.0 : p.F) ^ z);
        tc.Fb = __longlong_as_double(__double_as_longlong(p.const_F ? p.F : 0.0) ^ z);
        tc.Cr = __longlong_as_double(__double_as_longlong(p.Cr) ^ z);
        tc.lo = __longlong_as_double(__double_as_longlong(p.lo_s) ^ z);
        tc.hi = __longlong_as_double(__double_as_longlong(p.hi_s) ^ z);
        tc.gen = p.gen ^ (uint32_t)z;
        tc.cf = (p.const_F ^ (int)z) != 0;
        tc.crT = p.cr_thresh ^ (uint32_t)z;
        tc.crAll = (p.cr_all ^ (int)z) != 0;
        ld = p.ld ^ (int)z;
    }
    const double *const buf0 = p.buf0, *const buf1 = p.buf1;
    const bool vec_base = p.base_vec != nullptr, vec_dir = p.d1_vec != nullptr;

    // look-ahead cursor (the next item to be issued) and its row's sources
    int la_row = first;
    int la_leaf = 0;
    RowSrc la;
    la.P = la.B = la.R1 = la.R2 = nullptr;
    int f_ib = 0, f_r1 = 0, f_r2 = 0;
    uint8_t f_srow = 0;

    auto resolve1 = [&](int lrow) {                         // first-level loads of a row
        const int row = p.row0 + lrow;
        f_srow = p.sel_cur[row];
        if (MODE != MODE_EVAL) { f_ib = p.ib[row]; f_r1 = p.r1[row]; f_r2 = p.r2[row]; }
    };
    auto resolve2 = [&](int lrow) {                         // second level + base pointers
        const int row = p.row0 + lrow;
        la.P = (f_srow ? buf1 : buf0) + (int64_t)row * ld;
        if (MODE != MODE_EVAL) {
            // strategies (de_oo.py:186-227): the base row / the two direction rows are either gathered
            // per trial or one vector shared by every trial of the generation
            // The donor-index and selector loads are unconditional -- the index arrays always hold valid rows,
            // also under strategies that do not use them -- because predicating them on the strategy keeps the
            // compiler from scheduling them early (measured: 20 % on narrow rows).
            const uint8_t sb = p.sel_cur[f_ib], sr1 = p.sel_cur[f_r1], sr2 = p.sel_cur[f_r2];
            const double *gB = (sb ? buf1 : buf0) + (int64_t)f_ib * ld;
            const double *g1 = (sr1 ? buf1 : buf0) + (int64_t)f_r1 * ld;
            const double *g2 = (sr2 ? buf1 : buf0) + (int64_t)f_r2 * ld;
            la.B = vec_base ? p.base_vec : gB;
            la.R1 = vec_dir ? p.d1_vec : g1;
            la.R2 = vec_dir ? p.d2_vec : g2;
        }
    };
    auto issue = [&](int s) {                                // lane 0 only
        const int2 lf = p.leaves[la_leaf];
        int ng = lf.y + O::HALO;                             // genes needed, rounded up to 16 bytes
        ng = min((ng + 1) & ~1, ld - lf.x);
        const uint32_t bytes = (uint32_t)ng * 8u;
        const uint32_t full = bar0 + 8u * s, dst = dst0 + (uint32_t)s * TMA_STAGE_BYTES;
        mbar_arrive_expect_tx(full, MODE == MODE_EVAL ? bytes : 4u * bytes);
        bulk_g2s(dst, la.P + lf.x, bytes, full);
        if (MODE != MODE_EVAL) {
            bulk_g2s(dst + TMA_ROW_BYTES, la.B + lf.x, bytes, full);
            bulk_g2s(dst + 2 * TMA_ROW_BYTES, la.R1 + lf.x, bytes, full);
            bulk_g2s(dst + 3 * TMA_ROW_BYTES, la.R2 + lf.x, bytes, full);
        }
    };
    auto advance_la = [&]() {
        if (++la_leaf == nleaves) { la_leaf = 0; la_row += G; }
    };

    // prologue: fill all stages
    for (int s = 0; s < S; ++s) {
        if (la_row < p.nrows) {
            if (la_leaf == 0) { resolve1(la_row); resolve2(la_row); }
            if (lane == 0) issue(s);
            advance_la();
        }
    }

    unsigned nrep = 0;
    int s = 0;
    uint32_t parity = 0;
    for (int lrow = first; lrow < p.nrows; lrow += G) {
        const int row = p.row0 + lrow;
        double *Wrow = nullptr;
        if (MODE != MODE_EVAL) Wrow = const_cast<double *>(p.sel_cur[row] ? buf0 : buf1) + (int64_t)row * ld;
        const double *Uf = nullptr, *Uc = nullptr;
        if (MODE == MODE_REPLAY) {
            Uf = p.Uf + row * (int64_t)p.D;
            Uc = p.Uc + row * (int64_t)p.D;
        }
        const uint32_t row_lo = (uint32_t)row, row_hi = 0u;

        for (int leaf = 0; leaf < nleaves; ++leaf) {
            const uint32_t full = bar0 + 8u * s;
            double *st = my_stages + (size_t)s * (TMA_STAGE_BYTES / 8);
            const int2 lf = p.leaves[leaf];
            const int len = lf.y;                                            // terms of this leaf
            const int glen = len + ((O::HALO && leaf == nleaves - 1) ? (p.Dl - p.nT) : 0);   // genes to write
            const bool la_valid = la_row < p.nrows;
            const bool la_new = la_valid && la_leaf == 0;
            if (la_new) resolve1(la_row);

            mbar_wait(full, parity);
            if (len == 128 && glen == 128) {
                leaf_body<OBJ, MODE, UNI, true, CK>(p, tc, K, st, lf.x, 128, 128, Wrow, Uf, Uc, row_lo, row_hi, lane, nrep,
                                                    la_new, [&]() { resolve2(la_row); });
                leaf_sum<OBJ, true>(p, st, 128, lane, ((int64_t)lrow * p.part_stride + leaf) * NSL);
            } else {
                leaf_body<OBJ, MODE, UNI, false, CK>(p, tc, K, st, lf.x, len, glen, Wrow, Uf, Uc, row_lo, row_hi, lane,
                                                     nrep, la_new, [&]() { resolve2(la_row); });
                leaf_sum<OBJ, false>(p, st, len, lane, ((int64_t)lrow * p.part_stride + leaf) * NSL);
            }
            __syncwarp();                            // the stage is free: refill it
            if (la_valid) {
                if (lane == 0) {
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    issue(s);
                }
                advance_la();
            }
            if (++s == S) { s = 0; parity ^= 1u; }
        }
    }
    if (MODE != MODE_EVAL) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) nrep += __shfl_xor_sync(0xFFFFFFFFu, nrep, o);
        if (lane == 0 && nrep) atomicAdd(p.n_repaired, (unsigned long long)nrep);
    }
}

}  // namespace oode

// =============================================================================================
// Half-warp workers.  Same algorithm as de_trial_tma_kernel, but every HALF-WARP is a worker with
// its own row sequence, stages and barriers; the two workers of a warp walk rows r and r+1 in
// lockstep (same leaf at the same time, so they never diverge).  The per-item bookkeeping
// instructions (row pointers, barrier wait, refill, reduction, loop) are then issued once for two
// items, and each lane carries four gene pairs of a leaf instead of two, which is what narrow rows
// (few leaves per row, e.g. the <= 128-column slices of an 8-way column shard) need.
// =============================================================================================
namespace oode {

template <int NC, int S> struct Tma2Smem {
    static constexpr size_t stages = (size_t)NC * 2 * S * TMA_STAGE_BYTES;
    static constexpr size_t bars = (size_t)NC * 2 * S * sizeof(uint64_t);
    static constexpr size_t total = stages + bars;
};

// KIND 2: exactly 128 terms/genes (no predicate).  KIND 1: 96..128 terms, genes == terms: the first three
// 32-column blocks are full, only the fourth is predicated; a tail of len % 8 terms is added by every lane of
// the slot group in turn (uniform trip count, no divergence).  KIND 0: anything.
template <int OBJ, int MODE, bool UNI, int KIND, bool CK, typename MidFn>
__device__ __forceinline__ void leaf_body2(const GenParams &p, const TrialConsts &tc, const RegKeys &K, double *st, int off,
                                           int len, int glen, double *Wrow, const double *Uf, const double *Uc,
                                           uint32_t row_lo, int hl, unsigned &nrep, bool mid_flag, MidFn mid) {
    using O = Obj<OBJ>;
    constexpr int NSL = O::NS + O::NPR;
    constexpr bool FULL = KIND == 2;
    const int glen2 = (glen + 1) & ~1;
    double t[4][2][NSL];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int lc = 32 * k + 2 * hl;                              // column inside the leaf
        if (FULL || (KIND == 1 && k < 3) || lc < glen2) {
            const bool ok1 = FULL || (KIND == 1 && k < 3) || lc + 1 < glen;
            const double2 x = gene_pair_smem<MODE, UNI>(p, tc, K, st, lc, off + lc, row_lo, 0u, Uf, Uc, ok1, nrep);
            if (MODE != MODE_EVAL) *reinterpret_cast<double2 *>(Wrow + off + lc) = x;
            double xn1 = 0.0;
            if (O::HALO && lc + 1 < len) xn1 = gene_one_smem<MODE, UNI>(p, tc, K, st, lc + 2, off + lc + 2, row_lo, 0u, Uf, Uc);
            double auxa = 0.0, auxb = 0.0;
            if (O::AUX) { const double2 av = *reinterpret_cast<const double2 *>(p.aux + off + lc); auxa = av.x; auxb = av.y; }
            O::template term<CK>(x.x, x.y, auxa, p.col0 + off + lc, t[k][0]);
            O::template term<CK>(x.y, xn1, auxb, p.col0 + off + lc + 1, t[k][1]);
        }
        if (k == 1 && mid_flag) mid();
    }
    __syncwarp();                            // all inputs of the stage have been read
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int lc = 32 * k + 2 * hl;
        if (FULL || (KIND == 1 && k < 3) || lc < glen2) {
#pragma unroll
            for (int q = 0; q < NSL; ++q)
                *reinterpret_cast<double2 *>(st + q * TMA_ROW_DOUBLES + lc) = make_double2(t[k][0][q], t[k][1][q]);
        }
    }
    __syncwarp();
}

template <int OBJ, int KIND>
__device__ __forceinline__ void leaf_sum2(const GenParams &p, const double *st, int len, int lane, int64_t out_index) {
    using O = Obj<OBJ>;
    constexpr int NSL = O::NS + O::NPR;
    const int q = (lane & 15) >> 3, jn = lane & 7;
    if (q < NSL) {
        const double *tq = st + q * TMA_ROW_DOUBLES;
        const unsigned gm = 0xFFu << (lane & 24);
        if (KIND == 2) {
            double r = tq[jn];
#pragma unroll
            for (int k = 1; k < 16; ++k) r = slot_op<OBJ>(q, r, tq[8 * k + jn]);
            r = slot_op<OBJ>(q, r, __shfl_xor_sync(gm, r, 1));
            r = slot_op<OBJ>(q, r, __shfl_xor_sync(gm, r, 2));
            r = slot_op<OBJ>(q, r, __shfl_xor_sync(gm, r, 4));
            put_partial<8>(p, out_index + q, r, jn);
        } else if (KIND == 1) {                      // 12..16 full steps + a tail of len % 8 terms
            const int nfull = len >> 3;
            double r = tq[jn];
#pragma unroll
            for (int k = 1; k < 12; ++k) r = slot_op<OBJ>(q, r, tq[8 * k + jn]);
#pragma unroll
            for (int k = 12; k < 16; ++k)
                if (k < nfull) r = slot_op<OBJ>(q, r, tq[8 * k + jn]);
            r = slot_op<OBJ>(q, r, __shfl_xor_sync(gm, r, 1));
            r = slot_op<OBJ>(q, r, __shfl_xor_sync(gm, r, 2));
            r = slot_op<OBJ>(q, r, __shfl_xor_sync(gm, r, 4));
            for (int i = nfull * 8; i < len; ++i) r = slot_op<OBJ>(q, r, tq[i]);    // same value in all 8 lanes
            put_partial<8>(p, out_index + q, r, jn);
        } else {
            double res;
            const int nfull = len >> 3;
            int i = nfull * 8;
            if (nfull > 0) {
                double r = tq[jn];
                for (int k = 1; k < nfull; ++k) r = slot_op<OBJ>(q, r, tq[8 * k + jn]);
                r = slot_op<OBJ>(q, r, __shfl_xor_sync(gm, r, 1));
                r = slot_op<OBJ>(q, r, __shfl_xor_sync(gm, r, 2));
                r = slot_op<OBJ>(q, r, __shfl_xor_sync(gm, r, 4));
                res = r;
            } else {
                res = (q < O::NS) ? 0.0 : 1.0;     // numpy: n < 8 starts from 0. (1. for a product)
            }
            if (jn == 0) {
                for (; i < len; ++i) res = slot_op<OBJ>(q, res, tq[i]);
                put_partial<1>(p, out_index + q, res, 0);
            }
        }
    }
}

template <int OBJ, int MODE, bool UNI, bool CK, int NC, int S>
__global__ void __launch_bounds__(NC * 32, 1) de_trial_tma2_kernel(const GenParams p) {
    using O = Obj<OBJ>;
    constexpr int NSL = O::NS + O::NPR;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double *stage_base = reinterpret_cast<double *>(smem_raw);
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem_raw + Tma2Smem<NC, S>::stages);
    const int c = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int half = lane >> 4, hl = lane & 15;
    const int w = c * 2 + half;                                  // worker inside the CTA
    if (threadIdx.x == 0) {
        for (int i = 0; i < NC * 2 * S; ++i) mbar_init(smem_u32(bars + i), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    const int nleaves = p.nleaves;
    const int G = (int)gridDim.x * NC * 2;                      // rows advance by G per worker
    const int first = ((int)blockIdx.x * NC + c) * 2 + half;
    const uint32_t bar0 = smem_u32(bars + w * S);
    const uint32_t dst0 = smem_u32(stage_base) + (uint32_t)(w * S) * TMA_STAGE_BYTES;
    double *my_stages = stage_base + (size_t)(w * S) * (TMA_STAGE_BYTES / 8);
    RegKeys K;
    if (MODE == MODE_PHILOX) K.load(p.keys, (uint32_t)*p.n_repaired_zero);
    TrialConsts tc;
    int ld;
    {
        const long long z = (long long)*p.n_repaired_zero;
        tc.Fa = __longlong_as_double(__double_as_longlong(p.const_F ? 0.0 : p.F) ^ z);
        tc.Fb = __longlong_as_double(__double_as_longlong(p.const_F ? p.F : 0.0) ^ z);
        tc.Cr = __longlong_as_double(__double_as_longlong(p.Cr) ^ z);
        tc.lo = __longlong_as_double(__double_as_longlong(p.lo_s) ^ z);
        tc.hi = __longlong_as_double(__double_as_longlong(p.hi_s) ^ z);
        tc.gen = p.gen ^ (uint32_t)z;
        tc.cf = (p.const_F ^ (int)z) != 0;
        tc.crT = p.cr_thresh ^ (uint32_t)z;
        tc.crAll = (p.cr_all ^ (int)z) != 0;
        ld = p.ld ^ (int)z;
    }
    const double *const buf0 = p.buf0, *const buf1 = p.buf1;
    const bool vec_base = p.base_vec != nullptr, vec_dir = p.d1_vec != nullptr;

    int la_row = first, la_leaf = 0;
    RowSrc la;
    la.P = la.B = la.R1 = la.R2 = nullptr;
    int f_ib = 0, f_r1 = 0, f_r2 = 0;
    uint8_t f_srow = 0;

    auto resolve1 = [&](int lrow) {
        const int row = p.row0 + lrow;
        f_srow = p.sel_cur[row];
        if (MODE != MODE_EVAL) { f_ib = p.ib[row]; f_r1 = p.r1[row]; f_r2 = p.r2[row]; }
    };
    auto resolve2 = [&](int lrow) {
        const int row = p.row0 + lrow;
        la.P = (f_srow ? buf1 : buf0) + (int64_t)row * ld;
        if (MODE != MODE_EVAL) {
            // strategies (de_oo.py:186-227): the base row / the two direction rows are either gathered
            // per trial or one vector shared by every trial of the generation
            // The donor-index and selector loads are unconditional -- the index arrays always hold valid rows,
            // also under strategies that do not use them -- because predicating them on the strategy keeps the
            // compiler from scheduling them early (measured: 20 % on narrow rows).
            const uint8_t sb = p.sel_cur[f_ib], sr1 = p.sel_cur[f_r1], sr2 = p.sel_cur[f_r2];
            const double *gB = (sb ? buf1 : buf0) + (int64_t)f_ib * ld;
            const double *g1 = (sr1 ? buf1 : buf0) + (int64_t)f_r1 * ld;
            const double *g2 = (sr2 ? buf1 : buf0) + (int64_t)f_r2 * ld;
            la.B = vec_base ? p.base_vec : gB;
            la.R1 = vec_dir ? p.d1_vec : g1;
            la.R2 = vec_dir ? p.d2_vec : g2;
        }
    };
    auto issue = [&](int s) {                                // one lane per worker
        const int2 lf = p.leaves[la_leaf];
        int ng = lf.y + O::HALO;
        ng = min((ng + 1) & ~1, ld - lf.x);
        const uint32_t bytes = (uint32_t)ng * 8u;
        const uint32_t full = bar0 + 8u * s, dst = dst0 + (uint32_t)s * TMA_STAGE_BYTES;
        mbar_arrive_expect_tx(full, MODE == MODE_EVAL ? bytes : 4u * bytes);
        bulk_g2s(dst, la.P + lf.x, bytes, full);
        if (MODE != MODE_EVAL) {
            bulk_g2s(dst + TMA_ROW_BYTES, la.B + lf.x, bytes, full);
            bulk_g2s(dst + 2 * TMA_ROW_BYTES, la.R1 + lf.x, bytes, full);
            bulk_g2s(dst + 3 * TMA_ROW_BYTES, la.R2 + lf.x, bytes, full);
        }
    };
    auto advance_la = [&]() {
        if (++la_leaf == nleaves) { la_leaf = 0; la_row += G; }
    };

    for (int s = 0; s < S; ++s) {
        const bool v = la_row < p.nrows;
        if (v && la_leaf == 0) { resolve1(la_row); resolve2(la_row); }
        if (v && hl == 0) issue(s);
        if (v) advance_la();
    }

    unsigned nrep = 0;
    int s = 0;
    uint32_t parity = 0;
    for (int lrow = first; __any_sync(0xFFFFFFFFu, lrow < p.nrows); lrow += G) {
        const bool valid = lrow < p.nrows;
        const int row = p.row0 + (valid ? lrow : 0);
        double *Wrow = nullptr;
        if (MODE != MODE_EVAL) Wrow = const_cast<double *>(p.sel_cur[row] ? buf0 : buf1) + (int64_t)row * ld;
        const double *Uf = nullptr, *Uc = nullptr;
        if (MODE == MODE_REPLAY) {
            Uf = p.Uf + row * (int64_t)p.D;
            Uc = p.Uc + row * (int64_t)p.D;
        }
        const uint32_t row_lo = (uint32_t)row;

        for (int leaf = 0; leaf < nleaves; ++leaf) {
            const uint32_t full = bar0 + 8u * s;
            double *st = my_stages + (size_t)s * (TMA_STAGE_BYTES / 8);
            const int2 lf = p.leaves[leaf];
            const int len = valid ? lf.y : 0;
            const int glen = valid ? len + ((O::HALO && leaf == nleaves - 1) ? (p.Dl - p.nT) : 0) : 0;
            const bool la_valid = la_row < p.nrows;
            const bool la_new = la_valid && la_leaf == 0;
            if (la_new) resolve1(la_row);

            if (valid) mbar_wait(full, parity);
            const int64_t out_index = ((int64_t)(valid ? lrow : 0) * p.part_stride + leaf) * NSL;
            const bool both = __all_sync(0xFFFFFFFFu, valid);
            if (lf.y == 128 && !O::HALO && both) {
                leaf_body2<OBJ, MODE, UNI, 2, CK>(p, tc, K, st, lf.x, 128, 128, Wrow, Uf, Uc, row_lo, hl, nrep, la_new,
                                                  [&]() { resolve2(la_row); });
                leaf_sum2<OBJ, 2>(p, st, 128, lane, out_index);
            } else if (lf.y >= 96 && !O::HALO && both) {
                leaf_body2<OBJ, MODE, UNI, 1, CK>(p, tc, K, st, lf.x, len, len, Wrow, Uf, Uc, row_lo, hl, nrep, la_new,
                                                  [&]() { resolve2(la_row); });
                leaf_sum2<OBJ, 1>(p, st, len, lane, out_index);
            } else {
                leaf_body2<OBJ, MODE, UNI, 0, CK>(p, tc, K, st, lf.x, len, glen, Wrow, Uf, Uc, row_lo, hl, nrep, la_new,
                                                  [&]() { resolve2(la_row); });
                if (valid) leaf_sum2<OBJ, 0>(p, st, len, lane, out_index);
            }
            __syncwarp();                            // the stages are free: refill them
            if (la_valid) {
                if (hl == 0) {
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                    issue(s);
                }
                advance_la();
            }
            if (++s == S) { s = 0; parity ^= 1u; }
        }
    }
    if (MODE != MODE_EVAL) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) nrep += __shfl_xor_sync(0xFFFFFFFFu, nrep, o);
        if (lane == 0 && nrep) atomicAdd(p.n_repaired, (unsigned long long)nrep);
    }
}

}  // namespace oode
