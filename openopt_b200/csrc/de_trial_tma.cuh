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
//   every warp       owns S stages.  Per item: wait full[s]; lane L makes the genes of columns
//                    2L,2L+1 and 64+2L,64+2L+1 of the leaf (one Philox call per pair), writes the
//                    trial row as 16-byte vectors (512 contiguous bytes per warp instruction), puts
//                    the objective terms in shared memory; then lane 0 refills the drained stage
//                    with the warp's item S positions ahead (row pointers were resolved by loads
//                    issued before / in the middle of the compute phase), and 8 lanes per sum slot
//                    walk the terms in numpy's r[j] chain order (bit-identical pairwise sum).
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

// genes of one column pair from staged inputs (smem), lc = column inside the leaf (even)
template <int MODE, bool UNI>
__device__ __forceinline__ double2 gene_pair_smem(const GenParams &p, const double *st, int lc, int gcol_local,
                                                  uint32_t row_lo, uint32_t row_hi, const double *Uf, const double *Uc,
                                                  bool ok0, bool ok1, unsigned &nrep) {
    const double2 xo = *reinterpret_cast<const double2 *>(st + lc);
    if (MODE == MODE_EVAL) return xo;
    const double2 xb = *reinterpret_cast<const double2 *>(st + TMA_ROW_DOUBLES + lc);
    const double2 x1 = *reinterpret_cast<const double2 *>(st + 2 * TMA_ROW_DOUBLES + lc);
    const double2 x2 = *reinterpret_cast<const double2 *>(st + 3 * TMA_ROW_DOUBLES + lc);
    const int col = p.col0 + gcol_local;
    double uf0, uc0, uf1, uc1;
    if (MODE == MODE_PHILOX) {
        uint32_t w0 = (uint32_t)col >> 1, w1 = row_lo, w2 = row_hi | (TAG_ELEM << 24), w3 = p.gen;
        philox4x32_10(w0, w1, w2, w3, p.keys);
        uf0 = u32_to_unit(w0); uc0 = u32_to_unit(w1);
        uf1 = u32_to_unit(w2); uc1 = u32_to_unit(w3);
    } else {
        uf0 = ok0 ? Uf[col] : 0.0; uc0 = ok0 ? Uc[col] : 1.0;
        uf1 = ok1 ? Uf[col + 1] : 0.0; uc1 = ok1 ? Uc[col + 1] : 1.0;
    }
    double lo0, lo1, hi0, hi1;
    if (UNI) { lo0 = lo1 = p.lo_s; hi0 = hi1 = p.hi_s; }
    else {
        const double2 l = *reinterpret_cast<const double2 *>(p.lb + gcol_local);
        const double2 h = *reinterpret_cast<const double2 *>(p.ub + gcol_local);
        lo0 = l.x; lo1 = l.y; hi0 = h.x; hi1 = h.y;
    }
    unsigned r0 = 0, r1 = 0;
    double2 r;
    r.x = finish_gene(xo.x, xb.x, x1.x, x2.x, uf0, uc0, p.F, p.Cr, lo0, hi0, r0);
    r.y = finish_gene(xo.y, xb.y, x1.y, x2.y, uf1, uc1, p.F, p.Cr, lo1, hi1, r1);
    nrep += (ok0 ? r0 : 0u) + (ok1 ? r1 : 0u);
    return r;
}

// one gene from staged inputs (halo neighbour)
template <int MODE, bool UNI>
__device__ __forceinline__ double gene_one_smem(const GenParams &p, const double *st, int lc, int gcol_local,
                                                uint32_t row_lo, uint32_t row_hi, const double *Uf, const double *Uc) {
    if (MODE == MODE_EVAL) return st[lc];
    const int col = p.col0 + gcol_local;
    double uf, uc;
    if (MODE == MODE_PHILOX) {
        uint32_t w0 = (uint32_t)col >> 1, w1 = row_lo, w2 = row_hi | (TAG_ELEM << 24), w3 = p.gen;
        philox4x32_10(w0, w1, w2, w3, p.keys);
        uf = u32_to_unit((col & 1) ? w2 : w0);
        uc = u32_to_unit((col & 1) ? w3 : w1);
    } else { uf = Uf[col]; uc = Uc[col]; }
    const double lo = UNI ? p.lo_s : p.lb[gcol_local], hi = UNI ? p.hi_s : p.ub[gcol_local];
    unsigned dummy = 0;
    return finish_gene(st[lc], st[TMA_ROW_DOUBLES + lc], st[2 * TMA_ROW_DOUBLES + lc], st[3 * TMA_ROW_DOUBLES + lc], uf,
                       uc, p.F, p.Cr, lo, hi, dummy);
}

template <int NC, int S> struct TmaSmem {
    static constexpr size_t stages = (size_t)NC * S * TMA_STAGE_BYTES;
    static constexpr size_t terms = (size_t)NC * 2 * 128 * sizeof(double);
    static constexpr size_t bars = (size_t)NC * S * sizeof(uint64_t);
    static constexpr size_t total = stages + terms + bars;
};

// Row pointers of one work item, resolved in two dependent steps so that each step's global
// loads can be issued early and consumed a compute phase later.
struct ItemFetch {
    int64_t row;
    int off, bytes;
    int ib, r1, r2;
    uint8_t s_row, s_b, s_1, s_2;
    bool valid;
};

template <int OBJ, int MODE>
__device__ __forceinline__ void fetch_level1(const GenParams &p, int64_t nitems, int64_t item, ItemFetch &f) {
    f.valid = item < nitems;
    if (!f.valid) return;
    int64_t lrow;
    int leaf;
    split_item(p, nitems, item, lrow, leaf);
    f.row = p.row0 + lrow;
    const int2 lf = p.leaves[leaf];
    // genes needed: the leaf's terms, one more with a halo, rounded up to a 16-byte multiple
    int ng = lf.y + Obj<OBJ>::HALO;
    ng = min((ng + 1) & ~1, (int)(p.ld - lf.x));
    f.off = lf.x;
    f.bytes = ng * 8;
    f.s_row = p.sel_cur[f.row];
    if (MODE != MODE_EVAL) { f.ib = p.ib[f.row]; f.r1 = p.r1[f.row]; f.r2 = p.r2[f.row]; }
}

template <int MODE> __device__ __forceinline__ void fetch_level2(const GenParams &p, ItemFetch &f) {
    if (!f.valid || MODE == MODE_EVAL) return;
    f.s_b = p.sel_cur[f.ib];
    f.s_1 = p.sel_cur[f.r1];
    f.s_2 = p.sel_cur[f.r2];
}

template <int MODE>
__device__ __forceinline__ void issue_item(const GenParams &p, const ItemFetch &f, uint32_t dst, uint32_t full) {
    const uint32_t bytes = (uint32_t)f.bytes;
    mbar_arrive_expect_tx(full, MODE == MODE_EVAL ? bytes : 4u * bytes);
    bulk_g2s(dst, (f.s_row ? p.buf1 : p.buf0) + f.row * p.ld + f.off, bytes, full);
    if (MODE != MODE_EVAL) {
        bulk_g2s(dst + TMA_ROW_BYTES, (f.s_b ? p.buf1 : p.buf0) + (int64_t)f.ib * p.ld + f.off, bytes, full);
        bulk_g2s(dst + 2 * TMA_ROW_BYTES, (f.s_1 ? p.buf1 : p.buf0) + (int64_t)f.r1 * p.ld + f.off, bytes, full);
        bulk_g2s(dst + 3 * TMA_ROW_BYTES, (f.s_2 ? p.buf1 : p.buf0) + (int64_t)f.r2 * p.ld + f.off, bytes, full);
    }
}

// Every warp is a self-serving consumer: it owns S stages; after finishing an item it refills the
// stage it just drained with its item S positions ahead (lane 0 issues the four bulk copies).
template <int OBJ, int MODE, bool UNI, int NC, int S>
__global__ void __launch_bounds__(NC * 32, 1) de_trial_tma_kernel(const GenParams p) {
    using O = Obj<OBJ>;
    constexpr int NSL = O::NS + O::NPR;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    double *stage_base = reinterpret_cast<double *>(smem_raw);
    double *terms_base = reinterpret_cast<double *>(smem_raw + TmaSmem<NC, S>::stages);
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem_raw + TmaSmem<NC, S>::stages + TmaSmem<NC, S>::terms);
    const int c = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int i = 0; i < NC * S; ++i) mbar_init(smem_u32(bars + i), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    const int64_t nitems = p.nrows * p.nleaves;
    const int64_t G = (int64_t)gridDim.x * NC;
    const int64_t first = (int64_t)blockIdx.x * NC + c;
    double *terms = terms_base + (size_t)c * 2 * 128;
    const uint32_t bar0 = smem_u32(bars + c * S);
    const uint32_t dst0 = smem_u32(stage_base) + (uint32_t)(c * S) * TMA_STAGE_BYTES;

    // prologue: fill all stages
    for (int s = 0; s < S; ++s) {
        ItemFetch f;
        fetch_level1<OBJ, MODE>(p, nitems, first + (int64_t)s * G, f);
        fetch_level2<MODE>(p, f);
        if (lane == 0 && f.valid) issue_item<MODE>(p, f, dst0 + (uint32_t)s * TMA_STAGE_BYTES, bar0 + 8u * s);
    }

    unsigned nrep = 0;
    int j = 0;
    for (int64_t item = first; item < nitems; item += G, ++j) {
        const int s = j % S;
        const uint32_t parity = (uint32_t)(j / S) & 1u;
        const uint32_t full = bar0 + 8u * s;
        const double *st = stage_base + (size_t)(c * S + s) * (TMA_STAGE_BYTES / 8);
        int64_t lrow;
        int leaf;
        split_item(p, nitems, item, lrow, leaf);
        const int64_t row = p.row0 + lrow;
        const int2 lf = p.leaves[leaf];
        const int len = lf.y;                                            // terms of this leaf
        const bool last = (leaf == p.nleaves - 1);
        const int glen = len + ((O::HALO && last) ? (p.Dl - p.nT) : 0);  // genes this item must write
        double *W = nullptr;
        if (MODE != MODE_EVAL) W = (p.sel_cur[row] ? p.buf0 : p.buf1) + row * p.ld + lf.x;
        const double *Uf = nullptr, *Uc = nullptr;
        if (MODE == MODE_REPLAY) {
            Uf = p.Uf + row * (int64_t)p.D;
            Uc = p.Uc + row * (int64_t)p.D;
        }
        const uint32_t row_lo = (uint32_t)row, row_hi = (uint32_t)((uint64_t)row >> 32);

        // the item that will take over this stage: first-level loads now, second-level after half 0
        ItemFetch nf;
        fetch_level1<OBJ, MODE>(p, nitems, item + (int64_t)S * G, nf);

        mbar_wait(full, parity);
#pragma unroll
        for (int half = 0; half < 2; ++half) {
            const int lc = 64 * half + 2 * lane;                         // column inside the leaf
            if (lc < glen) {
                const bool ok1 = lc + 1 < glen;
                const double2 x = gene_pair_smem<MODE, UNI>(p, st, lc, lf.x + lc, row_lo, row_hi, Uf, Uc, true, ok1, nrep);
                if (MODE != MODE_EVAL) {
                    if (ok1 || lf.x + lc + 1 < (int)p.ld) *reinterpret_cast<double2 *>(W + lc) = x;   // pad column is scratch
                    else W[lc] = x.x;
                }
                double xn1 = 0.0;
                if (O::HALO && lc + 1 < len) xn1 = gene_one_smem<MODE, UNI>(p, st, lc + 2, lf.x + lc + 2, row_lo, row_hi, Uf, Uc);
                double auxa = 0.0, auxb = 0.0;
                if (O::AUX) { const double2 av = *reinterpret_cast<const double2 *>(p.aux + lf.x + lc); auxa = av.x; auxb = av.y; }
                double t0[NSL], t1[NSL];
                O::term(x.x, x.y, auxa, t0);
                O::term(x.y, xn1, auxb, t1);
#pragma unroll
                for (int q = 0; q < NSL; ++q) {
                    if (lc < len) terms[q * 128 + lc] = t0[q];
                    if (lc + 1 < len) terms[q * 128 + lc + 1] = t1[q];
                }
            }
            if (half == 0) fetch_level2<MODE>(p, nf);
        }
        __syncwarp();                                // every lane has finished reading the stage
        if (lane == 0 && nf.valid) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            issue_item<MODE>(p, nf, dst0 + (uint32_t)s * TMA_STAGE_BYTES, full);
        }
        // ---- numpy's pairwise leaf sum over terms[q][0..len): lanes 8q..8q+7 are the r[j] chains ----
        const int q = lane >> 3, jn = lane & 7;
        if (q < NSL) {
            const double *t = terms + q * 128;
            double res;
            const int nfull = len >> 3;
            int i = nfull * 8;
            if (nfull > 0) {
                double r = t[jn];
                for (int k = 1; k < nfull; ++k) r = slot_op<OBJ>(q, r, t[8 * k + jn]);
                const unsigned gm = 0xFFu << (lane & 24);
                r = slot_op<OBJ>(q, r, __shfl_xor_sync(gm, r, 1));
                r = slot_op<OBJ>(q, r, __shfl_xor_sync(gm, r, 2));
                r = slot_op<OBJ>(q, r, __shfl_xor_sync(gm, r, 4));
                res = r;
            } else {
                res = (q < O::NS) ? 0.0 : 1.0;         // numpy: n < 8 starts from 0. (1. for a product)
            }
            if (jn == 0) {
                for (; i < len; ++i) res = slot_op<OBJ>(q, res, t[i]);
                p.partials[(lrow * p.nleaves + leaf) * NSL + q] = res;
            }
        }
        __syncwarp();                               // terms are rewritten by the next item
    }
    if (MODE != MODE_EVAL) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) nrep += __shfl_xor_sync(0xFFFFFFFFu, nrep, o);
        if (lane == 0 && nrep) atomicAdd(p.n_repaired, (unsigned long long)nrep);
    }
}

}  // namespace oode
