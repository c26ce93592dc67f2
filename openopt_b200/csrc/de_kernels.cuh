// The generation kernels.
//
//   K0 de_init_kernel      pop = U*(ub-lb)+lb ; pop[0] = x0                  de_oo.py:147-150
//   K2a de_trial_kernel    mutation + crossover + bound repair + objective terms + leaf sums
//                                                                            de_oo.py:186-254, 292-297
//   K2b de_select_kernel   combine leaf sums -> f, NaN->inf, greedy selection, block-then-grid
//                          first-occurrence argmin, Best update, next generation's donor indices
//                                                                            de_oo.py:256-286, 301-305
//
// Data layout in HBM.  The population lives in TWO row-major slot arrays buf[0], buf[1] of
// [NP][ld] float64 plus one selector byte per row: the current version of row r is
// buf[sel[r]][r].  A trial for row i is written into the row's SPARE slot buf[sel[i]^1][i]; if the
// trial is accepted the selector flips, if not nothing else happens.  This is the reference's
// two-array scheme (old_pop is read-only during a generation, de_oo.py:180,277) without the
// full-matrix blend: per trial 4 row reads + 1 row write = 40*D bytes, nothing is copied.
//
// Work decomposition of K2a.  numpy's .sum() over a contiguous row is a pairwise tree whose
// leaves are blocks of <=128 consecutive terms, each summed with 8 interleaved accumulators
// (oracle/de_oracle.py::pairwise_sum).  A work item is one (row, leaf); it is owned by a group of
// 8 lanes, lane j accumulating terms off+j, off+j+8, ... -- exactly numpy's r[j] chains -- so the
// row reduction is bit-identical to numpy's with no staging buffer: each step the group reads
// four 64-byte row segments (parent and three donors; two full 32-byte sectors each) and writes
// one.  The 8 chain sums are combined with three xor-shuffles in numpy's
// ((r0+r1)+(r2+r3))+((r4+r5)+(r6+r7)) order, the tail terms added sequentially.  Leaf sums go to
// `partials`; K2b folds them in tree order (ReductionPlan::prog).
#pragma once
#include "de_device.cuh"

namespace oode {

enum { MODE_EVAL = 0, MODE_PHILOX = 1, MODE_REPLAY = 2 };

struct GenParams {
    // population slots and selectors
    double *buf0, *buf1;
    const uint8_t *sel_cur;
    int64_t ld;              // leading dimension (doubles) of the slot arrays
    // per-row donor indices of this generation (global row numbers)
    const int32_t *ib, *r1, *r2;
    // per-column data (local column slice)
    const double *lb, *ub, *aux;
    // replay draws, [NP][D] global row-major
    const double *Uf, *Uc;
    // leaf sums out: [(lrow*nleaves + leaf)*NSLOT + s]
    double *partials;
    const int2 *leaves;      // (local column offset, length) of every leaf of this rank
    int nleaves;
    int nT;                  // number of terms in this rank's slice (Dl, or Dl-1 with a halo)
    int Dl;                  // local number of columns
    int D;                   // global number of columns
    int col0;                // global index of local column 0
    int64_t row0;            // first row handled by this launch
    int64_t nrows;           // number of rows handled by this launch
    double F, Cr;
    uint32_t gen;
    PhiloxKeys keys;
    unsigned long long *n_repaired;
};

template <int OBJ, int MODE> struct GeneCtx {
    const double *P, *B, *R1, *R2;
    double *W;
    const double *Uf, *Uc;
    uint32_t row_lo, row_hi;
};

template <int OBJ, int MODE>
__device__ __forceinline__ double make_gene(const GeneCtx<OBJ, MODE> &c, const GenParams &p, int lc, unsigned &nrep) {
    if (MODE == MODE_EVAL) return c.P[lc];
    const double xo = c.P[lc], xb = c.B[lc], x1 = c.R1[lc], x2 = c.R2[lc];
    double uf, uc;
    if (MODE == MODE_PHILOX) {
        uint32_t w0 = (uint32_t)(p.col0 + lc), w1 = c.row_lo, w2 = c.row_hi | (TAG_ELEM << 24), w3 = p.gen;
        philox4x32_10(w0, w1, w2, w3, p.keys);
        uf = u52(w0, w1);
        uc = u52(w2, w3);
    } else {
        uf = c.Uf[p.col0 + lc];
        uc = c.Uc[p.col0 + lc];
    }
    // de_oo.py:236-250: Ft = U*F; mutant = beta + Ft*(x_r2 - x_r1); parent gene kept iff Uc > Cr
    const double v = (uc > p.Cr) ? xo : dadd(xb, dmul(dmul(uf, p.F), dsub(x2, x1)));
    return repair_gene(v, p.lb[lc], p.ub[lc], nrep);
}

template <int OBJ, int MODE>
__global__ void __launch_bounds__(256) de_trial_kernel(const GenParams p) {
    using O = Obj<OBJ>;
    constexpr int NSL = O::NS + O::NPR;
    const int lane8 = threadIdx.x & 7;
    const int gbase = threadIdx.x & 24;
    const unsigned gmask = 0xFFu << gbase;
    const int64_t ngroups = ((int64_t)gridDim.x * blockDim.x) >> 3;
    const int64_t nitems = p.nrows * p.nleaves;
    unsigned nrep = 0, nrep_dummy = 0;

    for (int64_t item = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 3; item < nitems; item += ngroups) {
        int64_t lrow;
        int leaf;
        if (p.nleaves == 1) { lrow = item; leaf = 0; }
        else { lrow = item / p.nleaves; leaf = (int)(item - lrow * p.nleaves); }
        const int64_t row = p.row0 + lrow;
        const int2 lf = p.leaves[leaf];

        GeneCtx<OBJ, MODE> c;
        const uint8_t s = p.sel_cur[row];
        c.P = (s ? p.buf1 : p.buf0) + row * p.ld;
        c.W = (s ? p.buf0 : p.buf1) + row * p.ld;
        if (MODE != MODE_EVAL) {
            const int64_t b = p.ib[row], a1 = p.r1[row], a2 = p.r2[row];
            c.B = (p.sel_cur[b] ? p.buf1 : p.buf0) + b * p.ld;
            c.R1 = (p.sel_cur[a1] ? p.buf1 : p.buf0) + a1 * p.ld;
            c.R2 = (p.sel_cur[a2] ? p.buf1 : p.buf0) + a2 * p.ld;
        }
        if (MODE == MODE_REPLAY) {
            c.Uf = p.Uf + row * (int64_t)p.D;
            c.Uc = p.Uc + row * (int64_t)p.D;
        }
        c.row_lo = (uint32_t)row;
        c.row_hi = (uint32_t)((uint64_t)row >> 32);

        double acc[NSL];
        const int nfull = lf.y >> 3;
#pragma unroll 2
        for (int k = 0; k < nfull; ++k) {
            const int lc = lf.x + 8 * k + lane8;
            const double x = make_gene<OBJ, MODE>(c, p, lc, nrep);
            if (MODE != MODE_EVAL) c.W[lc] = x;
            const double xn = O::HALO ? make_gene<OBJ, MODE>(c, p, lc + 1, nrep_dummy) : 0.0;
            double t[NSL];
            O::term(x, xn, p.col0 + lc, p.aux, t);
#pragma unroll
            for (int q = 0; q < NSL; ++q) acc[q] = (k == 0) ? t[q] : slot_op<OBJ>(q, acc[q], t[q]);
        }
        if (nfull > 0) {
#pragma unroll
            for (int q = 0; q < NSL; ++q) {
                acc[q] = slot_op<OBJ>(q, acc[q], __shfl_xor_sync(gmask, acc[q], 1));
                acc[q] = slot_op<OBJ>(q, acc[q], __shfl_xor_sync(gmask, acc[q], 2));
                acc[q] = slot_op<OBJ>(q, acc[q], __shfl_xor_sync(gmask, acc[q], 4));
            }
        } else {
#pragma unroll
            for (int q = 0; q < NSL; ++q) acc[q] = slot_id<OBJ>(q);
        }
        const int tail = lf.y & 7;
        if (tail) {
            double t[NSL];
#pragma unroll
            for (int q = 0; q < NSL; ++q) t[q] = slot_id<OBJ>(q);
            if (lane8 < tail) {
                const int lc = lf.x + 8 * nfull + lane8;
                const double x = make_gene<OBJ, MODE>(c, p, lc, nrep);
                if (MODE != MODE_EVAL) c.W[lc] = x;
                const double xn = O::HALO ? make_gene<OBJ, MODE>(c, p, lc + 1, nrep_dummy) : 0.0;
                O::term(x, xn, p.col0 + lc, p.aux, t);
            }
            for (int j = 0; j < tail; ++j) {
#pragma unroll
                for (int q = 0; q < NSL; ++q) acc[q] = slot_op<OBJ>(q, acc[q], __shfl_sync(gmask, t[q], gbase + j));
            }
        }
        // columns that carry no term of their own (Rosenbrock's last column)
        if (MODE != MODE_EVAL && O::HALO && leaf == p.nleaves - 1) {
            const int lc = p.nT + lane8;
            if (lc < p.Dl) c.W[lc] = make_gene<OBJ, MODE>(c, p, lc, nrep);
        }
        if (lane8 == 0) {
#pragma unroll
            for (int q = 0; q < NSL; ++q) p.partials[(lrow * p.nleaves + leaf) * NSL + q] = acc[q];
        }
    }
    if (MODE != MODE_EVAL) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) nrep += __shfl_xor_sync(0xFFFFFFFFu, nrep, o);
        if ((threadIdx.x & 31) == 0 && nrep) atomicAdd(p.n_repaired, (unsigned long long)nrep);
    }
}

// ---------------------------------------------------------------------------------------------
// Tree fold of the leaf sums.  prog is a post-order program: v >= 0 pushes leaf v, -1 combines
// the two topmost entries (left op right), mirroring pairwise_sum's recursion.
// ---------------------------------------------------------------------------------------------
template <int OBJ>
__device__ __forceinline__ double fold_slot(const double *part, int nsl, int q, const int *prog, int nprog) {
    if (nprog == 1) return part[q];
    double st[40];
    int sp = 0;
    for (int i = 0; i < nprog; ++i) {
        const int v = prog[i];
        if (v >= 0) st[sp++] = part[v * nsl + q];
        else { --sp; st[sp - 1] = slot_op<OBJ>(q, st[sp - 1], st[sp]); }
    }
    return st[0];
}

struct SelParams {
    const double *partials;
    const int *prog;
    int nprog, nleaves;
    int D;                   // global D (objective finalisation)
    int Dl;                  // local columns (Best.x copy)
    int invert;
    int init;                // 1: first evaluation of the initial population
    int do_indices;          // 1: draw the next generation's donor indices (Philox mode)
    int64_t NP;
    int64_t row0, nrows;     // rows this launch handles
    double *vals, *tvals;
    const uint8_t *sel_cur;
    uint8_t *sel_next;
    uint8_t *accept;
    // argmin / counters scratch
    double *blk_f;
    int64_t *blk_i;
    unsigned long long *blk_acc;
    unsigned int *ticket;
    unsigned long long *n_repaired;
    // Best state and history
    double *best_f;
    double *best_x;          // [Dl]
    oode_gen_record *rec;    // device record for this generation
    double *bestx_slot;      // history slot for this generation, [Dl]
    const double *buf0, *buf1;
    int64_t ld;
    int32_t generation;
    // next generation's donor indices
    int32_t *ib, *r1, *r2;
    uint32_t next_gen;
    PhiloxKeys keys;
};

__device__ __forceinline__ bool better_fi(double fa, int64_t ia, double fb, int64_t ib) {
    return fa < fb || (fa == fb && ia < ib);      // first-occurrence argmin (numpy argmin)
}

template <int OBJ>
__global__ void __launch_bounds__(256) de_select_kernel(const SelParams p) {
    using O = Obj<OBJ>;
    constexpr int NSL = O::NS + O::NPR;
    const int64_t lrow = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t row = p.row0 + lrow;
    double f = __longlong_as_double(0x7FF0000000000000LL);
    int64_t fi = INT64_MAX;
    unsigned acc = 0;
    if (lrow < p.nrows) {
        const double *part = p.partials + lrow * (int64_t)p.nleaves * NSL;
        double s[NSL];
#pragma unroll
        for (int q = 0; q < NSL; ++q) s[q] = fold_slot<OBJ>(part, NSL, q, p.prog, p.nprog);
        f = O::finalize(s, p.D);
        if (p.invert) f = -f;                                     // kernel/nonLinFuncs.py:297-298
        if (f != f) f = __longlong_as_double(0x7FF0000000000000LL); // de_oo.py:301
        fi = row;
        p.tvals[row] = f;
        if (p.init) {
            p.vals[row] = f;
        } else {
            // de_oo.py:260-282 with constr_vals == 0: keep the parent iff old_f < trial_f; vals are
            // blended arithmetically like the reference so inf*0 = NaN appears where it does there.
            const double old = p.vals[row];
            const bool keep = old < f;
            const double k = keep ? 1.0 : 0.0;
            p.vals[row] = dadd(dmul(old, k), dmul(f, dsub(1.0, k)));
            const uint8_t sc = p.sel_cur[row];
            p.sel_next[row] = keep ? sc : (uint8_t)(sc ^ 1);
            p.accept[row] = keep ? 0 : 1;
            acc = keep ? 0u : 1u;
        }
        if (p.do_indices) {
            uint32_t a0 = 0, a1 = (uint32_t)row, a2 = (uint32_t)((uint64_t)row >> 32) | (TAG_INDEX << 24), a3 = p.next_gen;
            uint32_t b0 = 1, b1 = a1, b2 = a2, b3 = a3;
            philox4x32_10(a0, a1, a2, a3, p.keys);
            philox4x32_10(b0, b1, b2, b3, p.keys);
            p.ib[row] = (int32_t)mulhi64_index(a0, a1, (uint64_t)p.NP);
            p.r1[row] = (int32_t)mulhi64_index(a2, a3, (uint64_t)p.NP);
            p.r2[row] = (int32_t)mulhi64_index(b0, b1, (uint64_t)p.NP);
        }
    }
    // ---- block argmin + accepted count ----
    __shared__ double sh_f[8];
    __shared__ int64_t sh_i[8];
    __shared__ unsigned sh_a[8];
    __shared__ bool sh_last;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const double of = __shfl_xor_sync(0xFFFFFFFFu, f, o);
        const int64_t oi = __shfl_xor_sync(0xFFFFFFFFu, fi, o);
        if (better_fi(of, oi, f, fi)) { f = of; fi = oi; }
        acc += __shfl_xor_sync(0xFFFFFFFFu, acc, o);
    }
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) { sh_f[warp] = f; sh_i[warp] = fi; sh_a[warp] = acc; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < 8; ++w) {
            if (better_fi(sh_f[w], sh_i[w], f, fi)) { f = sh_f[w]; fi = sh_i[w]; }
            acc += sh_a[w];
        }
        p.blk_f[blockIdx.x] = f;
        p.blk_i[blockIdx.x] = fi;
        p.blk_acc[blockIdx.x] = acc;
        __threadfence();
        const unsigned t = atomicAdd(p.ticket, 1u);
        sh_last = (t == gridDim.x - 1);
    }
    __syncthreads();
    if (!sh_last) return;
    // ---- grid stage: the last block folds the per-block results ----
    __threadfence();
    f = __longlong_as_double(0x7FF0000000000000LL);
    fi = INT64_MAX;
    unsigned long long tot = 0;
    for (unsigned b = threadIdx.x; b < gridDim.x; b += blockDim.x) {
        const double bf = ((volatile double *)p.blk_f)[b];
        const int64_t bi = ((volatile int64_t *)p.blk_i)[b];
        if (better_fi(bf, bi, f, fi)) { f = bf; fi = bi; }
        tot += ((volatile unsigned long long *)p.blk_acc)[b];
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const double of = __shfl_xor_sync(0xFFFFFFFFu, f, o);
        const int64_t oi = __shfl_xor_sync(0xFFFFFFFFu, fi, o);
        if (better_fi(of, oi, f, fi)) { f = of; fi = oi; }
        tot += __shfl_xor_sync(0xFFFFFFFFu, tot, o);
    }
    __shared__ unsigned long long sh_t[8];
    __shared__ int sh_changed;
    __syncthreads();
    if (lane == 0) { sh_f[warp] = f; sh_i[warp] = fi; sh_t[warp] = tot; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < 8; ++w) {
            if (better_fi(sh_f[w], sh_i[w], f, fi)) { f = sh_f[w]; fi = sh_i[w]; }
            tot += sh_t[w];
        }
        // Best tracking (de_oo.py:257,285-286; kernel/Point.py:309-333): Old_best survives only if
        // its f is strictly smaller, so a tie hands Best to the newer trial-best.
        const double old_best = *p.best_f;
        const int changed = p.init ? 1 : !(old_best < f);
        if (changed) *p.best_f = f;
        oode_gen_record r;
        r.trial_best_f = f;
        r.best_f = changed ? f : old_best;
        r.trial_best_idx = fi;
        r.n_accepted = (int64_t)tot;
        r.n_repaired = (int64_t)(*p.n_repaired);
        r.best_changed = changed;
        r.generation = p.generation;
        *p.rec = r;
        *p.n_repaired = 0ull;
        *p.ticket = 0u;
        sh_changed = changed;
        sh_i[0] = fi;
    }
    __syncthreads();
    // Best.x: the trial-best row sits in the spare slot of its row (trials are always written);
    // for the initial population it is the current slot.
    const int64_t bi = sh_i[0];
    const double *src;
    if (p.init) src = p.buf0 + bi * p.ld;
    else src = (p.sel_cur[bi] ? p.buf0 : p.buf1) + bi * p.ld;
    if (sh_changed) {
        for (int j = threadIdx.x; j < p.Dl; j += blockDim.x) {
            const double v = src[j];
            p.best_x[j] = v;
            p.bestx_slot[j] = v;
        }
    } else {
        for (int j = threadIdx.x; j < p.Dl; j += blockDim.x) p.bestx_slot[j] = p.best_x[j];
    }
}

// K0: initial population (de_oo.py:147-150).  U from Philox (generation 0) or from the caller.
struct InitParams {
    double *buf0;
    uint8_t *sel0, *sel1;
    int64_t ld, NP;
    int Dl, D, col0;
    const double *lb, *ub, *x0;   // x0 NULL if it has no finite entry
    const double *U0;             // replay: [NP][D]
    int use_philox;
    PhiloxKeys keys;
};

__global__ void __launch_bounds__(256) de_init_kernel(const InitParams p) {
    const int64_t n = p.NP * p.Dl;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x) {
        const int64_t row = e / p.Dl;
        const int lc = (int)(e - row * p.Dl);
        double u;
        if (p.use_philox) {
            uint32_t w0 = (uint32_t)(p.col0 + lc), w1 = (uint32_t)row, w2 = (uint32_t)((uint64_t)row >> 32) | (TAG_ELEM << 24), w3 = 0u;
            philox4x32_10(w0, w1, w2, w3, p.keys);
            u = u52(w0, w1);
        } else {
            u = p.U0[row * p.D + p.col0 + lc];
        }
        double v = dadd(dmul(u, dsub(p.ub[lc], p.lb[lc])), p.lb[lc]);
        if (row == 0 && p.x0) v = p.x0[lc];
        p.buf0[row * p.ld + lc] = v;
        if (lc == 0) { p.sel0[row] = 0; p.sel1[row] = 0; }
    }
}

// objective of arbitrary points (oode_eval_points): fold + finalize only
template <int OBJ>
__global__ void de_finalize_kernel(const double *partials, const int *prog, int nprog, int nleaves, int D,
                                   int invert, int64_t n, double *out) {
    using O = Obj<OBJ>;
    constexpr int NSL = O::NS + O::NPR;
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double s[NSL];
#pragma unroll
    for (int q = 0; q < NSL; ++q) s[q] = fold_slot<OBJ>(partials + i * (int64_t)nleaves * NSL, NSL, q, prog, nprog);
    double f = O::finalize(s, D);
    out[i] = invert ? -f : f;
}

}  // namespace oode
