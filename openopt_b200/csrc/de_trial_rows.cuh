// K2a, row-walking form (round 2): one kernel template for wide and narrow rows.
//
// Same data movement as de_trial_tma.cuh -- every worker walks whole rows leaf by leaf, the four
// <= 1 KB leaf segments of a trial (parent + three donors) arrive by TMA bulk copies in a 2-stage
// shared-memory ring per worker with mbarrier completion -- but the per-gene instruction stream is
// rebuilt around what ncu showed the old forms spend their issue slots on (profiles/r02a_*):
//
//   * WPW workers per warp (1 or 2; a worker is 32/WPW lanes and owns its own row sequence, stages and
//     barriers; the workers of a warp walk consecutive rows in lockstep), so the per-row bookkeeping
//     (donor-index chase, slot selectors, barrier wait, refill) is issued once for WPW rows.  (The
//     template also instantiates for 4 -- quarter-warp workers, a single-pair tail batch so that a
//     100-column row costs 56 pair slots -- but that form measured slowest on every shape and is not built.)
//   * genes are made in BATCHES of two column pairs (four genes): the Philox rounds, the mutation, the
//     bound repair and the cosine polynomial of the four genes are written step-major, so one load of a
//     round key / polynomial coefficient (LDCU -> uniform register) serves four genes instead of one,
//     and four independent dependency chains are in flight per lane.
//   * FAST instantiation: when the host can prove that a single reflection always lands inside the box
//     (|F| < 1, population inside [lb, ub]) the general (1+2^-k) repair loop is not instantiated at
//     all -- no range check, no call, no reconvergence barrier per gene -- and neither is the |y| <= 1e6
//     guard of the cosine.  Everything else runs the SAFE instantiation (same results, more checks).
//   * uniforms come from one I2F: Ft = fma(double(w), F*2^-32, 0) == (w*2^-32)*F rounded once (the
//     scaling by 2^-32 is exact), instead of assembling 1.m from the bits and subtracting 1.
//   * objective terms are written over the consumed inputs of the SAME lane's columns right away (no
//     second pass over registers, one warp barrier less per leaf); kernel constants are read from
//     the parameter bank where they are used (the XOR-with-zero register pinning of round 1 made the
//     compiler recompute the XOR at every use).
//   * compile-time leaf lengths 128 and 120 (rows_gene_batch<CLEN>), a three-row pipelined donor chase (load_idx /
//     load_sel / make_ptrs), all slots of a sum-only objective summed in one pass (rows_leaf_sum).
//   * multi-GPU: under the peer exchange a worker takes chunks of 8 consecutive rows and hands their leaf sums to every
//     rank as one contiguous run (rows_put / rows_flush); under row shards with single-leaf rows it takes K2b's
//     accept decision itself and stores an accepted row straight into the other ranks' replicas (RowPush).
//
// Objectives that couple neighbouring columns (Rosenbrock) keep the round-1 kernels.
#pragma once
#include "de_trial_tma.cuh"
#include "de_launch.h"

namespace oode {

// B Philox4x32-10 calls in lockstep: a round key is fetched once per round for all of them.
template <int B>
__device__ __forceinline__ void philox_batch(uint32_t (&c)[B][4], const PhiloxKeys &K) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t k0 = K.rk[2 * r], k1 = K.rk[2 * r + 1];
#pragma unroll
        for (int b = 0; b < B; ++b) {
            const uint64_t p0 = (uint64_t)PHILOX_M0 * c[b][0];
            const uint64_t p1 = (uint64_t)PHILOX_M1 * c[b][2];
            const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c[b][1] ^ k0;
            const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c[b][3] ^ k1;
            c[b][1] = (uint32_t)p1;
            c[b][3] = (uint32_t)p0;
            c[b][0] = n0;
            c[b][2] = n2;
        }
    }
}

// obj_cos (de_device.cuh) for N arguments, step-major: the same operations per element, in the same
// order -- bit-identical results -- with every coefficient fetched once per step.
template <bool CK, int N>
__device__ __forceinline__ void obj_cos_n(const double (&y)[N], double (&c)[N]) {
    if (CK) {
        bool slow = false;
#pragma unroll
        for (int i = 0; i < N; ++i) slow |= !(fabs(y[i]) <= COS_K[11]);
        if (slow) {
#pragma unroll
            for (int i = 0; i < N; ++i) c[i] = obj_cos<true>(y[i]);
            return;
        }
    }
    double kd[N], r[N], h[N], z[N], pl[N];
    int kb[N];
#pragma unroll
    for (int i = 0; i < N; ++i) kd[i] = __fma_rn(y[i], COS_K[0], COS_K[1]);
#pragma unroll
    for (int i = 0; i < N; ++i) { kb[i] = __double2loint(kd[i]); kd[i] = __dsub_rn(kd[i], COS_K[1]); }
#pragma unroll
    for (int i = 0; i < N; ++i) r[i] = __fma_rn(-kd[i], COS_K[2], y[i]);
#pragma unroll
    for (int i = 0; i < N; ++i) r[i] = __fma_rn(-kd[i], COS_K[3], r[i]);
#pragma unroll
    for (int i = 0; i < N; ++i) r[i] = __fma_rn(-kd[i], COS_K[4], r[i]);
#pragma unroll
    for (int i = 0; i < N; ++i) { h[i] = __dmul_rn(0.5, r[i]); z[i] = __dmul_rn(h[i], h[i]); }
#pragma unroll
    for (int i = 0; i < N; ++i) pl[i] = __fma_rn(z[i], COS_K[10], COS_K[9]);
#pragma unroll
    for (int i = 0; i < N; ++i) pl[i] = __fma_rn(z[i], pl[i], COS_K[8]);
#pragma unroll
    for (int i = 0; i < N; ++i) pl[i] = __fma_rn(z[i], pl[i], COS_K[7]);
#pragma unroll
    for (int i = 0; i < N; ++i) pl[i] = __fma_rn(z[i], pl[i], COS_K[6]);
#pragma unroll
    for (int i = 0; i < N; ++i) pl[i] = __fma_rn(z[i], pl[i], COS_K[5]);
#pragma unroll
    for (int i = 0; i < N; ++i) {
        const double sn = __fma_rn(__dmul_rn(h[i], z[i]), pl[i], h[i]);
        const double cv = __fma_rn(__dmul_rn(-2.0, sn), sn, 1.0);
        c[i] = __hiloint2double(__double2hiint(cv) ^ (kb[i] << 31), __double2loint(cv));
    }
}

// Objective terms of N genes (no neighbour coupling): Obj<>::term per gene, with the cosines of the
// N genes evaluated together.
template <int OBJ, bool CK, int N>
__device__ __forceinline__ void obj_terms_n(const double (&x)[N], const double (&aux)[N], const double (&raux)[N],
                                            const int (&col)[N], double (&t)[N][Obj<OBJ>::NS + Obj<OBJ>::NPR]) {
    if constexpr (OBJ == OODE_OBJ_RASTRIGIN || OBJ == OODE_OBJ_ACKLEY || OBJ == OODE_OBJ_GRIEWANK) {
        double y[N], c[N];
#pragma unroll
        for (int i = 0; i < N; ++i) {
            if (OBJ == OODE_OBJ_GRIEWANK) {
                // x / aux with the host's correctly rounded reciprocal r = RN(1/aux): q = RN(x r), then one residual
                // correction q + r * (x - q aux) (two FMAs) -- the quotient to within its last bit, three instructions
                // instead of the ~20 of a full division; the cosine that follows is itself good to 2.7e-16 only
                const double q = dmul(x[i], raux[i]);
                y[i] = __fma_rn(__fma_rn(-q, aux[i], x[i]), raux[i], q);
            } else {
                y[i] = dmul(TWO_PI, x[i]);
            }
        }
        obj_cos_n<CK, N>(y, c);
#pragma unroll
        for (int i = 0; i < N; ++i) {
            if (OBJ == OODE_OBJ_RASTRIGIN) t[i][0] = dadd(dsub(dmul(x[i], x[i]), dmul(10.0, c[i])), 10.0);
            else { t[i][0] = dmul(x[i], x[i]); t[i][Obj<OBJ>::NS + Obj<OBJ>::NPR - 1] = c[i]; }
        }
    } else {
#pragma unroll
        for (int i = 0; i < N; ++i) Obj<OBJ>::template term<CK>(x[i], 0.0, aux[i], col[i], t[i]);
    }
}

// B column pairs (2B genes) of one leaf for every worker of the warp: draws, mutation, crossover, bound repair,
// trial-row store, objective terms (written over the consumed inputs of this lane's own columns: no other
// lane reads them, so no barrier is needed before the write).  k0 = first pair pass of the batch.
// CLEN: the leaf's length when it is known at compile time and every worker of the warp is valid -- 128 (no
// predicates at all) or 120 (the other leaf size of a 1000-column row and of its 8-way column shards: only the
// last pair pass keeps a lane predicate, everything else folds) -- or 0: run-time length.
// `lim` = terms of this leaf for this worker (0 for a
// worker past the end of the population).
template <int OBJ, int MODE, bool UNI, bool FAST, int WPW, int CLEN, int B>
__device__ __forceinline__ void rows_gene_batch(const GenParams &p, double *st, int rs, int off, int ulen2_, int lim_, int hl, int k0,
                                                double *Wp, const double *Uf, const double *Uc, uint32_t row_lo,
                                                unsigned &nrep, bool keepx) {
    using O = Obj<OBJ>;
    constexpr int NSL = O::NS + O::NPR;
    constexpr int LW = 32 / WPW;
    constexpr bool FULL = CLEN == 128;
    const int lim = CLEN ? CLEN : lim_, ulen2 = CLEN ? CLEN : ulen2_;
    const int glim = (lim + 1) & ~1;         // genes to store: an odd count drags the pad column along
    int lc[B], ll[B];                        // column of the pair inside the leaf; ll: the same, clamped for loads
    bool a0[B], a1[B];
#pragma unroll
    for (int b = 0; b < B; ++b) {
        lc[b] = 2 * LW * (k0 + b) + 2 * hl;
        // lanes past the end of a short leaf compute on a clamped column (results masked): no load may leave the
        // leaf's columns (ulen2 = its length rounded up to even), neither in the stage nor in the per-column arrays
        ll[b] = FULL ? lc[b] : min(lc[b], ulen2 - 2);
        OODE_ASSERT(ll[b] >= 0 && ll[b] + 1 < rs + (rs & 1) && ll[b] + 1 < ulen2 + 1 && off + ll[b] + 1 <= p.Dl + 1);
        a0[b] = FULL || lc[b] < lim;
        a1[b] = FULL || lc[b] + 1 < lim;
    }
    double x[2 * B];
    if (MODE == MODE_EVAL) {
#pragma unroll
        for (int b = 0; b < B; ++b) {
            const double2 xo = *reinterpret_cast<const double2 *>(st + ll[b]);
            x[2 * b] = xo.x; x[2 * b + 1] = xo.y;
        }
    } else {
        double ft[2 * B];
        bool keep[2 * B];
        if (MODE == MODE_PHILOX) {
            uint32_t w[B][4];
#pragma unroll
            for (int b = 0; b < B; ++b) {
                w[b][0] = (uint32_t)(p.col0 + off + lc[b]) >> 1;
                w[b][1] = row_lo;
                w[b][2] = (TAG_ELEM << 24);
                w[b][3] = p.gen;
            }
            philox_batch<B>(w, p.keys);
#pragma unroll
            for (int b = 0; b < B; ++b) {
                // Ft = (w * 2^-32) * F rounded once (de_oo.py:236-237): double(w) is exact and Fa32 = F * 2^-32 is an
                // exact scaling; (Fa32, Fb) = (0, F) is the 'constant' difference factor (:238-239)
                ft[2 * b] = __fma_rn(__uint2double_rn(w[b][0]), p.Fa32, p.Fb);
                ft[2 * b + 1] = __fma_rn(__uint2double_rn(w[b][2]), p.Fa32, p.Fb);
                keep[2 * b] = keep_parent(w[b][1], p.cr_thresh, p.cr_all != 0);
                keep[2 * b + 1] = keep_parent(w[b][3], p.cr_thresh, p.cr_all != 0);
            }
        } else {
#pragma unroll
            for (int b = 0; b < B; ++b) {
                const int col = p.col0 + off + lc[b];
                const double uc0 = a0[b] ? Uc[col] : 2.0, uc1 = a1[b] ? Uc[col + 1] : 2.0;
                const double uf0 = (a0[b] && !p.const_F) ? Uf[col] : 0.0, uf1 = (a1[b] && !p.const_F) ? Uf[col + 1] : 0.0;
                keep[2 * b] = uc0 > p.Cr; keep[2 * b + 1] = uc1 > p.Cr;
                ft[2 * b] = __fma_rn(uf0, p.Fa, p.Fb); ft[2 * b + 1] = __fma_rn(uf1, p.Fa, p.Fb);
            }
        }
        double v[2 * B], lo[2 * B], hi[2 * B];
#pragma unroll
        for (int b = 0; b < B; ++b) {
            const double2 xo = *reinterpret_cast<const double2 *>(st + ll[b]);
            const double2 xb = *reinterpret_cast<const double2 *>(st + rs + ll[b]);
            const double2 x1 = *reinterpret_cast<const double2 *>(st + 2 * rs + ll[b]);
            const double2 x2 = *reinterpret_cast<const double2 *>(st + 3 * rs + ll[b]);
            // de_oo.py:241,250: mutant = beta + Ft*delta; the parent's gene is kept iff Uc > Cr
            v[2 * b] = keep[2 * b] ? xo.x : dadd(xb.x, dmul(ft[2 * b], dsub(x2.x, x1.x)));
            v[2 * b + 1] = keep[2 * b + 1] ? xo.y : dadd(xb.y, dmul(ft[2 * b + 1], dsub(x2.y, x1.y)));
            if (UNI) { lo[2 * b] = lo[2 * b + 1] = p.lo_s; hi[2 * b] = hi[2 * b + 1] = p.hi_s; }
            else {
                const double2 l = *reinterpret_cast<const double2 *>(p.lb + off + ll[b]);
                const double2 u = *reinterpret_cast<const double2 *>(p.ub + off + ll[b]);
                lo[2 * b] = l.x; lo[2 * b + 1] = l.y; hi[2 * b] = u.x; hi[2 * b + 1] = u.y;
            }
        }
        // _correct_box_constraints (de_oo.py:346-375): one reflection about the violated bound (repair_gene, de_device.cuh)
        bool bad = false;
#pragma unroll
        for (int i = 0; i < 2 * B; ++i) {
            const bool act = (i & 1) ? a1[i >> 1] : a0[i >> 1];
            const bool lt = v[i] < lo[i], gt = v[i] > hi[i];
            const double bnd = lt ? lo[i] : hi[i];
            const double refl = __fma_rn(-2.0, dsub(v[i], bnd), v[i]);
            const bool viol = lt || gt;
            x[i] = viol ? refl : v[i];
            nrep += (viol && act) ? 1u : 0u;
            if (!FAST) bad |= act && !((x[i] >= lo[i]) && (x[i] <= hi[i]));
        }
        if (!FAST && bad) {                    // rare: re-run the reference's general loops from the original value
#pragma unroll
            for (int i = 0; i < 2 * B; ++i) {
                const bool act = (i & 1) ? a1[i >> 1] : a0[i >> 1];
                if (act && !((x[i] >= lo[i]) && (x[i] <= hi[i]))) x[i] = repair_gene_slow(v[i], lo[i], hi[i]);
            }
        }
#pragma unroll
        for (int b = 0; b < B; ++b)
            if (FULL || lc[b] < glim) {
                OODE_ASSERT(off + lc[b] + 1 < p.ld);
                *reinterpret_cast<double2 *>(Wp + off + 2 * LW * (k0 + b)) = make_double2(x[2 * b], x[2 * b + 1]);
                // fused row push: the genes stay in the stage's fourth row (its inputs, this lane's own columns, are
                // consumed), from where an accepted row is copied to the peers after the leaf sum
                if (keepx) *reinterpret_cast<double2 *>(st + 3 * rs + lc[b]) = make_double2(x[2 * b], x[2 * b + 1]);
            }
    }
    double aux[2 * B], raux[2 * B];
#pragma unroll
    for (int b = 0; b < B; ++b) {
        if (O::AUX) {
            const double2 av = *reinterpret_cast<const double2 *>(p.aux + off + ll[b]);
            aux[2 * b] = av.x; aux[2 * b + 1] = av.y;
        } else { aux[2 * b] = aux[2 * b + 1] = 0.0; }
        if (OBJ == OODE_OBJ_GRIEWANK) {
            const double2 rv = *reinterpret_cast<const double2 *>(p.raux + off + ll[b]);
            raux[2 * b] = rv.x; raux[2 * b + 1] = rv.y;
        } else { raux[2 * b] = raux[2 * b + 1] = 0.0; }
    }
    double t[2 * B][NSL];
    int col[2 * B];
#pragma unroll
    for (int i = 0; i < 2 * B; ++i) col[i] = p.col0 + off + lc[i >> 1] + (i & 1);
    obj_terms_n<OBJ, !FAST, 2 * B>(x, aux, raux, col, t);
#pragma unroll
    for (int b = 0; b < B; ++b)
        if (FULL || lc[b] < glim) {
            OODE_ASSERT(lc[b] + 1 < rs + (rs & 1));
#pragma unroll
            for (int q = 0; q < NSL; ++q)
                *reinterpret_cast<double2 *>(st + q * rs + lc[b]) = make_double2(t[2 * b][q], t[2 * b + 1][q]);
        }
}

// All column pairs of one leaf: 64 / (32/WPW) pair passes per lane, two at a time.  ulen = terms of the leaf
// (the same for every worker of the warp).  Quarter-warp workers finish an odd number of passes with a
// single-pair batch, so a 100-column row costs 7 passes of 8 lanes (56 pair slots for 50 pairs).
template <int OBJ, int MODE, bool UNI, bool FAST, int WPW, int CLEN>
__device__ __forceinline__ void rows_leaf_body(const GenParams &p, double *st, int rs, int off, int ulen, int lim, int hl,
                                               double *Wp, const double *Uf, const double *Uc, uint32_t row_lo,
                                               unsigned &nrep, bool keepx) {
    constexpr int LW = 32 / WPW, PP = 64 / LW;
    if (CLEN) {                                  // 128 or 120 columns: every pair pass holds columns
#pragma unroll
        for (int k0 = 0; k0 < PP; k0 += 2)
            rows_gene_batch<OBJ, MODE, UNI, FAST, WPW, CLEN, 2>(p, st, rs, off, CLEN, CLEN, hl, k0, Wp, Uf, Uc, row_lo, nrep, keepx);
    } else {
        const int npass = (ulen + 2 * LW - 1) / (2 * LW);           // pair passes that hold columns of this leaf
        const int ulen2 = (ulen + 1) & ~1;
#pragma unroll
        for (int k0 = 0; k0 < PP; k0 += 2) {
            if (k0 + 1 < npass || (WPW != 4 && k0 < npass))
                rows_gene_batch<OBJ, MODE, UNI, FAST, WPW, 0, 2>(p, st, rs, off, ulen2, lim, hl, k0, Wp, Uf, Uc, row_lo, nrep, keepx);
            else if (WPW == 4 && k0 < npass)
                rows_gene_batch<OBJ, MODE, UNI, FAST, WPW, 0, 1>(p, st, rs, off, ulen2, lim, hl, k0, Wp, Uf, Uc, row_lo, nrep, keepx);
        }
    }
}

// Where a finished leaf sum goes.  Single GPU / NCCL exchange: the local partials array.  Peer exchange: the worker's
// small shared-memory buffer xb (indexed like the destination, relative to the first row of the worker's row chunk),
// flushed to every rank in contiguous >= 128-byte runs when the chunk is done (rows_flush): storing each 8-byte sum
// to 8 peers as it appears costs one NVLink write packet per 16 ... 32 bytes, and at 1M rows the packet rate -- not
// the byte rate -- delayed the slowest rank's epoch flag by ~0.1 ms (profiles/r02l_*).
__device__ __forceinline__ void rows_put(const GenParams &p, int64_t out, double v, int jn, double *xb) {
    OODE_ASSERT(!xb || (out >= 0 && out < XBUF_DOUBLES));
    if (xb) { if (jn == 0) xb[out] = v; }
    else put_partial<8>(p, out, v, jn);
}

// Flush a worker's exchange buffer: n doubles that are contiguous in every rank's copy, starting at `first`.
template <int LW>
__device__ __forceinline__ void rows_flush(const GenParams &p, const double *xb, int64_t first, int n, int hl) {
    OODE_ASSERT(n > 0 && n <= XBUF_DOUBLES && first >= 0 && first + n <= p.part_limit);
    for (int g = 0; g < p.peers.n; ++g) {
        double *dst = p.peers.part[g] + first;
        for (int t = hl; t < n; t += LW) dst[t] = xb[t];
    }
}

// numpy's pairwise leaf sum over the terms in the stage: 8 lanes per (worker, slot) are the r[j] chains.
// The slot index is a compile-time constant inside the unrolled loop, so slot_op is ONE add or multiply (with a
// run-time slot it is an add, a multiply and a select per step).
template <int OBJ, int Q, int CLEN>
__device__ __forceinline__ void rows_slot_sum(const GenParams &p, const double *tq, int len, int jn, unsigned gm, int64_t out,
                                              double *xb, double *fs, int qrt) {
    using O = Obj<OBJ>;
    if (CLEN) {                                   // 16 (128 terms) or 15 (120 terms) full steps, no tail
        double r = tq[jn];
#pragma unroll
        for (int k = 1; k < CLEN / 8; ++k) r = slot_op<OBJ>(Q, r, tq[8 * k + jn]);
        r = slot_op<OBJ>(Q, r, __shfl_xor_sync(gm, r, 1));
        r = slot_op<OBJ>(Q, r, __shfl_xor_sync(gm, r, 2));
        r = slot_op<OBJ>(Q, r, __shfl_xor_sync(gm, r, 4));
        rows_put(p, out, r, jn, xb);
        if (fs && jn == 0) fs[qrt] = r;
    } else if (len >= 96) {                       // 12..16 full steps + a tail of len % 8 terms
        const int nfull = len >> 3;
        double r = tq[jn];
#pragma unroll
        for (int k = 1; k < 12; ++k) r = slot_op<OBJ>(Q, r, tq[8 * k + jn]);
#pragma unroll
        for (int k = 12; k < 16; ++k)
            if (k < nfull) r = slot_op<OBJ>(Q, r, tq[8 * k + jn]);
        r = slot_op<OBJ>(Q, r, __shfl_xor_sync(gm, r, 1));
        r = slot_op<OBJ>(Q, r, __shfl_xor_sync(gm, r, 2));
        r = slot_op<OBJ>(Q, r, __shfl_xor_sync(gm, r, 4));
        for (int i = nfull * 8; i < len; ++i) r = slot_op<OBJ>(Q, r, tq[i]);    // same value in all 8 lanes
        rows_put(p, out, r, jn, xb);
        if (fs && jn == 0) fs[qrt] = r;
    } else {
        const int nfull = len >> 3;
        double r = (Q < O::NS) ? 0.0 : 1.0;      // numpy: n < 8 starts from 0. (1. for a product)
        if (nfull > 0) {
            r = tq[jn];
            for (int k = 1; k < nfull; ++k) r = slot_op<OBJ>(Q, r, tq[8 * k + jn]);
            r = slot_op<OBJ>(Q, r, __shfl_xor_sync(gm, r, 1));
            r = slot_op<OBJ>(Q, r, __shfl_xor_sync(gm, r, 2));
            r = slot_op<OBJ>(Q, r, __shfl_xor_sync(gm, r, 4));
        }
        for (int i = nfull * 8; i < len; ++i) r = slot_op<OBJ>(Q, r, tq[i]);
        rows_put(p, out, r, jn, xb);
        if (fs && jn == 0) fs[qrt] = r;
    }
}

template <int OBJ, int Q, int NSL, int WPW, int CLEN> struct RowsSumLoop {
    static __device__ __forceinline__ void run(const GenParams &p, const double *st, int rs, int len, int lane, int hl, bool valid,
                                               int64_t out_index, double *xb, double *fs) {
        constexpr int QP = (32 / WPW) / 8;                  // 8-lane groups per worker: group g sums slots g, g+QP, ...
        if ((hl >> 3) == (Q % QP) && valid)
            rows_slot_sum<OBJ, Q, CLEN>(p, st + Q * rs, len, hl & 7, 0xFFu << (lane & 24), out_index + Q, xb, fs, Q);
        RowsSumLoop<OBJ, Q + 1, NSL, WPW, CLEN>::run(p, st, rs, len, lane, hl, valid, out_index, xb, fs);
    }
};
template <int OBJ, int NSL, int WPW, int CLEN> struct RowsSumLoop<OBJ, NSL, NSL, WPW, CLEN> {
    static __device__ __forceinline__ void run(const GenParams &, const double *, int, int, int, int, bool, int64_t, double *, double *) {}
};

template <int OBJ, int WPW, int CLEN>
__device__ __forceinline__ void rows_leaf_sum(const GenParams &p, const double *st, int rs, int len, int lane, int hl,
                                              bool valid, int64_t out_index, double *xb, double *fs) {
    using O = Obj<OBJ>;
    constexpr int NSL = O::NS + O::NPR;
    constexpr int QP = (32 / WPW) / 8;
    if constexpr (O::NPR == 0 && NSL <= QP) {
        // every slot is a sum and a worker has an 8-lane group per slot: all slots in ONE pass (group g sums slot g;
        // the slot index may be a run-time value because the operation is an add for every slot)
        const int q = hl >> 3;
        if (q < NSL && valid)
            rows_slot_sum<OBJ, 0, CLEN>(p, st + q * rs, len, hl & 7, 0xFFu << (lane & 24), out_index + q, xb, fs, q);
    } else {
        // sums and products mixed (or fewer groups than slots): one pass per slot with the slot index -- hence the
        // operation -- known at compile time
        RowsSumLoop<OBJ, 0, NSL, WPW, CLEN>::run(p, st, rs, len, lane, hl, valid, out_index, xb, fs);
    }
}

// MAXW: most warps a CTA is launched with (register budget); the launch picks warps = blockDim.x / 32 <= MAXW so
// that the stages of all workers fit the SM's shared memory (rows_smem_bytes).
template <int OBJ, int MODE, bool UNI, bool FAST, int WPW, int MAXW>
__global__ void __launch_bounds__(MAXW * 32, 1) de_trial_rows_kernel(const GenParams p) {
    using O = Obj<OBJ>;
    static_assert(O::HALO == 0, "neighbour-coupled objectives use the round-1 kernels");
    constexpr int NSL = O::NS + O::NPR;
    constexpr int LW = 32 / WPW, S = ROWS_S;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int NC = (int)blockDim.x >> 5;
    const int rs = p.rs;
    const uint32_t stage_bytes = 32u * (uint32_t)rs;
    const int c = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int sub = lane / LW, hl = lane % LW;
    const int w = c * WPW + sub;                                 // worker inside the CTA
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem_raw + (size_t)NC * WPW * S * stage_bytes);
    if (threadIdx.x == 0) {
        for (int i = 0; i < NC * WPW * S; ++i) mbar_init(smem_u32(bars + i), 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    const int nleaves = p.nleaves;
    // A worker takes runs of R consecutive rows (row chunks w, w + G', ... with G' workers in the grid).  R = 1 -- rows
    // w, w + G', ... -- except under the peer exchange, where the leaf sums of a chunk are buffered and flushed to the
    // peers as one contiguous run (rows_put / rows_flush): R = p.xrows (8 when a row's sums are 16 bytes).
    const int R = p.xrows;
    const int Gw = (int)gridDim.x * NC * WPW;                   // workers in the grid
    const int first = (((int)blockIdx.x * NC + c) * WPW + sub) * R;
    const int jump = (Gw - 1) * R + 1;                          // from the last row of a chunk to the first of the next
    auto nxt = [&](int r) { return ((r + 1) % R) ? r + 1 : r + jump; };
    // Row shards with single-leaf rows: accepted trial rows are pushed to the peers from here (RowPush).  fs: where the
    // worker's slot sums of the row are left for the acceptance test (the tail of its exchange buffer).
    const bool rpush = MODE != MODE_EVAL && p.rpush.n > 1 && nleaves == 1;
    double *fs = rpush ? reinterpret_cast<double *>(smem_raw + (size_t)NC * WPW * S * (stage_bytes + 8)) + w * XBUF_DOUBLES + (XBUF_DOUBLES - 4)
                       : nullptr;
    double *xb = nullptr;                                       // this worker's exchange buffer (peer exchange only)
    if (p.peers.n && R * p.part_stride * NSL <= XBUF_DOUBLES)
        xb = reinterpret_cast<double *>(smem_raw + (size_t)NC * WPW * S * (stage_bytes + 8)) + w * XBUF_DOUBLES;
    const uint32_t bar0 = smem_u32(bars + w * S);
    const uint32_t dst0 = smem_u32(smem_raw) + (uint32_t)(w * S) * stage_bytes;
    double *my_stages = reinterpret_cast<double *>(smem_raw + (size_t)(w * S) * stage_bytes);
    const int ld = p.ld;
    const double *const buf0 = p.buf0, *const buf1 = p.buf1;
    const bool vec_base = p.base_vec != nullptr, vec_dir = p.d1_vec != nullptr;

    // Look-ahead: the cursor (la_row, la_leaf) is the next item to be issued, S items ahead of the one being computed.
    // A row's sources sit behind two dependent loads -- its donor indices, then the donors' slot selectors -- so the
    // chase is pipelined over THREE rows of the cursor's sequence: while row R is being issued, the selector loads of
    // row R+G and the index loads of row R+2G are in flight, and each gets a whole row period (>= one item) to land.
    // With one leaf per row (<= 128 columns) the chase would otherwise stall every refill for a full memory latency.
    struct RowIdx { int ib, r1, r2; uint8_t srow; };
    struct RowSel { uint8_t b, r1, r2; };
    int la_row = first, la_leaf = 0;
    RowSrc la;
    la.P = la.B = la.R1 = la.R2 = nullptr;
    RowIdx f1{0, 0, 0, 0}, f2{0, 0, 0, 0};
    RowSel s1{0, 0, 0};

    auto load_idx = [&](int lrow, RowIdx &f) {              // first level: a row's selector and donor indices
        if (lrow < p.nrows) {
            const int row = p.row0 + lrow;
            f.srow = p.sel_cur[row];
            // unconditional on the strategy (the index arrays always hold valid rows): predicating these loads keeps
            // the compiler from scheduling them early
            if (MODE != MODE_EVAL) { f.ib = p.ib[row]; f.r1 = p.r1[row]; f.r2 = p.r2[row]; }
        }
    };
    auto load_sel = [&](const RowIdx &f, RowSel &sl) {      // second level: the donors' slot selectors
        if (MODE != MODE_EVAL) { sl.b = p.sel_cur[f.ib]; sl.r1 = p.sel_cur[f.r1]; sl.r2 = p.sel_cur[f.r2]; }
    };
    auto make_ptrs = [&](int lrow, const RowIdx &f, const RowSel &sl) {
        const int row = p.row0 + lrow;
        OODE_ASSERT(f.ib >= 0 && f.ib < p.NP && f.r1 >= 0 && f.r1 < p.NP && f.r2 >= 0 && f.r2 < p.NP && sl.b < 2 && sl.r1 < 2 &&
                    sl.r2 < 2 && f.srow < 2);
        la.P = (f.srow ? buf1 : buf0) + (int64_t)row * ld;
        if (MODE != MODE_EVAL) {
            // strategies (de_oo.py:186-227): the base row / the two direction rows are either gathered per trial or
            // one vector shared by every trial of the generation
            const double *gB = (sl.b ? buf1 : buf0) + (int64_t)f.ib * ld;
            const double *g1 = (sl.r1 ? buf1 : buf0) + (int64_t)f.r1 * ld;
            const double *g2 = (sl.r2 ? buf1 : buf0) + (int64_t)f.r2 * ld;
            la.B = vec_base ? p.base_vec : gB;
            la.R1 = vec_dir ? p.d1_vec : g1;
            la.R2 = vec_dir ? p.d2_vec : g2;
        }
    };
    auto issue = [&](int s) {                                // one lane per worker
        const int2 lf = p.leaves[la_leaf];
        const int ng = min((lf.y + 1) & ~1, ld - lf.x);      // genes needed, rounded up to 16 bytes
        OODE_ASSERT(la_row >= 0 && la_row < p.nrows && p.row0 + la_row < p.NP && ng > 0 && ng <= rs && lf.x >= 0 &&
                    lf.x + ng <= ld && s >= 0 && s < S);
        const uint32_t bytes = (uint32_t)ng * 8u;
        const uint32_t full = bar0 + 8u * s, dst = dst0 + (uint32_t)s * stage_bytes;
        mbar_arrive_expect_tx(full, MODE == MODE_EVAL ? bytes : 4u * bytes);
        bulk_g2s(dst, la.P + lf.x, bytes, full);
        if (MODE != MODE_EVAL) {
            bulk_g2s(dst + 8u * rs, la.B + lf.x, bytes, full);
            bulk_g2s(dst + 16u * rs, la.R1 + lf.x, bytes, full);
            bulk_g2s(dst + 24u * rs, la.R2 + lf.x, bytes, full);
        }
    };
    auto advance_la = [&]() {
        if (++la_leaf == nleaves) {
            la_leaf = 0;
            la_row = nxt(la_row);
            make_ptrs(la_row, f1, s1);                       // loaded one row period ago
            f1 = f2;
            load_sel(f1, s1);
            load_idx(nxt(nxt(la_row)), f2);
        }
    };

    {                                                        // prologue: the first row's chase is not hidden
        RowIdx f0{0, 0, 0, 0};
        RowSel s0{0, 0, 0};
        load_idx(la_row, f0);
        load_idx(nxt(la_row), f1);
        load_idx(nxt(nxt(la_row)), f2);
        load_sel(f0, s0);
        load_sel(f1, s1);
        make_ptrs(la_row, f0, s0);
    }
    for (int s = 0; s < S; ++s) {                            // fill all stages
        const bool v = la_row < p.nrows;
        if (v && hl == 0) issue(s);
        if (v) advance_la();
    }

    unsigned nrep = 0;
    int s = 0;
    uint32_t parity = 0;
    for (int lrow = first; (WPW == 1) ? (lrow < p.nrows) : __any_sync(0xFFFFFFFFu, lrow < p.nrows); lrow = nxt(lrow)) {
        const bool valid = lrow < p.nrows;
        const int row = p.row0 + (valid ? lrow : 0);
        OODE_ASSERT(row >= 0 && row < p.NP);
        double *Wp = nullptr;                                // this lane's first column pair of the trial row
        if (MODE != MODE_EVAL) Wp = const_cast<double *>(p.sel_cur[row] ? buf0 : buf1) + (int64_t)row * ld + 2 * hl;
        const double *Uf = nullptr, *Uc = nullptr;
        if (MODE == MODE_REPLAY) {
            Uf = p.Uf + row * (int64_t)p.D;
            Uc = p.Uc + row * (int64_t)p.D;
        }
        const uint32_t row_lo = (uint32_t)row;

        for (int leaf = 0; leaf < nleaves; ++leaf) {
            const uint32_t full = bar0 + 8u * s;
            double *st = my_stages + (size_t)s * (4 * rs);
            const int2 lf = p.leaves[leaf];
            const int len = lf.y;
            const bool la_valid = la_row < p.nrows;

            if (valid) mbar_wait(full, parity);
            // leaf sums: index into the destination array, or (buffered) relative to the chunk's first row
            const int64_t out_index = ((int64_t)(valid ? (xb ? lrow % R : lrow) : 0) * p.part_stride + leaf) * NSL;
            const bool all_valid = (WPW == 1) ? true : __all_sync(0xFFFFFFFFu, valid);
            if (len == 128 && all_valid) {
                rows_leaf_body<OBJ, MODE, UNI, FAST, WPW, 128>(p, st, rs, lf.x, 128, 128, hl, Wp, Uf, Uc, row_lo, nrep, rpush);
                __syncwarp();                        // every lane's terms are in the stage
                rows_leaf_sum<OBJ, WPW, 128>(p, st, rs, 128, lane, hl, true, out_index, xb, fs);
            } else if (len == 120 && all_valid) {
                rows_leaf_body<OBJ, MODE, UNI, FAST, WPW, 120>(p, st, rs, lf.x, 120, 120, hl, Wp, Uf, Uc, row_lo, nrep, rpush);
                __syncwarp();
                rows_leaf_sum<OBJ, WPW, 120>(p, st, rs, 120, lane, hl, true, out_index, xb, fs);
            } else {
                rows_leaf_body<OBJ, MODE, UNI, FAST, WPW, 0>(p, st, rs, lf.x, len, valid ? len : 0, hl, Wp, Uf, Uc, row_lo, nrep, rpush);
                __syncwarp();
                rows_leaf_sum<OBJ, WPW, 0>(p, st, rs, len, lane, hl, valid, out_index, xb, fs);
            }
            if (rpush) {
                // K2b's decision for this row, taken here (same f, same comparison): an accepted trial goes into the
                // same spare slot of every other rank's replica now, from the genes kept in the stage's fourth row
                __syncwarp();                        // the slot sums are in fs
                if (valid) {
                    double sums[NSL];
#pragma unroll
                    for (int q = 0; q < NSL; ++q) sums[q] = fs[q];
                    double f = O::finalize(sums, p.rpush.D);
                    if (p.rpush.invert) f = -f;
                    if (f != f) f = __longlong_as_double(0x7FF0000000000000LL);
                    if (!(p.rpush.vals[row] < f)) {
                        const bool spare1 = p.sel_cur[row] == 0;                 // the trial sits in the row's spare slot
                        const int ng = (len + 1) & ~1;
                        for (int g = 0; g < p.rpush.n; ++g) {
                            if (g == p.rpush.rank) continue;
                            double *dst = (spare1 ? p.rpush.buf1[g] : p.rpush.buf0[g]) + (int64_t)row * ld + lf.x;
                            for (int j = 2 * hl; j < ng; j += 2 * LW)
                                *reinterpret_cast<double2 *>(dst + j) = *reinterpret_cast<const double2 *>(st + 3 * rs + j);
                        }
                    }
                }
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
        if (xb) {                                            // end of a row chunk (or of the population): hand its sums over
            const bool endc = (lrow + 1) % R == 0;           // the same for every worker of the warp (chunks are aligned)
            const bool last = valid && lrow + 1 >= p.nrows;
            if (endc || ((WPW == 1) ? last : __any_sync(0xFFFFFFFFu, last))) {
                __syncwarp();
                if (valid && (endc || last)) {
                    const int c0 = lrow - lrow % R;
                    rows_flush<LW>(p, xb, (int64_t)c0 * p.part_stride * NSL, (lrow - c0 + 1) * p.part_stride * NSL, hl);
                }
                __syncwarp();
            }
        }
    }
    if (MODE != MODE_EVAL) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) nrep += __shfl_xor_sync(0xFFFFFFFFu, nrep, o);
        if (lane == 0 && nrep) atomicAdd(p.n_repaired, (unsigned long long)nrep);
    }
    peer_publish_tail(p);
}

}  // namespace oode
