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
// (oracle/de_oracle.py::pairwise_sum).  A work item is one (row, leaf); it is owned by a quad of
// lanes, lane m accumulating the terms of columns off+8k+2m and off+8k+2m+1 -- exactly numpy's
// r[2m], r[2m+1] chains -- so the row reduction is bit-identical to numpy's with no staging
// buffer: each step the quad reads four 64-byte row segments (parent and three donors; two full
// 32-byte sectors each, as 16-byte vectors) and writes one.  The 8 chain sums are combined in
// numpy's ((r0+r1)+(r2+r3))+((r4+r5)+(r6+r7)) order (one in-lane add, two xor-shuffles), the tail
// terms added sequentially.  Leaf sums go to `partials`; K2b folds them in tree order
// (ReductionPlan::prog).  In native mode one Philox4x32-10 call serves both genes of a lane's
// column pair.
#pragma once
#include "de_device.cuh"

namespace oode {

enum { MODE_EVAL = 0, MODE_PHILOX = 1, MODE_REPLAY = 2 };

// Column shards, peer exchange: every rank keeps a copy of ALL ranks' leaf sums, and the trial kernel
// of rank r stores each of its leaf sums straight into block r of every rank's copy -- its own through
// local memory, the others' through NVLink (CUDA-IPC mapped peer memory).  part[g] points at block r of
// rank g's copy for the current epoch.  n == 0: single GPU or NCCL exchange, `GenParams::partials` is used.
constexpr int MAX_PEERS = 8;
struct PeerOut {
    double *part[MAX_PEERS];
    int n;
    // Publication of this rank's leaf sums, fused into the trial kernel's tail (peer_publish_tail): the last CTA to
    // finish appends the rank's bound-repair counter to its block on every rank and raises its epoch flag there.
    unsigned long long *flags[MAX_PEERS];     // rank g's flag array [MAX_PEERS] (own entry = local memory); null: no publication
    unsigned int *ticket;                     // CTAs of this launch that have finished
    int64_t counter_slot;                     // index of the repair counter inside a block
    int rank;
    unsigned long long epoch;
};

// Row shards, single-leaf rows: the trial kernel itself sends an ACCEPTED trial row to the other ranks' replicas while it
// still holds the row in shared memory -- the NVLink transfer rides under the compute instead of following it in a
// separate pass (de_rows_push_kernel then only delivers the per-row results).  The decision is the one K2b will take:
// the row has one leaf, so f = finalize(leaf sums), and the trial replaces its parent unless vals[row] < f
// (de_oo.py:260-273, box-bounded problems).
struct RowPush {
    double *buf0[MAX_PEERS], *buf1[MAX_PEERS];   // every rank's slot arrays as mapped here (own entry unused)
    const double *vals;                          // f of the current population (the parents)
    int n, rank;                                 // n == 0: no fused push
    int invert, D;
};

struct GenParams {
    // population slots and selectors
    double *buf0, *buf1;
    const uint8_t *sel_cur;
    int ld;                  // leading dimension (doubles) of the slot arrays; even
    // per-row donor indices of this generation (global row numbers)
    const int32_t *ib, *r1, *r2;
    // per-column data (local column slice)
    const double *lb, *ub, *aux;
    const double *raux;      // 1 / aux, rounded once on the host (row kernel, Griewank's x / aux)
    double lo_s, hi_s;       // the bounds when every column has the same ones (UNI kernels)
    // replay draws, [NP][D] global row-major
    const double *Uf, *Uc;
    // leaf sums out: [(lrow*nleaves + leaf)*NSLOT + s]
    double *partials;
    const int2 *leaves;      // (local column offset, length) of every leaf of this rank
    int nleaves;
    int part_stride;         // leaf slots per row in `partials` (nleaves, or the widest shard's when column-sharded)
    int nT;                  // number of terms in this rank's slice (Dl, or Dl-1 with a halo)
    int Dl;                  // local number of columns
    int D;                   // global number of columns
    int col0;                // global index of local column 0 (even)
    int row0;                // first row handled by this launch
    int nrows;               // number of rows handled by this launch (NP < 2^31)
    double F, Cr;
    uint32_t gen;
    PhiloxKeys keys;
    unsigned long long *n_repaired;
    const unsigned long long *n_repaired_zero;   // points at a device word that is always 0
    PeerOut peers;
    // strategy parameters (de_oo.py:71-89); all null / 1 / 0 = the reference's defaults
    const double *base_vec;          // base vector 'best': local slice of the last trial-best row, used for
                                     // every trial (null: row ib[i] of the old population)            :186-193
    const double *d1_vec, *d2_vec;   // directed search: the worst / best barycentre, one (D,) vector each,
                                     // used for every trial (null: per-trial donors r1, r2)          :209-227
    int hndvi;                       // rows per barycentre; r1, r2 are [hndvi][NP]                   :195-208
    int const_F;                     // 1: Ft = F ('constant' difference factor)                      :238-239
    uint32_t cr_thresh;              // floor(Cr * 2^32): a Philox word w keeps the parent's gene iff w > cr_thresh
    int cr_all;                      // 1: Cr < 0, every gene is kept from the parent
    int NP;                          // population size (stride of r1, r2)
    // de_trial_rows_kernel (de_trial_rows.cuh)
    double Fa, Fa32, Fb;             // Ft = fma(uf, Fa, Fb) (replay draws) = fma(double(w), Fa32, Fb) (Philox word w), with
                                     // (Fa, Fb) = (F, 0), Fa32 = F * 2^-32; 'constant' difference factor: (0, F), Fa32 = 0
    int rs;                          // doubles per input row of a shared-memory stage (even, >= the widest leaf)
    int xrows;                       // rows per worker chunk (1; peer exchange: up to 8, de_trial_rows.cuh)
    int64_t part_limit;              // doubles in a leaf-sum destination (checked under OODE_DEBUG_BOUNDS)
    RowPush rpush;                   // row shards: accepted rows leave from the trial kernel (de_trial_rows.cuh)
    int count_tail;                  // neighbour-coupled objectives: count the bound repairs of the columns behind the
                                     // last term (0 on a column shard whose last column is the next shard's first)
};

// One leaf sum -> where K2b will read it.  W lanes (j = 0..W-1) of the caller hold the same value v;
// with a peer exchange they share the up-to-8 remote stores between them (lane j serves ranks j,
// j+W, ...), otherwise lane 0 stores it locally.  The part[] entries are picked with compile-time
// indices (selects) so the parameter block is never indexed dynamically.
template <int W>
__device__ __forceinline__ void put_partial(const GenParams &p, int64_t idx, double v, int j) {
    OODE_ASSERT(idx >= 0 && idx < p.part_limit);
    if (p.peers.n == 0) {
        if (j == 0) p.partials[idx] = v;
        return;
    }
#pragma unroll
    for (int r = 0; r < MAX_PEERS / W; ++r) {
        double *dst = nullptr;
#pragma unroll
        for (int w = 0; w < W; ++w)
            if (w == j) dst = p.peers.part[r * W + w];
        if (r * W + j < p.peers.n) dst[idx] = v;
    }
}

// End of a trial kernel on a column-sharded handle with the peer exchange: every CTA fences its remote stores at
// system scope and takes a ticket; the last one knows that ALL leaf sums of this rank have been performed, so it
// stores the repair counter into its block on every rank and then raises this rank's epoch flag there
// (st.release.sys).  K2b's prologue spins on the local flags (peer_wait).  Epochs only grow and the blocks are
// double-buffered by epoch parity: a rank can be at most one epoch ahead of the slowest one, so nothing a reader
// still needs is overwritten.  Call with all threads of the CTA.
__device__ __forceinline__ void peer_publish_tail(const GenParams &p) {
    if (p.peers.n == 0 || p.peers.flags[0] == nullptr) return;
    __syncthreads();                                   // every warp's stores and repair-count atomics are issued
    if (threadIdx.x == 0) {
        __threadfence_system();
        const unsigned t = atomicAdd(p.peers.ticket, 1u);
        if (t == gridDim.x - 1) {
            __threadfence_system();
            const unsigned long long c = *((volatile unsigned long long *)p.n_repaired);
#pragma unroll
            for (int g = 0; g < MAX_PEERS; ++g)
                if (g < p.peers.n) p.peers.part[g][p.peers.counter_slot] = __longlong_as_double((long long)c);
            *p.n_repaired = 0ull;
            *p.peers.ticket = 0u;
            __threadfence_system();
#pragma unroll
            for (int g = 0; g < MAX_PEERS; ++g)
                if (g < p.peers.n)
                    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p.peers.flags[g] + p.peers.rank), "l"(p.peers.epoch) : "memory");
        }
    }
}

constexpr long long PEER_WAIT_CYCLES = 30ll << 30;   // ~15 s: a rank that never arrives is an error, not a hang

// K2b prologue on such a handle: wait until every rank's flag has reached the epoch (ld.acquire.sys), i.e. until all
// leaf sums of the generation are in this rank's copy.  Thread 0 of block 0 adds the time it waited to *wait_ns.
struct PeerWait {
    const unsigned long long *flags;   // this rank's flag array, or null: nothing to wait for
    int n;
    unsigned long long epoch;
    unsigned long long *err;           // set to the epoch if a rank did not arrive within the time limit
    unsigned long long *wait_ns;       // accumulated wait time of block 0 (diagnostic: inter-rank skew)
};

__device__ __forceinline__ void peer_wait(const PeerWait &w) {
    if (w.flags == nullptr) return;
    unsigned long long t_begin = 0;
    if (blockIdx.x == 0 && threadIdx.x == 0) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_begin));
    if ((int)threadIdx.x < w.n) {
        const long long t0 = clock64();
        for (;;) {
            unsigned long long v;
            asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(w.flags + threadIdx.x) : "memory");
            if (v >= w.epoch) break;
            if (clock64() - t0 > PEER_WAIT_CYCLES) { *w.err = w.epoch; break; }
            __nanosleep(100);
        }
    }
    __syncthreads();                                   // every rank's flag has been seen
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        unsigned long long t_end;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_end));
        atomicAdd(w.wait_ns, t_end - t_begin);
    }
}

// Per-item state: where the four input rows and the output row of this trial live.
struct RowPtrs {
    const double *P, *B, *R1, *R2;
    double *W;
    const double *Uf, *Uc;
    uint32_t row_lo, row_hi;
    bool multi;              // hndvi > 1 with per-trial donors: R1, R2 are unused, see barycentres()
};

// 'random' direction with hndvi = h > 1 (de_oo.py:198-208,229-231): barycentre_k = (sum over the h rows
// r_k[0..h-1][i], added in that order as numpy's axis-0 reduction does) / h, for the genes of columns
// lc, lc+1 of trial `row`.
__device__ __forceinline__ void barycentres(const GenParams &p, int64_t row, int lc, bool two, double2 &x1, double2 &x2) {
    double2 s1 = make_double2(0.0, 0.0), s2 = s1;
    for (int k = 0; k < p.hndvi; ++k) {
        const int64_t a1 = p.r1[(int64_t)k * p.NP + row], a2 = p.r2[(int64_t)k * p.NP + row];
        const double *q1 = (p.sel_cur[a1] ? p.buf1 : p.buf0) + a1 * (int64_t)p.ld + lc;
        const double *q2 = (p.sel_cur[a2] ? p.buf1 : p.buf0) + a2 * (int64_t)p.ld + lc;
        const double2 v1 = make_double2(q1[0], two ? q1[1] : 0.0), v2 = make_double2(q2[0], two ? q2[1] : 0.0);
        if (k == 0) { s1 = v1; s2 = v2; }
        else {
            s1 = make_double2(dadd(s1.x, v1.x), dadd(s1.y, v1.y));
            s2 = make_double2(dadd(s2.x, v2.x), dadd(s2.y, v2.y));
        }
    }
    const double hd = (double)p.hndvi;
    x1 = make_double2(ddiv(s1.x, hd), ddiv(s1.y, hd));
    x2 = make_double2(ddiv(s2.x, hd), ddiv(s2.y, hd));
}

// de_oo.py:236-250 for one gene: Ft = U*F; mutant = beta + Ft*(x_r2 - x_r1); the parent's gene is
// kept iff Uc > Cr; then _correct_box_constraints (:254).
// Ft = fma(uf, Fa, Fb) with (Fa, Fb) = (F, 0) is uf*F rounded once, exactly the reference's Rand*F (:236-237);
// with (Fa, Fb) = (0, F) it is F, the 'constant' difference factor (:238-239) -- no select in the gene loop.
__device__ __forceinline__ double finish_gene(double xo, double xb, double x1, double x2, double uf, bool keep,
                                              double Fa, double Fb, double lo, double hi, unsigned &nrep) {
    const double v = keep ? xo : dadd(xb, dmul(__fma_rn(uf, Fa, Fb), dsub(x2, x1)));
    return repair_gene(v, lo, hi, nrep);
}

// Crossover decision from a Philox word without converting it: uc = w * 2^-32 exactly, so
// uc > Cr  <=>  w > Cr * 2^32  <=>  w > floor(Cr * 2^32) = cr_thresh (w is an integer).  Cr >= 1 or NaN never
// keeps (threshold 0xFFFFFFFF); Cr < 0 always keeps (cr_all).
__device__ __forceinline__ bool keep_parent(uint32_t w, uint32_t cr_thresh, bool cr_all) { return w > cr_thresh || cr_all; }

// One gene by the scalar path (tails, halo columns).  lc = local column.
template <int MODE, bool UNI>
__device__ __forceinline__ double gene_scalar(const RowPtrs &c, const GenParams &p, int lc, unsigned &nrep) {
    if (MODE == MODE_EVAL) return c.P[lc];
    const int col = p.col0 + lc;
    double uf;
    bool keep;
    if (MODE == MODE_PHILOX) {
        uint32_t w0 = (uint32_t)col >> 1, w1 = c.row_lo, w2 = c.row_hi | (TAG_ELEM << 24), w3 = p.gen;
        philox4x32_10(w0, w1, w2, w3, p.keys);
        uf = u32_to_unit((col & 1) ? w2 : w0);
        keep = keep_parent((col & 1) ? w3 : w1, p.cr_thresh, p.cr_all != 0);
    } else {
        uf = p.const_F ? 0.0 : c.Uf[col];
        keep = c.Uc[col] > p.Cr;
    }
    const double lo = UNI ? p.lo_s : p.lb[lc], hi = UNI ? p.hi_s : p.ub[lc];
    double x1, x2;
    if (c.multi) {
        double2 b1, b2;
        barycentres(p, ((int64_t)c.row_hi << 32) | c.row_lo, lc, false, b1, b2);
        x1 = b1.x; x2 = b2.x;
    } else { x1 = c.R1[lc]; x2 = c.R2[lc]; }
    return finish_gene(c.P[lc], c.B[lc], x1, x2, uf, keep, p.const_F ? 0.0 : p.F, p.const_F ? p.F : 0.0, lo, hi, nrep);
}

struct PairIn {
    double2 xo, xb, x1, x2;
};

template <int MODE>
__device__ __forceinline__ void load_pair(const RowPtrs &c, const GenParams &p, int lc, PairIn &in) {
    in.xo = *reinterpret_cast<const double2 *>(c.P + lc);
    if (MODE != MODE_EVAL) {
        in.xb = *reinterpret_cast<const double2 *>(c.B + lc);
        if (c.multi) barycentres(p, ((int64_t)c.row_hi << 32) | c.row_lo, lc, true, in.x1, in.x2);
        else {
            in.x1 = *reinterpret_cast<const double2 *>(c.R1 + lc);
            in.x2 = *reinterpret_cast<const double2 *>(c.R2 + lc);
        }
    }
}

// Two adjacent genes (lc even) from one Philox call.
template <int MODE, bool UNI>
__device__ __forceinline__ double2 gene_pair(const RowPtrs &c, const GenParams &p, int lc, const PairIn &in,
                                             unsigned &nrep) {
    if (MODE == MODE_EVAL) return in.xo;
    const int col = p.col0 + lc;
    double uf0, uf1;
    bool keep0, keep1;
    if (MODE == MODE_PHILOX) {
        uint32_t w0 = (uint32_t)col >> 1, w1 = c.row_lo, w2 = c.row_hi | (TAG_ELEM << 24), w3 = p.gen;
        philox4x32_10(w0, w1, w2, w3, p.keys);
        uf0 = u32_to_unit(w0); keep0 = keep_parent(w1, p.cr_thresh, p.cr_all != 0);
        uf1 = u32_to_unit(w2); keep1 = keep_parent(w3, p.cr_thresh, p.cr_all != 0);
    } else {
        keep0 = c.Uc[col] > p.Cr; keep1 = c.Uc[col + 1] > p.Cr;
        uf0 = p.const_F ? 0.0 : c.Uf[col];
        uf1 = p.const_F ? 0.0 : c.Uf[col + 1];
    }
    double lo0, lo1, hi0, hi1;
    if (UNI) { lo0 = lo1 = p.lo_s; hi0 = hi1 = p.hi_s; }
    else {
        const double2 l = *reinterpret_cast<const double2 *>(p.lb + lc);
        const double2 h = *reinterpret_cast<const double2 *>(p.ub + lc);
        lo0 = l.x; lo1 = l.y; hi0 = h.x; hi1 = h.y;
    }
    double2 r;
    const double Fa = p.const_F ? 0.0 : p.F, Fb = p.const_F ? p.F : 0.0;
    r.x = finish_gene(in.xo.x, in.xb.x, in.x1.x, in.x2.x, uf0, keep0, Fa, Fb, lo0, hi0, nrep);
    r.y = finish_gene(in.xo.y, in.xb.y, in.x1.y, in.x2.y, uf1, keep1, Fa, Fb, lo1, hi1, nrep);
    return r;
}

// K2a.  A QUAD of lanes owns one (row, leaf); lane m handles columns off+8k+2m, +1 of every
// 8-column step k, i.e. numpy's accumulator chains r[2m], r[2m+1]; inputs and the trial row move
// as 16-byte vectors (a quad reads/writes 64 contiguous bytes per row per step).
template <int OBJ, int MODE, bool UNI>
__global__ void __launch_bounds__(256, 2) de_trial_kernel(const GenParams p) {
    using O = Obj<OBJ>;
    constexpr int NSL = O::NS + O::NPR;
    const int lane4 = threadIdx.x & 3;
    const int qbase = threadIdx.x & 28;
    const unsigned qmask = 0xFu << qbase;
    const int64_t nquads = ((int64_t)gridDim.x * blockDim.x) >> 2;
    const int64_t nitems = (int64_t)p.nrows * p.nleaves;
    unsigned nrep = 0, nrep_dummy = 0;

    for (int64_t item = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 2; item < nitems; item += nquads) {
        int64_t lrow;
        int leaf;
        if (p.nleaves == 1) { lrow = item; leaf = 0; }
        else if (nitems < 0x7FFFFFFFll) {
            const unsigned q = (unsigned)item / (unsigned)p.nleaves;
            lrow = q; leaf = (int)((unsigned)item - q * (unsigned)p.nleaves);
        } else { lrow = item / p.nleaves; leaf = (int)(item - lrow * p.nleaves); }
        const int64_t row = p.row0 + lrow;
        const int2 lf = p.leaves[leaf];

        RowPtrs c;
        const uint8_t s = p.sel_cur[row];
        c.P = (s ? p.buf1 : p.buf0) + row * (int64_t)p.ld;
        c.W = (s ? p.buf0 : p.buf1) + row * (int64_t)p.ld;
        c.multi = false;
        if (MODE != MODE_EVAL) {
            // unconditional index / selector loads (the index arrays always hold valid rows), then pointer
            // selects: predicating the loads on the strategy costs their early scheduling
            const int64_t b = p.ib[row], a1 = p.r1[row], a2 = p.r2[row];
            const double *gB = (p.sel_cur[b] ? p.buf1 : p.buf0) + b * (int64_t)p.ld;
            const double *g1 = (p.sel_cur[a1] ? p.buf1 : p.buf0) + a1 * (int64_t)p.ld;
            const double *g2 = (p.sel_cur[a2] ? p.buf1 : p.buf0) + a2 * (int64_t)p.ld;
            c.B = p.base_vec ? p.base_vec : gB;
            c.R1 = p.d1_vec ? p.d1_vec : g1;
            c.R2 = p.d2_vec ? p.d2_vec : g2;
            c.multi = p.hndvi > 1 && !p.d1_vec;
        }
        if (MODE == MODE_REPLAY) {
            c.Uf = p.Uf + row * (int64_t)p.D;
            c.Uc = p.Uc + row * (int64_t)p.D;
        }
        c.row_lo = (uint32_t)row;
        c.row_hi = (uint32_t)((uint64_t)row >> 32);

        double a0[NSL], a1v[NSL];          // chains r[2m], r[2m+1]
#pragma unroll
        for (int q = 0; q < NSL; ++q) { a0[q] = slot_id<OBJ>(q); a1v[q] = slot_id<OBJ>(q); }
        const int nfull = lf.y >> 3;
        int lc = lf.x + 2 * lane4;

        auto step = [&](const PairIn &in, int lcs) {
            const double2 x = gene_pair<MODE, UNI>(c, p, lcs, in, nrep);
            if (MODE != MODE_EVAL) *reinterpret_cast<double2 *>(c.W + lcs) = x;
            double xn1 = 0.0;
            if (O::HALO) xn1 = gene_scalar<MODE, UNI>(c, p, lcs + 2, nrep_dummy);
            double auxa = 0.0, auxb = 0.0;
            if (O::AUX) { const double2 av = *reinterpret_cast<const double2 *>(p.aux + lcs); auxa = av.x; auxb = av.y; }
            double t0[NSL], t1[NSL];
            O::term(x.x, x.y, auxa, p.col0 + lcs, t0);
            O::term(x.y, xn1, auxb, p.col0 + lcs + 1, t1);
#pragma unroll
            for (int q = 0; q < NSL; ++q) {
                a0[q] = slot_op<OBJ>(q, a0[q], t0[q]);
                a1v[q] = slot_op<OBJ>(q, a1v[q], t1[q]);
            }
        };
        // two steps per trip on ping-pong register sets: the loads of step k+1 are in flight while
        // step k is computed
        PairIn A, B;
        int k = 0;
        if (nfull > 0) load_pair<MODE>(c, p, lc, A);
        for (; k + 1 < nfull; k += 2, lc += 16) {
            load_pair<MODE>(c, p, lc + 8, B);
            step(A, lc);
            if (k + 2 < nfull) load_pair<MODE>(c, p, lc + 16, A);
            step(B, lc + 8);
        }
        if (k < nfull) step(A, lc);
        // numpy: ((r0+r1)+(r2+r3))+((r4+r5)+(r6+r7)); lane m holds r[2m], r[2m+1].  With no full
        // step every chain still holds the identity, and so does the result.
        double acc[NSL];
#pragma unroll
        for (int q = 0; q < NSL; ++q) {
            acc[q] = slot_op<OBJ>(q, a0[q], a1v[q]);
            acc[q] = slot_op<OBJ>(q, acc[q], __shfl_xor_sync(qmask, acc[q], 1));
            acc[q] = slot_op<OBJ>(q, acc[q], __shfl_xor_sync(qmask, acc[q], 2));
        }
        const int tail = lf.y & 7;
        if (tail) {
            double t0[NSL], t1[NSL];
#pragma unroll
            for (int q = 0; q < NSL; ++q) { t0[q] = slot_id<OBJ>(q); t1[q] = slot_id<OBJ>(q); }
            const int tc = lf.x + 8 * nfull + 2 * lane4;
            if (2 * lane4 < tail) {
                const double x = gene_scalar<MODE, UNI>(c, p, tc, nrep);
                if (MODE != MODE_EVAL) c.W[tc] = x;
                const double xn = O::HALO ? gene_scalar<MODE, UNI>(c, p, tc + 1, nrep_dummy) : 0.0;
                O::term(x, xn, O::AUX ? p.aux[tc] : 0.0, p.col0 + tc, t0);
            }
            if (2 * lane4 + 1 < tail) {
                const double x = gene_scalar<MODE, UNI>(c, p, tc + 1, nrep);
                if (MODE != MODE_EVAL) c.W[tc + 1] = x;
                const double xn = O::HALO ? gene_scalar<MODE, UNI>(c, p, tc + 2, nrep_dummy) : 0.0;
                O::term(x, xn, O::AUX ? p.aux[tc + 1] : 0.0, p.col0 + tc + 1, t1);
            }
            for (int j = 0; j < tail; ++j) {
#pragma unroll
                for (int q = 0; q < NSL; ++q)
                    acc[q] = slot_op<OBJ>(q, acc[q], __shfl_sync(qmask, (j & 1) ? t1[q] : t0[q], qbase + (j >> 1)));
            }
        }
        // columns that carry no term of their own (Rosenbrock's last column)
        if (MODE != MODE_EVAL && O::HALO && leaf == p.nleaves - 1) {
            const int hc = p.nT + lane4;
            if (hc < p.Dl) c.W[hc] = gene_scalar<MODE, UNI>(c, p, hc, p.count_tail ? nrep : nrep_dummy);
        }
#pragma unroll
        for (int q = 0; q < NSL; ++q) put_partial<4>(p, (lrow * p.part_stride + leaf) * NSL + q, acc[q], lane4);
    }
    if (MODE != MODE_EVAL) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) nrep += __shfl_xor_sync(0xFFFFFFFFu, nrep, o);
        if ((threadIdx.x & 31) == 0 && nrep) atomicAdd(p.n_repaired, (unsigned long long)nrep);
    }
    peer_publish_tail(p);
}

// ---------------------------------------------------------------------------------------------
// Tree fold of the leaf sums.  prog is a post-order program: v >= 0 pushes leaf v, -1 combines
// the two topmost entries (left op right), mirroring pairwise_sum's recursion.
// ---------------------------------------------------------------------------------------------
// Where the leaf sums of a row live.  Single GPU: [NP][nleaves][NSL].  Column shards: rows are
// exchanged in C chunks of R rows; chunk c is an all-gather result [G][rows_c][Lmax][NSL] (the
// last chunk carries one extra slot per rank: that rank's bound-repair counter).
struct PartLayout {
    const double *base;
    const int *leaf_rank;        // owner rank of every global leaf
    const int *leaf_li;          // its index among the owner's leaves
    int R, C, G, stride;         // rows per chunk, chunks, ranks, doubles per row (Lmax*NSL)
    int64_t count_full, count_last;   // doubles per rank in a full chunk / in the last chunk
    // Fast fold (host-verified, fold_is_levelwise in oode.cu): the tree has <= 8 leaves and equals the level-wise
    // pairwise reduction (l0+l1),(l2+l3),.. then (0,2),(4,6), then (0,4) -- true for every balanced power-of-two
    // tree, e.g. the 8 leaves of a 1000-column row.  Then the leaf sums are folded in registers instead of by the
    // post-order interpreter (whose dynamically indexed stack lives in local memory).
    int fast_leaves;                  // 0: use the interpreter
    int8_t fast_rank[8], fast_li[8];  // owner rank / index among the owner's leaves of leaf 0..7
};

__device__ __forceinline__ const double *part_row(const PartLayout &L, int64_t row, int64_t &rank_stride) {
    const int c = (L.C == 1) ? 0 : (int)(row / L.R);
    const int64_t lr = row - (int64_t)c * L.R;
    rank_stride = (c == L.C - 1) ? L.count_last : L.count_full;
    return L.base + (int64_t)L.G * c * L.count_full + lr * L.stride;
}

template <int OBJ>
__device__ __forceinline__ double fold_slot(const PartLayout &L, const double *part, int64_t rank_stride, int nsl, int q,
                                            const int *prog, int nprog) {
    if (nprog == 1) return part[q];
    if (L.fast_leaves) {
        double v[8];
#pragma unroll
        for (int l = 0; l < 8; ++l)
            if (l < L.fast_leaves) v[l] = part[L.fast_rank[l] * rank_stride + L.fast_li[l] * nsl + q];
#pragma unroll
        for (int w = 1; w < 8; w <<= 1) {
#pragma unroll
            for (int i = 0; i + w < 8; i += 2 * w)
                if (i + w < L.fast_leaves) v[i] = slot_op<OBJ>(q, v[i], v[i + w]);
        }
        return v[0];
    }
    double st[40];
    int sp = 0;
    for (int i = 0; i < nprog; ++i) {
        const int v = prog[i];
        if (v >= 0) st[sp++] = part[L.leaf_rank[v] * rank_stride + L.leaf_li[v] * nsl + q];
        else { --sp; st[sp - 1] = slot_op<OBJ>(q, st[sp - 1], st[sp]); }
    }
    return st[0];
}

// Row shards (de_rows.cuh): K2b runs over this rank's row block only; instead of the Best bookkeeping its
// last block hands the block's result to every rank: summary = {trial-best f, its global row, accepted
// count, repaired genes} and the trial-best row itself (Best.x / next base vector candidates).
struct RowsOut {
    double *summary[MAX_PEERS];   // rank g's summary table [G][4] for this epoch; entry `rank` is ours
    double *cand[MAX_PEERS];      // rank g's candidate rows [G][ldc]
    int n, rank, ldc;
};

struct SelParams {
    PartLayout part;
    const int *prog;
    int nprog, nleaves;
    int n_rep_parts;             // column-sharded: number of per-rank repair counters to add up (else 0)
    const double *rep_parts;     // first counter (bit pattern of an u64 in a double slot)
    int64_t rep_stride;          // doubles between the counters of consecutive ranks
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
    int hndvi;               // r1, r2 hold hndvi index rows of NP entries each
    double *tbest_x;         // [Dl] x of this generation's trial-best row (base vector 'best'), or null
    RowsOut rows;            // row shards: where this rank's summary and candidate row go (n == 0 otherwise)
    // general constraints (null / unused on a box-bounded problem)
    const double *tcvals;    // [NP] constr_val of every trial row (de_constr_kernel)
    double *cvals;           // [NP] constr_val of the current population
    double contol;
    int *blk_cls;            // per-block class of the block's best row (0 feasible, 1 not)
    PeerWait wait;           // column shards, peer exchange: every rank's leaf sums must have arrived first
};

__device__ __forceinline__ bool better_fi(double fa, int64_t ia, double fb, int64_t ib) {
    return fa < fb || (fa == fb && ia < ib);      // first-occurrence argmin (numpy argmin)
}
// _eval_pop's best row under general constraints (de_oo.py:318-338): any row with constr_val < contol
// (class 0, keyed by f) beats every other row (class 1, keyed by constr_val); first occurrence within a class.
__device__ __forceinline__ bool better_key(int ca, double fa, int64_t ia, int cb, double fb, int64_t ib) {
    return ca < cb || (ca == cb && better_fi(fa, ia, fb, ib));
}

// CON: general constraints are attached (tcvals / cvals valid); the box-bounded instantiation carries none
// of the class bookkeeping.
template <int OBJ, bool CON>
__global__ void __launch_bounds__(256) de_select_kernel(const SelParams p) {
    using O = Obj<OBJ>;
    constexpr int NSL = O::NS + O::NPR;
    const int64_t lrow = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int64_t row = p.row0 + lrow;
    double f = __longlong_as_double(0x7FF0000000000000LL);
    int64_t fi = INT64_MAX;
    int cls = CON ? 1 : 0;
    unsigned acc = 0;
    peer_wait(p.wait);
    if (lrow < p.nrows) {
        int64_t rank_stride;
        const double *part = part_row(p.part, lrow, rank_stride);
        double s[NSL];
#pragma unroll
        for (int q = 0; q < NSL; ++q) s[q] = fold_slot<OBJ>(p.part, part, rank_stride, NSL, q, p.prog, p.nprog);
        f = O::finalize(s, p.D);
        if (p.invert) f = -f;                                     // kernel/nonLinFuncs.py:297-298
        if (f != f) f = __longlong_as_double(0x7FF0000000000000LL); // de_oo.py:301
        fi = row;
        p.tvals[row] = f;
        const double nc = CON ? p.tcvals[row] : 0.0;
        if (p.init) {
            p.vals[row] = f;
            if (CON) p.cvals[row] = nc;
        } else {
            // de_oo.py:260-282.  constr_vals == 0 on both sides (always so without general constraints): keep
            // the parent iff old_f < trial_f; if either side violates anything: iff old_c < trial_c.  vals (and
            // constr_vals) are blended arithmetically like the reference so inf*0 = NaN appears where it does there.
            const double old = p.vals[row];
            bool keep = old < f;
            if (CON) {
                const double oc = p.cvals[row];
                if (nc > 0.0 || oc > 0.0) keep = oc < nc;
                const double kc = keep ? 1.0 : 0.0;
                p.cvals[row] = dadd(dmul(oc, kc), dmul(nc, dsub(1.0, kc)));
            }
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
            for (int k = 1; k < p.hndvi; ++k) {           // further index rows of a wider barycentre
                uint32_t c0 = 1u + (uint32_t)k, c1 = (uint32_t)row, c2 = (uint32_t)((uint64_t)row >> 32) | (TAG_INDEX << 24),
                         c3 = p.next_gen;
                philox4x32_10(c0, c1, c2, c3, p.keys);
                p.r1[(int64_t)k * p.NP + row] = (int32_t)mulhi64_index(c0, c1, (uint64_t)p.NP);
                p.r2[(int64_t)k * p.NP + row] = (int32_t)mulhi64_index(c2, c3, (uint64_t)p.NP);
            }
        }
    }
    // ---- block argmin + accepted count ----
    // The comparison key of a row is its f, or its constr_val if it is not feasible (class 1).
    if (CON && cls == 1 && lrow < p.nrows) {
        const double nc = p.tcvals[row];
        if (nc < p.contol) cls = 0;
        else f = nc;
    }
    __shared__ double sh_f[8];
    __shared__ int64_t sh_i[8];
    __shared__ int sh_c[8];
    __shared__ unsigned sh_a[8];
    __shared__ bool sh_last;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const double of = __shfl_xor_sync(0xFFFFFFFFu, f, o);
        const int64_t oi = __shfl_xor_sync(0xFFFFFFFFu, fi, o);
        const int oc = CON ? __shfl_xor_sync(0xFFFFFFFFu, cls, o) : 0;
        if (better_key(oc, of, oi, cls, f, fi)) { f = of; fi = oi; cls = oc; }
        acc += __shfl_xor_sync(0xFFFFFFFFu, acc, o);
    }
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) { sh_f[warp] = f; sh_i[warp] = fi; sh_c[warp] = cls; sh_a[warp] = acc; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < 8; ++w) {
            if (better_key(sh_c[w], sh_f[w], sh_i[w], cls, f, fi)) { f = sh_f[w]; fi = sh_i[w]; cls = sh_c[w]; }
            acc += sh_a[w];
        }
        p.blk_f[blockIdx.x] = f;
        p.blk_i[blockIdx.x] = fi;
        if (CON) p.blk_cls[blockIdx.x] = cls;
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
    cls = CON ? 1 : 0;
    unsigned long long tot = 0;
    for (unsigned b = threadIdx.x; b < gridDim.x; b += blockDim.x) {
        const double bf = ((volatile double *)p.blk_f)[b];
        const int64_t bi = ((volatile int64_t *)p.blk_i)[b];
        const int bc = CON ? ((volatile int *)p.blk_cls)[b] : 0;
        if (better_key(bc, bf, bi, cls, f, fi)) { f = bf; fi = bi; cls = bc; }
        tot += ((volatile unsigned long long *)p.blk_acc)[b];
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const double of = __shfl_xor_sync(0xFFFFFFFFu, f, o);
        const int64_t oi = __shfl_xor_sync(0xFFFFFFFFu, fi, o);
        const int oc = CON ? __shfl_xor_sync(0xFFFFFFFFu, cls, o) : 0;
        if (better_key(oc, of, oi, cls, f, fi)) { f = of; fi = oi; cls = oc; }
        tot += __shfl_xor_sync(0xFFFFFFFFu, tot, o);
    }
    __shared__ unsigned long long sh_t[8];
    __shared__ int sh_changed;
    __syncthreads();
    if (lane == 0) { sh_f[warp] = f; sh_i[warp] = fi; sh_c[warp] = cls; sh_t[warp] = tot; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int w = 1; w < 8; ++w) {
            if (better_key(sh_c[w], sh_f[w], sh_i[w], cls, f, fi)) { f = sh_f[w]; fi = sh_i[w]; cls = sh_c[w]; }
            tot += sh_t[w];
        }
        if (p.rows.n) {                       // row shards: de_rows_finish_kernel does the Best bookkeeping
            const unsigned long long nrep = *p.n_repaired;
            for (int g = 0; g < p.rows.n; ++g) {
                double *sm = p.rows.summary[g] + 4 * p.rows.rank;
                sm[0] = f;
                sm[1] = __longlong_as_double(fi);
                sm[2] = __longlong_as_double((long long)tot);
                sm[3] = __longlong_as_double((long long)nrep);
            }
            *p.n_repaired = 0ull;
            *p.ticket = 0u;
            sh_changed = -1;
            sh_i[0] = fi;
        } else {
        // Best tracking (de_oo.py:257,285-286; kernel/Point.py:309-333): Old_best survives only if
        // its f is strictly smaller, so a tie hands Best to the newer trial-best.
        // With general constraints the winner's key may be its constr_val: fetch its f; Best itself is then
        // tracked by the host (Point.betterThan needs the host's max-residual), so every generation "changes".
        double win_c = 0.0;
        if (CON) {
            f = ((volatile double *)p.tvals)[fi];
            win_c = ((volatile const double *)p.tcvals)[fi];
        }
        const double old_best = *p.best_f;
        const int changed = (p.init || CON) ? 1 : !(old_best < f);
        if (changed) *p.best_f = f;
        oode_gen_record r;
        r.trial_best_f = f;
        r.best_f = changed ? f : old_best;
        r.trial_best_idx = fi;
        r.trial_best_c = win_c;
        r.trial_best_feasible = (cls == 0) ? 1 : 0;
        r.reserved = 0;
        r.n_accepted = (int64_t)tot;
        unsigned long long nrep_total = *p.n_repaired;
        for (int g = 0; g < p.n_rep_parts; ++g) nrep_total += (unsigned long long)__double_as_longlong(p.rep_parts[g * p.rep_stride]);
        r.n_repaired = (int64_t)nrep_total;
        r.best_changed = changed;
        r.generation = p.generation;
        *p.rec = r;
        *p.n_repaired = 0ull;
        *p.ticket = 0u;
        sh_changed = changed;
        sh_i[0] = fi;
        }
    }
    __syncthreads();
    // Best.x: the trial-best row sits in the spare slot of its row (trials are always written);
    // for the initial population it is the current slot.
    const int64_t bi = sh_i[0];
    const double *src;
    if (p.init) src = p.buf0 + bi * p.ld;
    else src = (p.sel_cur[bi] ? p.buf0 : p.buf1) + bi * p.ld;
    if (sh_changed < 0) {                     // row shards: this block's trial-best row, to every rank
        for (int g = 0; g < p.rows.n; ++g) {
            double *dst = p.rows.cand[g] + (int64_t)p.rows.rank * p.rows.ldc;
            for (int j = threadIdx.x; j < p.Dl; j += blockDim.x) dst[j] = src[j];
        }
        return;
    }
    if (sh_changed) {
        for (int j = threadIdx.x; j < p.Dl; j += blockDim.x) {
            const double v = src[j];
            p.best_x[j] = v;
            p.bestx_slot[j] = v;
        }
    } else {
        for (int j = threadIdx.x; j < p.Dl; j += blockDim.x) p.bestx_slot[j] = p.best_x[j];
    }
    // best[2] of this _eval_pop (de_oo.py:153,256): next generation's base vector under 'best' (:192)
    if (p.tbest_x)
        for (int j = threadIdx.x; j < p.Dl; j += blockDim.x) p.tbest_x[j] = src[j];
}

// objective of arbitrary points (oode_eval_points): fold + finalize only
template <int OBJ>
__global__ void de_finalize_kernel(const PartLayout part, const int *prog, int nprog, int D, int invert, int64_t n, double *out) {
    using O = Obj<OBJ>;
    constexpr int NSL = O::NS + O::NPR;
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    double s[NSL];
    int64_t rank_stride;
    const double *pr = part_row(part, i, rank_stride);
#pragma unroll
    for (int q = 0; q < NSL; ++q) s[q] = fold_slot<OBJ>(part, pr, rank_stride, NSL, q, prog, nprog);
    double f = O::finalize(s, D);
    out[i] = invert ? -f : f;
}

}  // namespace oode
