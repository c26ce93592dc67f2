// Row shards (OODE_SHARD_ROWS): for rows too narrow to column-shard (fewer pairwise-sum leaves than GPUs,
// e.g. the 100-column rows of BASELINE configs[4]) and for objectives that couple neighbouring columns.
//
// Every rank holds the whole population (both slot arrays, selectors, vals) -- donors come from anywhere --
// but builds and evaluates only the trials of ITS contiguous row block (K2a, K2b with row0/nrows).  What the
// other ranks need afterwards is exactly what changed: the rows whose trial was accepted.  There is no
// all-gather: de_rows_push_kernel copies each accepted trial row from the local spare slot straight into
// the SAME spare slot of every other rank's replica (CUDA-IPC mapped peer memory, NVLink), and each row's
// accept flag / new f / trial f into a per-epoch mailbox on every rank.  After the epoch-flag barrier
// (de_publish_kernel / de_wait_kernel) de_rows_finish_kernel applies the mailbox to the rows a rank does
// not own (selector flip, vals) and reduces the per-rank summaries into the generation record and Best.
//
// Replicas stay bit-identical: draws are keyed by the global row, a row's spare slot is the same slot on
// every rank, and a rank only ever writes the spare slots of its own rows (nobody reads those before the
// barrier).  Mailboxes are double-buffered by epoch parity like the column shards' leaf-sum blocks.
#pragma once
#include "de_kernels.cuh"

namespace oode {

struct RowsMailbox {              // one epoch's mailbox inside a rank's arena
    double *vals;                 // [NP] post-selection f
    double *tvals;                // [NP] trial f
    uint8_t *acc;                 // [NP] 1 = trial accepted
    double *summary;              // [G][4]
    double *cand;                 // [G][ldc]
};

struct PushParams {
    RowsMailbox mb[MAX_PEERS];    // every rank's mailbox (this epoch)
    double *peer_buf0[MAX_PEERS]; // every rank's slot arrays (own entry = local)
    double *peer_buf1[MAX_PEERS];
    int n, rank;
    const double *buf0, *buf1;
    const uint8_t *sel_cur;       // selectors BEFORE this generation's flips
    const uint8_t *accept;
    const double *vals, *tvals;
    int64_t ld;
    int Dl;
    int64_t row0, nrows;
    int init;                     // 1: initial population -- nothing was accepted, only vals travel
    int rows_sent;                // 1: the trial kernel has already pushed the accepted rows (RowPush): results only
};

// One warp per row of this rank's block.
__global__ void __launch_bounds__(256) de_rows_push_kernel(const PushParams p) {
    const int lane = threadIdx.x & 31;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t lrow = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; lrow < p.nrows; lrow += nwarps) {
        const int64_t row = p.row0 + lrow;
        const uint8_t a = p.init ? (uint8_t)0 : p.accept[row];
        if (lane < p.n && lane != p.rank) {
            const RowsMailbox &m = p.mb[lane];
            m.vals[row] = p.vals[row];
            m.tvals[row] = p.tvals[row];
            m.acc[row] = a;
        }
        if (a && !p.rows_sent) {
            const uint8_t s = p.sel_cur[row];
            const double2 *src = reinterpret_cast<const double2 *>((s ? p.buf0 : p.buf1) + row * p.ld);   // spare slot
            for (int j = lane; 2 * j < p.Dl; j += 32) {
                const double2 v = src[j];
                for (int g = 0; g < p.n; ++g) {
                    if (g == p.rank) continue;
                    reinterpret_cast<double2 *>((s ? p.peer_buf0[g] : p.peer_buf1[g]) + row * p.ld)[j] = v;
                }
            }
        }
    }
}

struct FinishParams {
    RowsMailbox mb;               // local mailbox (this epoch)
    int G, rank, ldc;
    int64_t NP, row0, nrows;      // own block: [row0, row0 + nrows)
    int Dl, init;
    double *vals, *tvals;
    uint8_t *accept;
    const uint8_t *sel_cur;
    uint8_t *sel_next;
    double *best_f, *best_x, *bestx_slot, *tbest_x;
    oode_gen_record *rec;
    int32_t generation;
};

__global__ void __launch_bounds__(256) de_rows_finish_kernel(const FinishParams p) {
    // rows owned by the other ranks: take over their decisions
    for (int64_t row = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; row < p.NP; row += (int64_t)gridDim.x * blockDim.x) {
        if (row >= p.row0 && row < p.row0 + p.nrows) continue;
        const uint8_t a = p.mb.acc[row];
        p.vals[row] = p.mb.vals[row];
        p.tvals[row] = p.mb.tvals[row];
        p.accept[row] = a;
        if (!p.init) p.sel_next[row] = (uint8_t)(p.sel_cur[row] ^ a);
    }
    if (blockIdx.x != 0) return;
    // block 0: the generation record from the G block summaries (first-occurrence argmin: lowest row wins)
    __shared__ int sh_win, sh_changed;
    if (threadIdx.x == 0) {
        double f = p.mb.summary[0];
        int64_t fi = __double_as_longlong(p.mb.summary[1]);
        long long acc = __double_as_longlong(p.mb.summary[2]), rep = __double_as_longlong(p.mb.summary[3]);
        int win = 0;
        for (int g = 1; g < p.G; ++g) {
            const double gf = p.mb.summary[4 * g];
            const int64_t gi = __double_as_longlong(p.mb.summary[4 * g + 1]);
            if (better_fi(gf, gi, f, fi)) { f = gf; fi = gi; win = g; }
            acc += __double_as_longlong(p.mb.summary[4 * g + 2]);
            rep += __double_as_longlong(p.mb.summary[4 * g + 3]);
        }
        const double old_best = *p.best_f;
        const int changed = p.init ? 1 : !(old_best < f);     // de_oo.py:285-286: a tie goes to the newer trial-best
        if (changed) *p.best_f = f;
        oode_gen_record r;
        r.trial_best_f = f;
        r.best_f = changed ? f : old_best;
        r.trial_best_idx = fi;
        r.n_accepted = acc;
        r.n_repaired = rep;
        r.best_changed = changed;
        r.generation = p.generation;
        r.trial_best_c = 0.0;
        r.trial_best_feasible = 1;
        r.reserved = 0;
        *p.rec = r;
        sh_win = win;
        sh_changed = changed;
    }
    __syncthreads();
    const double *src = p.mb.cand + (int64_t)sh_win * p.ldc;
    for (int j = threadIdx.x; j < p.Dl; j += blockDim.x) {
        const double v = src[j];
        if (sh_changed) { p.best_x[j] = v; p.bestx_slot[j] = v; }
        else p.bestx_slot[j] = p.best_x[j];
        if (p.tbest_x) p.tbest_x[j] = v;
    }
}

}  // namespace oode
