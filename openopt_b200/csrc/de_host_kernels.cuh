// Kernels without an objective template parameter: directed-search barycentres, general constraints,
// population init, the peer-exchange flag kernels and the row gather of the verification taps.
// Included by oode.cu only (they are ordinary __global__ functions: one definition per library).
#pragma once
#include "de_kernels.cuh"

namespace oode {

// ---------------------------------------------------------------------------------------------
// Directed search (searchDirectionStrategy 'best', de_oo.py:209-231): ONE set of 2h row indices for
// the whole population; the rows are ordered by (constr_val, val, index) -- numpy's structured sort
// with order=['constr_val','val'] breaks ties with the remaining field, NaN sorts last -- the first h
// are the 'best', the next h the 'worst'; barycentre = (their rows added in that order) / h, and
// divided by h once more when h != 1 (the reference does both :222-223 and :229-231).  The two
// (D,) vectors are what every trial of the generation uses as x_r1 (worst) and x_r2 (best).
// One CTA; constr_val == 0 on this (box-bounded) path.
// ---------------------------------------------------------------------------------------------
struct DirectParams {
    const int32_t *r;            // [2h] indices (replay / CPython stream), or null: draw them with Philox
    const double *vals;          // post-selection f of the old population
    const double *cvals;         // its constr_vals, or null (all zero)
    const uint8_t *sel_cur;
    const double *buf0, *buf1;
    int64_t ld, NP;
    int Dl, h;
    uint32_t gen;
    PhiloxKeys keys;
    double *d1_vec, *d2_vec;     // out: worst / best barycentre, [Dl]
};

// DOUBLE_compare: NaN is larger than everything, two NaNs are equal.  Returns -1, 0, 1.
__device__ __forceinline__ int direct_cmp(double a, double b) {
    if (a < b) return -1;
    if (a > b) return 1;
    if (a == b) return 0;
    const bool na = a != a, nb = b != b;
    if (na && nb) return 0;
    return nb ? -1 : 1;          // exactly one is NaN: the other one is smaller
}
__device__ __forceinline__ bool direct_less(double ca, double fa, int ia, double cb, double fb, int ib) {
    int c = direct_cmp(ca, cb);                   // order = ['constr_val', 'val'], then the index field
    if (c == 0) c = direct_cmp(fa, fb);
    return c < 0 || (c == 0 && ia < ib);
}

__global__ void __launch_bounds__(256) de_direct_kernel(const DirectParams p) {
    __shared__ int s_idx[2 * OODE_MAX_HNDVI];
    __shared__ double s_val[2 * OODE_MAX_HNDVI];
    __shared__ double s_con[2 * OODE_MAX_HNDVI];
    const int n = 2 * p.h;
    for (int m = threadIdx.x; m < n; m += blockDim.x) {
        int j;
        if (p.r) j = p.r[m];
        else {
            uint32_t w0 = 0u, w1 = (uint32_t)m, w2 = (TAG_DIRECT << 24), w3 = p.gen;
            philox4x32_10(w0, w1, w2, w3, p.keys);
            j = (int)mulhi64_index(w0, w1, (uint64_t)p.NP);
        }
        s_idx[m] = j;
        s_val[m] = p.vals[j];
        s_con[m] = p.cvals ? p.cvals[j] : 0.0;
    }
    __syncthreads();
    if (threadIdx.x == 0) {                      // insertion sort of <= 1024 records
        for (int a = 1; a < n; ++a) {
            const int ji = s_idx[a];
            const double jf = s_val[a], jc = s_con[a];
            int b = a - 1;
            while (b >= 0 && direct_less(jc, jf, ji, s_con[b], s_val[b], s_idx[b])) {
                s_idx[b + 1] = s_idx[b];
                s_val[b + 1] = s_val[b];
                s_con[b + 1] = s_con[b];
                --b;
            }
            s_idx[b + 1] = ji;
            s_val[b + 1] = jf;
            s_con[b + 1] = jc;
        }
    }
    __syncthreads();
    const double hd = (double)p.h;
    for (int c = threadIdx.x; c < p.Dl; c += blockDim.x) {
        double sb = 0.0, sw = 0.0;
        for (int k = 0; k < p.h; ++k) {
            const int64_t jb = s_idx[k], jw = s_idx[p.h + k];
            const double xb = (p.sel_cur[jb] ? p.buf1 : p.buf0)[jb * p.ld + c];
            const double xw = (p.sel_cur[jw] ? p.buf1 : p.buf0)[jw * p.ld + c];
            sb = k ? dadd(sb, xb) : xb;
            sw = k ? dadd(sw, xw) : xw;
        }
        sb = ddiv(sb, hd);
        sw = ddiv(sw, hd);
        if (p.h != 1) { sb = ddiv(sb, hd); sw = ddiv(sw, hd); }
        p.d1_vec[c] = sw;
        p.d2_vec[c] = sb;
    }
}

// ---------------------------------------------------------------------------------------------
// General constraints: constr_val of every trial row (de_oo.py:314-316; kernel/Point.py:128-167,363-374):
//   max(0, A x - b, c(x), |Aeq x - beq|, |h(x)|)  (NaN residuals skipped, as nanmax does)
//   + 1e10 * (number of NaN among c(x), h(x)),   c, h = diagonal quadratics sum_j Q[k][j] x_j^2 - r[k].
// One warp per row; a residual is accumulated by the 32 lanes over strided columns (fma) and folded with
// shuffles.  The reference forms the linear residuals with a BLAS product whose summation order is
// implementation defined, so these values are reproducible to rounding (1e-12 relative in the tests), and
// exactly whenever a constraint row has at most two non-zero coefficients.
// ---------------------------------------------------------------------------------------------
struct ConstrParams {
    const double *buf0, *buf1;
    const uint8_t *sel_cur;
    int64_t ld, NP;
    int D, init;                 // init: evaluate the current slot (initial population), else the spare slot
    int m[4];                    // rows of A, Aeq, Qc, Qh
    const double *M[4];          // the four coefficient matrices, row-major [m][D]
    const double *rhs[4];        // b, beq, rc, rh
    double *out;                 // [NP]
};

__global__ void __launch_bounds__(256) de_constr_kernel(const ConstrParams p) {
    const int lane = threadIdx.x & 31;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; row < p.NP; row += nwarps) {
        const uint8_t s = p.sel_cur[row];
        const double *x = (p.init ? p.buf0 : (s ? p.buf0 : p.buf1)) + row * p.ld;
        double r = 0.0;
        int nnan = 0;
#pragma unroll
        for (int kind = 0; kind < 4; ++kind) {           // 0: A x <= b, 1: Aeq x = beq, 2: c(x) <= 0, 3: h(x) = 0
            for (int k = 0; k < p.m[kind]; ++k) {
                const double *coef = p.M[kind] + (int64_t)k * p.D;
                double acc = 0.0;
                for (int j = lane; j < p.D; j += 32) {
                    const double xj = x[j];
                    acc = __fma_rn(coef[j], kind >= 2 ? dmul(xj, xj) : xj, acc);
                }
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) acc = dadd(acc, __shfl_xor_sync(0xFFFFFFFFu, acc, o));
                double res = dsub(acc, p.rhs[kind][k]);
                if (kind & 1) res = fabs(res);
                if (res != res) { if (kind >= 2) ++nnan; }
                else if (r < res) r = res;
            }
        }
        if (lane == 0) p.out[row] = dadd(r, dmul(1.0e10, (double)nnan));
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

// One thread per COLUMN PAIR: one Philox call serves both genes (w0 -> even column, w2 -> odd column, the same
// words the generation draws use for Uf), and the pair is stored as one 16-byte vector.
__global__ void __launch_bounds__(256) de_init_kernel(const InitParams p) {
    const int pairs = (p.Dl + 1) >> 1;
    const int64_t n = p.NP * pairs;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x) {
        int64_t row;
        int pr;
        if (n < 0x7FFFFFFFll) { const unsigned q = (unsigned)e / (unsigned)pairs; row = q; pr = (int)((unsigned)e - q * (unsigned)pairs); }
        else { row = e / pairs; pr = (int)(e - row * pairs); }
        const int lc = 2 * pr;
        const bool two = lc + 1 < p.Dl;
        double u0, u1;
        if (p.use_philox) {
            const int col = p.col0 + lc;                       // col0 is even
            uint32_t w0 = (uint32_t)col >> 1, w1 = (uint32_t)row, w2 = (uint32_t)((uint64_t)row >> 32) | (TAG_ELEM << 24), w3 = 0u;
            philox4x32_10(w0, w1, w2, w3, p.keys);
            u0 = u32_to_unit(w0);
            u1 = u32_to_unit(w2);
        } else {
            const double *U = p.U0 + row * p.D + p.col0 + lc;
            u0 = U[0];
            u1 = two ? U[1] : 0.0;
        }
        const int lc1 = two ? lc + 1 : lc;
        double v0 = dadd(dmul(u0, dsub(p.ub[lc], p.lb[lc])), p.lb[lc]);          // de_oo.py:147 (mul, then add)
        double v1 = dadd(dmul(u1, dsub(p.ub[lc1], p.lb[lc1])), p.lb[lc1]);
        if (row == 0 && p.x0) { v0 = p.x0[lc]; v1 = p.x0[lc1]; }                  // de_oo.py:149-150
        double *dst = p.buf0 + row * p.ld + lc;
        if (two) *reinterpret_cast<double2 *>(dst) = make_double2(v0, v1);
        else dst[0] = v0;
        if (pr == 0) { p.sel0[row] = 0; p.sel1[row] = 0; }
    }
}

// NCCL exchange: moves this rank's repair counter into the tail of its all-gather payload and clears it
__global__ void de_pack_counter_kernel(unsigned long long *counter, double *tail) {
    *tail = __longlong_as_double((long long)*counter);
    *counter = 0ull;
}

// ---------------------------------------------------------------------------------------------
// Peer exchange, end of a rank's K2a: hand the leaf sums over.  The trial kernel has already stored
// them into every rank's copy; this one-warp kernel (next in stream order, so all of those stores
// have been performed) appends this rank's bound-repair counter to its block on every rank and then
// raises this rank's epoch flag there (release at system scope).  de_wait_kernel, launched right
// before K2b, spins until every rank's flag has reached the epoch (acquire at system scope).
// Epochs only grow and the blocks are double-buffered by epoch parity: a rank can be at most one
// epoch ahead of the slowest one, so nothing a reader still needs is ever overwritten.
// ---------------------------------------------------------------------------------------------
struct PeerSync {
    double *part[MAX_PEERS];                  // this rank's block on rank g (current epoch's buffer)
    unsigned long long *flags[MAX_PEERS];     // rank g's flag array [MAX_PEERS]
    int64_t counter_slot;                     // index of the repair counter inside a block
    int n, rank;
    unsigned long long epoch;
};

__global__ void de_publish_kernel(const PeerSync s, unsigned long long *counter) {
    const int j = threadIdx.x;
    const unsigned long long c = *counter;
    __syncwarp();
    if (j < s.n) {
        if (s.counter_slot >= 0) s.part[j][s.counter_slot] = __longlong_as_double((long long)c);
        __threadfence_system();
        asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(s.flags[j] + s.rank), "l"(s.epoch) : "memory");
    }
    if (j == 0 && s.counter_slot >= 0) *counter = 0ull;
}


__global__ void de_wait_kernel(const unsigned long long *flags, int n, unsigned long long epoch,
                               unsigned long long *err) {
    const int j = threadIdx.x;
    if (j < n) {
        const long long t0 = clock64();
        for (;;) {
            unsigned long long v;
            asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(flags + j) : "memory");
            if (v >= epoch) break;
            if (clock64() - t0 > PEER_WAIT_CYCLES) { *err = epoch; break; }
            __nanosleep(200);
        }
    }
    __threadfence_system();
}

// Column shards, peer exchange: ONE warp waits for every rank's flag (peer_wait, de_kernels.cuh) in front of K2b and
// accounts the time it waited.  (Waiting inside K2b's prologue instead -- every one of its ~4000 CTAs polling the
// flags at system scope -- was measured to DELAY the arrival of the flags: the slower rank's remote stores, flag
// included, got through 0.14 ms late at two GPUs while the faster rank's SMs were all spinning.)
__global__ void de_peer_wait_kernel(const PeerWait w) { peer_wait(w); }

// oode_get_rows: a warp copies one requested row (picked slot) into a dense staging array.
// which: 0 current, 1 spare, 2 last trial (accepted ? current : spare), 3 last parent (accepted ? spare : current)
__global__ void __launch_bounds__(256) de_gather_rows_kernel(const double *buf0, const double *buf1, const uint8_t *sel,
                                                             const uint8_t *accept, int64_t ld, int Dl, int which,
                                                             const int64_t *rows, int64_t n, double *out) {
    const int lane = threadIdx.x & 31;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t k = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5; k < n; k += nwarps) {
        const int64_t r = rows[k];
        int slot = sel[r];
        const int acc = accept[r];
        if (which == 1 || (which == 2 && !acc) || (which == 3 && acc)) slot ^= 1;
        const double *src = (slot ? buf1 : buf0) + r * ld;
        for (int j = lane; j < Dl; j += 32) out[k * Dl + j] = src[j];
    }
}

}  // namespace oode
