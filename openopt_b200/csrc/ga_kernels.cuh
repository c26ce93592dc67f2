// Device side of the `galileo` genetic algorithm (include/ooga.h): one generation of
// openopt/solvers/Standalone/galileo_oo.py:73-92 on a population resident in HBM.  All line numbers below refer to that
// file; oracle/galileo_oracle.py is the CPU restatement these kernels are checked against.
//
//   ga_init_kernel      prepPopulation :237-253          genes = lb + (ub-lb)*u
//   ga_scan_*_kernel    evaluate :255-276                running fitness sums IN LIST ORDER (the reference accumulates left
//                                                        to right, and the roulette compares against those very sums), first
//                                                        strictly-largest fitness; large populations: the serial chain
//                                                        re-expressed as integer prefix sums per binade, bit-exact
//   ga_select_kernel    select :290-301 + select_Roulette :326-346, the pairs' crossover draws :360-363 (a thread per spin)
//   ga_children_kernel  crossover_OnePoint :354-369 + mutate_Default :425-441 for the fresh children (one warp per pair, a
//                                                        lane per column pair; row hashes summed on the way)
//   ga_alias_kernel     mutate_Default on entries that were not crossed: they are the current chromosomes themselves and
//                                                        are changed in place (last writer in list order wins)
//   ga_table_insert_kernel + ga_dedupe_kernel            the isIdentical scan of replace_SteadyStateNoDuplicates :465-471 as a
//                                                        hash join with an exact row comparison (row hashes: from
//                                                        ga_children_kernel for the entries, carried along by
//                                                        ga_gather_kernel for the survivors, ga_hash_kernel after ooga_init
//                                                        and after in-place mutations)
//   ga_keys_kernel + ga_bitonic_* (+ ga_merge_rank_kernel) + ga_gather_kernel    sort / reverse / truncate :472-474
//   ga_record_kernel    what __solver__ reports :86-89
//
// HBM-bound byte work; no tensor cores: nothing here is a contraction.  Objective values come from the de library's
// row evaluator (MODE_EVAL of the trial kernels, numpy's pairwise summation order), launched by the host code
// (ooga_host.cuh) between the mutation and the duplicate test.
#pragma once
#include <cstdint>
#include <climits>
#include <cuda_runtime.h>
#include "de_device.cuh"

namespace ooga {

using oode::PhiloxKeys;
using oode::philox4x32_10;
using oode::u32_to_unit;

constexpr uint32_t TAG_GA = 3u;            // Philox counter tag of the per-spin / per-pair draws (tags 0..2 belong to de)
constexpr int ENTRY_ALIAS = 0, ENTRY_FRESH = 1;

// CPython's random(): 53 bits from two 32-bit words
__device__ __forceinline__ double u53(uint32_t a, uint32_t b) {
    return __dmul_rn(__dadd_rn(__dmul_rn((double)(a >> 5), 67108864.0), (double)(b >> 6)), 1.0 / 9007199254740992.0);
}

// uniform(lo, hi) = lo + (hi - lo) * random()        (Lib/random.py; :127-133, :439)
__device__ __forceinline__ double uniform_gene(double lo, double hi, double u) {
    return __dadd_rn(lo, __dmul_rn(__dsub_rn(hi, lo), u));
}

struct GaParams {
    int64_t N;              // population (numChromosomes = replacementSize, :57)
    int64_t P;              // pairs per generation = ceil(N/2) (:295, :309)
    int D, ld;              // genes per chromosome, row pitch in doubles
    const double *lb, *ub;
    double cr, mr;          // crossoverRate, mutationRate
    uint32_t gen;
    int philox;             // 1: draws from Philox; 0: from the arrays below (the library's CPython stream)
    PhiloxKeys keys;
    const double *wheel_u;  // [2P]
    const int32_t *cut_in;  // [P], -1 = not crossed
    const double *mut_u;    // [N][D]: the uniform behind a mutated gene, < 0 where the gene is not mutated
    // state
    double *cur;            // [N][ld] current generation in list order
    const double *fit;      // [N] cached fitness
    const double *lmax;     // [N] running maximum of the running fitness sums, restarting at every SCAN_CHUNK entries
    const double *crun;     // [nchunks] running maximum of the running sums through the end of each chunk
    int nchunks;
    const double *scal;     // [0] sumFitness, [1] maxFitness; as int64: [2] best index
    double *child;          // [2P][ld]
    int32_t *picks;         // [2P] selected current rows
    int32_t *cut;           // [P]
    uint8_t *state;         // [2P] ENTRY_*
    int32_t *owner;         // [N][D] alias mutation: 1 + the last entry that rewrites a gene (only when cr < 1)
    uint64_t *hash;         // [N + 2P] row hashes by entry number: [0, N) current chromosomes, [N, N + 2P) this generation's entries
};

// the (mutate?, uniform) draw of gene j of entry i (Philox: the words of column pair j >> 1, as de's gene stream)
__device__ __forceinline__ bool mutation_draw(const GaParams &p, int64_t i, int j, double &u) {
    if (p.philox) {
        uint32_t w0 = (uint32_t)j >> 1, w1 = (uint32_t)i, w2 = (uint32_t)((uint64_t)i >> 32) | (oode::TAG_ELEM << 24), w3 = p.gen;
        philox4x32_10(w0, w1, w2, w3, p.keys);
        const double prob = u32_to_unit((j & 1) ? w2 : w0);
        u = u32_to_unit((j & 1) ? w3 : w1);
        return prob <= p.mr;
    }
    u = p.mut_u[i * p.D + j];
    return u >= 0.0;
}

// both genes of column pair jp (columns 2jp, 2jp+1) of entry i: one Philox call serves the pair.  m0 / m1: mutate?
__device__ __forceinline__ void mutation_draw_pair(const GaParams &p, int64_t i, int jp, bool &m0, double &u0, bool &m1, double &u1) {
    if (p.philox) {
        uint32_t w0 = (uint32_t)jp, w1 = (uint32_t)i, w2 = (uint32_t)((uint64_t)i >> 32) | (oode::TAG_ELEM << 24), w3 = p.gen;
        philox4x32_10(w0, w1, w2, w3, p.keys);
        m0 = u32_to_unit(w0) <= p.mr; u0 = u32_to_unit(w1);
        m1 = u32_to_unit(w2) <= p.mr; u1 = u32_to_unit(w3);
        return;
    }
    const int j = 2 * jp;
    u0 = p.mut_u[i * p.D + j];
    u1 = j + 1 < p.D ? p.mut_u[i * p.D + j + 1] : -1.0;
    m0 = u0 >= 0.0;
    m1 = u1 >= 0.0;
}

// ---- prepPopulation (:237-253): genes = uniform(lb, ub) ----
__global__ void ga_init_kernel(GaParams p, const double *U) {
    const int64_t total = p.N * p.D;
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
        const int64_t i = t / p.D;
        const int j = (int)(t - i * p.D);
        double u;
        if (U) u = U[t];
        else {
            uint32_t w0 = (uint32_t)j >> 1, w1 = (uint32_t)i, w2 = (uint32_t)((uint64_t)i >> 32) | (oode::TAG_ELEM << 24), w3 = 0u;
            philox4x32_10(w0, w1, w2, w3, p.keys);
            u = u32_to_unit((j & 1) ? w2 : w0);
        }
        p.cur[i * p.ld + j] = uniform_gene(p.lb[j], p.ub[j], u);
    }
}

// ---- evaluate (:255-276) ------------------------------------------------------------------------------------------
// sumFitness and the roulette's running sums are SEQUENTIALLY rounded sums in list order (:268, :340-343):
// s[k] = RN(s[k-1] + fit[k]).  Outputs, for chunks of SCAN_CHUNK consecutive entries: prefix[N] (the sums), lmax[N]
// (their running maximum inside the chunk), crun[nchunks] (running maximum through the end of each chunk), scal =
// {sumFitness, maxFitness, index of the FIRST strictly largest fitness (:269-271)}.
//
// Small populations: one CTA, one thread walks the chain (ga_scan_small_kernel).  Large populations: the chain is made
// parallel WITHOUT changing a single rounding.  While the running sum stays inside one binade, s = S*u with an integer
// S in [2^52, 2^53) and u = ulp(s), and RN(s + f) = (S + rint(f/u))*u unless f/u falls exactly between two integers: the
// chain is then an INTEGER prefix sum, which is associative.  So
//   1. ga_scan_totals_kernel    approximate chunk totals (any order) and each chunk's best fitness;
//   2. ga_scan_predict_kernel   approximate carry into each chunk -> the binade the chunk is expected to run in;
//   3. ga_scan_quantize_kernel  per chunk: q = rint(f/u) and their int64 prefix sums, min / max / total; the chunk is
//                               marked HARD if any f/u is a tie, is inexact or is too large;
//   4. ga_scan_carry_kernel     one thread walks the CHUNKS with the exact carry: a chunk whose predicted binade is the
//                               carry's and whose integer prefixes stay inside [2^52+1, 2^53-1] is taken in O(1); a hard
//                               chunk is walked element by element (staged through shared memory by the CTA);
//   5. ga_scan_finish_kernel    sums of the easy chunks from their integers, running maxima.
// (tests: all N running sums against numpy's sequential cumsum at N = 10^6, tests/test_gpu_galileo.py)
constexpr int SCAN_CHUNK = 1024, SCAN_THREADS = 256, SCAN_PER_THREAD = SCAN_CHUNK / SCAN_THREADS;
constexpr long long MANT_LO = (1ll << 52) + 1, MANT_HI = (1ll << 53) - 1;
constexpr int HARD_CHUNK = INT_MIN;

struct ScanArrays {
    const double *fit;
    int64_t N;
    int nchunks;
    double *prefix, *lmax, *crun, *scal;
    // large populations only
    double *tot, *approx_in, *bestv, *cmax;
    long long *besti, *qpre, *qsum, *qmin, *qmax, *carry;
    int *cexp;
    unsigned char *direct;
};

__device__ __forceinline__ double pow2(int e) { return __longlong_as_double((long long)(e + 1023) << 52); }   // -1022 <= e <= 1023

struct OpAddLL { __device__ long long operator()(long long a, long long b) const { return a + b; } };
struct OpMaxD { __device__ double operator()(double a, double b) const { return fmax(a, b); } };

// inclusive scan over the SCAN_CHUNK values of a CTA (thread t holds entries 4t .. 4t+3); returns the CTA's total
template <class T, class Op> __device__ __forceinline__ T block_scan(T (&v)[SCAN_PER_THREAD], Op op, T identity, T *warp_tot) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int i = 1; i < SCAN_PER_THREAD; ++i) v[i] = op(v[i - 1], v[i]);
    T run = v[SCAN_PER_THREAD - 1];
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const T up = __shfl_up_sync(0xFFFFFFFFu, run, o);
        if (lane >= o) run = op(up, run);
    }
    T excl = __shfl_up_sync(0xFFFFFFFFu, run, 1);
    if (lane == 0) excl = identity;
    if (lane == 31) warp_tot[warp] = run;
    __syncthreads();
    T off = identity, total = identity;
#pragma unroll
    for (int w = 0; w < SCAN_THREADS / 32; ++w) {
        if (w < warp) off = op(off, warp_tot[w]);
        total = op(total, warp_tot[w]);
    }
    off = op(off, excl);
#pragma unroll
    for (int i = 0; i < SCAN_PER_THREAD; ++i) v[i] = op(off, v[i]);
    __syncthreads();                        // warp_tot may be reused by the caller
    return total;
}

__global__ void __launch_bounds__(SCAN_CHUNK) ga_scan_small_kernel(ScanArrays a) {
    __shared__ double sf[SCAN_CHUNK], sp[SCAN_CHUNK], sm[SCAN_CHUNK];
    double s = 0.0, run = 0.0, best = 0.0;
    int64_t best_i = 0;
    for (int c = 0; c < a.nchunks; ++c) {
        const int64_t base = (int64_t)c * SCAN_CHUNK, k = base + threadIdx.x;
        if (k < a.N) sf[threadIdx.x] = a.fit[k];
        __syncthreads();
        if (threadIdx.x == 0) {
            const int n = (int)min((int64_t)SCAN_CHUNK, a.N - base);
            double m = 0.0;
            for (int t = 0; t < n; ++t) {
                const double f = sf[t];
                s = __dadd_rn(s, f);                                  // :268
                m = t == 0 ? s : fmax(m, s);
                sp[t] = s;
                sm[t] = m;
                if (base + t == 0 || f > best) { best = f; best_i = base + t; }     // :269-271 (strict: the first one wins)
            }
            run = c == 0 ? m : fmax(run, m);
            a.crun[c] = run;
        }
        __syncthreads();
        if (k < a.N) { a.prefix[k] = sp[threadIdx.x]; a.lmax[k] = sm[threadIdx.x]; }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        a.scal[0] = s;
        a.scal[1] = best;
        reinterpret_cast<long long *>(a.scal)[2] = best_i;
    }
}

__global__ void __launch_bounds__(SCAN_THREADS) ga_scan_totals_kernel(ScanArrays a) {
    __shared__ double ws[SCAN_THREADS / 32], wb[SCAN_THREADS / 32];
    __shared__ long long wi[SCAN_THREADS / 32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int c = blockIdx.x; c < a.nchunks; c += gridDim.x) {
        const int64_t k0 = (int64_t)c * SCAN_CHUNK + threadIdx.x * SCAN_PER_THREAD;
        double sum = 0.0, bv = -INFINITY;
        long long bi = LLONG_MAX;
#pragma unroll
        for (int i = 0; i < SCAN_PER_THREAD; ++i)
            if (k0 + i < a.N) {
                const double f = a.fit[k0 + i];
                sum += f;
                if (bi == LLONG_MAX || f > bv) { bv = f; bi = k0 + i; }
            }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            sum += __shfl_xor_sync(0xFFFFFFFFu, sum, o);
            const double ov = __shfl_xor_sync(0xFFFFFFFFu, bv, o);
            const long long oi = __shfl_xor_sync(0xFFFFFFFFu, bi, o);
            if (oi != LLONG_MAX && (bi == LLONG_MAX || ov > bv || (ov == bv && oi < bi))) { bv = ov; bi = oi; }
        }
        if (lane == 0) { ws[warp] = sum; wb[warp] = bv; wi[warp] = bi; }
        __syncthreads();
        if (threadIdx.x == 0) {
            for (int w = 1; w < SCAN_THREADS / 32; ++w) {
                sum += ws[w];
                if (wi[w] != LLONG_MAX && (bi == LLONG_MAX || wb[w] > bv || (wb[w] == bv && wi[w] < bi))) { bv = wb[w]; bi = wi[w]; }
            }
            a.tot[c] = sum;
            a.bestv[c] = bv;
            a.besti[c] = bi;
        }
        __syncthreads();
    }
}

// single CTA of SCAN_CHUNK threads: inclusive Hillis-Steele scan of one value per thread in shared memory
template <class Op> __device__ __forceinline__ double cta_scan1024(double v, Op op, double *sh) {
    sh[threadIdx.x] = v;
    __syncthreads();
    for (int o = 1; o < SCAN_CHUNK; o <<= 1) {
        const double up = threadIdx.x >= o ? sh[threadIdx.x - o] : v;
        __syncthreads();
        if (threadIdx.x >= o) { v = op(up, v); sh[threadIdx.x] = v; }
        __syncthreads();
    }
    return v;
}
struct OpAddD { __device__ double operator()(double a, double b) const { return a + b; } };

// approximate carry into every chunk (any summation order will do: it only predicts the binade) and the first chunk
// holding the largest fitness
__global__ void __launch_bounds__(SCAN_CHUNK) ga_scan_predict_kernel(ScanArrays a) {
    __shared__ double sh[SCAN_CHUNK];
    __shared__ long long si[SCAN_CHUNK];
    double carry = 0.0, bv = -INFINITY;
    long long bi = LLONG_MAX;
    for (int base = 0; base < a.nchunks; base += SCAN_CHUNK) {
        const int c = base + threadIdx.x;
        const double v = c < a.nchunks ? a.tot[c] : 0.0;
        const double incl = cta_scan1024(v, OpAddD(), sh);
        if (c < a.nchunks) {
            a.approx_in[c] = carry + (incl - v);
            const double f = a.bestv[c];
            if (bi == LLONG_MAX || f > bv) { bv = f; bi = a.besti[c]; }     // this thread's chunks come in ascending order
        }
        carry += sh[SCAN_CHUNK - 1];
        __syncthreads();
    }
    sh[threadIdx.x] = bv;
    si[threadIdx.x] = bi;
    __syncthreads();
    for (int o = SCAN_CHUNK / 2; o > 0; o >>= 1) {
        if (threadIdx.x < o) {
            const double ov = sh[threadIdx.x + o];
            const long long oi = si[threadIdx.x + o];
            const long long mi = si[threadIdx.x];
            if (oi != LLONG_MAX && (mi == LLONG_MAX || ov > sh[threadIdx.x] || (ov == sh[threadIdx.x] && oi < mi))) {
                sh[threadIdx.x] = ov;
                si[threadIdx.x] = oi;
            }
        }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        a.scal[1] = sh[0];
        reinterpret_cast<long long *>(a.scal)[2] = si[0];
    }
}

__global__ void __launch_bounds__(SCAN_THREADS) ga_scan_quantize_kernel(ScanArrays a) {
    __shared__ long long wt[SCAN_THREADS / 32], wmin[SCAN_THREADS / 32], wmax[SCAN_THREADS / 32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int c = blockIdx.x; c < a.nchunks; c += gridDim.x) {
        const double ap = a.approx_in[c];
        const int be = (int)((__double_as_longlong(ap) >> 52) & 0x7FF);
        const int E = be - 1023;
        bool hard = be == 0 || be == 0x7FF || E < -900 || E > 900;       // zero / subnormal / inf / NaN / extreme carry
        const double u = hard ? 1.0 : pow2(E - 52), inv_u = hard ? 1.0 : pow2(52 - E);
        const int64_t k0 = (int64_t)c * SCAN_CHUNK + threadIdx.x * SCAN_PER_THREAD;
        long long q[SCAN_PER_THREAD];
#pragma unroll
        for (int i = 0; i < SCAN_PER_THREAD; ++i) {
            q[i] = 0;
            if (k0 + i < a.N) {
                const double x = a.fit[k0 + i];
                const double t = __dmul_rn(x, inv_u);                                   // exact unless it over- or underflows
                const bool exact = __dmul_rn(t, u) == x && fabs(t) < 2251799813685248.0; // |t| < 2^51 (NaN / inf fail too)
                const bool tie = fabs(__dsub_rn(t, trunc(t))) == 0.5;                    // RN would look at the parity of S + q
                if (!exact || tie) hard = true;
                else q[i] = (long long)rint(t);
            }
        }
        const long long total = block_scan(q, OpAddLL(), 0ll, wt);
        long long mn = q[0], mx = q[0];
#pragma unroll
        for (int i = 0; i < SCAN_PER_THREAD; ++i) {
            if (k0 + i < a.N) a.qpre[k0 + i] = q[i];
            mn = min(mn, q[i]);                      // entries past N repeat the last prefix: harmless for min / max
            mx = max(mx, q[i]);
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            mn = min(mn, __shfl_xor_sync(0xFFFFFFFFu, mn, o));
            mx = max(mx, __shfl_xor_sync(0xFFFFFFFFu, mx, o));
        }
        if (lane == 0) { wmin[warp] = mn; wmax[warp] = mx; }
        const bool any_hard = __syncthreads_or(hard ? 1 : 0) != 0;
        if (threadIdx.x == 0) {
            for (int w = 1; w < SCAN_THREADS / 32; ++w) { mn = min(mn, wmin[w]); mx = max(mx, wmax[w]); }
            a.cexp[c] = any_hard ? HARD_CHUNK : E;
            a.qsum[c] = total;
            a.qmin[c] = mn;
            a.qmax[c] = mx;
        }
        __syncthreads();
    }
}

__global__ void __launch_bounds__(SCAN_THREADS) ga_scan_carry_kernel(ScanArrays a) {
    __shared__ double sx[SCAN_CHUNK], sp[SCAN_CHUNK];
    __shared__ long long t_qsum[SCAN_THREADS], t_qmin[SCAN_THREADS], t_qmax[SCAN_THREADS];
    __shared__ int t_exp[SCAN_THREADS];
    __shared__ int next_stop;
    double s = 0.0;                 // the exact running sum (thread 0)
    int c = 0;
    for (;;) {
        // the summaries of the next SCAN_THREADS chunks, staged so that thread 0 walks them at shared-memory latency
        const int tile0 = c, tile_end = min(c + SCAN_THREADS, a.nchunks);
        if (tile0 + (int)threadIdx.x < tile_end) {
            const int k = tile0 + threadIdx.x;
            t_exp[threadIdx.x] = a.cexp[k];
            t_qsum[threadIdx.x] = a.qsum[k];
            t_qmin[threadIdx.x] = a.qmin[k];
            t_qmax[threadIdx.x] = a.qmax[k];
        }
        __syncthreads();
        if (threadIdx.x == 0) {
            while (c < tile_end) {              // take easy chunks in O(1) each
                const int E = t_exp[c - tile0];
                const int be = (int)((__double_as_longlong(s) >> 52) & 0x7FF);
                if (E == HARD_CHUNK || be - 1023 != E) break;
                const long long S = (long long)__dmul_rn(s, pow2(52 - E));               // exact: |S| in [2^52, 2^53)
                const long long lo = S + t_qmin[c - tile0], hi = S + t_qmax[c - tile0];
                const bool inside = S > 0 ? (lo >= MANT_LO && hi <= MANT_HI) : (-hi >= MANT_LO && -lo <= MANT_HI);
                if (!inside) break;
                a.carry[c] = S;
                a.direct[c] = 0;
                s = __dmul_rn((double)(S + t_qsum[c - tile0]), pow2(E - 52));            // exact
                ++c;
            }
            next_stop = c;
        }
        __syncthreads();
        c = next_stop;
        if (c >= a.nchunks) break;
        if (c == tile_end) continue;            // the whole tile was easy: stage the next one
        // a hard chunk: walked element by element
        const int64_t base = (int64_t)c * SCAN_CHUNK;
        const int n = (int)min((int64_t)SCAN_CHUNK, a.N - base);
        for (int t = threadIdx.x; t < SCAN_CHUNK; t += SCAN_THREADS) sx[t] = t < n ? a.fit[base + t] : 0.0;
        __syncthreads();
        if (threadIdx.x == 0) {
            for (int t0 = 0; t0 < n; t0 += 8) {     // operands first, then the dependent adds: the chain is 8 DADDs back to back
                double x[8], r[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) x[i] = sx[t0 + i];
#pragma unroll
                for (int i = 0; i < 8; ++i) { s = __dadd_rn(s, x[i]); r[i] = s; }
#pragma unroll
                for (int i = 0; i < 8; ++i) sp[t0 + i] = r[i];
            }
            s = sp[n - 1];                          // entries past n were padded with 0.0 (s + 0.0 == s, but be exact about it)
            a.direct[c] = 1;
        }
        __syncthreads();
        for (int t = threadIdx.x; t < n; t += SCAN_THREADS) a.prefix[base + t] = sp[t];
        ++c;
        __syncthreads();                            // the tile arrays and sx are rewritten at the top of the loop
    }
    if (threadIdx.x == 0) a.scal[0] = s;
}

__global__ void __launch_bounds__(SCAN_THREADS) ga_scan_finish_kernel(ScanArrays a) {
    __shared__ double wt[SCAN_THREADS / 32];
    for (int c = blockIdx.x; c < a.nchunks; c += gridDim.x) {
        const int64_t k0 = (int64_t)c * SCAN_CHUNK + threadIdx.x * SCAN_PER_THREAD;
        const bool direct = a.direct[c] != 0;
        const long long S = direct ? 0 : a.carry[c];
        const double u = direct ? 1.0 : pow2(a.cexp[c] - 52);
        double v[SCAN_PER_THREAD];
#pragma unroll
        for (int i = 0; i < SCAN_PER_THREAD; ++i) {
            v[i] = -INFINITY;
            if (k0 + i < a.N) {
                if (direct) v[i] = a.prefix[k0 + i];
                else {
                    v[i] = __dmul_rn((double)(S + a.qpre[k0 + i]), u);      // exact: the integer is below 2^53, u a power of two
                    a.prefix[k0 + i] = v[i];
                }
            }
        }
        const double cm = block_scan(v, OpMaxD(), (double)-INFINITY, wt);
#pragma unroll
        for (int i = 0; i < SCAN_PER_THREAD; ++i)
            if (k0 + i < a.N) a.lmax[k0 + i] = v[i];
        if (threadIdx.x == 0) a.cmax[c] = cm;
    }
}

__global__ void __launch_bounds__(SCAN_CHUNK) ga_scan_crun_kernel(ScanArrays a) {
    __shared__ double sh[SCAN_CHUNK];
    double carry = -INFINITY;
    for (int base = 0; base < a.nchunks; base += SCAN_CHUNK) {
        const int c = base + threadIdx.x;
        const double v = c < a.nchunks ? a.cmax[c] : -INFINITY;
        const double incl = fmax(carry, cta_scan1024(v, OpMaxD(), sh));
        if (c < a.nchunks) a.crun[c] = incl;
        carry = fmax(carry, sh[SCAN_CHUNK - 1]);
        __syncthreads();
    }
}

// select_Roulette (:326-346): the first chromosome whose running sum is >= wheel, else the last one.  The running
// maximum of the sums is non-decreasing and crosses `wheel` at the same index: find the chunk, then the entry.
__device__ __forceinline__ int roulette(const GaParams &p, double u) {
    const double wheel = __dadd_rn(0.0, __dmul_rn(__dsub_rn(p.scal[0], 0.0), u));       // uniform(0, sumFitness) :337
    int lo = 0, hi = p.nchunks;                                                          // first chunk with crun[c] >= wheel
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (p.crun[mid] >= wheel) hi = mid; else lo = mid + 1;
    }
    if (lo >= p.nchunks) return (int)(p.N - 1);                                          // :346
    const int64_t base = (int64_t)lo * SCAN_CHUNK;
    int64_t a = base, b = min(base + SCAN_CHUNK, p.N);                                   // first entry of the chunk with lmax >= wheel
    while (a < b) {
        const int64_t mid = (a + b) >> 1;
        if (p.lmax[mid] >= wheel) b = mid; else a = mid + 1;
    }
    return (int)min(a, p.N - 1);
}

// ---- row hashes: a commutative sum of mixed (gene bits, column) words, so the lanes of a warp can split the row.
// x + 0.0 maps -0.0 to +0.0 (Python compares floats by value, :197). ----
__device__ __forceinline__ uint64_t mix64(uint64_t z) {
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
__device__ __forceinline__ uint64_t gene_hash(double x, int j) {
    return mix64((uint64_t)__double_as_longlong(__dadd_rn(x, 0.0)) ^ ((uint64_t)(j + 1) * 0x9E3779B97F4A7C15ull));
}
__device__ __forceinline__ uint64_t warp_sum(uint64_t h) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) h += __shfl_xor_sync(0xFFFFFFFFu, h, o);
    return h;
}

// ---- select (:290-301): one thread per roulette spin; the even spins also draw their pair's crossover (:360-363).
// (A thread per spin keeps ~10^6 binary searches in flight; inside the children kernel each warp waited for its own.) ----
__global__ void ga_select_kernel(GaParams p) {
    for (int64_t m = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; m < 2 * p.P; m += (int64_t)gridDim.x * blockDim.x) {
        double u;
        if (p.philox) {
            uint32_t w0 = 0u, w1 = (uint32_t)m, w2 = (uint32_t)((uint64_t)m >> 32) | (TAG_GA << 24), w3 = p.gen;
            philox4x32_10(w0, w1, w2, w3, p.keys);
            u = u53(w0, w1);
        } else u = p.wheel_u[m];
        p.picks[m] = roulette(p, u);
        if ((m & 1) == 0) {
            const int64_t pr = m >> 1;
            int cut;
            if (p.philox) {
                uint32_t w0 = 1u, w1 = (uint32_t)pr, w2 = (uint32_t)((uint64_t)pr >> 32) | (TAG_GA << 24), w3 = p.gen;
                philox4x32_10(w0, w1, w2, w3, p.keys);
                cut = (u53(w0, w1) <= p.cr) ? (int)oode::mulhi64_index(w2, w3, (uint64_t)p.D) : -1;
            } else cut = p.cut_in[pr];
            p.cut[pr] = cut;
            p.state[m] = p.state[m + 1] = (uint8_t)(cut >= 0 ? ENTRY_FRESH : ENTRY_ALIAS);
        }
    }
}

// ---- crossover + mutate for the fresh children: one warp per pair of entries (2p, 2p+1), a lane per column pair; the
// children's row hashes are summed while the genes are in registers ----
__global__ void ga_children_kernel(GaParams p) {
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    const int npairs = (p.D + 1) >> 1;
    for (int64_t pr = warp0; pr < p.P; pr += nwarps) {
        const int cut = p.cut[pr];
        if (cut < 0) continue;              // the pair stays aliased to the current chromosomes (:368-369)
        const int64_t i0 = 2 * pr, i1 = 2 * pr + 1;
        const int a = p.picks[i0], b = p.picks[i1];
        const double *ra = p.cur + (int64_t)a * p.ld, *rb = p.cur + (int64_t)b * p.ld;
        double *c0 = p.child + i0 * p.ld, *c1 = p.child + i1 * p.ld;
        uint64_t h0 = 0, h1 = 0;
        for (int jp = lane; jp < npairs; jp += 32) {
            const int j = 2 * jp;
            const bool two = j + 1 < p.D;   // the pitch is even: column D of an odd row is padding (kept zero)
            const double2 xa = *reinterpret_cast<const double2 *>(ra + j), xb = *reinterpret_cast<const double2 *>(rb + j);
            double2 x0, x1;
            x0.x = j < cut ? xa.x : xb.x;           // chromo1[:cut] + chromo2[cut:]   :364
            x1.x = j < cut ? xb.x : xa.x;           // chromo2[:cut] + chromo1[cut:]   :365
            x0.y = j + 1 < cut ? xa.y : xb.y;
            x1.y = j + 1 < cut ? xb.y : xa.y;
            bool m0, m1;
            double u0, u1;
            if (i0 < p.N) {                         // :287 only the first N entries are mutated
                mutation_draw_pair(p, i0, jp, m0, u0, m1, u1);
                if (m0) x0.x = uniform_gene(p.lb[j], p.ub[j], u0);
                if (m1 && two) x0.y = uniform_gene(p.lb[j + 1], p.ub[j + 1], u1);
            }
            if (i1 < p.N) {
                mutation_draw_pair(p, i1, jp, m0, u0, m1, u1);
                if (m0) x1.x = uniform_gene(p.lb[j], p.ub[j], u0);
                if (m1 && two) x1.y = uniform_gene(p.lb[j + 1], p.ub[j + 1], u1);
            }
            if (!two) x0.y = x1.y = 0.0;
            *reinterpret_cast<double2 *>(c0 + j) = x0;
            *reinterpret_cast<double2 *>(c1 + j) = x1;
            h0 += gene_hash(x0.x, j);
            h1 += gene_hash(x1.x, j);
            if (two) { h0 += gene_hash(x0.y, j + 1); h1 += gene_hash(x1.y, j + 1); }
        }
        h0 = warp_sum(h0);
        h1 = warp_sum(h1);
        if (lane == 0) {
            p.hash[p.N + i0] = h0 ? h0 : 1ull;      // 0 marks an empty table slot
            p.hash[p.N + i1] = h1 ? h1 : 1ull;
        }
    }
}

// ---- mutate_Default on aliased entries: the entry IS current chromosome picks[i], so the write goes into the current
// generation.  Entries are processed in list order by the reference, so the last entry that rewrites a gene wins:
// PHASE 0 records that entry per (row, gene), PHASE 1 lets exactly that entry write. ----
template <int PHASE> __global__ void ga_alias_kernel(GaParams p) {
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t i = warp0; i < p.N; i += nwarps) {
        if (p.state[i] != ENTRY_ALIAS) continue;
        const int64_t r = p.picks[i];
        for (int j = lane; j < p.D; j += 32) {
            double u;
            if (!mutation_draw(p, i, j, u)) continue;
            int32_t *own = p.owner + r * p.D + j;
            if (PHASE == 0) atomicMax(own, (int32_t)(i + 1));
            else if (*own == (int32_t)(i + 1)) p.cur[r * p.ld + j] = uniform_gene(p.lb[j], p.ub[j], u);
        }
    }
}

// ---- the isIdentical scan (:465-471) as a hash join ----
struct RowSet {             // logical entries e: [0, N) = current rows, [N, N + 2P) = this generation's entries
    const double *cur, *child;
    int64_t N, n;
    int D, ld;
    const uint8_t *state;
    __device__ __forceinline__ const double *row(int64_t e) const { return e < N ? cur + e * ld : child + (e - N) * ld; }
    __device__ __forceinline__ bool live(int64_t e) const { return e < N || state[e - N] == ENTRY_FRESH; }
};

// hashes of rows [0, s.n): the current chromosomes after ooga_init and after in-place mutations (the entries' hashes come
// from ga_children_kernel, the survivors' are carried along by ga_gather_kernel)
__global__ void ga_hash_kernel(RowSet s, uint64_t *hash) {
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t e = warp0; e < s.n; e += nwarps) {
        if (!s.live(e)) continue;
        const double *r = s.row(e);
        uint64_t h = 0;
        for (int j = lane; j < s.D; j += 32) h += gene_hash(r[j], j);
        h = warp_sum(h);
        if (lane == 0) hash[e] = h ? h : 1ull;          // 0 marks an empty table slot
    }
}

// open-addressing table: key = row hash, value = the smallest entry number with that hash
__global__ void ga_table_insert_kernel(RowSet s, const uint64_t *hash, unsigned long long *tkey, uint32_t *tmin, uint64_t mask) {
    for (int64_t e = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; e < s.n; e += (int64_t)gridDim.x * blockDim.x) {
        if (!s.live(e)) continue;
        const unsigned long long h = hash[e];
        uint64_t slot = h & mask;
        for (;;) {
            const unsigned long long prev = atomicCAS(tkey + slot, 0ull, h);
            if (prev == 0ull || prev == h) { atomicMin(tmin + slot, (uint32_t)e); break; }
            slot = (slot + 1) & mask;
        }
    }
}

__device__ __forceinline__ bool rows_identical(const double *x, const double *y, int D, int lane) {
    bool same = true;
    for (int j = lane; j < D; j += 32) same = same && (x[j] == y[j]);
    return __all_sync(0xFFFFFFFFu, same);
}

// keep[i] = 1 if fresh entry i equals no chromosome already in the list: no current row and no earlier entry (:465-471).
// An earlier identical entry that was itself dropped was dropped because of a still earlier identical one, so "any
// earlier live entry with the same genes" is the same test.
__global__ void ga_dedupe_kernel(RowSet s, const uint64_t *hash, const unsigned long long *tkey, const uint32_t *tmin,
                                 uint64_t mask, uint8_t *keep) {
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t e = s.N + warp0; e < s.n; e += nwarps) {
        if (!s.live(e)) { if (lane == 0) keep[e - s.N] = 0; continue; }
        const unsigned long long h = hash[e];
        uint64_t slot = h & mask;
        while (tkey[slot] != h) slot = (slot + 1) & mask;
        const int64_t first = tmin[slot];
        bool dup = false;
        if (first < e) {
            dup = rows_identical(s.row(first), s.row(e), s.D, lane);
            if (!dup) {
                // two different rows with one 64-bit hash: look at every earlier live entry with this hash
                for (int64_t base = 0; base < e && !dup; base += 32) {
                    const int64_t c = base + lane;
                    const bool cand = c < e && c != first && s.live(c) && hash[c] == h;
                    unsigned hits = __ballot_sync(0xFFFFFFFFu, cand);
                    while (hits && !dup) {
                        const int l = __ffs(hits) - 1;
                        hits &= hits - 1;
                        dup = rows_identical(s.row(base + l), s.row(e), s.D, lane);
                    }
                }
            }
        }
        if (lane == 0) keep[e - s.N] = dup ? 0 : 1;
    }
}

// ---- sort / reverse / truncate (:472-474): descending cached fitness, ties towards the LATER list position (a stable
// ascending sort followed by reverse()).  Entries become (order-preserving fitness key, list position) pairs. ----
struct SortEntry {
    unsigned long long key;     // larger = sorts earlier; 0 = not in the list
    unsigned long long pos;
};

__device__ __forceinline__ unsigned long long fitness_key(double f) {
    if (f != f) f = -INFINITY;                     // a NaN objective counts as the worst possible fitness
    f = __dadd_rn(f, 0.0);                         // -0.0 and 0.0 compare equal (:188-191: s1 - s2)
    const unsigned long long b = (unsigned long long)__double_as_longlong(f);
    return (b >> 63) ? ~b : (b | 0x8000000000000000ull);
}

// fitness of a child from the objective value the evaluator produced (:39 evalFunc = -p.f)
__device__ __forceinline__ double child_fitness(double f) { return (f != f) ? -INFINITY : -f; }

// ent[idx] for the entries first + idx, idx in [0, count): first = 0 covers the whole list, first = N this generation's
// entries only (ga_merge_rank_kernel then ranks them against the current generation)
__global__ void ga_keys_kernel(RowSet s, const double *fit, const double *fchild, const uint8_t *keep, SortEntry *ent,
                               int64_t first, int64_t count, unsigned long long *counts) {
    unsigned long long kept = 0, alias = 0, dups = 0;
    for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < count; idx += (int64_t)gridDim.x * blockDim.x) {
        const int64_t e = first + idx;
        SortEntry t{0ull, 0ull};
        if (e < s.N) t = SortEntry{fitness_key(fit[e]), (unsigned long long)e};
        else if (e < s.n) {
            if (s.state[e - s.N] != ENTRY_FRESH) ++alias;
            else if (!keep[e - s.N]) ++dups;
            else { ++kept; t = SortEntry{fitness_key(child_fitness(fchild[e - s.N])), (unsigned long long)e}; }
        }
        ent[idx] = t;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        kept += __shfl_xor_sync(0xFFFFFFFFu, kept, o);
        alias += __shfl_xor_sync(0xFFFFFFFFu, alias, o);
        dups += __shfl_xor_sync(0xFFFFFFFFu, dups, o);
    }
    if ((threadIdx.x & 31) == 0) {
        if (kept) atomicAdd(counts + 0, kept);
        if (alias) atomicAdd(counts + 1, alias);
        if (dups) atomicAdd(counts + 2, dups);
    }
}

__device__ __forceinline__ bool sorts_before(const SortEntry &a, const SortEntry &b) {
    return a.key > b.key || (a.key == b.key && a.pos > b.pos);
}

// Bitonic network over n2 = 2^k entries, "sorts_before" order.  (key, pos) pairs are distinct for live entries, so the
// result does not depend on the network.  Steps with partner distance < TILE/2 run in shared memory.
constexpr int SORT_TILE = 2048, SORT_THREADS = 1024;

__device__ __forceinline__ void bitonic_exchange(SortEntry &a, SortEntry &b, bool first_goes_low) {
    // a sits at the lower index.  first_goes_low: the element that sorts before belongs at the lower index
    if (sorts_before(b, a) == first_goes_low) { const SortEntry t = a; a = b; b = t; }
}

// all steps (k, j) with j <= j_start inside one tile, for k from k_start up to k_end (k_end <= SORT_TILE means the tile is
// sorted from scratch; otherwise only the tail of stage k_start == k_end)
__global__ void __launch_bounds__(SORT_THREADS) ga_bitonic_tile_kernel(SortEntry *ent, int64_t k_start, int64_t k_end, int j_start) {
    __shared__ SortEntry sh[SORT_TILE];
    const int64_t base = (int64_t)blockIdx.x * SORT_TILE;
    for (int t = threadIdx.x; t < SORT_TILE; t += SORT_THREADS) sh[t] = ent[base + t];
    __syncthreads();
    for (int64_t k = k_start; k <= k_end; k <<= 1) {
        for (int j = (int)min((int64_t)j_start, k >> 1); j > 0; j >>= 1) {
            const int t = threadIdx.x;
            const int lo = ((t & ~(j - 1)) << 1) | (t & (j - 1));     // index of the lower element of pair t
            const bool low_first = (((base + lo) & k) == 0);
            bitonic_exchange(sh[lo], sh[lo | j], low_first);
            __syncthreads();
        }
    }
    for (int t = threadIdx.x; t < SORT_TILE; t += SORT_THREADS) ent[base + t] = sh[t];
}

// one step (k, j) with j >= SORT_TILE / 2... in global memory
__global__ void ga_bitonic_step_kernel(SortEntry *ent, int64_t n2, int64_t k, int64_t j) {
    for (int64_t t = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; t < (n2 >> 1); t += (int64_t)gridDim.x * blockDim.x) {
        const int64_t lo = ((t & ~(j - 1)) << 1) | (t & (j - 1));
        SortEntry a = ent[lo], b = ent[lo | j];
        const bool low_first = ((lo & k) == 0);
        if (sorts_before(b, a) == low_first) { ent[lo] = b; ent[lo | j] = a; }
    }
}

// From the second iteration on the current generation is already in descending fitness order (it is the previous
// truncate step's output), so only this generation's entries are sorted and the two sequences are merged by RANK: an
// element's place in the reference's sort-then-reverse order is the number of elements that sort before it.
//   current chromosome k, key K: the chromosomes with a larger key, those of its own tie run [a, b] that come LATER in the
//     list (ties go to the later position: the run is reversed), and every kept entry with key >= K (entries sit later in
//     the list than any chromosome);
//   sorted entry t, key K: the t entries before it and the chromosomes with key > K.
// top[rank] = (key, list position) for rank < N -- exactly what the full sort leaves in ent[0, N).
__device__ __forceinline__ int64_t first_cur_below(const double *fit, int64_t N, unsigned long long K, bool or_equal) {
    int64_t lo = 0, hi = N;                         // first i with key(i) < K (or <= K): keys are non-increasing
    while (lo < hi) {
        const int64_t mid = (lo + hi) >> 1;
        const unsigned long long km = fitness_key(fit[mid]);
        if (or_equal ? km <= K : km < K) hi = mid; else lo = mid + 1;
    }
    return lo;
}

__global__ void ga_merge_rank_kernel(RowSet s, const double *fit, const SortEntry *cent, int64_t n2c, SortEntry *top) {
    const int64_t total = s.N + n2c;
    for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
        if (idx < s.N) {
            const int64_t k = idx;
            const unsigned long long K = fitness_key(fit[k]);
            const int64_t a = first_cur_below(fit, s.N, K, true), b1 = first_cur_below(fit, s.N, K, false);
            int64_t lo = 0, hi = n2c;               // first sorted entry with key < K (unused slots have key 0)
            while (lo < hi) {
                const int64_t mid = (lo + hi) >> 1;
                if (cent[mid].key < K) hi = mid; else lo = mid + 1;
            }
            const int64_t rank = a + (b1 - 1 - k) + lo;
            if (rank < s.N) top[rank] = SortEntry{K, (unsigned long long)k};
        } else {
            const int64_t t = idx - s.N;
            const SortEntry c = cent[t];
            if (c.key == 0ull) continue;
            const int64_t rank = t + first_cur_below(fit, s.N, c.key, true);
            if (rank < s.N) top[rank] = c;
        }
    }
}

// the first N sorted entries become the next current generation (list order = sorted order)
__global__ void ga_gather_kernel(RowSet s, const SortEntry *ent, const double *fit, const double *fchild, const uint64_t *hash,
                                 double *next, double *next_fit, uint64_t *next_hash) {
    const int lane = threadIdx.x & 31;
    const int64_t warp0 = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
    const int64_t nwarps = ((int64_t)gridDim.x * blockDim.x) >> 5;
    for (int64_t k = warp0; k < s.N; k += nwarps) {
        const int64_t e = (int64_t)ent[k].pos;
        const double *src = s.row(e);
        double *dst = next + k * s.ld;
        for (int j = lane; j < s.D; j += 32) dst[j] = src[j];
        if (lane == 0) {
            next_fit[k] = e < s.N ? fit[e] : child_fitness(fchild[e - s.N]);
            next_hash[k] = hash[e];
        }
    }
}

struct GaRecord {           // = ooga_gen_record (include/ooga.h)
    double best_f, sum_fitness;
    int64_t best_idx, n_appended, n_aliased, n_duplicates;
    int32_t generation, reserved;
};

// what __solver__ reports for this iteration (:86-89): -maxFitness and the genes of bestFitIndividual, both from the
// evaluate step at the START of the iteration; the genes are read now, after the in-place mutations
__global__ void ga_record_kernel(const double *cur, int ld, int D, const double *scal, const unsigned long long *counts,
                                 int32_t generation, GaRecord *rec, double *best_x) {
    const int64_t bi = reinterpret_cast<const long long *>(scal)[2];
    for (int j = threadIdx.x; j < D; j += blockDim.x) best_x[j] = cur[bi * ld + j];
    if (threadIdx.x == 0) {
        rec->best_f = -scal[1];
        rec->sum_fitness = scal[0];
        rec->best_idx = bi;
        rec->n_appended = (int64_t)counts[0];
        rec->n_aliased = (int64_t)counts[1];
        rec->n_duplicates = (int64_t)counts[2];
        rec->generation = generation;
        rec->reserved = 0;
    }
}

}  // namespace ooga
