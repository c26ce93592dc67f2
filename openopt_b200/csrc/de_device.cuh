// Device-side building blocks of the DE generation: exact-rounding arithmetic helpers,
// Philox4x32-10, the reference's bound repair and the registered objectives.
//
// Everything that must round exactly like numpy (one IEEE rounding per operation, never an FMA)
// goes through dadd/dsub/dmul/ddiv below: the __d*_rn intrinsics are never contracted by nvcc.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>
#include "../../include/oode.h"

// -DOODE_DEBUG_BOUNDS: device-side assertions on every index the row kernel and the selection pass form (row and
// donor indices, stage / exchange-buffer / trial-row / leaf-sum offsets).  compute-sanitizer is closed on the GPU pool
// this was developed on, so the parity suite is run once per round against a library built this way instead
// (scripts/ab.py build dbg -D OODE_DEBUG_BOUNDS; profiles/r02*_bounds_assert_build.txt).  Off in the product build.
#ifdef OODE_DEBUG_BOUNDS
#include <cassert>
#define OODE_ASSERT(c) assert(c)
#else
#define OODE_ASSERT(c) ((void)0)
#endif

namespace oode {

__device__ __forceinline__ double dadd(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double dsub(double a, double b) { return __dsub_rn(a, b); }
__device__ __forceinline__ double dmul(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double ddiv(double a, double b) { return __ddiv_rn(a, b); }

// ---------------------------------------------------------------------------------------------
// Philox4x32-10 (Salmon et al. SC'11).  The ten round keys depend only on the seed, so the host
// precomputes them (rk[2r], rk[2r+1]) and they live in kernel-parameter constant memory.
// ---------------------------------------------------------------------------------------------
constexpr uint32_t PHILOX_M0 = 0xD2511F53u;
constexpr uint32_t PHILOX_M1 = 0xCD9E8D57u;
constexpr uint32_t PHILOX_W0 = 0x9E3779B9u;
constexpr uint32_t PHILOX_W1 = 0xBB67AE85u;
constexpr uint32_t TAG_ELEM = 0u;    // per-gene draws (generation 0 = initial population)
constexpr uint32_t TAG_INDEX = 1u;   // per-row donor indices
constexpr uint32_t TAG_DIRECT = 2u;  // the 2*hndvi population-wide indices of the directed search

struct PhiloxKeys {
    uint32_t rk[20];
};

__device__ __forceinline__ void philox4x32_10(uint32_t &c0, uint32_t &c1, uint32_t &c2, uint32_t &c3,
                                              const PhiloxKeys &K) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint64_t p0 = (uint64_t)PHILOX_M0 * c0;
        const uint64_t p1 = (uint64_t)PHILOX_M1 * c2;
        const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ K.rk[2 * r];
        const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ K.rk[2 * r + 1];
        c1 = (uint32_t)p1;
        c3 = (uint32_t)p0;
        c0 = n0;
        c2 = n2;
    }
}

// one 32-bit word -> double in [0,1): exactly w * 2^-32 (build 1.m from the bits, subtract 1)
__device__ __forceinline__ double u32_to_unit(uint32_t w) {
    return dsub(__hiloint2double((int)(0x3FF00000u | (w >> 12)), (int)(w << 20)), 1.0);
}

__device__ __forceinline__ int64_t mulhi64_index(uint32_t whi, uint32_t wlo, uint64_t n) {
    return (int64_t)__umul64hi(((uint64_t)whi << 32) | wlo, n);
}

// ---------------------------------------------------------------------------------------------
// _correct_box_constraints (de_oo.py:346-375), per element: reflect with factor (1+2^-k) while
// the gene violates lb, then ub, then clamp.  See oracle/de_oracle.py::correct_box_constraints.
// ---------------------------------------------------------------------------------------------
constexpr int MAX_REPAIR_PASSES = 4096;

static __device__ __noinline__ double repair_gene_slow(double v, double lo, double hi) {
    double d = dsub(v, lo);
    if (d < 0.0) {
        double scale = 1.0;
        int it = 0;
        do {
            v = dsub(v, dmul(dadd(1.0, scale), d));
            d = dsub(v, lo);
            scale = dmul(scale, 0.5);
        } while (d < 0.0 && ++it < MAX_REPAIR_PASSES);
    }
    d = dsub(v, hi);
    if (d > 0.0) {
        double scale = 1.0;
        int it = 0;
        do {
            v = dsub(v, dmul(dadd(1.0, scale), d));
            d = dsub(v, hi);
            scale = dmul(scale, 0.5);
        } while (d > 0.0 && ++it < MAX_REPAIR_PASSES);
    }
    if (v < lo) v = lo;
    if (v > hi) v = hi;
    return v;
}

// Fast path: a gene violates at most one bound (lb < ub); one reflection about that bound, x = fl(v - 2d)
// with d = v - b -- the same single rounding as the reference's v - (1+1)*d because 2d is exact, done here
// as fma(-2, d, v) -- and if the result lies inside [lb, ub] that is all the reference does (second loop and
// clamps are no-ops).  That is always the case when the mutation step is not longer than the box (F <= 1).
// Anything else -- reflection leaving the box on the other side, a bound still violated after its own
// reflection (rounding artefact), NaN/inf -- re-runs the general loops from the original value.  The
// repaired-gene count is the reference's (out != in).sum(): exactly the genes that were outside.
__device__ __forceinline__ double repair_gene(double v, double lo, double hi, unsigned &nrep) {
    const bool lt = v < lo, gt = v > hi;
    const double b = lt ? lo : hi;
    const double refl = __fma_rn(-2.0, dsub(v, b), v);
    const bool viol = lt || gt;
    double x = viol ? refl : v;
    nrep += viol ? 1u : 0u;
    if (!((x >= lo) && (x <= hi))) x = repair_gene_slow(v, lo, hi);
    return x;
}

// ---------------------------------------------------------------------------------------------
// cos for the objectives.  The objectives only ever SUM (or multiply O(1) factors of) cosines, so
// what matters is absolute accuracy; that allows one branch-free polynomial instead of the
// library's quadrant-selected pair fed from a coefficient table:
//   k = rint(y/pi) (magic-number rounding), r = y - k*pi by three FMAs (pi split in three doubles),
//   cos(y) = (-1)^k * (1 - 2*sin(r/2)^2), sin on [-pi/4, pi/4] with the fdlibm k_sin polynomial.
// Max absolute error 2.7e-16 over |y| <= 1e6 (checked in exact arithmetic against a 50-digit
// reference; numpy's cos: 5.5e-17), far inside the 1e-12 relative parity bar on objective values.
// Larger |y|, inf and NaN take the library cos.  Coefficients live in constant memory so they are
// instruction operands (no register traffic, no table loads).
// ---------------------------------------------------------------------------------------------
static __constant__ double COS_K[12] = {
    0.3183098861837907,          // 0: 1/pi
    6755399441055744.0,          // 1: 1.5 * 2^52 (round-to-nearest-integer magic number)
    3.141592653589793,           // 2: pi, first 53 bits
    1.2246467991473532e-16,      // 3: next 53 bits
    -2.9947698097183397e-33,     // 4: next 53 bits
    -1.66666666666666324348e-01, // 5: S1 .. S6 (fdlibm __kernel_sin)
    8.33333333332248946124e-03,
    -1.98412698298579493134e-04,
    2.75573137070700676789e-06,
    -2.50507602534068634195e-08,
    1.58969099521155010221e-10,
    1.0e6                        // 11: fast-path limit on |y|
};

template <bool CHECKED = true> __device__ __forceinline__ double obj_cos(double y) {
    if (CHECKED && !(fabs(y) <= COS_K[11])) return cos(y);
    double kd = __fma_rn(y, COS_K[0], COS_K[1]);
    const int kbits = __double2loint(kd);
    kd = __dsub_rn(kd, COS_K[1]);
    double r = __fma_rn(-kd, COS_K[2], y);
    r = __fma_rn(-kd, COS_K[3], r);
    r = __fma_rn(-kd, COS_K[4], r);
    const double h = __dmul_rn(0.5, r);
    const double z = __dmul_rn(h, h);
    double p = __fma_rn(z, COS_K[10], COS_K[9]);
    p = __fma_rn(z, p, COS_K[8]);
    p = __fma_rn(z, p, COS_K[7]);
    p = __fma_rn(z, p, COS_K[6]);
    p = __fma_rn(z, p, COS_K[5]);
    const double sn = __fma_rn(__dmul_rn(h, z), p, h);
    const double c = __fma_rn(__dmul_rn(-2.0, sn), sn, 1.0);
    return __hiloint2double(__double2hiint(c) ^ (kbits << 31), __double2loint(c));
}

constexpr double TWO_PI = 6.283185307179586;    // python: 2*pi
constexpr double EULER_E = 2.718281828459045;   // python: math.e / numpy.e

// Obj<k>: NS sum slots then NPR product slots per row; term(x, x_next, aux, col, t) writes the NS+NPR terms of the
// gene in (global) column col; finalize(s, D) turns the folded slots into f.  HALO = 1: the term also reads the next
// column's gene.  AUX: a per-column parameter array is used.
template <int OBJ> struct Obj;

template <> struct Obj<OODE_OBJ_SPHERE> {
    static constexpr int NS = 1, NPR = 0, HALO = 0;
    static constexpr bool AUX = false;
    template <bool CK = true> __device__ static __forceinline__ void term(double x, double, double, int, double *t) { t[0] = dmul(x, x); }
    __device__ static __forceinline__ double finalize(const double *s, int) { return s[0]; }
};

template <> struct Obj<OODE_OBJ_SHIFTED_SPHERE> {
    static constexpr int NS = 1, NPR = 0, HALO = 0;
    static constexpr bool AUX = true;
    template <bool CK = true> __device__ static __forceinline__ void term(double x, double, double aux, int, double *t) {
        const double d = dsub(x, aux);
        t[0] = dmul(d, d);
    }
    __device__ static __forceinline__ double finalize(const double *s, int) { return s[0]; }
};

template <> struct Obj<OODE_OBJ_ROSENBROCK> {
    static constexpr int NS = 1, NPR = 0, HALO = 1;
    // 100.0*(x[i+1]-x[i]**2)**2 + (1.0-x[i])**2
    static constexpr bool AUX = false;
    template <bool CK = true> __device__ static __forceinline__ void term(double x, double xn, double, int, double *t) {
        const double a = dsub(xn, dmul(x, x));
        const double b = dsub(1.0, x);
        t[0] = dadd(dmul(100.0, dmul(a, a)), dmul(b, b));
    }
    __device__ static __forceinline__ double finalize(const double *s, int) { return s[0]; }
};

template <> struct Obj<OODE_OBJ_RASTRIGIN> {
    static constexpr int NS = 1, NPR = 0, HALO = 0;
    // x*x - 10*cos(2*pi*x) + 10
    static constexpr bool AUX = false;
    template <bool CK = true> __device__ static __forceinline__ void term(double x, double, double, int, double *t) {
        const double c = obj_cos<CK>(dmul(TWO_PI, x));
        t[0] = dadd(dsub(dmul(x, x), dmul(10.0, c)), 10.0);
    }
    __device__ static __forceinline__ double finalize(const double *s, int) { return s[0]; }
};

template <> struct Obj<OODE_OBJ_ACKLEY> {
    static constexpr int NS = 2, NPR = 0, HALO = 0;
    static constexpr bool AUX = false;
    template <bool CK = true> __device__ static __forceinline__ void term(double x, double, double, int, double *t) {
        t[0] = dmul(x, x);
        t[1] = obj_cos<CK>(dmul(TWO_PI, x));
    }
    // -20.0*exp(-0.2*sqrt(s0/D)) - exp(s1/D) + 20.0 + e
    __device__ static __forceinline__ double finalize(const double *s, int D) {
        const double n = (double)D;
        const double a = dmul(-20.0, exp(dmul(-0.2, __dsqrt_rn(ddiv(s[0], n)))));
        const double b = exp(ddiv(s[1], n));
        return dadd(dadd(dsub(a, b), 20.0), EULER_E);
    }
};

template <> struct Obj<OODE_OBJ_GRIEWANK> {
    static constexpr int NS = 1, NPR = 1, HALO = 0;
    static constexpr bool AUX = true;
    template <bool CK = true> __device__ static __forceinline__ void term(double x, double, double aux, int, double *t) {
        t[0] = dmul(x, x);
        t[1] = obj_cos<CK>(ddiv(x, aux));           // aux[j] = sqrt(j+1), computed by numpy on the host
    }
    // 1.0 + s0/4000.0 - prod
    __device__ static __forceinline__ double finalize(const double *s, int) {
        return dsub(dadd(1.0, ddiv(s[0], 4000.0)), s[1]);
    }
};

// An objective registered at run time (openopt_b200/expr.py generates the specialisation; trial_obj.cu is compiled
// once more with -DOODE_OBJ=OODE_OBJ_CUSTOM -DOODE_CUSTOM_HEADER=... into its own shared object, oode_register_objective)
#ifdef OODE_CUSTOM_HEADER
#include OODE_CUSTOM_HEADER
#endif

template <int OBJ, int S> __device__ __forceinline__ double slot_identity() {
    return S < Obj<OBJ>::NS ? 0.0 : 1.0;
}
template <int OBJ> __device__ __forceinline__ double slot_op(int s, double a, double b) {
    return s < Obj<OBJ>::NS ? dadd(a, b) : dmul(a, b);
}
// identities that are exact in IEEE arithmetic: (-0.0) + t == t for every t (including both zeros)
template <int OBJ> __device__ __forceinline__ double slot_id(int s) { return s < Obj<OBJ>::NS ? -0.0 : 1.0; }

}  // namespace oode
