/*
 * de_oracle.c -- plain-C restatement of the reference's differential-evolution generation loop.
 * TEST INFRASTRUCTURE, NOT PRODUCT CODE: only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / `--impl reference` legs load it.  The product (openopt_b200) never does.
 *
 * It follows oracle/de_oracle.py function by function (which is pinned bit-exactly to the
 * unmodified reference by tests/golden/make_golden.py) and cites the same reference lines
 * (openopt/solvers/UkrOpt/de_oo.py).  Differences from the numpy oracle: rows are processed by
 * OpenMP threads, and cos/exp/sqrt come from glibc instead of numpy's SIMD kernels (so the
 * exp-based Ackley value can differ from numpy's in the last ulp; everything without exp is
 * bit-identical on this image -- tests/test_oracle_c.py).
 *
 * Covers the whole solver surface like the numpy oracle: every strategy parameter (de_oo.py:71-89) and
 * general constraints (linear A/b, Aeq/beq; diagonal-quadratic c / h, openopt_b200/constraints.py).  Linear
 * residuals are accumulated left to right here while the reference uses a BLAS product, so they agree
 * bit for bit only when the order cannot matter (<= 2 non-zero coefficients per row; tests say which).
 *
 * Build: make -C oracle   (gcc -O2 -fopenmp -ffp-contract=off; no FMA contraction).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

enum { OBJ_SPHERE = 0, OBJ_ROSENBROCK = 1, OBJ_RASTRIGIN = 2, OBJ_ACKLEY = 3, OBJ_GRIEWANK = 4, OBJ_SHIFTED_SPHERE = 5 };
enum { STREAM_PHILOX = 0, STREAM_CPYTHON = 1 };

/* ------------------------------------------------------------------ CPython random (MT19937) */
/* de_oo.py:16-53: Seed/Rand/Randint are random.seed / random.random() / random.randint(0,low-1),
 * one element at a time.  Algorithm: CPython Modules/_randommodule.c + Lib/random.py. */
typedef struct { uint32_t mt[624]; int mti; } mt_t;

static void mt_init_genrand(mt_t *m, uint32_t s) {
    m->mt[0] = s;
    for (m->mti = 1; m->mti < 624; m->mti++)
        m->mt[m->mti] = 1812433253u * (m->mt[m->mti - 1] ^ (m->mt[m->mti - 1] >> 30)) + (uint32_t)m->mti;
}

static void mt_seed(mt_t *m, uint64_t seed) {
    uint32_t key[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
    int klen = key[1] ? 2 : 1, i = 1, j = 0, k;
    mt_init_genrand(m, 19650218u);
    for (k = 624; k; k--) {
        m->mt[i] = (m->mt[i] ^ ((m->mt[i - 1] ^ (m->mt[i - 1] >> 30)) * 1664525u)) + key[j] + (uint32_t)j;
        i++; j++;
        if (i >= 624) { m->mt[0] = m->mt[623]; i = 1; }
        if (j >= klen) j = 0;
    }
    for (k = 623; k; k--) {
        m->mt[i] = (m->mt[i] ^ ((m->mt[i - 1] ^ (m->mt[i - 1] >> 30)) * 1566083941u)) - (uint32_t)i;
        i++;
        if (i >= 624) { m->mt[0] = m->mt[623]; i = 1; }
    }
    m->mt[0] = 0x80000000u;
}

static uint32_t mt_u32(mt_t *m) {
    static const uint32_t mag01[2] = {0u, 0x9908b0dfu};
    uint32_t y;
    if (m->mti >= 624) {
        int kk;
        for (kk = 0; kk < 624 - 397; kk++) {
            y = (m->mt[kk] & 0x80000000u) | (m->mt[kk + 1] & 0x7fffffffu);
            m->mt[kk] = m->mt[kk + 397] ^ (y >> 1) ^ mag01[y & 1u];
        }
        for (; kk < 623; kk++) {
            y = (m->mt[kk] & 0x80000000u) | (m->mt[kk + 1] & 0x7fffffffu);
            m->mt[kk] = m->mt[kk + (397 - 624)] ^ (y >> 1) ^ mag01[y & 1u];
        }
        y = (m->mt[623] & 0x80000000u) | (m->mt[0] & 0x7fffffffu);
        m->mt[623] = m->mt[396] ^ (y >> 1) ^ mag01[y & 1u];
        m->mti = 0;
    }
    y = m->mt[m->mti++];
    y ^= (y >> 11);
    y ^= (y << 7) & 0x9d2c5680u;
    y ^= (y << 15) & 0xefc60000u;
    y ^= (y >> 18);
    return y;
}

static double mt_random(mt_t *m) {
    uint32_t a = mt_u32(m) >> 5, b = mt_u32(m) >> 6;
    return (a * 67108864.0 + b) * (1.0 / 9007199254740992.0);
}

static uint32_t mt_randbelow(mt_t *m, uint32_t n) {
    int k = 0;
    uint32_t t, r;
    for (t = n; t; t >>= 1) k++;
    r = mt_u32(m) >> (32 - k);
    while (r >= n) r = mt_u32(m) >> (32 - k);
    return r;
}

/* ------------------------------------------------------------------ Philox4x32-10 */
static void philox(uint32_t c[4], uint32_t k0, uint32_t k1) {
    int r;
    for (r = 0; r < 10; r++) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c[0], p1 = (uint64_t)0xCD9E8D57u * c[2];
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c[1] ^ k0, n2 = (uint32_t)(p0 >> 32) ^ c[3] ^ k1;
        c[1] = (uint32_t)p1; c[3] = (uint32_t)p0; c[0] = n0; c[2] = n2;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
}

static double u32_unit(uint32_t w) { return (double)w * (1.0 / 4294967296.0); }

static int64_t mulhi64(uint32_t hi, uint32_t lo, uint64_t n) {
    return (int64_t)(((unsigned __int128)(((uint64_t)hi << 32) | lo) * n) >> 64);
}

/* ------------------------------------------------------------------ numpy pairwise sum */
static double pairwise_sum(const double *a, int64_t n) {
    if (n < 8) {
        double res = 0.;
        int64_t i;
        for (i = 0; i < n; i++) res += a[i];
        return res;
    } else if (n <= 128) {
        double r[8], res;
        int64_t i;
        int j;
        for (j = 0; j < 8; j++) r[j] = a[j];
        for (i = 8; i < n - (n % 8); i += 8)
            for (j = 0; j < 8; j++) r[j] += a[i + j];
        res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
        for (; i < n; i++) res += a[i];
        return res;
    } else {
        int64_t n2 = n / 2;
        n2 -= n2 % 8;
        return pairwise_sum(a, n2) + pairwise_sum(a + n2, n - n2);
    }
}

/* ------------------------------------------------------------------ state */
typedef struct {
    int obj, D, stream, invert;
    int64_t NP;
    double F, Cr;
    uint64_t seed;
    double *lb, *ub, *x0, *aux;
    int have_x0;
    double *pop, *trial, *vals, *tvals;
    uint8_t *accept;
    int64_t *ib, *r1, *r2;
    double *Uf, *Uc;            /* CPython stream: materialised draws of one generation */
    mt_t mt;
    int64_t gen, trial_best_idx, n_accepted, n_repaired;
    double trial_best_f, best_f;
    double *best_x;
    int best_changed;
    /* strategy parameters (de_oo.py:71-89); zero-initialised = the defaults */
    int base_best, dir_best, const_F, hndvi;
    int64_t *rdir;              /* [2h] indices of the directed search */
    double *last_best_x;        /* best[2] of the last _eval_pop (base vector 'best', de_oo.py:192) */
    double *b1, *b2;            /* directed search: worst / best barycentre */
    /* general constraints (de_oo.py:306-338) */
    int constrained, cm[4];     /* rows of A, Aeq, Qc, Qh */
    double *cM[4], *crhs[4], contol;
    double *cvals, *tcvals;
    int trial_best_feasible;
    double trial_best_c;
} orc_t;

static const double TWO_PI = 6.283185307179586;

/* canonical objectives, one contiguous row at a time (kernel/nonLinFuncs.py:190-200) */
static double objective(const orc_t *s, const double *x, double *t) {
    const int D = s->D;
    int j;
    switch (s->obj) {
        case OBJ_SPHERE:
            for (j = 0; j < D; j++) t[j] = x[j] * x[j];
            return pairwise_sum(t, D);
        case OBJ_SHIFTED_SPHERE:
            for (j = 0; j < D; j++) { double d = x[j] - s->aux[j]; t[j] = d * d; }
            return pairwise_sum(t, D);
        case OBJ_ROSENBROCK:
            for (j = 0; j < D - 1; j++) {
                double a = x[j + 1] - x[j] * x[j], b = 1.0 - x[j];
                t[j] = 100.0 * (a * a) + b * b;
            }
            return pairwise_sum(t, D - 1);
        case OBJ_RASTRIGIN:
            for (j = 0; j < D; j++) t[j] = (x[j] * x[j] - 10 * cos(TWO_PI * x[j])) + 10;
            return pairwise_sum(t, D);
        case OBJ_ACKLEY: {
            double s0, s1;
            for (j = 0; j < D; j++) t[j] = x[j] * x[j];
            s0 = pairwise_sum(t, D);
            for (j = 0; j < D; j++) t[j] = cos(TWO_PI * x[j]);
            s1 = pairwise_sum(t, D);
            return ((-20.0 * exp(-0.2 * sqrt(s0 / D)) - exp(s1 / D)) + 20.0) + 2.718281828459045;
        }
        default: { /* griewank: numpy .prod() is sequential */
            double s0, pr = 1.0;
            for (j = 0; j < D; j++) t[j] = x[j] * x[j];
            s0 = pairwise_sum(t, D);
            for (j = 0; j < D; j++) pr *= cos(x[j] / s->aux[j]);
            return (1.0 + s0 / 4000.0) - pr;
        }
    }
}

/* _correct_box_constraints (de_oo.py:346-375), per element */
static double repair(double v, double lo, double hi, int64_t *nrep) {
    int changed = 0, it;
    double d = v - lo, scale;
    if (d < 0.0) {
        changed = 1; scale = 1.0; it = 0;
        do { v = v - (1 + scale) * d; d = v - lo; scale /= 2; } while (d < 0.0 && ++it < 4096);
    }
    d = v - hi;
    if (d > 0.0) {
        changed = 1; scale = 1.0; it = 0;
        do { v = v - (1 + scale) * d; d = v - hi; scale /= 2; } while (d > 0.0 && ++it < 4096);
    }
    if (v < lo) { v = lo; changed = 1; }
    if (v > hi) { v = hi; changed = 1; }
    *nrep += changed;
    return v;
}

static void gene_draws(const orc_t *s, int64_t gen, int64_t row, int col, double *uf, double *uc) {
    uint32_t c[4] = {(uint32_t)col >> 1, (uint32_t)row, (uint32_t)((uint64_t)row >> 32), (uint32_t)gen};
    philox(c, (uint32_t)s->seed, (uint32_t)(s->seed >> 32));
    *uf = u32_unit((col & 1) ? c[2] : c[0]);
    *uc = u32_unit((col & 1) ? c[3] : c[1]);
}

orc_t *orc_create(int obj, int D, int64_t NP, const double *lb, const double *ub, const double *x0, double F, double Cr,
                  uint64_t seed, int stream, int invert, const double *aux) {
    orc_t *s = (orc_t *)calloc(1, sizeof(orc_t));
    int j;
    s->obj = obj; s->D = D; s->NP = NP; s->F = F; s->Cr = Cr; s->seed = seed; s->stream = stream; s->invert = invert;
    s->lb = (double *)malloc(sizeof(double) * D); memcpy(s->lb, lb, sizeof(double) * D);
    s->ub = (double *)malloc(sizeof(double) * D); memcpy(s->ub, ub, sizeof(double) * D);
    s->aux = (double *)calloc(D, sizeof(double));
    if (aux) memcpy(s->aux, aux, sizeof(double) * D);
    s->x0 = (double *)calloc(D, sizeof(double));
    if (x0) { memcpy(s->x0, x0, sizeof(double) * D); for (j = 0; j < D; j++) s->have_x0 |= isfinite(x0[j]) != 0; }
    s->pop = (double *)malloc(sizeof(double) * NP * D);
    s->trial = (double *)malloc(sizeof(double) * NP * D);
    s->vals = (double *)malloc(sizeof(double) * NP);
    s->tvals = (double *)malloc(sizeof(double) * NP);
    s->accept = (uint8_t *)calloc(NP, 1);
    s->ib = (int64_t *)malloc(sizeof(int64_t) * NP);
    s->r1 = (int64_t *)malloc(sizeof(int64_t) * NP);
    s->r2 = (int64_t *)malloc(sizeof(int64_t) * NP);
    s->best_x = (double *)malloc(sizeof(double) * D);
    s->hndvi = 1;
    s->rdir = (int64_t *)malloc(sizeof(int64_t) * 2);
    s->last_best_x = (double *)calloc(D, sizeof(double));
    s->b1 = (double *)calloc(D, sizeof(double));
    s->b2 = (double *)calloc(D, sizeof(double));
    s->cvals = (double *)calloc(NP, sizeof(double));
    s->tcvals = (double *)calloc(NP, sizeof(double));
    s->trial_best_feasible = 1;
    if (stream == STREAM_CPYTHON) {
        s->Uf = (double *)malloc(sizeof(double) * NP * D);
        s->Uc = (double *)malloc(sizeof(double) * NP * D);
        mt_seed(&s->mt, seed);
    }
    return s;
}

/* call before orc_init */
void orc_set_strategy(orc_t *s, int base_best, int dir_best, int const_F, int hndvi) {
    s->base_best = base_best; s->dir_best = dir_best; s->const_F = const_F; s->hndvi = hndvi > 0 ? hndvi : 1;
    free(s->r1); free(s->r2); free(s->rdir);
    s->r1 = (int64_t *)malloc(sizeof(int64_t) * s->NP * s->hndvi);
    s->r2 = (int64_t *)malloc(sizeof(int64_t) * s->NP * s->hndvi);
    s->rdir = (int64_t *)malloc(sizeof(int64_t) * 2 * s->hndvi);
}

/* call before orc_init; matrices row-major [m][D] in the order A, Aeq, Qc, Qh */
void orc_set_constraints(orc_t *s, const int *m, const double *const *M, const double *const *rhs, double contol) {
    int k;
    for (k = 0; k < 4; k++) {
        s->cm[k] = m[k];
        if (!m[k]) continue;
        s->cM[k] = (double *)malloc(sizeof(double) * m[k] * s->D);
        memcpy(s->cM[k], M[k], sizeof(double) * m[k] * s->D);
        s->crhs[k] = (double *)malloc(sizeof(double) * m[k]);
        memcpy(s->crhs[k], rhs[k], sizeof(double) * m[k]);
        s->constrained = 1;
    }
    s->contol = contol;
}

void orc_destroy(orc_t *s) {
    int k;
    if (!s) return;
    free(s->rdir); free(s->last_best_x); free(s->b1); free(s->b2); free(s->cvals); free(s->tcvals);
    for (k = 0; k < 4; k++) { free(s->cM[k]); free(s->crhs[k]); }
    free(s->lb); free(s->ub); free(s->aux); free(s->x0); free(s->pop); free(s->trial); free(s->vals); free(s->tvals);
    free(s->accept); free(s->ib); free(s->r1); free(s->r2); free(s->best_x); free(s->Uf); free(s->Uc); free(s);
}

/* vals[isnan] = inf; first-occurrence argmin (de_oo.py:301-305) */
static int64_t eval_rows(orc_t *s, const double *X, double *out) {
    const int64_t NP = s->NP;
    const int D = s->D;
    int64_t i, bi = 0;
#pragma omp parallel
    {
        double *t = (double *)malloc(sizeof(double) * (D + 1));
#pragma omp for schedule(static)
        for (i = 0; i < NP; i++) {
            double f = objective(s, X + i * D, t);
            if (s->invert) f = -f;
            if (isnan(f)) f = INFINITY;
            out[i] = f;
        }
        free(t);
    }
    for (i = 1; i < NP; i++) if (out[i] < out[bi]) bi = i;
    return bi;
}

/* P.mr(checkBoxBounds=False) + nanPenalty*P.nNaNs() of one row (de_oo.py:314-316; kernel/Point.py:128-167,
 * 363-374): max(0, A x - b, c(x), |Aeq x - beq|, |h(x)|) with NaN residuals skipped, + 1e10 per NaN of c, h. */
static double constr_val(const orc_t *s, const double *x, double *t) {
    const int D = s->D;
    double r = 0.0;
    int kind, k, j, nnan = 0;
    for (kind = 0; kind < 4; kind++)
        for (k = 0; k < s->cm[kind]; k++) {
            const double *co = s->cM[kind] + (int64_t)k * D;
            double res;
            if (kind < 2) {                         /* linear: left-to-right accumulation */
                double acc = 0.0;
                for (j = 0; j < D; j++) acc += co[j] * x[j];
                res = acc - s->crhs[kind][k];
            } else {                                /* (Q * (x*x)).sum(-1) - r: numpy pairwise sum of the products */
                for (j = 0; j < D; j++) t[j] = co[j] * (x[j] * x[j]);
                res = pairwise_sum(t, D) - s->crhs[kind][k];
            }
            if (kind & 1) res = fabs(res);
            if (isnan(res)) { if (kind >= 2) nnan++; }
            else if (r < res) r = res;
        }
    return r + 1e10 * nnan;
}

static void eval_constraints(orc_t *s, const double *X, double *out) {
    int64_t i;
#pragma omp parallel
    {
        double *t = (double *)malloc(sizeof(double) * (s->D + 1));
#pragma omp for schedule(static)
        for (i = 0; i < s->NP; i++) out[i] = constr_val(s, X + i * s->D, t);
        free(t);
    }
}

/* _eval_pop's best row under general constraints (de_oo.py:318-338) */
static int64_t constrained_best(orc_t *s, const double *vals, const double *c, int *feasible) {
    int64_t i, bi = -1;
    for (i = 0; i < s->NP; i++)
        if (c[i] < s->contol && (bi < 0 || vals[i] < vals[bi])) bi = i;
    *feasible = bi >= 0;
    if (bi >= 0) return bi;
    bi = 0;
    for (i = 1; i < s->NP; i++) if (c[i] < c[bi]) bi = i;
    return bi;
}

/* numpy structured sort with order=['constr_val','val'], ties by the index field; NaN sorts last */
typedef struct { int64_t i; double c, f; } dir_rec;
static int dcmp(double a, double b) {
    if (a < b) return -1;
    if (a > b) return 1;
    if (a == b) return 0;
    if (isnan(a) && isnan(b)) return 0;
    return isnan(b) ? -1 : 1;
}
static int dir_cmp(const void *pa, const void *pb) {
    const dir_rec *a = (const dir_rec *)pa, *b = (const dir_rec *)pb;
    int c = dcmp(a->c, b->c);
    if (!c) c = dcmp(a->f, b->f);
    if (!c) c = (a->i > b->i) - (a->i < b->i);
    return c;
}

void orc_init(orc_t *s) {
    const int64_t NP = s->NP;
    const int D = s->D;
    int64_t i, bi;
    int j;
    if (s->stream == STREAM_CPYTHON) {
        for (i = 0; i < NP * D; i++) s->pop[i] = mt_random(&s->mt);            /* Rand(NP,D), de_oo.py:147 */
    } else {
#pragma omp parallel for private(j) schedule(static)
        for (i = 0; i < NP; i++)
            for (j = 0; j < D; j++) { double uf, uc; gene_draws(s, 0, i, j, &uf, &uc); s->pop[i * D + j] = uf; }
    }
    for (i = 0; i < NP; i++)
        for (j = 0; j < D; j++) s->pop[i * D + j] = s->pop[i * D + j] * (s->ub[j] - s->lb[j]) + s->lb[j];
    if (s->have_x0) memcpy(s->pop, s->x0, sizeof(double) * D);                   /* de_oo.py:149-150 */
    bi = eval_rows(s, s->pop, s->vals);
    memcpy(s->tvals, s->vals, sizeof(double) * NP);
    if (s->constrained) {
        eval_constraints(s, s->pop, s->cvals);
        memcpy(s->tcvals, s->cvals, sizeof(double) * NP);
        bi = constrained_best(s, s->vals, s->cvals, &s->trial_best_feasible);
    }
    s->trial_best_c = s->cvals[bi];
    s->gen = 0; s->trial_best_idx = bi; s->trial_best_f = s->best_f = s->vals[bi];
    memcpy(s->best_x, s->pop + bi * D, sizeof(double) * D);
    memcpy(s->last_best_x, s->pop + bi * D, sizeof(double) * D);
    s->best_changed = 1; s->n_accepted = 0; s->n_repaired = 0;
}

void orc_step(orc_t *s, int ngens) {
    const int64_t NP = s->NP;
    const int D = s->D;
    int g;
    for (g = 0; g < ngens; g++) {
        int64_t i, nrep = 0, nacc = 0, bi;
        const int64_t gen = ++s->gen;
        const int h = s->hndvi;
        int hk;
        if (s->stream == STREAM_CPYTHON) {
            /* draw order of de_oo.py: :188 (only 'random' base), :200-201 two (h,NP) blocks or :210 one block of
             * 2h, :236 (only 'random' factor), :245 */
            if (!s->base_best) for (i = 0; i < NP; i++) s->ib[i] = mt_randbelow(&s->mt, (uint32_t)NP);
            if (s->dir_best) for (i = 0; i < 2 * h; i++) s->rdir[i] = mt_randbelow(&s->mt, (uint32_t)NP);
            else {
                for (i = 0; i < NP * h; i++) s->r1[i] = mt_randbelow(&s->mt, (uint32_t)NP);
                for (i = 0; i < NP * h; i++) s->r2[i] = mt_randbelow(&s->mt, (uint32_t)NP);
            }
            if (!s->const_F) for (i = 0; i < NP * D; i++) s->Uf[i] = mt_random(&s->mt);
            for (i = 0; i < NP * D; i++) s->Uc[i] = mt_random(&s->mt);
        } else {
            /* oracle/de_oracle.py::PhiloxStream: counter[0] = 0 -> ib, r1[0]; 1 -> r2[0]; 1+k -> r1[k], r2[k];
             * directed search: counter = [0, m, TAG_DIRECT << 24, generation] */
#pragma omp parallel for schedule(static)
            for (i = 0; i < NP; i++) {
                uint32_t a[4] = {0u, (uint32_t)i, (uint32_t)((uint64_t)i >> 32) | (1u << 24), (uint32_t)gen};
                uint32_t b[4] = {1u, a[1], a[2], a[3]};
                int k;
                philox(a, (uint32_t)s->seed, (uint32_t)(s->seed >> 32));
                philox(b, (uint32_t)s->seed, (uint32_t)(s->seed >> 32));
                s->ib[i] = mulhi64(a[0], a[1], (uint64_t)NP);
                s->r1[i] = mulhi64(a[2], a[3], (uint64_t)NP);
                s->r2[i] = mulhi64(b[0], b[1], (uint64_t)NP);
                for (k = 1; k < h; k++) {
                    uint32_t c[4] = {1u + (uint32_t)k, (uint32_t)i, (uint32_t)((uint64_t)i >> 32) | (1u << 24), (uint32_t)gen};
                    philox(c, (uint32_t)s->seed, (uint32_t)(s->seed >> 32));
                    s->r1[(int64_t)k * NP + i] = mulhi64(c[0], c[1], (uint64_t)NP);
                    s->r2[(int64_t)k * NP + i] = mulhi64(c[2], c[3], (uint64_t)NP);
                }
            }
            for (i = 0; s->dir_best && i < 2 * h; i++) {
                uint32_t c[4] = {0u, (uint32_t)i, 2u << 24, (uint32_t)gen};
                philox(c, (uint32_t)s->seed, (uint32_t)(s->seed >> 32));
                s->rdir[i] = mulhi64(c[0], c[1], (uint64_t)NP);
            }
        }
        if (s->dir_best) {      /* de_oo.py:209-231: one delta vector for the whole population */
            dir_rec *rec = (dir_rec *)malloc(sizeof(dir_rec) * 2 * h);
            int j;
            for (i = 0; i < 2 * h; i++) { rec[i].i = s->rdir[i]; rec[i].c = s->cvals[s->rdir[i]]; rec[i].f = s->vals[s->rdir[i]]; }
            qsort(rec, 2 * h, sizeof(dir_rec), dir_cmp);
            for (j = 0; j < D; j++) {
                double sb = 0.0, sw = 0.0;
                for (hk = 0; hk < h; hk++) {
                    const double xb = s->pop[rec[hk].i * D + j], xw = s->pop[rec[h + hk].i * D + j];
                    sb = hk ? sb + xb : xb;
                    sw = hk ? sw + xw : xw;
                }
                sb /= h; sw /= h;
                if (h != 1) { sb /= h; sw /= h; }          /* divided twice: reference quirk :222-223,229-231 */
                s->b1[j] = sw; s->b2[j] = sb;
            }
            free(rec);
        }
#pragma omp parallel reduction(+ : nrep)
        {
            double *t = (double *)malloc(sizeof(double) * (D + 1));
#pragma omp for schedule(static)
            for (i = 0; i < NP; i++) {
                const double *xo = s->pop + i * D, *xb = s->base_best ? s->last_best_x : s->pop + s->ib[i] * D;
                double *tr = s->trial + i * D, f;
                int j, k;
                for (j = 0; j < D; j++) {
                    double uf = 0.0, uc, v, d1, d2, Ft;
                    if (s->stream == STREAM_CPYTHON) { if (!s->const_F) uf = s->Uf[i * D + j]; uc = s->Uc[i * D + j]; }
                    else gene_draws(s, gen, i, j, &uf, &uc);
                    if (s->dir_best) { d1 = s->b1[j]; d2 = s->b2[j]; }
                    else {      /* barycentres of h gathered rows, added in order, / h when h != 1  :198-208,229-231 */
                        d1 = s->pop[s->r1[i] * D + j]; d2 = s->pop[s->r2[i] * D + j];
                        for (k = 1; k < h; k++) {
                            d1 += s->pop[s->r1[(int64_t)k * NP + i] * D + j];
                            d2 += s->pop[s->r2[(int64_t)k * NP + i] * D + j];
                        }
                        if (h != 1) { d1 /= h; d2 /= h; }
                    }
                    /* de_oo.py:235-250: Ft = U*F (or F); mutant = beta + Ft*delta; parent gene kept iff Uc > Cr */
                    Ft = s->const_F ? s->F : uf * s->F;
                    v = (uc > s->Cr) ? xo[j] : xb[j] + Ft * (d2 - d1);
                    tr[j] = repair(v, s->lb[j], s->ub[j], &nrep);
                }
                f = objective(s, tr, t);
                if (s->invert) f = -f;
                if (isnan(f)) f = INFINITY;
                s->tvals[i] = f;
            }
            free(t);
        }
        bi = 0;
        for (i = 1; i < NP; i++) if (s->tvals[i] < s->tvals[bi]) bi = i;
        s->trial_best_feasible = 1;
        if (s->constrained) {
            eval_constraints(s, s->trial, s->tcvals);
            bi = constrained_best(s, s->tvals, s->tcvals, &s->trial_best_feasible);
        }
        s->trial_best_c = s->constrained ? s->tcvals[bi] : 0.0;
        memcpy(s->last_best_x, s->trial + bi * D, sizeof(double) * D);
        /* greedy selection (de_oo.py:260-282), vals (and constr_vals) blended arithmetically like the reference:
         * if either side violates anything the smaller constr_val wins, otherwise the smaller f; ties -> trial */
#pragma omp parallel for reduction(+ : nacc) schedule(static)
        for (i = 0; i < NP; i++) {
            int keep = s->vals[i] < s->tvals[i];
            double k;
            if (s->constrained) {
                const double oc = s->cvals[i], nc = s->tcvals[i];
                double kc;
                if (nc > 0.0 || oc > 0.0) keep = oc < nc;
                kc = keep ? 1.0 : 0.0;
                s->cvals[i] = oc * kc + nc * (1.0 - kc);
            }
            k = keep ? 1.0 : 0.0;
            s->vals[i] = s->vals[i] * k + s->tvals[i] * (1.0 - k);
            s->accept[i] = (uint8_t)!keep;
            if (!keep) { memcpy(s->pop + i * D, s->trial + i * D, sizeof(double) * D); nacc++; }
        }
        s->trial_best_idx = bi; s->trial_best_f = s->tvals[bi];
        /* de_oo.py:285-286; with general constraints Best is the host's business (Point.betterThan) */
        s->best_changed = s->constrained ? 1 : !(s->best_f < s->tvals[bi]);
        if (s->best_changed) { s->best_f = s->tvals[bi]; memcpy(s->best_x, s->trial + bi * D, sizeof(double) * D); }
        s->n_accepted = nacc; s->n_repaired = nrep;
    }
}

const double *orc_pop(const orc_t *s) { return s->pop; }
const double *orc_vals(const orc_t *s) { return s->vals; }
const double *orc_tvals(const orc_t *s) { return s->tvals; }
const uint8_t *orc_accept(const orc_t *s) { return s->accept; }
const double *orc_best_x(const orc_t *s) { return s->best_x; }
double orc_best_f(const orc_t *s) { return s->best_f; }
double orc_trial_best_f(const orc_t *s) { return s->trial_best_f; }
int64_t orc_trial_best_idx(const orc_t *s) { return s->trial_best_idx; }
int64_t orc_n_accepted(const orc_t *s) { return s->n_accepted; }
int64_t orc_n_repaired(const orc_t *s) { return s->n_repaired; }
int orc_best_changed(const orc_t *s) { return s->best_changed; }
const double *orc_cvals(const orc_t *s) { return s->cvals; }
int orc_trial_best_feasible(const orc_t *s) { return s->trial_best_feasible; }
double orc_trial_best_c(const orc_t *s) { return s->trial_best_c; }
int orc_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
/* bench.py: torchrun exports OMP_NUM_THREADS=1 to every rank, which would silently make the CPU baseline
 * single-threaded; the caller sets the thread count it wants (and reports it). */
void orc_set_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#else
    (void)n;
#endif
}
void orc_eval(orc_t *s, const double *X, int64_t n, double *out) {
    int64_t i;
    double *t = (double *)malloc(sizeof(double) * (s->D + 1));
    for (i = 0; i < n; i++) { double f = objective(s, X + i * s->D, t); out[i] = s->invert ? -f : f; }
    free(t);
}
