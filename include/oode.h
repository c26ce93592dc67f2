/*
 * oode.h -- C ABI of liboode: the B200-native differential-evolution generation loop.
 *
 * This library replaces ONE path of OpenOpt 0.43: the body of
 *     de.__solver__(p)                      openopt/solvers/UkrOpt/de_oo.py:118-290
 * i.e. population initialisation (:147-155), the per-generation mutation / binomial crossover /
 * bound repair / objective evaluation / greedy selection (:178-282, :292-305, :346-375) and the
 * trial-best bookkeeping the solver hands to p.iterfcn (:257, :285-288).
 *
 * The reference has no FFI of its own (it is pure Python); the boundary it offers is the solver
 * plugin contract (kernel/baseSolver.py:5-33, kernel/nonOptMisc.py:89-130).  The plugin
 * openopt_b200/solvers/cuda/de_oo.py keeps that contract and calls the functions below through
 * ctypes; INTEGRATION.md shows the same stub dropped into the reference tree.
 *
 * Conventions: plain C, no exceptions cross the boundary.  Every function returns 0 on success or
 * a negative OODE_E* code; oode_last_error() gives the message (thread-local).  The caller owns
 * every host buffer; the library owns device buffers, streams and NCCL communicators.  One host
 * thread per handle.  All floating point data is IEEE float64, as in the reference.
 */
#ifndef OODE_H
#define OODE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define OODE_ABI_VERSION 3

enum {
    OODE_OK = 0,
    OODE_EINVAL = -1,    /* bad argument / configuration                       */
    OODE_ECUDA = -2,     /* CUDA runtime failure (message has the CUDA error)  */
    OODE_ENCCL = -3,     /* NCCL failure                                       */
    OODE_ESTATE = -4,    /* call out of order (e.g. step before init)          */
    OODE_ENOMEM = -5
};

/* Registered device-side objectives.  The host-callable numpy form of each one is the parity
 * definition (openopt_b200/objectives.py); the reference evaluates the
 * user's f row by row through kernel/nonLinFuncs.py:190-200. */
enum {
    OODE_OBJ_SPHERE = 0,          /* (x*x).sum()                                                  */
    OODE_OBJ_ROSENBROCK = 1,      /* (100.0*(x[1:]-x[:-1]**2)**2 + (1.0-x[:-1])**2).sum()         */
    OODE_OBJ_RASTRIGIN = 2,       /* (x*x - 10*cos(2*pi*x) + 10).sum()   openopt/tests/rastrigin.py:4 */
    OODE_OBJ_ACKLEY = 3,          /* -20*exp(-0.2*sqrt((x*x).sum()/D)) - exp(cos(2*pi*x).sum()/D) + 20 + e */
    OODE_OBJ_GRIEWANK = 4,        /* 1 + (x*x).sum()/4000 - cos(x/aux).prod(),  aux[j] = sqrt(j+1) */
    OODE_OBJ_SHIFTED_SPHERE = 5,  /* ((x-aux)**2).sum()                  examples/glp_3.py:6      */
    OODE_OBJ_COUNT = 6            /* built-in objectives; ids >= OODE_OBJ_COUNT come from oode_register_objective */
};

/* Where the random draws of de_oo.py:147,188,200-201,236,245 come from. */
enum {
    OODE_RNG_PHILOX = 0,   /* Philox4x32-10 on the device, keyed (seed, generation, row, column)    */
    OODE_RNG_REPLAY = 1,   /* the caller supplies every draw (bit-exact replay of the reference)    */
    OODE_RNG_CPYTHON = 2   /* the library regenerates CPython's random (MT19937) stream on the host:
                              same draws as the reference's Rand/Randint (de_oo.py:16-53)           */
};

/* The solver's strategy parameters (de_oo.py:71-89,157-176).  All zero = the reference's defaults. */
enum {
    OODE_BASE_RANDOM = 0,      /* baseVectorStrategy 'random': beta = old_pop[Randint(NP,size=NP)]   :188   */
    OODE_BASE_BEST = 1         /* 'best': every row of beta is the best row of the last evaluated
                                  (trial) population                                          :191-193 */
};
enum {
    OODE_DIR_RANDOM = 0,       /* searchDirectionStrategy 'random': barycentres of hndvi random rows
                                  per trial                                                   :198-208 */
    OODE_DIR_BEST = 1          /* 'best' (directed): one set of 2*hndvi rows for the whole population,
                                  sorted by (constr_val, val); delta = best - worst barycentre :209-227 */
};
enum {
    OODE_FACTOR_RANDOM = 0,    /* differenceFactorStrategy 'random': Ft = Rand(NP,D)*F        :235-237 */
    OODE_FACTOR_CONSTANT = 1   /* 'constant': Ft = F                                          :238-239 */
};
#define OODE_MAX_HNDVI 512

/* How the population is split when world_size > 1 (one process per GPU). */
enum {
    OODE_SHARD_NONE = 0,
    OODE_SHARD_ROWS = 1,   /* contiguous row blocks, full replica on every GPU; each generation the
                              accepted rows are pushed into the peers' replicas (peer memory)       */
    OODE_SHARD_COLS = 2    /* contiguous column blocks cut on the pairwise-sum tree; only per-row
                              partial sums cross the GPUs                                           */
};

/* How a sharded handle moves its per-row partial sums between the GPUs (oode_exchange_mode). */
enum {
    OODE_XCHG_NONE = 0,    /* single GPU                                                             */
    OODE_XCHG_NCCL = 1,    /* NCCL all-gather in row chunks on a second stream                       */
    OODE_XCHG_PEER = 2     /* the trial kernel stores them straight into every rank's copy over
                              NVLink (CUDA-IPC mapped peer memory) and the ranks meet at a flag
                              barrier; default whenever the ranks can map each other's memory
                              (environment OODE_EXCHANGE=nccl|p2p forces one)                        */
};

typedef struct oode_handle oode_handle;

typedef struct oode_config {
    int32_t abi_version;      /* OODE_ABI_VERSION                                                   */
    int32_t dim;              /* D = p.n                                   de_oo.py:128             */
    int64_t population;       /* NP                                        de_oo.py:130-133         */
    const double *lb;         /* [D] finite                                de_oo.py:122,127         */
    const double *ub;         /* [D] finite                                                         */
    const double *x0;         /* [D] or NULL; injected as individual 0 if any entry is finite :149  */
    double diff_factor;       /* F  = de.differenceFactor (0.8)            de_oo.py:87,135          */
    double cross_rate;        /* Cr = de.crossoverRate (0.5)               de_oo.py:88,137          */
    uint64_t seed;            /* de.seed (150880)                          de_oo.py:90,144          */
    int32_t objective;        /* OODE_OBJ_*                                                         */
    int32_t invert;           /* 1: minimise -f (goal='max')               kernel/nonLinFuncs.py:297 */
    const double *obj_aux;    /* [D] per-column objective data or NULL                               */
    int32_t rng;              /* OODE_RNG_*                                                         */
    int32_t device;           /* CUDA device ordinal                                                */
    int32_t world_size;       /* 1 = single GPU                                                     */
    int32_t rank;
    int32_t shard;            /* OODE_SHARD_*                                                       */
    int32_t max_batch;        /* most generations one oode_step call may run (history ring size);
                                 0 = default (64)                                                   */
    const void *nccl_id;      /* world_size > 1: the 128-byte ncclUniqueId shared by all ranks      */
    int32_t base_strategy;    /* OODE_BASE_*    de.baseVectorStrategy        de_oo.py:71            */
    int32_t direction_strategy; /* OODE_DIR_*   de.searchDirectionStrategy   de_oo.py:77            */
    int32_t factor_strategy;  /* OODE_FACTOR_*  de.differenceFactorStrategy  de_oo.py:84            */
    int32_t hndvi;            /* de.hndvi (half of the rows in a difference vector); 0 means 1  :89 */
} oode_config;

/* What the solver needs from one generation to call p.iterfcn(Best) (de_oo.py:256-288).
 * With general constraints (oode_set_constraints) the trial-best is _eval_pop's constrained choice
 * (de_oo.py:318-338) and Best is NOT tracked by the library: Point.betterThan needs the host's own
 * max-residual of the point (kernel/Point.py:280-340), so best_f mirrors trial_best_f, best_changed
 * is always 1 and oode_get_best_x returns the trial-best row of that generation. */
typedef struct oode_gen_record {
    double trial_best_f;      /* f of the best row of the trial population      de_oo.py:301-305    */
    double best_f;            /* f of Best after `if Old_best.betterThan(Best)`  de_oo.py:285-286    */
    int64_t trial_best_idx;   /* first-occurrence argmin                        de_oo.py:304        */
    int64_t n_accepted;       /* trials that replaced their parent              de_oo.py:260-277    */
    int64_t n_repaired;       /* genes changed by _correct_box_constraints      de_oo.py:346-375    */
    int32_t best_changed;     /* 1 if Best switched to this generation's trial-best                 */
    int32_t generation;       /* 0 = initial population                                             */
    double trial_best_c;      /* constr_val of that row (0 without general constraints)  :316       */
    int32_t trial_best_feasible; /* 1 if some row had constr_val < contol       de_oo.py:318-320    */
    int32_t reserved;
} oode_gen_record;

/* General constraints (de_oo.py:306-338; kernel/Point.py:128-167).  Every row of the population gets
 *   constr_val = max(0, A x - b, c(x), |Aeq x - beq|, |h(x)|) + 1e10 * (number of NaN among c(x), h(x))
 * (NaN residuals are skipped by the max), and selection prefers the smaller constr_val whenever either
 * side violates anything (de_oo.py:260-273).  Nonlinear c / h must be registered device constraints; the one
 * family provided is the diagonal quadratic  g_k(x) = sum_j Q[k][j]*x[j]^2 - r[k]
 * (openopt_b200/constraints.py).  All matrices are row-major [rows][D]; a count of 0 leaves its pointers unused. */
typedef struct oode_constraints {
    int32_t n_lin_ineq;       /* rows of A      A x <= b        kernel/residuals.py:33-41           */
    int32_t n_lin_eq;         /* rows of Aeq    Aeq x = beq     kernel/residuals.py:43-51           */
    int32_t n_quad_ineq;      /* rows of Qc     c(x) <= 0                                           */
    int32_t n_quad_eq;        /* rows of Qh     h(x) = 0                                            */
    const double *A, *b;
    const double *Aeq, *beq;
    const double *Qc, *rc;
    const double *Qh, *rh;
    double contol;            /* p.contol as the solver sees it (the driver has already scaled it)  */
} oode_constraints;

/* Draws of ONE generation in OODE_RNG_REPLAY mode, in the reference's draw order.  A strategy that
 * does not draw an item leaves its pointer NULL (h = hndvi). */
typedef struct oode_replay_draws {
    const int64_t *ib;        /* [NP]   Randint(NP, size=NP)       de_oo.py:188   NULL: base 'best' */
    const int64_t *r1;        /* [h*NP] Randint(NP, size=(h,NP))   de_oo.py:200   (row-major);
                                 directed search: [2h] Randint(NP, size=(2*h))    de_oo.py:210      */
    const int64_t *r2;        /* [h*NP] Randint(NP, size=(h,NP))   de_oo.py:201   NULL: directed    */
    const double *Uf;         /* [NP*D] Rand(NP,D), row-major      de_oo.py:236   NULL: constant F  */
    const double *Uc;         /* [NP*D] Rand(NP,D), row-major      de_oo.py:245                     */
} oode_replay_draws;

typedef struct oode_counters {
    double kernel_ms;         /* device time of the generations run so far (CUDA events)            */
    int64_t generations;
    int64_t trial_evals;      /* objective evaluations on the device (NP per generation + init)     */
    int64_t gpu_launches;     /* kernels launched by this handle                                    */
    int64_t bytes_algorithmic;/* (40*D+16) per trial evaluation, SURVEY.md section 8d               */
    int64_t h2d_bytes;        /* host->device bytes copied by this handle (config, draws, points)   */
    int64_t d2h_bytes;        /* device->host bytes copied (records, Best.x, population reads)      */
    double trial_kernel_ms;   /* with oode_set_profiling(h,1): device time inside de_trial_kernel   */
    double select_kernel_ms;  /*                               and inside de_select_kernel          */
    int64_t profiled_generations;
    double exchange_wait_ms;  /* peer exchange, profiling on.  Column shards: part of trial_kernel_ms spent
                                 after this rank's trial kernel had finished (publish + waiting for the
                                 slowest rank).  Row shards: time after K2b (push of the accepted rows,
                                 publish, wait, finish) -- not part of the other two */
} oode_counters;

int oode_abi_version(void);
int oode_device_count(void);
const char *oode_last_error(void);

/* NCCL bootstrap for world_size > 1: rank 0 calls this and ships the 128 bytes to the others. */
int oode_nccl_unique_id(void *id128);

/* Host-only: which columns [col0, col0+ncols) and which leaves [leaf0, leaf0+nleaves) of the row's
 * pairwise-sum tree rank `rank` of `world_size` owns under OODE_SHARD_COLS (no GPU needed). */
int oode_shard_layout(int32_t dim, int32_t objective, int32_t world_size, int32_t rank, int32_t *col0, int32_t *ncols,
                      int32_t *leaf0, int32_t *nleaves);

int oode_create(const oode_config *cfg, oode_handle **out);
void oode_destroy(oode_handle *h);

/* oode_destroy keeps the handle's device blocks in a process-wide cache for the next handle that asks
 * for the same sizes on the same device; this returns them to the CUDA driver. */
void oode_release_cached_memory(void);

/* Attach general constraints; call between oode_create and oode_init.  Not available on a sharded
 * handle. */
int oode_set_constraints(oode_handle *h, const oode_constraints *c);

/* Verification tap: constr_vals of the current population [NP] (zeros without general constraints). */
int oode_get_constr_vals(oode_handle *h, double *cvals);

/* Population init + first evaluation (de_oo.py:147-155).  u0: REPLAY mode only, the [NP*D]
 * uniforms of Rand(NP,D) at :147 (NULL otherwise).  Writes record 0. */
int oode_init(oode_handle *h, const double *u0, oode_gen_record *out);

/* Run n_gens generations (de_oo.py:178-290 minus the host-side p.iterfcn).  draws: REPLAY mode
 * only, n_gens entries.  out: n_gens records.  n_gens <= max_batch. */
int oode_step(oode_handle *h, int32_t n_gens, const oode_replay_draws *draws, oode_gen_record *out);

/* x of Best after generation `generation` (must lie within the last oode_step batch, or be the
 * current generation).  x: [D].  Collective on a sharded handle (every rank must call it). */
int oode_get_best_x(oode_handle *h, int64_t generation, double *x);

/* Current population (post-selection), row-major [NP*D], and its f values [NP]; either may be NULL.
 * Checkpoint / resume: oode_set_population restores both.  On a column-sharded handle each rank
 * reads / writes only ITS columns of the [NP*D] host array (the others are left untouched). */
int oode_get_population(oode_handle *h, double *pop, double *vals);
int oode_set_population(oode_handle *h, const double *pop, const double *vals);

/* Checkpoint / resume (the reference has none: a problem instance cannot even be re-run,
 * kernel/runProbSolver.py:49-52).  oode_save_state writes everything the NEXT generation depends on -- the
 * current population, f and constr_vals, Best, the last trial-best row, the next donor indices, the generation
 * number and the CPython generator state -- into a caller-owned buffer of oode_state_size(h) bytes;
 * oode_load_state restores it into a handle created with the same configuration (and the same constraints),
 * after which oode_step continues bit for bit as the saved run would have.  Not available on a sharded handle. */
int64_t oode_state_size(oode_handle *h);
int oode_save_state(oode_handle *h, void *buf);
int oode_load_state(oode_handle *h, const void *buf);

/* Verification taps for the last generation run: accept mask [NP] (1 = trial replaced parent)
 * and the trial population's f [NP]. */
int oode_get_accept_mask(oode_handle *h, uint8_t *mask);
int oode_get_trial_vals(oode_handle *h, double *tvals);

/* Verification taps that stay cheap at full size (NP = 1M .. 16M rows): n rows picked by global row number,
 * out = [n][D] row-major (a column-sharded handle fills its own column slice of every row).
 *   which = OODE_ROWS_CURRENT  the row's current version (pop[r] of de_oo.py:180)
 *   which = OODE_ROWS_SPARE    the row's other slot: the last trial if it was rejected, else the replaced parent
 *   which = OODE_ROWS_TRIAL    the trial built for the row by the last generation (de_oo.py:241-254)
 *   which = OODE_ROWS_PARENT   the row as that generation read it (old_pop[r], de_oo.py:180)
 * oode_get_donor_indices copies the donor rows [NP] the NEXT generation will use (Philox: drawn by the previous
 * selection pass; replay / CPython: the last ones uploaded), so a test can fetch exactly the rows a trial is
 * built from (de_oo.py:188,200-204) before running it.  r1, r2: first index row of an hndvi > 1 block. */
enum { OODE_ROWS_CURRENT = 0, OODE_ROWS_SPARE = 1, OODE_ROWS_TRIAL = 2, OODE_ROWS_PARENT = 3 };
int oode_get_rows(oode_handle *h, int32_t which, const int64_t *rows, int64_t n, double *out);
int oode_get_donor_indices(oode_handle *h, int64_t *ib, int64_t *r1, int64_t *r2);

/* Open objective registry.  The reference calls an arbitrary Python `f` one row at a time
 * (kernel/nonLinFuncs.py:190-200); here an objective is device code: a specialisation of the kernels' Obj<>
 * template (csrc/de_device.cuh).  openopt_b200/expr.py generates one from a declarative expression (per-gene term
 * expressions, sum / product slots, a finalising expression), compiles csrc/trial_obj.cu once more against it into
 * its own shared object and registers that here.  `library` = path of the shared object (it exports
 * oode_custom_objective_v1); returns the new objective id (>= OODE_OBJ_COUNT) to put into oode_config.objective,
 * or a negative error code.  Registering the same path twice returns the same id. */
int oode_register_objective(const char *library);

/* Evaluate the registered objective on n host points ([n*D] row-major) with the device kernels. */
int oode_eval_points(oode_handle *h, const double *x, int64_t n, double *f);

int oode_get_counters(oode_handle *h, oode_counters *out);

/* OODE_XCHG_* of this handle (negative on a NULL handle). */
int oode_exchange_mode(oode_handle *h);

/* Name of the trial-kernel form this handle launches for its (local) row width and bounds:
 * "de_trial_tma_kernel", "de_trial_tma2_kernel" or "de_trial_kernel" (static string). */
const char *oode_trial_kernel_name(oode_handle *h);

/* on != 0: bracket every kernel of oode_step with CUDA events on the handle's stream and
 * accumulate per-kernel device time into oode_counters (used by bench.py's roofline line). */
int oode_set_profiling(oode_handle *h, int32_t on);

/* Host-only (needs no GPU): how the Philox mode takes the crossover decision of de_oo.py:245-250 without
 * converting the draw: the parent's gene is kept iff the raw 32-bit word w > *threshold, or always if *keep_all
 * (uc = w * 2^-32 exactly, so this is `uc > crossoverRate` bit for bit). */
int oode_crossover_threshold(double cross_rate, uint32_t *threshold, int32_t *keep_all);

/* Host-only diagnostic (needs no GPU): the CPython `random` stream the OODE_RNG_CPYTHON mode feeds
 * the kernels with -- after random.seed(seed): n_ints draws of random.randint(0, below-1), then
 * n_doubles draws of random.random()  (de_oo.py:16-53). */
int oode_cpython_stream(uint64_t seed, uint32_t below, int64_t n_ints, int64_t *ints, int64_t n_doubles,
                        double *doubles);

#ifdef __cplusplus
}
#endif
#endif /* OODE_H */
