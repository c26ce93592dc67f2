/*
 * ooga.h -- C ABI of the `galileo` genetic algorithm in liboode (the sibling population
 * solver behind the same boundary as de, include/oode.h).
 *
 * Replaces the body of
 *     galileo.__solver__(p)                 openopt/solvers/Standalone/galileo_oo.py:27-92
 * i.e. the hard-wired configuration of Donald Goodman's GA classes in the same file (:35-66): Population.prepPopulation
 * (:237-253), evaluate (:255-276), select / select_Roulette (:290-301, :326-346), crossover / crossover_OnePoint
 * (:303-312, :354-369), mutate / mutate_Default (:281-288, :425-441) and replace_SteadyStateNoDuplicates (:458-474).
 * The reference has no FFI; the plugin openopt_b200/solvers/cuda/galileo_oo.py keeps its plugin contract and calls
 * these functions through ctypes.
 *
 * Conventions are those of oode.h: plain C, 0 or a negative OODE_E* code, the message from oode_last_error(), the caller
 * owns host buffers, one host thread per handle, IEEE float64 throughout.  Objectives are the registered device
 * objectives of oode.h (OODE_OBJ_* or an id from oode_register_objective); there is no CPU fallback.
 *
 * Reference behaviour kept on purpose (each pinned by a recorded run of the reference, the .npz files under tests/golden/galileo/):
 *   - fitness = -p.f(x) is cached per chromosome and never refreshed (:143-160);
 *   - sumFitness and the roulette's running sums are accumulated left to right in list order (:268, :340-343); with
 *     negative fitness values the wheel therefore lands on the first or the last chromosome only;
 *   - an odd population selects and crosses ceil(N/2) pairs but mutates only the first N entries (:287, :295, :309);
 *   - a pair that is not crossed (prob > crossoverRate) stays ALIASED to the current chromosomes: its mutations change
 *     the current generation in place, the cached fitness goes stale, and the entry is never appended (:299-300, :368);
 *   - replacement drops an entry whose genes equal those of any chromosome already in the list, then sorts by cached
 *     fitness (stable ascending, reversed): ties go to the LATER list position (:465-474);
 *   - the values reported per iteration are those of the evaluate step at its START (:86-89).
 */
#ifndef OOGA_H
#define OOGA_H

#include <stdint.h>
#include "oode.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct ooga_handle ooga_handle;

typedef struct ooga_config {
    int32_t abi_version;      /* OODE_ABI_VERSION                                                            */
    int32_t dim;              /* D = p.n = len(chromoMinValues)                       galileo_oo.py:42-45    */
    int64_t population;       /* galileo.population (15)                              :20, :36               */
    const double *lb;         /* [D] finite: P.chromoMinValues                        :42                    */
    const double *ub;         /* [D] finite: P.chromoMaxValues                        :45                    */
    double crossover_rate;    /* galileo.crossoverRate (1.0)                          :21, :51               */
    double mutation_rate;     /* galileo.mutationRate (0.05)                          :22, :54               */
    uint64_t seed;            /* the reference's generator is unseeded (:223); runs here are reproducible     */
    int32_t objective;        /* OODE_OBJ_* / registered id                                                  */
    int32_t invert;           /* 1: p.f is the negated user objective (goal='max')                            */
    const double *obj_aux;    /* [D] per-column objective data or NULL                                       */
    int32_t rng;              /* OODE_RNG_PHILOX: device draws; OODE_RNG_CPYTHON: the library regenerates the
                                 stream of random.Random(seed) in the reference's order of consumption        */
    int32_t device;           /* CUDA device ordinal                                                         */
    int32_t max_batch;        /* most generations per ooga_step call; 0 = default (64)                       */
    int32_t reserved;
} ooga_config;

/* One iteration of the reference's loop (:73-92). */
typedef struct ooga_gen_record {
    double best_f;            /* fval = -P.maxFitness of the evaluate step                      :86          */
    double sum_fitness;       /* P.sumFitness                                                   :268         */
    int64_t best_idx;         /* list position of P.bestFitIndividual                           :269-271     */
    int64_t n_appended;       /* entries appended by replace = objective evaluations of the iteration :470-471 */
    int64_t n_aliased;        /* entries left aliased by crossover (never appended)             :368-369     */
    int64_t n_duplicates;     /* fresh entries dropped by the isIdentical scan                  :467-469     */
    int32_t generation;       /* 1-based iteration number                                                    */
    int32_t reserved;
} ooga_gen_record;

int ooga_create(const ooga_config *cfg, ooga_handle **out);
void ooga_destroy(ooga_handle *h);

/* prepPopulation (:237-253) and the first evaluation of every chromosome (:255-266). */
int ooga_init(ooga_handle *h);

/* n_gens iterations of the loop (:75-84).  out[n_gens]; best_x[n_gens][D] (or NULL) receives the genes of
 * bestFitIndividual as __solver__ passes them to p.iterfcn (:89). */
int ooga_step(ooga_handle *h, int32_t n_gens, ooga_gen_record *out, double *best_x);

/* currentGeneration in list order: genes [N][D] (or NULL) and cached fitness [N] (or NULL). */
int ooga_get_population(ooga_handle *h, double *genes, double *fitness);

/* Replace currentGeneration: genes [N][D] in list order; fitness [N] = the cached values to keep, or NULL to evaluate
 * -p.f(genes) (:39).  The list order is taken as given (the next replacement sorts the whole list again). */
int ooga_set_population(ooga_handle *h, const double *genes, const double *fitness);

/* Exact checkpoint / resume (the reference has none: a solve is one-shot, kernel/runProbSolver.py:49-52): everything the
 * next iteration depends on -- genes, cached fitness, iteration number, whether the list is in sorted order, and the
 * state of the CPython generator in OODE_RNG_CPYTHON mode.  A handle of the same configuration that loads the blob
 * (after ooga_init) continues bit for bit. */
int64_t ooga_state_size(ooga_handle *h);
int ooga_save_state(ooga_handle *h, void *buf);
int ooga_load_state(ooga_handle *h, const void *buf, int64_t size);

/* Verification tap: the last iteration's roulette picks [2*ceil(N/2)] and cut points [ceil(N/2)] (-1 = not crossed). */
int ooga_get_selection(ooga_handle *h, int64_t *picks, int64_t *cut);

/* Verification tap of the evaluate step (:255-276) on caller-supplied fitness values [N]: the running sums in list
 * order (:268; [N] or NULL), their total and the index of the first strictly largest value (:269-271).  The handle's
 * population is not touched. */
int ooga_scan_values(ooga_handle *h, const double *values, double *running_sums, double *sum, int64_t *best_idx);

/* Kernels launched on behalf of this handle so far (the objective evaluator's included). */
int64_t ooga_gpu_launches(ooga_handle *h);

/* Device time.  total_ms: CUDA events on the handle's stream around the generations of every ooga_step call.  With
 * profiling on, every generation is also timed phase by phase (an event after each phase; the generation is then
 * synchronised, so profiled steps are slower): */
enum {
    OOGA_PHASE_EVALUATE = 0,   /* running fitness sums, best chromosome                       :255-276           */
    OOGA_PHASE_CHILDREN = 1,   /* roulette, one-point crossover, mutation of the fresh entries :290-312, :425-441 */
    OOGA_PHASE_ALIASES = 2,    /* in-place mutation of the entries that were not crossed      :368, :425-441     */
    OOGA_PHASE_HASH = 3,       /* row hashes of the current chromosomes and the entries       :465-469           */
    OOGA_PHASE_OBJECTIVE = 4,  /* -p.f of the fresh entries                                   :39, :143-149      */
    OOGA_PHASE_DUPLICATES = 5, /* hash join + exact comparison, sort keys                     :465-471           */
    OOGA_PHASE_SORT = 6,       /* sort / reverse                                              :472-473           */
    OOGA_PHASE_TRUNCATE = 7,   /* the first N entries become currentGeneration; the iteration's report :474, :86-89 */
    OOGA_PHASES = 8
};
typedef struct ooga_timing {
    double total_ms;
    double phase_ms[OOGA_PHASES];
    int64_t generations;            /* generations behind total_ms                                              */
    int64_t profiled_generations;   /* generations behind phase_ms                                              */
    int64_t gpu_launches;
} ooga_timing;
int ooga_set_profiling(ooga_handle *h, int32_t on);
int ooga_get_timing(ooga_handle *h, ooga_timing *out);

#ifdef __cplusplus
}
#endif
#endif /* OOGA_H */
