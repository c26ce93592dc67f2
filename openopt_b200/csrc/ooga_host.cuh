// Host side of the galileo C ABI (include/ooga.h).  Included at the end of oode.cu: it uses that file's error
// plumbing (fail, CU), its device-block cache and the de library's row evaluator (eval_device_rows) for the objective.
// The kernels are in ga_kernels.cuh; the algorithm they restate is openopt/solvers/Standalone/galileo_oo.py:27-92
// (oracle/galileo_oracle.py is the CPU restatement the parity tests compare with).
#pragma once
#include "../../include/ooga.h"
#include "ga_kernels.cuh"

struct ooga_handle {
    ooga_config cfg{};
    oode_handle *ev = nullptr;          // a 4-row de handle: objective registry entry, reduction plan, stream, scratch
    int D = 0;
    int64_t N = 0, P = 0, n = 0, n2 = 0, ld = 0;
    uint64_t tmask = 0;
    int max_batch = 64;
    bool alias_path = false;            // crossoverRate < 1: entries can stay aliased to current chromosomes
    bool inited = false;
    int64_t gen = 0;
    int64_t launches = 0;
    int cur = 0;
    // device memory (all from the evaluator handle's block cache)
    double *lb = nullptr, *ub = nullptr;
    double *pop[2] = {nullptr, nullptr}, *fit[2] = {nullptr, nullptr};
    double *scal = nullptr, *child = nullptr, *fchild = nullptr, *bestx = nullptr;
    int32_t *picks = nullptr, *cut = nullptr, *owner = nullptr;
    uint8_t *state = nullptr, *keep = nullptr;
    uint64_t *hash[2] = {nullptr, nullptr};   // row hashes by entry number, ping-ponged with the population
    unsigned long long *tkey = nullptr, *counts = nullptr;
    uint32_t *tmin = nullptr;
    ooga::SortEntry *ent = nullptr, *top = nullptr;
    int64_t n2c = 0;                    // sort size of this generation's entries alone (merge-by-rank path)
    bool merge_sort = true;             // OOGA_SORT=full: sort the whole list every iteration (diagnostics)
    bool sorted = false;                // the current generation is in descending fitness order (after the first truncate)
    ooga::GaRecord *rec = nullptr;
    ooga::ScanArrays scan{};            // evaluate: running sums, their maxima, scratch of the large-population path
    bool scan_large = false;
    // OODE_RNG_CPYTHON: the reference's generator, consumed on the host in the reference's order
    MT19937CPython *mt = nullptr;
    double *d_wheel = nullptr, *d_mutu = nullptr;
    int32_t *d_cutin = nullptr;
    std::vector<double> h_wheel, h_mutu;
    std::vector<int32_t> h_cut;
    // timing (ooga_get_timing): device time of the generations of every ooga_step call; per phase when profiling is on
    bool profiling = false;
    cudaEvent_t pev[OOGA_PHASES + 1] = {}, t0 = nullptr, t1 = nullptr;
    double phase_ms[OOGA_PHASES] = {}, total_ms = 0.0;
    int64_t timed_gens = 0, profiled_gens = 0;
};

static_assert(sizeof(ooga::GaRecord) == sizeof(ooga_gen_record), "GaRecord must mirror ooga_gen_record");

static int ga_grid(const ooga_handle *g, int64_t items, int per_block) {
    const int64_t want = (items + per_block - 1) / per_block;
    return (int)std::max<int64_t>(1, std::min<int64_t>(want, (int64_t)g->ev->sm_count * 8));
}

static ooga::GaParams ga_params(const ooga_handle *g, uint32_t gen) {
    ooga::GaParams p{};
    p.N = g->N; p.P = g->P; p.D = g->D; p.ld = (int)g->ld;
    p.lb = g->lb; p.ub = g->ub;
    p.cr = g->cfg.crossover_rate; p.mr = g->cfg.mutation_rate;
    p.gen = gen;
    p.philox = g->cfg.rng == OODE_RNG_PHILOX ? 1 : 0;
    p.keys = g->ev->keys;
    p.wheel_u = g->d_wheel; p.cut_in = g->d_cutin; p.mut_u = g->d_mutu;
    p.cur = g->pop[g->cur]; p.fit = g->fit[g->cur];
    p.lmax = g->scan.lmax; p.crun = g->scan.crun; p.nchunks = g->scan.nchunks; p.scal = g->scal;
    p.child = g->child; p.picks = g->picks; p.cut = g->cut; p.state = g->state; p.owner = g->owner; p.hash = g->hash[g->cur];
    return p;
}

static ooga::RowSet ga_rowset(const ooga_handle *g) {
    ooga::RowSet s{};
    s.cur = g->pop[g->cur]; s.child = g->child;
    s.N = g->N; s.n = g->n; s.D = g->D; s.ld = (int)g->ld;
    s.state = g->state;
    return s;
}

__global__ void ga_fitness_kernel(const double *f, int64_t n, double *fit) {       // evalFunc = -p.f (:39)
    for (int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; k < n; k += (int64_t)gridDim.x * blockDim.x)
        fit[k] = ooga::child_fitness(f[k]);
}

static int ga_create_impl(const ooga_config *cfg, ooga_handle *g) {
    oode_config dc{};
    dc.abi_version = OODE_ABI_VERSION;
    dc.dim = cfg->dim;
    dc.population = 4;                  // the evaluator never runs a de generation; its scratch grows with the rows it is given
    dc.lb = cfg->lb; dc.ub = cfg->ub; dc.x0 = nullptr;
    dc.diff_factor = 0.8; dc.cross_rate = 0.5;
    dc.seed = cfg->seed;
    dc.objective = cfg->objective; dc.invert = cfg->invert; dc.obj_aux = cfg->obj_aux;
    dc.rng = OODE_RNG_PHILOX;
    dc.device = cfg->device;
    dc.world_size = 1; dc.rank = 0; dc.shard = OODE_SHARD_NONE; dc.max_batch = 1;
    dc.hndvi = 1;
    int rc = oode_create(&dc, &g->ev);
    if (rc) return rc;
    oode_handle *h = g->ev;
    g->cfg = *cfg;
    g->cfg.lb = g->cfg.ub = g->cfg.obj_aux = nullptr;      // the caller's arrays are not kept
    g->D = cfg->dim;
    g->N = cfg->population;
    g->P = (g->N + 1) / 2;
    g->n = g->N + 2 * g->P;
    g->ld = h->ld;
    g->n2 = ooga::SORT_TILE;
    while (g->n2 < g->n) g->n2 <<= 1;
    uint64_t tsize = 1024;
    while (tsize < 2 * (uint64_t)g->n) tsize <<= 1;
    g->tmask = tsize - 1;
    g->max_batch = cfg->max_batch > 0 ? cfg->max_batch : 64;
    g->alias_path = !(cfg->crossover_rate >= 1.0);
    CU(cudaSetDevice(cfg->device));
    DA(g->lb, g->D); DA(g->ub, g->D);
    CU(cudaMemcpyAsync(g->lb, cfg->lb, g->D * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    CU(cudaMemcpyAsync(g->ub, cfg->ub, g->D * sizeof(double), cudaMemcpyHostToDevice, h->stream));
    for (int b = 0; b < 2; ++b) { DA(g->pop[b], (size_t)g->N * g->ld); DA(g->fit[b], g->N); }
    DA(g->scal, 4);
    {
        ooga::ScanArrays &a = g->scan;
        a.N = g->N;
        a.nchunks = (int)((g->N + ooga::SCAN_CHUNK - 1) / ooga::SCAN_CHUNK);
        a.scal = g->scal;
        DA(a.prefix, g->N); DA(a.lmax, g->N); DA(a.crun, a.nchunks);
        const char *force = getenv("OOGA_SCAN");                       // diagnostics: OOGA_SCAN=small|large forces a path
        g->scan_large = force ? !strcmp(force, "large") : g->N > 2048;
        if (g->scan_large) {
            DA(a.tot, a.nchunks); DA(a.approx_in, a.nchunks); DA(a.bestv, a.nchunks); DA(a.cmax, a.nchunks);
            DA(a.besti, a.nchunks); DA(a.qpre, g->N); DA(a.qsum, a.nchunks); DA(a.qmin, a.nchunks); DA(a.qmax, a.nchunks);
            DA(a.carry, a.nchunks); DA(a.cexp, a.nchunks); DA(a.direct, a.nchunks);
        }
    }
    DA(g->child, (size_t)2 * g->P * g->ld); DA(g->fchild, 2 * g->P);
    DA(g->picks, 2 * g->P); DA(g->cut, g->P); DA(g->state, 2 * g->P); DA(g->keep, 2 * g->P);
    if (g->alias_path) DA(g->owner, (size_t)g->N * g->D);
    DA(g->hash[0], g->n); DA(g->hash[1], g->n); DA(g->tkey, tsize); DA(g->tmin, tsize);
    DA(g->ent, g->n2); DA(g->top, g->N); DA(g->counts, 4);
    g->n2c = ooga::SORT_TILE;
    while (g->n2c < 2 * g->P) g->n2c <<= 1;
    {
        const char *force = getenv("OOGA_SORT");
        g->merge_sort = !(force && !strcmp(force, "full"));
    }
    DA(g->rec, g->max_batch); DA(g->bestx, (size_t)g->max_batch * g->D);
    // rows of both buffers are read whole by the evaluator (pitch ld >= D) and entries that stay aliased are evaluated
    // along with the fresh ones: everything must hold finite numbers from the start
    for (int b = 0; b < 2; ++b) CU(cudaMemsetAsync(g->pop[b], 0, (size_t)g->N * g->ld * sizeof(double), h->stream));
    CU(cudaMemsetAsync(g->child, 0, (size_t)2 * g->P * g->ld * sizeof(double), h->stream));
    CU(cudaMemsetAsync(g->picks, 0, 2 * g->P * sizeof(int32_t), h->stream));
    CU(cudaMemsetAsync(g->cut, 0xFF, g->P * sizeof(int32_t), h->stream));
    if (cfg->rng == OODE_RNG_CPYTHON) {
        g->mt = new MT19937CPython(cfg->seed);
        DA(g->d_wheel, 2 * g->P); DA(g->d_cutin, g->P); DA(g->d_mutu, (size_t)g->N * g->D);
        g->h_wheel.resize(2 * g->P); g->h_cut.resize(g->P); g->h_mutu.resize((size_t)g->N * g->D);
    }
    for (int k = 0; k <= OOGA_PHASES; ++k) CU(cudaEventCreate(&g->pev[k]));
    CU(cudaEventCreate(&g->t0));
    CU(cudaEventCreate(&g->t1));
    CU(cudaStreamSynchronize(h->stream));
    return OODE_OK;
}

// the draws of one iteration in the reference's order of consumption: 2P wheel spins (:295-297, :337), per pair the
// crossover probability and, if it fires, the cut point (:360-363), then per gene of the first N entries the mutation
// probability and, if it fires, the uniform behind the new value (:430-439)
static int ga_cpython_draws(ooga_handle *g) {
    MT19937CPython &mt = *g->mt;
    for (int64_t m = 0; m < 2 * g->P; ++m) g->h_wheel[m] = mt.random();
    for (int64_t p = 0; p < g->P; ++p) {
        const double prob = mt.random();
        g->h_cut[p] = prob <= g->cfg.crossover_rate ? (int32_t)mt.randbelow((uint32_t)g->D) : -1;
    }
    const size_t nd = (size_t)g->N * g->D;
    for (size_t t = 0; t < nd; ++t) {
        const double prob = mt.random();
        g->h_mutu[t] = prob <= g->cfg.mutation_rate ? mt.random() : -1.0;
    }
    cudaStream_t st = g->ev->stream;
    CU(cudaMemcpyAsync(g->d_wheel, g->h_wheel.data(), 2 * g->P * sizeof(double), cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(g->d_cutin, g->h_cut.data(), g->P * sizeof(int32_t), cudaMemcpyHostToDevice, st));
    CU(cudaMemcpyAsync(g->d_mutu, g->h_mutu.data(), nd * sizeof(double), cudaMemcpyHostToDevice, st));
    return OODE_OK;
}

static int ga_sort(ooga_handle *g, int64_t n2) {
    cudaStream_t st = g->ev->stream;
    const int tiles = (int)(n2 / ooga::SORT_TILE);
    ooga::ga_bitonic_tile_kernel<<<tiles, ooga::SORT_THREADS, 0, st>>>(g->ent, 2, ooga::SORT_TILE, ooga::SORT_TILE / 2);
    g->launches++;
    for (int64_t k = 2 * ooga::SORT_TILE; k <= n2; k <<= 1) {
        for (int64_t j = k >> 1; j >= ooga::SORT_TILE; j >>= 1) {
            ooga::ga_bitonic_step_kernel<<<ga_grid(g, n2 / 2, 256), 256, 0, st>>>(g->ent, n2, k, j);
            g->launches++;
        }
        ooga::ga_bitonic_tile_kernel<<<tiles, ooga::SORT_THREADS, 0, st>>>(g->ent, k, k, ooga::SORT_TILE / 2);
        g->launches++;
    }
    CU(cudaGetLastError());
    return OODE_OK;
}

// evaluate (:255-276): running sums in list order, their running maxima, sumFitness, the best chromosome
static int ga_scan(ooga_handle *g, const double *fit) {
    cudaStream_t st = g->ev->stream;
    ooga::ScanArrays a = g->scan;
    a.fit = fit;
    if (!g->scan_large) {
        ooga::ga_scan_small_kernel<<<1, ooga::SCAN_CHUNK, 0, st>>>(a);
        g->launches++;
    } else {
        const int grid = (int)std::min<int64_t>(a.nchunks, (int64_t)g->ev->sm_count * 8);
        ooga::ga_scan_totals_kernel<<<grid, ooga::SCAN_THREADS, 0, st>>>(a);
        ooga::ga_scan_predict_kernel<<<1, ooga::SCAN_CHUNK, 0, st>>>(a);
        ooga::ga_scan_quantize_kernel<<<grid, ooga::SCAN_THREADS, 0, st>>>(a);
        ooga::ga_scan_carry_kernel<<<1, ooga::SCAN_THREADS, 0, st>>>(a);
        ooga::ga_scan_finish_kernel<<<grid, ooga::SCAN_THREADS, 0, st>>>(a);
        ooga::ga_scan_crun_kernel<<<1, ooga::SCAN_CHUNK, 0, st>>>(a);
        g->launches += 6;
    }
    CU(cudaGetLastError());
    return OODE_OK;
}

// profiling: an event after each phase of a generation; the generation is then synchronised and its phase times added up
#define GA_MARK(k)                                                     \
    do {                                                               \
        if (g->profiling) CU(cudaEventRecord(g->pev[k], st));           \
    } while (0)

static int ga_generation(ooga_handle *g, int slot) {
    cudaStream_t st = g->ev->stream;
    const uint32_t gen = (uint32_t)(g->gen + 1);
    if (g->cfg.rng == OODE_RNG_CPYTHON) { const int rc = ga_cpython_draws(g); if (rc) return rc; }
    ooga::GaParams p = ga_params(g, gen);
    GA_MARK(0);
    // evaluate (:255-276)
    {
        const int rc = ga_scan(g, g->fit[g->cur]);
        if (rc) return rc;
    }
    CU(cudaMemsetAsync(g->counts, 0, 4 * sizeof(unsigned long long), st));
    GA_MARK(1);
    // select, crossover, mutate (:290-312, :281-288)
    ooga::ga_select_kernel<<<ga_grid(g, 2 * g->P, 256), 256, 0, st>>>(p);
    ooga::ga_children_kernel<<<ga_grid(g, g->P, 8), 256, 0, st>>>(p);
    g->launches += 2;
    GA_MARK(2);
    const ooga::RowSet s = ga_rowset(g);
    if (g->alias_path) {
        CU(cudaMemsetAsync(g->owner, 0, (size_t)g->N * g->D * sizeof(int32_t), st));
        ooga::ga_alias_kernel<0><<<ga_grid(g, g->N, 8), 256, 0, st>>>(p);
        ooga::ga_alias_kernel<1><<<ga_grid(g, g->N, 8), 256, 0, st>>>(p);
        g->launches += 2;
    }
    GA_MARK(3);
    // replace (:458-474): duplicates, the objective of the fresh entries, sort, truncate
    if (g->alias_path) {                 // current chromosomes may have changed in place: their hashes are recomputed
        ooga::RowSet cur_rows = s;
        cur_rows.n = g->N;
        ooga::ga_hash_kernel<<<ga_grid(g, g->N, 8), 256, 0, st>>>(cur_rows, g->hash[g->cur]);
        g->launches++;
    }
    CU(cudaGetLastError());
    GA_MARK(4);
    {
        const int64_t before = g->ev->ctr.gpu_launches;
        const int rc = eval_device_rows(g->ev, g->child, 2 * g->P, g->fchild);
        if (rc) return rc;
        g->launches += g->ev->ctr.gpu_launches - before;
    }
    GA_MARK(5);
    CU(cudaMemsetAsync(g->tkey, 0, (g->tmask + 1) * sizeof(unsigned long long), st));
    CU(cudaMemsetAsync(g->tmin, 0xFF, (g->tmask + 1) * sizeof(uint32_t), st));
    ooga::ga_table_insert_kernel<<<ga_grid(g, g->n, 256), 256, 0, st>>>(s, g->hash[g->cur], g->tkey, g->tmin, g->tmask);
    ooga::ga_dedupe_kernel<<<ga_grid(g, 2 * g->P, 8), 256, 0, st>>>(s, g->hash[g->cur], g->tkey, g->tmin, g->tmask, g->keep);
    // sort / reverse (:472-473).  First iteration: the whole list.  Later ones: the current generation is already in
    // order, so only the entries are sorted and the two are merged by rank (ga_merge_rank_kernel).
    const bool merge = g->merge_sort && g->sorted;
    const int64_t nsort = merge ? g->n2c : g->n2;
    ooga::ga_keys_kernel<<<ga_grid(g, nsort, 256), 256, 0, st>>>(s, g->fit[g->cur], g->fchild, g->keep, g->ent, merge ? g->N : 0,
                                                                 nsort, g->counts);
    g->launches += 3;
    CU(cudaGetLastError());
    GA_MARK(6);
    {
        const int rc = ga_sort(g, nsort);
        if (rc) return rc;
    }
    const ooga::SortEntry *order = g->ent;
    if (merge) {
        ooga::ga_merge_rank_kernel<<<ga_grid(g, g->N + g->n2c, 256), 256, 0, st>>>(s, g->fit[g->cur], g->ent, g->n2c, g->top);
        g->launches++;
        order = g->top;
    }
    GA_MARK(7);
    // what __solver__ reports (:86-89), then the next current generation
    ooga::ga_record_kernel<<<1, 256, 0, st>>>(g->pop[g->cur], (int)g->ld, g->D, g->scal, g->counts, (int32_t)gen,
                                              g->rec + slot, g->bestx + (size_t)slot * g->D);
    ooga::ga_gather_kernel<<<ga_grid(g, g->N, 8), 256, 0, st>>>(s, order, g->fit[g->cur], g->fchild, g->hash[g->cur],
                                                                g->pop[g->cur ^ 1], g->fit[g->cur ^ 1], g->hash[g->cur ^ 1]);
    g->launches += 2;
    CU(cudaGetLastError());
    GA_MARK(8);
    if (g->profiling) {
        CU(cudaEventSynchronize(g->pev[OOGA_PHASES]));
        for (int k = 0; k < OOGA_PHASES; ++k) {
            float ms = 0.f;
            CU(cudaEventElapsedTime(&ms, g->pev[k], g->pev[k + 1]));
            g->phase_ms[k] += ms;
        }
        g->profiled_gens++;
    }
    g->cur ^= 1;
    g->gen++;
    g->sorted = true;
    return OODE_OK;
}

extern "C" {

int ooga_create(const ooga_config *cfg, ooga_handle **out) {
    NvtxRange nvtx_range("ooga_create");
    if (!cfg || !out) return fail(OODE_EINVAL, "ooga_create: NULL argument");
    *out = nullptr;
    if (cfg->abi_version != OODE_ABI_VERSION) return fail(OODE_EINVAL, "ooga_create: ABI version mismatch");
    if (cfg->dim < 1 || cfg->population < 1) return fail(OODE_EINVAL, "ooga_create: dim and population must be >= 1");
    if (cfg->population >= (1ll << 30)) return fail(OODE_EINVAL, "ooga_create: population must be < 2^30");
    if (!cfg->lb || !cfg->ub) return fail(OODE_EINVAL, "ooga_create: lb and ub are required");
    if (cfg->rng != OODE_RNG_PHILOX && cfg->rng != OODE_RNG_CPYTHON)
        return fail(OODE_EINVAL, "ooga_create: rng must be OODE_RNG_PHILOX or OODE_RNG_CPYTHON");
    if (!(cfg->crossover_rate == cfg->crossover_rate) || !(cfg->mutation_rate == cfg->mutation_rate))
        return fail(OODE_EINVAL, "ooga_create: crossover_rate / mutation_rate is NaN");
    ooga_handle *g = new ooga_handle();
    const int rc = ga_create_impl(cfg, g);
    if (rc) {
        const std::string keep = g_err;
        ooga_destroy(g);
        g_err = keep;
        return rc;
    }
    *out = g;
    return OODE_OK;
}

void ooga_destroy(ooga_handle *g) {
    if (!g) return;
    delete g->mt;
    for (int k = 0; k <= OOGA_PHASES; ++k)
        if (g->pev[k]) cudaEventDestroy(g->pev[k]);
    if (g->t0) cudaEventDestroy(g->t0);
    if (g->t1) cudaEventDestroy(g->t1);
    if (g->ev) oode_destroy(g->ev);     // every device block of this handle came from the evaluator's cache list
    delete g;
}

int ooga_init(ooga_handle *g) {
    NvtxRange nvtx_range("ooga_init");
    if (!g) return fail(OODE_EINVAL, "ooga_init: NULL handle");
    oode_handle *h = g->ev;
    CU(cudaSetDevice(g->cfg.device));
    cudaStream_t st = h->stream;
    double *dU = nullptr;
    if (g->cfg.rng == OODE_RNG_CPYTHON) {          // Chromosome.randomInit (:112-134): one random() per gene, list order
        delete g->mt;
        g->mt = new MT19937CPython(g->cfg.seed);
        const size_t nd = (size_t)g->N * g->D;
        g->mt->fill_random(g->h_mutu.data(), nd);
        dU = g->d_mutu;
        CU(cudaMemcpyAsync(dU, g->h_mutu.data(), nd * sizeof(double), cudaMemcpyHostToDevice, st));
    }
    g->cur = 0;
    g->gen = 0;
    g->sorted = false;
    ooga::GaParams p = ga_params(g, 0);
    ooga::ga_init_kernel<<<ga_grid(g, g->N * g->D, 256), 256, 0, st>>>(p, dU);
    g->launches++;
    CU(cudaGetLastError());
    {
        const int64_t before = h->ctr.gpu_launches;
        const int rc = eval_device_rows(h, g->pop[0], g->N, g->fchild);      // fchild holds >= N doubles
        if (rc) return rc;
        g->launches += h->ctr.gpu_launches - before;
    }
    ga_fitness_kernel<<<ga_grid(g, g->N, 256), 256, 0, st>>>(g->fchild, g->N, g->fit[0]);
    {
        ooga::RowSet cur_rows = ga_rowset(g);
        cur_rows.n = g->N;
        ooga::ga_hash_kernel<<<ga_grid(g, g->N, 8), 256, 0, st>>>(cur_rows, g->hash[0]);
    }
    g->launches += 2;
    CU(cudaGetLastError());
    CU(cudaStreamSynchronize(st));
    g->inited = true;
    return OODE_OK;
}

int ooga_step(ooga_handle *g, int32_t n_gens, ooga_gen_record *out, double *best_x) {
    NvtxRange nvtx_range("ooga_step");
    if (!g || !out) return fail(OODE_EINVAL, "ooga_step: NULL argument");
    if (!g->inited) return fail(OODE_ESTATE, "ooga_step: call ooga_init first");
    if (n_gens < 1 || n_gens > g->max_batch) return fail(OODE_EINVAL, "ooga_step: n_gens must be in [1, max_batch]");
    CU(cudaSetDevice(g->cfg.device));
    cudaStream_t st = g->ev->stream;
    CU(cudaEventRecord(g->t0, st));
    for (int k = 0; k < n_gens; ++k) {
        const int rc = ga_generation(g, k);
        if (rc) return rc;
    }
    CU(cudaEventRecord(g->t1, st));
    CU(cudaMemcpyAsync(out, g->rec, n_gens * sizeof(ooga_gen_record), cudaMemcpyDeviceToHost, st));
    if (best_x) CU(cudaMemcpyAsync(best_x, g->bestx, (size_t)n_gens * g->D * sizeof(double), cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    {
        float ms = 0.f;
        CU(cudaEventElapsedTime(&ms, g->t0, g->t1));
        g->total_ms += ms;
        g->timed_gens += n_gens;
    }
    return OODE_OK;
}

int ooga_set_profiling(ooga_handle *g, int32_t on) {
    if (!g) return fail(OODE_EINVAL, "ooga_set_profiling: NULL handle");
    g->profiling = on != 0;
    return OODE_OK;
}

int ooga_get_timing(ooga_handle *g, ooga_timing *out) {
    if (!g || !out) return fail(OODE_EINVAL, "ooga_get_timing: NULL argument");
    out->total_ms = g->total_ms;
    out->generations = g->timed_gens;
    out->profiled_generations = g->profiled_gens;
    for (int k = 0; k < OOGA_PHASES; ++k) out->phase_ms[k] = g->phase_ms[k];
    out->gpu_launches = g->launches;
    return OODE_OK;
}

int ooga_get_population(ooga_handle *g, double *genes, double *fitness) {
    if (!g) return fail(OODE_EINVAL, "ooga_get_population: NULL handle");
    if (!g->inited) return fail(OODE_ESTATE, "ooga_get_population: call ooga_init first");
    CU(cudaSetDevice(g->cfg.device));
    cudaStream_t st = g->ev->stream;
    if (genes)
        CU(cudaMemcpy2DAsync(genes, g->D * sizeof(double), g->pop[g->cur], g->ld * sizeof(double), g->D * sizeof(double), g->N,
                             cudaMemcpyDeviceToHost, st));
    if (fitness) CU(cudaMemcpyAsync(fitness, g->fit[g->cur], g->N * sizeof(double), cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    return OODE_OK;
}

int ooga_get_selection(ooga_handle *g, int64_t *picks, int64_t *cut) {
    if (!g || !picks || !cut) return fail(OODE_EINVAL, "ooga_get_selection: NULL argument");
    if (!g->inited) return fail(OODE_ESTATE, "ooga_get_selection: call ooga_init first");
    CU(cudaSetDevice(g->cfg.device));
    std::vector<int32_t> a(2 * g->P), b(g->P);
    CU(cudaMemcpyAsync(a.data(), g->picks, a.size() * sizeof(int32_t), cudaMemcpyDeviceToHost, g->ev->stream));
    CU(cudaMemcpyAsync(b.data(), g->cut, b.size() * sizeof(int32_t), cudaMemcpyDeviceToHost, g->ev->stream));
    CU(cudaStreamSynchronize(g->ev->stream));
    for (size_t i = 0; i < a.size(); ++i) picks[i] = a[i];
    for (size_t i = 0; i < b.size(); ++i) cut[i] = b[i];
    return OODE_OK;
}

int64_t ooga_gpu_launches(ooga_handle *g) { return g ? g->launches : 0; }

// ---- population I/O and exact checkpoint / resume (the reference has neither: kernel/runProbSolver.py:49-52) ----
int ooga_set_population(ooga_handle *g, const double *genes, const double *fitness) {
    if (!g || !genes) return fail(OODE_EINVAL, "ooga_set_population: NULL argument");
    if (!g->inited) return fail(OODE_ESTATE, "ooga_set_population: call ooga_init first");
    oode_handle *h = g->ev;
    CU(cudaSetDevice(g->cfg.device));
    cudaStream_t st = h->stream;
    CU(cudaMemcpy2DAsync(g->pop[g->cur], g->ld * sizeof(double), genes, g->D * sizeof(double), g->D * sizeof(double), g->N,
                         cudaMemcpyHostToDevice, st));
    if (fitness) CU(cudaMemcpyAsync(g->fit[g->cur], fitness, g->N * sizeof(double), cudaMemcpyHostToDevice, st));
    else {                                   // cached fitness = -p.f(genes) (:39, :143-149)
        const int64_t before = h->ctr.gpu_launches;
        const int rc = eval_device_rows(h, g->pop[g->cur], g->N, g->fchild);
        if (rc) return rc;
        ga_fitness_kernel<<<ga_grid(g, g->N, 256), 256, 0, st>>>(g->fchild, g->N, g->fit[g->cur]);
        g->launches += h->ctr.gpu_launches - before + 1;
    }
    ooga::RowSet cur_rows = ga_rowset(g);
    cur_rows.n = g->N;
    ooga::ga_hash_kernel<<<ga_grid(g, g->N, 8), 256, 0, st>>>(cur_rows, g->hash[g->cur]);
    g->launches++;
    CU(cudaGetLastError());
    CU(cudaStreamSynchronize(st));
    g->sorted = false;                       // the caller's list order is taken as it is
    return OODE_OK;
}

struct GaStateHeader {
    uint64_t magic;
    int64_t N, gen;
    int32_t D, rng, objective, sorted;
    double cr, mr;
    uint64_t seed;
};
constexpr uint64_t GA_STATE_MAGIC = 0x4F4F474153544154ull;       // "OOGASTAT"

int64_t ooga_state_size(ooga_handle *g) {
    if (!g) return fail(OODE_EINVAL, "ooga_state_size: NULL handle");
    return (int64_t)sizeof(GaStateHeader) + g->N * g->D * 8 + g->N * 8 +
           (g->cfg.rng == OODE_RNG_CPYTHON ? (int64_t)MT19937CPython::STATE_WORDS * 4 : 0);
}

int ooga_save_state(ooga_handle *g, void *buf) {
    if (!g || !buf) return fail(OODE_EINVAL, "ooga_save_state: NULL argument");
    if (!g->inited) return fail(OODE_ESTATE, "ooga_save_state: call ooga_init first");
    unsigned char *p = (unsigned char *)buf;
    GaStateHeader hd{};
    hd.magic = GA_STATE_MAGIC; hd.N = g->N; hd.gen = g->gen; hd.D = g->D; hd.rng = g->cfg.rng; hd.objective = g->cfg.objective;
    hd.sorted = g->sorted ? 1 : 0; hd.cr = g->cfg.crossover_rate; hd.mr = g->cfg.mutation_rate; hd.seed = g->cfg.seed;
    memcpy(p, &hd, sizeof(hd));
    double *genes = (double *)(p + sizeof(hd)), *fit = genes + g->N * g->D;
    const int rc = ooga_get_population(g, genes, fit);
    if (rc) return rc;
    if (g->cfg.rng == OODE_RNG_CPYTHON) g->mt->save((uint32_t *)(fit + g->N));
    return OODE_OK;
}

int ooga_load_state(ooga_handle *g, const void *buf, int64_t size) {
    if (!g || !buf) return fail(OODE_EINVAL, "ooga_load_state: NULL argument");
    if (!g->inited) return fail(OODE_ESTATE, "ooga_load_state: call ooga_init first");
    if (size != ooga_state_size(g)) return fail(OODE_EINVAL, "ooga_load_state: the buffer does not have this handle's state size");
    const unsigned char *p = (const unsigned char *)buf;
    GaStateHeader hd;
    memcpy(&hd, p, sizeof(hd));
    if (hd.magic != GA_STATE_MAGIC || hd.N != g->N || hd.D != g->D || hd.rng != g->cfg.rng || hd.objective != g->cfg.objective ||
        hd.cr != g->cfg.crossover_rate || hd.mr != g->cfg.mutation_rate || hd.seed != g->cfg.seed)
        return fail(OODE_EINVAL, "ooga_load_state: the state was saved by a handle with another configuration");
    if (hd.gen < 0) return fail(OODE_EINVAL, "ooga_load_state: corrupt state (negative generation)");
    const double *genes = (const double *)(p + sizeof(hd)), *fit = genes + g->N * g->D;
    const int rc = ooga_set_population(g, genes, fit);
    if (rc) return rc;
    if (g->cfg.rng == OODE_RNG_CPYTHON) g->mt->load((const uint32_t *)(fit + g->N));
    g->gen = hd.gen;
    g->sorted = hd.sorted != 0;
    return OODE_OK;
}

int ooga_scan_values(ooga_handle *g, const double *values, double *running_sums, double *sum, int64_t *best_idx) {
    if (!g || !values) return fail(OODE_EINVAL, "ooga_scan_values: NULL argument");
    if (!g->inited) return fail(OODE_ESTATE, "ooga_scan_values: call ooga_init first");
    CU(cudaSetDevice(g->cfg.device));
    cudaStream_t st = g->ev->stream;
    double *tmp = g->fit[g->cur ^ 1];              // the spare fitness array: rewritten by the next generation's truncate step
    CU(cudaMemcpyAsync(tmp, values, g->N * sizeof(double), cudaMemcpyHostToDevice, st));
    const int rc = ga_scan(g, tmp);
    if (rc) return rc;
    double sc[4];
    if (running_sums) CU(cudaMemcpyAsync(running_sums, g->scan.prefix, g->N * sizeof(double), cudaMemcpyDeviceToHost, st));
    CU(cudaMemcpyAsync(sc, g->scal, sizeof(sc), cudaMemcpyDeviceToHost, st));
    CU(cudaStreamSynchronize(st));
    if (sum) *sum = sc[0];
    if (best_idx) memcpy(best_idx, &sc[2], sizeof(int64_t));
    return OODE_OK;
}

}  // extern "C"
