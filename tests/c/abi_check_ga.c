/* include/ooga.h used from plain C (c99, -pedantic): the boundary a non-Python host binds for the galileo solver.
 * Without a CUDA device ooga_create must fail loudly (no CPU fallback); with one, a 20-chromosome population is
 * created, initialised, advanced by three iterations, checkpointed and destroyed. */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "ooga.h"

int main(void) {
    double lb[3] = {-1, -1, -1}, ub[3] = {1, 2, 3}, genes[20 * 3], fit[20], bx[3 * 3];
    ooga_config cfg;
    ooga_handle *h = NULL;
    ooga_gen_record rec[3];
    ooga_timing tm;
    int rc, k;

    memset(&cfg, 0, sizeof cfg);
    cfg.abi_version = OODE_ABI_VERSION;
    cfg.dim = 3; cfg.population = 20; cfg.lb = lb; cfg.ub = ub;
    cfg.crossover_rate = 1.0; cfg.mutation_rate = 0.05; cfg.seed = 150880u; cfg.objective = OODE_OBJ_SPHERE;
    cfg.rng = OODE_RNG_CPYTHON;
    if (ooga_create(NULL, &h) != OODE_EINVAL || ooga_create(&cfg, NULL) != OODE_EINVAL) return 1;
    cfg.population = 0;
    if (ooga_create(&cfg, &h) != OODE_EINVAL || h != NULL) return 2;
    cfg.population = 20;
    cfg.rng = OODE_RNG_REPLAY;                         /* galileo has no replay mode */
    if (ooga_create(&cfg, &h) != OODE_EINVAL) return 3;
    cfg.rng = OODE_RNG_CPYTHON;
    rc = ooga_create(&cfg, &h);
    if (oode_device_count() == 0) {
        if (rc != OODE_ECUDA || h != NULL || !strstr(oode_last_error(), "no CPU fallback")) return 4;
        if (ooga_init(NULL) != OODE_EINVAL || ooga_gpu_launches(NULL) != 0) return 5;
    } else {
        void *blob;
        if (rc != OODE_OK) return 6;
        if (ooga_step(h, 1, rec, NULL) != OODE_ESTATE) return 7;          /* before ooga_init */
        if (ooga_init(h) != OODE_OK || ooga_step(h, 3, rec, bx) != OODE_OK) return 8;
        for (k = 0; k < 3; ++k)
            if (rec[k].generation != k + 1 || !(rec[k].best_f >= 0.0) || rec[k].n_appended + rec[k].n_aliased + rec[k].n_duplicates != 20) return 9;
        if (ooga_get_population(h, genes, fit) != OODE_OK) return 10;
        for (k = 1; k < 20; ++k)
            if (fit[k] > fit[k - 1]) return 11;                           /* list order = descending cached fitness */
        blob = malloc((size_t)ooga_state_size(h));
        if (!blob || ooga_save_state(h, blob) != OODE_OK || ooga_load_state(h, blob, ooga_state_size(h)) != OODE_OK) return 12;
        if (ooga_load_state(h, blob, 8) != OODE_EINVAL) return 13;
        free(blob);
        if (ooga_get_timing(h, &tm) != OODE_OK || tm.generations != 3 || ooga_gpu_launches(h) <= 0) return 14;
        ooga_destroy(h);
    }
    printf("c abi (galileo) ok: %d devices\n", oode_device_count());
    return 0;
}
