/* include/oode.h used from plain C (c99, -pedantic): the boundary a non-Python host binds.
 * Host-only entry points are exercised; oode_create must fail loudly without a CUDA device
 * (or succeed and be destroyed when one is present). */
#include <stdio.h>
#include <string.h>
#include <math.h>
#include "oode.h"

int main(void) {
    int32_t col0, ncols, leaf0, nleaves, r, covered = 0;
    int64_t ints[4];
    double dbl[2], lb[3] = {-1, -1, -1}, ub[3] = {1, 1, 1};
    oode_config cfg;
    oode_handle *h = NULL;
    int rc;

    if (oode_abi_version() != OODE_ABI_VERSION) { printf("abi version mismatch\n"); return 1; }
    /* 8 column shards of a 1000-column row: leaves of 120/128 columns, contiguous, covering the row */
    for (r = 0; r < 8; ++r) {
        if (oode_shard_layout(1000, OODE_OBJ_ACKLEY, 8, r, &col0, &ncols, &leaf0, &nleaves) != OODE_OK) return 2;
        if (col0 != covered || nleaves != 1 || leaf0 != r || (ncols != 120 && ncols != 128)) return 3;
        covered += ncols;
    }
    if (covered != 1000) return 4;
    if (oode_shard_layout(100, OODE_OBJ_RASTRIGIN, 2, 0, &col0, &ncols, &leaf0, &nleaves) != OODE_EINVAL) return 5;
    if (!strstr(oode_last_error(), "fewer pairwise-sum leaves")) return 6;
    /* CPython's random stream after random.seed(150880): first randint(0, 99) values and random() draws */
    if (oode_cpython_stream(150880u, 100u, 4, ints, 2, dbl) != OODE_OK) return 7;
    if (ints[0] < 0 || ints[0] > 99 || !(dbl[0] >= 0.0 && dbl[0] < 1.0)) return 8;
    memset(&cfg, 0, sizeof cfg);
    cfg.abi_version = OODE_ABI_VERSION;
    cfg.dim = 3; cfg.population = 8; cfg.lb = lb; cfg.ub = ub;
    cfg.diff_factor = 0.8; cfg.cross_rate = 0.5; cfg.seed = 150880u; cfg.objective = OODE_OBJ_SPHERE;
    rc = oode_create(&cfg, &h);
    if (oode_device_count() == 0) {
        if (rc != OODE_ECUDA || h != NULL || !strstr(oode_last_error(), "no CPU fallback")) return 9;
    } else {
        oode_gen_record rec;
        if (rc != OODE_OK || oode_init(h, NULL, &rec) != OODE_OK || !(rec.best_f >= 0.0)) return 10;
        oode_destroy(h);
    }
    cfg.hndvi = OODE_MAX_HNDVI + 1;
    if (oode_create(&cfg, &h) != OODE_EINVAL) return 11;
    printf("c abi ok: %d devices, first draws %lld %.17g\n", oode_device_count(), (long long)ints[0], dbl[0]);
    return 0;
}
