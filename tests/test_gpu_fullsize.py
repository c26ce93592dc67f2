"""Oracle checks AT THE BENCHMARKED SIZES (BASELINE configs[2]: Ackley 1000-D x 1M rows; configs[4]: Rastrigin
100-D x 16M rows), through the C ABI.

What is specific to full size -- byte offsets beyond 2^32 in the slot arrays, the 64-bit branch of the item
split, 148-CTA grids with every warp busy, the widening row addressing -- is not exercised by the small parity
cases, so here the run itself is compared with the oracle:

  * sampled rows (numpy oracle, de_oracle.py): for 4096 random rows plus rows next to every 2^32-byte boundary
    of the slot arrays, the parent and the three donor rows are fetched BEFORE a generation (oode_get_rows,
    oode_get_donor_indices), the trial is rebuilt on the host from them with the oracle's Philox draws
    (make_trial: mutation, crossover, bound repair) and compared bit for bit with the trial row the kernel
    wrote; its objective value within 1e-12 relative.  Donor indices of ALL rows are compared with the
    oracle's index draws.  A wrong-but-in-range donor row fails this test.
  * whole population (C port, c_oracle.py; needs ~40 GB of host memory, skipped otherwise): accept masks,
    trial-best index, accepted / repaired counts bit-exact, trial and population objective values within 1e-12,
    the final population bit-exact, after the initial evaluation and two generations.
"""
import numpy as np
import pytest

import c_oracle
import de_oracle as orc
from openopt_b200 import _lib, objectives

pytestmark = pytest.mark.gpu

RTOL = 1e-12
SEED = 150880
CASES = [('ackley', 1000, 1000000, 32.768), ('rastrigin', 100, 1 << 24, 5.12)]


def _device_or_skip(NP, D):
    import torch
    need = 2 * NP * D * 8 + 64 * NP
    if torch.cuda.mem_get_info()[0] < need + (2 << 30):
        pytest.skip('not enough free device memory')


def _sample_rows(NP, ld, n=4096, seed=7):
    """Random rows + the rows on both sides of every 2^32-byte boundary of a slot array + the first / last rows."""
    rng = np.random.RandomState(seed)
    rows = set(rng.randint(0, NP, size=n).tolist()) | {0, 1, NP - 2, NP - 1}
    k = 1
    while (k << 32) // (ld * 8) < NP:
        r = (k << 32) // (ld * 8)
        rows |= {r - 1, r, min(r + 1, NP - 1)}
        k += 1
    return np.array(sorted(rows), dtype=np.int64)


def _index_draws(gen, NP):
    """The oracle's donor-index draws of one generation for ALL rows (PhiloxStream.generation_draws without the
    per-gene uniforms, which would be NP x D doubles)."""
    rows = np.arange(NP, dtype=np.uint64)
    hi = (rows >> np.uint64(32)) | np.uint64(orc.TAG_INDEX << 24)
    k0, k1 = SEED & 0xFFFFFFFF, SEED >> 32
    a0, a1, a2, a3 = orc.philox4x32_10(np.uint64(0), rows & orc._M32, hi, np.uint64(gen), k0, k1)
    b0, b1, _, _ = orc.philox4x32_10(np.uint64(1), rows & orc._M32, hi, np.uint64(gen), k0, k1)
    return orc.mulhi64(a0, a1, NP), orc.mulhi64(a2, a3, NP), orc.mulhi64(b0, b1, NP)


@pytest.mark.parametrize('obj,D,NP,hw', CASES)
def test_full_size_sampled_trials_match_numpy_oracle(obj, D, NP, hw):
    _device_or_skip(NP, D)
    lb, ub = -hw * np.ones(D), hw * np.ones(D)
    x0 = lb + 0.25 * (ub - lb)
    dobj = objectives.get(obj)
    f = orc.OBJECTIVES[obj]
    stream = orc.PhiloxStream(SEED)
    F, Cr = 0.8, 0.5
    with _lib.DeviceDE(dobj.device_id, lb, ub, x0=x0, population=NP, diff_factor=F, cross_rate=Cr, seed=SEED,
                       rng='philox', max_batch=2) as dev:
        rec = dev.init()
        S = _sample_rows(NP, D + (D & 1))
        n = S.size
        # initial population (de_oo.py:147-150) and its evaluation, sampled rows
        pop0 = stream.init_uniforms(NP, D, rows=S) * (ub - lb) + lb
        pop0[S == 0] = x0
        assert np.array_equal(dev.rows(S, dev.ROWS_CURRENT), pop0), 'initial population rows'
        vals = dev.population_vals()
        assert np.allclose(vals[S], f(pop0), rtol=RTOL, atol=0), 'initial objective values'
        assert rec.trial_best_idx == int(np.argmin(vals)) and rec.best_f == vals.min()
        for gen in (1, 2):
            ib, r1, r2 = dev.donor_indices()
            oib, or1, or2 = _index_draws(gen, NP)
            assert np.array_equal(ib, oib) and np.array_equal(r1, or1) and np.array_equal(r2, or2), 'donor index draws'
            # the four source rows of every sampled trial, as the generation will read them
            src = np.concatenate([dev.rows(S, dev.ROWS_CURRENT), dev.rows(ib[S], dev.ROWS_CURRENT),
                                  dev.rows(r1[S], dev.ROWS_CURRENT), dev.rows(r2[S], dev.ROWS_CURRENT)])
            old_vals = vals
            rec = dev.step(1)[0]
            _, _, _, Uf, Uc = stream.generation_draws(gen, NP, D, rows=S)
            k = np.arange(n)
            # make_trial gathers old_pop[ib], old_pop[r1], old_pop[r2]: here the gathered rows are stacked behind the parents
            trial, _ = orc.make_trial(src, np.concatenate([n + k, k, k, k]), np.concatenate([2 * n + k, k, k, k]),
                                      np.concatenate([3 * n + k, k, k, k]), np.vstack([Uf] * 4), np.vstack([Uc] * 4), F, Cr, lb, ub)
            trial = trial[:n]
            assert np.array_equal(dev.rows(S, dev.ROWS_PARENT), src[:n]), 'parent rows, generation %d' % gen
            got = dev.rows(S, dev.ROWS_TRIAL)
            bad = np.nonzero((got != trial).any(axis=1))[0]
            assert bad.size == 0, 'trial rows differ from the oracle, generation %d: rows %s' % (gen, S[bad][:8])
            tv = dev.trial_vals()
            assert np.allclose(tv[S], f(trial), rtol=RTOL, atol=0), 'trial objective values, generation %d' % gen
            # selection, argmin and counters over ALL rows (de_oo.py:260-282, 304)
            acc = dev.accept_mask().astype(bool)
            vals = dev.population_vals()
            assert np.array_equal(acc, ~(old_vals < tv)), 'accept mask'
            assert np.array_equal(vals, np.where(acc, tv, old_vals)), 'post-selection vals'
            assert rec.n_accepted == int(acc.sum()) and rec.trial_best_idx == int(np.argmin(tv)) and rec.trial_best_f == tv.min()
            assert rec.best_f == vals.min()
            cur = dev.rows(S, dev.ROWS_CURRENT)
            assert np.array_equal(cur, np.where(acc[S][:, None], trial, src[:n])), 'selected rows'
        assert np.array_equal(dev.best_x(), dev.rows([int(np.argmin(vals))], dev.ROWS_CURRENT)[0])
    _lib.release_cached_memory()


@pytest.mark.parametrize('obj,D,NP,hw', CASES)
def test_full_size_run_matches_c_port(obj, D, NP, hw):
    import psutil
    _device_or_skip(NP, D)
    need_host = 3 * NP * D * 8 + (4 << 30)            # the port's two arrays + the downloaded population
    if psutil.virtual_memory().available < need_host:
        pytest.skip('not enough host memory for the C port at this size (%.0f GB)' % (need_host / 2 ** 30))
    lb, ub = -hw * np.ones(D), hw * np.ones(D)
    x0 = lb + 0.25 * (ub - lb)
    dobj = objectives.get(obj)
    c_oracle.set_threads()
    c = c_oracle.COracleDE(obj, lb, ub, x0=x0, population=NP, F=0.8, Cr=0.5, seed=SEED, stream='philox')
    view = lambda name, shape: np.ctypeslib.as_array(getattr(c.lib, name)(c.h), shape=shape)      # no copy
    with _lib.DeviceDE(dobj.device_id, lb, ub, x0=x0, population=NP, diff_factor=0.8, cross_rate=0.5, seed=SEED,
                       rng='philox', max_batch=2) as dev:
        rec = dev.init()
        assert np.allclose(dev.population_vals(), view('orc_vals', (NP,)), rtol=RTOL, atol=0)
        assert rec.trial_best_idx == c.trial_best_idx
        for gen in (1, 2):
            c.step(1)
            rec = dev.step(1)[0]
            assert np.array_equal(dev.accept_mask(), view('orc_accept', (NP,))), 'accept mask, generation %d' % gen
            assert (rec.trial_best_idx, rec.n_accepted, rec.n_repaired) == (c.trial_best_idx, c.n_accepted, c.n_repaired)
            assert np.allclose(dev.trial_vals(), view('orc_tvals', (NP,)), rtol=RTOL, atol=0)
            assert np.isclose(rec.best_f, c.best_f, rtol=RTOL, atol=0)
        pop, vals = dev.population()
        assert np.array_equal(pop, view('orc_pop', (NP, D))), 'final population'
        assert np.allclose(vals, view('orc_vals', (NP,)), rtol=RTOL, atol=0)
        assert np.array_equal(dev.best_x(), c.best_x)
    c.close()
    _lib.release_cached_memory()
