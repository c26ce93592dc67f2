"""GPU parity tests of the galileo solver (include/ooga.h, csrc/ga_kernels.cuh) through the C ABI: against the fixtures
recorded from the unmodified reference (CPython stream), against the oracle's native (Philox) schedule at sizes that
exercise every kernel path (multi-tile sort, aliased entries, odd populations), through the `p.solve('galileo')` API,
and through size-independent properties at a large population.

Bars: genes, list positions, selection picks, cut points and the appended / aliased / duplicate counts bit-exact
(genes are only ever copied or re-drawn); fitness values bit-exact for objectives without transcendental functions and
within 1e-12 relative for the cos / exp based ones."""
import os

import numpy as np
import pytest

import de_oracle as orc
import galileo_oracle as gorc
import helpers
from test_galileo_oracle import GA_CASES, GA_DIR, oracle_for
from openopt_b200 import GLP, _lib, objectives

pytestmark = pytest.mark.gpu

RTOL = 1e-12


def f_close(obj, got, want, what=''):
    got, want = np.asarray(got, dtype=np.float64), np.asarray(want, dtype=np.float64)
    if obj in helpers.EXACT_OBJECTIVES:
        assert np.array_equal(got, want), what
    else:
        assert np.allclose(got, want, rtol=RTOL, atol=0.0), (what, np.max(np.abs(got - want) / np.abs(want)))


def device_for(obj_name, lb, ub, **kw):
    obj = helpers.device_objective(obj_name)
    return _lib.DeviceGalileo(obj.device_id, lb, ub, obj_aux=obj.aux(np.size(lb)), **kw)


@pytest.mark.parametrize('batch', [1, 7])
@pytest.mark.parametrize('name', GA_CASES)
def test_cpython_stream_mode_matches_the_recorded_reference_run(name, batch):
    g = np.load(os.path.join(GA_DIR, name + '.npz'))
    obj, N, iters = str(g['obj']), int(g['N']), int(g['iters'])
    with device_for(obj, g['lb'], g['ub'], population=N, crossover_rate=float(g['cr']), mutation_rate=float(g['mr']),
                    seed=int(g['seed']), rng='cpython', invert=(str(g['goal']) == 'max'), max_batch=batch) as dev:
        dev.init()
        done = 0
        while done < iters:
            k = min(batch, iters - done)
            recs, xs = dev.step(k)
            for j, rec in enumerate(recs):
                it = done + j
                f_close(obj, rec.best_f, g['best_f'][it], (name, it, 'best_f'))
                assert np.array_equal(xs[j], g['best_x'][it]), (name, it, 'best_x')
                assert rec.best_idx == g['best_idx'][it] and rec.generation == it + 1
                assert rec.n_appended == g['n_appended'][it] and rec.n_aliased == g['n_aliased'][it], (name, it)
                assert rec.n_appended + rec.n_aliased + rec.n_duplicates == 2 * ((N + 1) // 2)
            done += k
        pop, fit = dev.population()
        assert np.array_equal(pop, g['final_pop']), name
        f_close(obj, fit, g['final_fit'], (name, 'final fitness'))
        assert dev.gpu_launches() > 10 * iters


NATIVE = [
    # obj, D, N, cr, mr, goal, steps
    ('sphere', 5, 15, 1.0, 0.05, 'min', 12),
    ('rosenbrock', 33, 1025, 0.7, 0.1, 'min', 6),          # odd population just over one sort tile, aliased entries
    ('sphere', 64, 3000, 1.0, 0.05, 'max', 5),             # positive fitness: the roulette spreads over the population
    ('ackley', 130, 4097, 1.0, 0.02, 'min', 4),            # three sort tiles -> 8192 entries, ragged row
    ('rastrigin', 1, 700, 0.9, 0.3, 'min', 8),             # one gene: almost every entry is a duplicate
    ('griewank', 200, 512, 0.5, 0.05, 'max', 5),
]


@pytest.mark.parametrize('obj,D,N,cr,mr,goal,steps', NATIVE)
def test_native_mode_matches_the_oracle(obj, D, N, cr, mr, goal, steps):
    lb = -5.12 * np.ones(D) if obj != 'griewank' else -600.0 * np.ones(D)
    ub = np.linspace(2.0, 5.12, D) if obj != 'griewank' else np.linspace(100.0, 900.0, D)
    o = gorc.OracleGalileo(orc.OBJECTIVES[obj], lb, ub, population=N, crossover_rate=cr, mutation_rate=mr, seed=20261020,
                           stream='philox', invert=(goal == 'max'))
    with device_for(obj, lb, ub, population=N, crossover_rate=cr, mutation_rate=mr, seed=20261020, rng='philox',
                    invert=(goal == 'max'), max_batch=4) as dev:
        dev.init()
        pop, fit = dev.population()
        assert np.array_equal(pop, o.pop), 'initial population'
        f_close(obj, fit, o.fit, 'initial fitness')
        for it in range(steps):
            f, x = o.step()
            recs, xs = dev.step(1)
            picks, cut = dev.selection()
            assert np.array_equal(picks, o.last_picks), (it, 'roulette picks')
            assert np.array_equal(cut, o.last_cut), (it, 'cut points')
            rec = recs[0]
            f_close(obj, rec.best_f, f, (it, 'best_f'))
            f_close(obj, rec.sum_fitness, o.history['sum_fitness'][-1], (it, 'sum_fitness'))
            assert np.array_equal(xs[0], x) and rec.best_idx == o.history['best_idx'][-1], it
            assert rec.n_appended == o.history['n_appended'][-1] and rec.n_aliased == o.history['n_aliased'][-1], it
            pop, fit = dev.population()
            assert np.array_equal(pop, o.pop), (it, 'population')
            f_close(obj, fit, o.fit, (it, 'fitness'))


@pytest.mark.parametrize('obj,D,N,cr', [('sphere', 3, 5000, 1.0), ('rosenbrock', 20, 2500, 0.6), ('sphere', 1, 300, 1.0)])
def test_merge_by_rank_equals_the_full_sort(obj, D, N, cr, monkeypatch):
    """From the second iteration on only the new entries are sorted and merged by rank with the (already ordered) current
    generation; OOGA_SORT=full sorts the whole list every time.  Same populations, same order."""
    lb, ub = -2.0 * np.ones(D), 3.0 * np.ones(D)
    runs = []
    for mode in ('full', 'merge'):
        monkeypatch.setenv('OOGA_SORT', mode)
        with device_for(obj, lb, ub, population=N, crossover_rate=cr, mutation_rate=0.1, seed=9, rng='philox', max_batch=4) as dev:
            dev.init()
            recs = []
            for _ in range(3):
                r, _x = dev.step(4)
                recs += [(q.best_f, q.best_idx, q.n_appended, q.n_aliased, q.n_duplicates) for q in r]
            runs.append((recs, dev.population()))
    assert runs[0][0] == runs[1][0]
    assert np.array_equal(runs[0][1][0], runs[1][1][0]) and np.array_equal(runs[0][1][1], runs[1][1][1])


def test_batched_steps_equal_single_steps():
    lb, ub = -2.0 * np.ones(40), 3.0 * np.ones(40)
    runs = []
    for batch in (1, 5):
        with device_for('rosenbrock', lb, ub, population=333, crossover_rate=0.8, mutation_rate=0.05, seed=5, rng='philox',
                        max_batch=batch) as dev:
            dev.init()
            fs = []
            for _ in range(10 // batch):
                recs, xs = dev.step(batch)
                fs += [(r.best_f, r.n_appended, r.n_aliased, r.n_duplicates) for r in recs]
            runs.append((fs, dev.population()))
    assert runs[0][0] == runs[1][0]
    assert np.array_equal(runs[0][1][0], runs[1][1][0]) and np.array_equal(runs[0][1][1], runs[1][1][1])


@pytest.mark.parametrize('name', ['rastrigin_d10_np15', 'rosenbrock_d7_np20_cr07_mr01', 'shifted_sphere_d8_np40_max',
                                  'sphere_d1_np5'])
def test_solve_api_reproduces_the_reference_run(name):
    """GLP(...).solve('galileo', rng='cpython') end to end against the reference's recorded result."""
    g = np.load(os.path.join(GA_DIR, name + '.npz'))
    obj = str(g['obj'])
    p = GLP(helpers.device_objective(obj), lb=g['lb'], ub=g['ub'], maxIter=int(g['iters']) + 1, maxFunEvals=10 ** 15,
            maxNonSuccess=10 ** 9, iprint=-1, goal=str(g['goal']))
    r = p.solve('galileo', population=int(g['N']), crossoverRate=float(g['cr']), mutationRate=float(g['mr']),
                seed=int(g['seed']), rng='cpython', gensPerCall=16)
    assert r.istop == int(g['istop']) and r.evals['f'] == int(g['n_evals'])
    f_close(obj, r.iterValues.f, g['iter_f'], name)
    assert np.array_equal(r.xf, g['xf'])
    f_close(obj, r.ff, g['ff'], name)
    assert p.solver.gpu_launches > 0


@pytest.mark.parametrize('rng,cr,N', [('philox', 1.0, 333), ('cpython', 0.7, 40), ('philox', 0.5, 2500)])
def test_checkpoint_resume_continues_bit_for_bit(rng, cr, N):
    lb, ub = -2.0 * np.ones(12), 3.0 * np.ones(12)
    kw = dict(population=N, crossover_rate=cr, mutation_rate=0.1, seed=31, rng=rng, max_batch=4)

    def tail(dev, n):
        out = []
        for _ in range(n):
            recs, xs = dev.step(1)
            out.append((recs[0].best_f, recs[0].best_idx, recs[0].n_appended, recs[0].n_aliased, xs[0].tobytes()))
        return out

    for first in (0, 5):                      # a checkpoint of the unsorted initial list, and of a sorted one
        with device_for('rosenbrock', lb, ub, **kw) as a:
            a.init()
            if first:
                a.step(4), a.step(1)
            blob = a.save_state()
            want = tail(a, 6)
            want_pop = a.population()
        with device_for('rosenbrock', lb, ub, **kw) as b:
            b.init()
            b.step(2)                          # the resumed handle had moved on: the blob overrides all of it
            b.load_state(blob)
            assert b.generation == first
            got = tail(b, 6)
            got_pop = b.population()
        assert got == want, (rng, first)
        assert np.array_equal(got_pop[0], want_pop[0]) and np.array_equal(got_pop[1], want_pop[1])
    with device_for('rosenbrock', lb, ub, **dict(kw, population=N + 1)) as c:
        c.init()
        with pytest.raises(_lib.LibraryError):
            c.load_state(blob)                 # another configuration


def test_set_population_evaluates_and_takes_the_list_order_as_given():
    D, N = 6, 50
    lb, ub = -np.ones(D), np.ones(D)
    rng = np.random.default_rng(3)
    genes = rng.uniform(-1, 1, (N, D))
    o = gorc.OracleGalileo(orc.OBJECTIVES['sphere'], lb, ub, population=N, seed=12, stream='philox')
    o.pop, o.fit = genes.copy(), -orc.OBJECTIVES['sphere'](genes)
    with device_for('sphere', lb, ub, population=N, seed=12, rng='philox', max_batch=2) as dev:
        dev.init()
        dev.step(2)
        dev.set_population(genes)
        pop, fit = dev.population()
        assert np.array_equal(pop, genes) and np.array_equal(fit, o.fit)
        o.gen = dev.generation                  # the draws continue with the handle's iteration counter
        for _ in range(4):
            f, x = o.step()
            recs, xs = dev.step(1)
            assert recs[0].best_f == f and np.array_equal(xs[0], x)
        pop, fit = dev.population()
        assert np.array_equal(pop, o.pop) and np.array_equal(fit, o.fit)


def test_the_references_own_call_site():
    """openopt/doc/solverParams.py:8-14 as written there -- its 4-gene objective (an expression objective here), its
    problem arguments and its solver arguments -- against the reference's recorded run of exactly that call."""
    g = np.load(os.path.join(GA_DIR, 'doc_solverparams_glp1_np15.npz'))
    f = helpers.device_objective('glp1')
    lb, ub = -np.ones(4), np.ones(4)
    p = GLP(f, lb=lb, ub=ub, maxIter=1e3, maxCPUTime=3, maxFunEvals=1e5, fEnough=80, iprint=-1)
    r = p.solve('galileo', crossoverRate=0.80, maxTime=3, population=15, mutationRate=0.15, rng='cpython', seed=int(g['seed']))
    assert r.istop == int(g['istop']) == 11 and r.evals['f'] == int(g['n_evals'])
    f_close('glp1', r.iterValues.f, g['iter_f'], 'iterValues.f')
    assert np.array_equal(r.xf, g['xf'])
    f_close('glp1', r.ff, g['ff'], 'ff')


def test_error_paths():
    lb, ub = -np.ones(3), np.ones(3)
    with pytest.raises(_lib.LibraryError):
        device_for('sphere', lb, ub, population=0)
    with pytest.raises(_lib.LibraryError):
        device_for('sphere', lb, np.array([1.0, np.inf, 1.0]), population=5)
    with device_for('sphere', lb, ub, population=5, max_batch=2) as dev:
        with pytest.raises(_lib.LibraryError):
            dev.step(1)                      # before init
        dev.init()
        with pytest.raises(_lib.LibraryError):
            dev.step(3)                      # more than max_batch


SCAN_N = 1_000_000


def scan_inputs():
    rng = np.random.default_rng(0)
    n = SCAN_N
    return {
        'positive': rng.random(n) * 100,
        'negative': -(rng.random(n) * 4000 + 100),                       # a minimised positive objective
        'mixed_signs': rng.standard_normal(n),                           # the sum wanders through zero: cancellations, many binades
        'integers': rng.integers(-5, 50, n).astype(np.float64),          # exact sums
        'quarters': rng.integers(0, 50, n) * 0.5 + 0.25,                 # f/ulp falls on .5: round-half-even decides
        'tiny': rng.random(n) * 1e-300, 'huge': rng.random(n) * 1e300,
        'growing': np.exp(rng.random(n) * 40),                           # magnitudes over 17 decades
        'cancel': np.concatenate([rng.random(n // 3) * 1e6, -rng.random(n // 3) * 1e6, rng.random(n - 2 * (n // 3))]),
        'equal_maxima': np.where(rng.random(n) < 1e-4, 7.0, rng.random(n)),   # the FIRST largest value is the best
        'with_inf': np.where(np.arange(n) == n // 2, np.inf, rng.random(n)),
    }


@pytest.fixture(scope='module')
def scan_device():
    dev = device_for('sphere', -np.ones(2), np.ones(2), population=SCAN_N, seed=1, rng='philox', max_batch=1)
    dev.init()
    yield dev
    dev.close()


@pytest.mark.parametrize('kind', sorted(scan_inputs()))
def test_running_sums_are_the_sequentially_rounded_sums(scan_device, kind):
    """evaluate (:268) and the roulette (:340-343) accumulate the fitness left to right; the large-population scan
    re-expresses that serial chain as integer prefix sums per binade -- every one of the 10^6 running sums must be the
    sequentially rounded one (numpy's cumsum is the sequential loop)."""
    f = scan_inputs()[kind]
    with np.errstate(all='ignore'):
        want = np.cumsum(f)
    got, total, best = scan_device.scan_values(f)
    assert np.array_equal(got, want, equal_nan=True), (kind, int(np.argmax(got != want)))
    assert total == want[-1] or (np.isnan(total) and np.isnan(want[-1]))
    assert best == int(np.argmax(f)), kind


@pytest.mark.parametrize('path', ['small', 'large'])
def test_both_scan_paths_agree_with_the_sequential_sum(path, monkeypatch):
    monkeypatch.setenv('OOGA_SCAN', path)
    rng = np.random.default_rng(5)
    for n in (1, 2, 1023, 1024, 1025, 5000):
        with device_for('sphere', -np.ones(2), np.ones(2), population=n, seed=1, rng='philox', max_batch=1) as dev:
            dev.init()
            for f in (rng.random(n) * 10, -rng.random(n), rng.standard_normal(n), np.full(n, 0.1)):
                got, total, best = dev.scan_values(f)
                assert np.array_equal(got, np.cumsum(f)) and total == np.cumsum(f)[-1] and best == int(np.argmax(f)), (path, n)
            recs, _ = dev.step(1)                                    # and the roulette on top of it still runs
            assert recs[0].generation == 1


def test_roulette_picks_at_a_large_population():
    """All 200 000 picks of one iteration against numpy: first index whose running sum reaches the wheel position."""
    N, D = 200_000, 16
    lb, ub = -2.0 * np.ones(D), 2.0 * np.ones(D)
    stream = gorc.GAPhiloxStream(77)
    with device_for('sphere', lb, ub, population=N, seed=77, rng='philox', invert=True, max_batch=1) as dev:
        dev.init()
        for it in range(1, 4):
            _, fit = dev.population()
            prefix = np.cumsum(fit)
            wheel_u, cut = stream.selection_draws(it, N, D, 1.0)
            wheel = 0.0 + (prefix[-1] - 0.0) * wheel_u
            want = np.minimum(np.searchsorted(np.maximum.accumulate(prefix), wheel, side='left'), N - 1)
            recs, _ = dev.step(1)
            picks, got_cut = dev.selection()
            assert recs[0].sum_fitness == prefix[-1]
            assert np.array_equal(picks, want), it
            assert np.array_equal(got_cut, cut), it
            assert len(np.unique(picks)) > N // 4                     # positive fitness: the wheel spreads


def test_whole_population_against_the_oracle_at_200k():
    """The shape of bench.py's galileo workload (100 genes, crossoverRate 1, mutationRate 0.05) at 200 000 chromosomes,
    EVERY pick, cut point, gene and cached fitness of two iterations against the oracle's vectorised step (an objective
    without transcendental functions, so that fitness -- and with it every roulette and sort decision -- is bit-exact):
    sort over 2^19 entries, then the merge-by-rank path, ~200 chunks of the running-sum scan."""
    N, D = 200_000, 100
    lb, ub = -5.12 * np.ones(D), 5.12 * np.ones(D)
    o = gorc.OracleGalileo(orc.OBJECTIVES['sphere'], lb, ub, population=N, seed=2026, stream='philox', invert=True)
    with device_for('sphere', lb, ub, population=N, seed=2026, rng='philox', invert=True, max_batch=1) as dev:
        dev.init()
        pop, fit = dev.population()
        assert np.array_equal(pop, o.pop) and np.array_equal(fit, o.fit), 'initial population'
        for it in range(2):
            f, x = o.step_fast()
            recs, xs = dev.step(1)
            picks, cut = dev.selection()
            assert np.array_equal(picks, o.last_picks) and np.array_equal(cut, o.last_cut), it
            rec = recs[0]
            assert rec.best_f == f and np.array_equal(xs[0], x) and rec.best_idx == o.history['best_idx'][-1]
            assert rec.sum_fitness == o.history['sum_fitness'][-1]
            assert rec.n_appended == o.history['n_appended'][-1] and rec.n_aliased == 0
            pop, fit = dev.population()
            assert np.array_equal(fit, o.fit), (it, 'cached fitness')
            assert np.array_equal(pop, o.pop), (it, 'population')


@pytest.mark.parametrize('obj,goal', [('rastrigin', 'min'), ('sphere', 'max')])
def test_large_population_properties(obj, goal):
    """N = 200 000 chromosomes of 100 genes (sort over 2^19 entries): what must hold at any size."""
    N, D = 200_000, 100
    lb, ub = -5.12 * np.ones(D), 5.12 * np.ones(D)
    sign = -1.0 if goal == 'max' else 1.0
    with device_for(obj, lb, ub, population=N, seed=11, rng='philox', invert=(goal == 'max'), max_batch=3) as dev:
        dev.init()
        pop0, fit0 = dev.population()
        f_close(obj, fit0, -sign * orc.OBJECTIVES[obj](pop0), 'initial fitness')
        best = np.inf
        total_appended = 0
        for _ in range(2):
            recs, xs = dev.step(3)
            for rec, x in zip(recs, xs):
                assert rec.best_f <= best                        # crossoverRate 1: the best chromosome never gets worse
                best = rec.best_f
                f_close(obj, rec.best_f, sign * orc.OBJECTIVES[obj](x), 'best_f is f(best_x)')
                assert rec.n_appended + rec.n_duplicates == N and rec.n_aliased == 0
                total_appended += rec.n_appended
        pop, fit = dev.population()
        assert np.all(np.diff(fit) <= 0), 'list order is descending cached fitness'
        assert np.all(pop >= lb) and np.all(pop <= ub)
        f_close(obj, fit, -sign * orc.OBJECTIVES[obj](pop), 'cached fitness is -f(genes)')
        assert len(np.unique(pop, axis=0)) == N, 'replacement without duplicates'
        assert fit[0] >= fit0.max() and fit[-1] >= np.sort(fit0)[0]
        if goal == 'max':
            assert total_appended > N                            # the roulette spreads: most entries are new chromosomes
