"""The galileo oracle against the LIVE unmodified reference (openopt/solvers/Standalone/galileo_oo.py behind
oracle/galileo_ref_shim.py) on randomly drawn small problems: population, row length, rates, seed, objective, goal and
bounds are all random, so shapes the fixed fixtures do not hold (odd / tiny populations with aliased entries, rates at 0
and 1, rows around the 8- and 128-term boundaries of numpy's summation) are crossed as well.

CPU only and only where /root/reference is mounted (the build container)."""
import os
import sys
import warnings

import numpy as np
import pytest

import de_oracle as orc
import galileo_oracle as gorc

REF = '/root/reference'
pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, 'openopt')), reason='reference tree not mounted')


def run_reference(f, lb, ub, N, iters, seed, cr, mr, goal):
    from galileo_ref_shim import load_galileo
    oo, _ = load_galileo(seed)
    rec = dict(x=[], f=[])
    p = oo.GLP(f, lb=lb, ub=ub, maxIter=iters, maxFunEvals=10 ** 15, maxNonSuccess=10 ** 9, iprint=-1, goal=goal)
    orig = p.iterfcn

    def iterfcn(*a, **k):
        loc = sys._getframe(1).f_locals
        if 'P' in loc:
            P = loc['P']
            rec['x'].append(np.array(a[0], dtype=np.float64).copy())
            rec['f'].append(float(np.ravel(a[1])[0]))
            rec['pop'] = np.array([c.genes for c in P.currentGeneration], dtype=np.float64)
            rec['fit'] = np.array([float(np.ravel(c.fitness)[0]) for c in P.currentGeneration])
        return orig(*a, **k)

    p.iterfcn = iterfcn
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        p.solve('galileo', population=N, crossoverRate=cr, mutationRate=mr)
    return rec


@pytest.mark.parametrize('case', range(24))
def test_oracle_equals_the_live_reference_on_random_problems(case):
    rng = np.random.default_rng(1000 + case)
    obj = ['sphere', 'rosenbrock', 'rastrigin', 'ackley', 'griewank'][case % 5]
    D = int(rng.choice([1, 2, 3, 7, 8, 9, 17, 33, 127, 129, 130])) if obj != 'rosenbrock' else int(rng.choice([2, 3, 8, 9, 40]))
    N = int(rng.choice([1, 2, 3, 4, 7, 15, 16, 31, 50]))
    cr = float(rng.choice([1.0, 1.0, 0.9, 0.5, 0.0]))
    mr = float(rng.choice([0.05, 0.05, 0.3, 0.0, 1.0]))
    goal = 'max' if rng.random() < 0.3 else 'min'
    seed = int(rng.integers(1, 2 ** 31))
    lb = -rng.random(D) * 5 - 0.5
    ub = rng.random(D) * 5 + 0.5
    iters = 25
    rec = run_reference(orc.OBJECTIVES[obj], lb, ub, N, iters, seed, cr, mr, goal)
    o = gorc.OracleGalileo(orc.OBJECTIVES[obj], lb, ub, population=N, crossover_rate=cr, mutation_rate=mr, seed=seed,
                           stream='cpython', invert=(goal == 'max'))
    o.run(len(rec['f']))
    what = (obj, D, N, cr, mr, goal, seed)
    assert np.array_equal(np.array(o.history['best_f']), np.array(rec['f'])), what
    assert np.array_equal(np.array(o.history['best_x']), np.array(rec['x'])), what
    assert np.array_equal(o.pop, rec['pop']) and np.array_equal(o.fit, rec['fit']), what
