"""Generate tests/golden/galileo/*.npz by running the UNMODIFIED reference solver
openopt/solvers/Standalone/galileo_oo.py (in the build container, where /root/reference is mounted):

    python tests/golden/make_golden_galileo.py

galileo_oo.py is Python-2 code with an unseeded generator; oracle/galileo_ref_shim.py restores the three Python-2
behaviours it relies on at run time and seeds the generator (nothing in the reference is edited).  A recorder wrapped
around ``p.iterfcn`` reads what ``__solver__`` passes in (the best chromosome's genes and -maxFitness, :86-89), the
evaluation counter, and the solver's own local ``P`` (currentGeneration genes and cached fitness) on every iteration.
The same run is then repeated with oracle/galileo_oracle.py and must be bit-identical before a fixture is written,
so the oracle is pinned to the reference here and the fixtures let that be re-checked anywhere
(tests/test_galileo_oracle.py).  All draws come from CPython's ``random.Random(seed)``, deterministic across
CPython >= 3.2, so the fixtures hold outputs only.
"""
import hashlib
import os
import sys
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, 'oracle'))

import de_oracle as orc                    # noqa: E402
import galileo_oracle as gorc              # noqa: E402
from galileo_ref_shim import load_galileo  # noqa: E402

OUT = os.path.join(HERE, 'galileo')
SEED = 150880


def cases():
    one, lin = np.ones, np.linspace
    return {
        # the solver's defaults (population 15, crossoverRate 1.0, mutationRate 0.05)
        'sphere_d5_np15': dict(obj='sphere', lb=-one(5), ub=2 * one(5), N=15, iters=40),
        'rastrigin_d10_np15': dict(obj='rastrigin', lb=-5.12 * one(10), ub=5.12 * one(10), N=15, iters=60),
        'rosenbrock_d10_np100': dict(obj='rosenbrock', lb=-5 * one(10), ub=5 * one(10), N=100, iters=60),
        # crossoverRate < 1: entries stay aliased to current chromosomes and are mutated in place
        'rosenbrock_d7_np20_cr07_mr01': dict(obj='rosenbrock', lb=-2 * one(7), ub=2 * one(7), N=20, iters=60, cr=0.7, mr=0.1),
        'rastrigin_d3_np16_cr05_mr03_max': dict(obj='rastrigin', lb=-5.12 * one(3), ub=5.12 * one(3), N=16, iters=50, cr=0.5,
                                                mr=0.3, goal='max'),
        'sphere_d9_np24_cr0': dict(obj='sphere', lb=-3 * one(9), ub=lin(1, 2, 9), N=24, iters=30, cr=0.0, mr=0.2),
        # edge cases: one gene (every crossover copies a parent: duplicates), odd populations, 1 and 2 chromosomes,
        # no mutation / always mutation, a ragged row just over numpy's 128-term summation block, per-column bounds
        'sphere_d1_np5': dict(obj='sphere', lb=np.array([-2.0]), ub=np.array([3.0]), N=5, iters=30),
        'ackley_d130_np33': dict(obj='ackley', lb=-32.768 * one(130), ub=32.768 * one(130), N=33, iters=25),
        'sphere_d6_np1': dict(obj='sphere', lb=-one(6), ub=one(6), N=1, iters=25, mr=0.3),
        'sphere_d6_np2': dict(obj='sphere', lb=-one(6), ub=one(6), N=2, iters=25, mr=0.3),
        'rastrigin_d12_np31_mr0': dict(obj='rastrigin', lb=-5.12 * one(12), ub=5.12 * one(12), N=31, iters=40, mr=0.0),
        'rastrigin_d12_np30_mr1': dict(obj='rastrigin', lb=-5.12 * one(12), ub=5.12 * one(12), N=30, iters=20, mr=1.0),
        'griewank_d200_np64': dict(obj='griewank', lb=-600 * one(200), ub=lin(100, 900, 200), N=64, iters=20),
        'rastrigin_d100_np300': dict(obj='rastrigin', lb=-5.12 * one(100), ub=5.12 * one(100), N=300, iters=12),
        # fitness values of both signs: the roulette really spreads over the population (max of 10 - x.x around 0)
        'shifted_sphere_d8_np40_max': dict(obj='sphere', lb=-2 * one(8), ub=2 * one(8), N=40, iters=40, goal='max', mr=0.1),
        # the reference's OWN galileo call site, openopt/doc/solverParams.py:8-14, with its problem and solver arguments as
        # written there (stops on GLP's default maxNonSuccess long before the 3 s time limits)
        'doc_solverparams_glp1_np15': dict(obj='glp1', lb=-one(4), ub=one(4), N=15, cr=0.80, mr=0.15, doc=True),
    }


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()[:16]


def run_reference(c):
    oo, _ = load_galileo(SEED)
    f = orc.OBJECTIVES[c['obj']]
    rec = dict(x=[], f=[], nevals=[], pop_sha=[], fit_sha=[])
    if c.get('doc'):
        p = oo.GLP(f, lb=c['lb'], ub=c['ub'], maxIter=1e3, maxCPUTime=3, maxFunEvals=1e5, fEnough=80, iprint=-1)
    else:
        p = oo.GLP(f, lb=c['lb'], ub=c['ub'], maxIter=c['iters'], maxFunEvals=10 ** 15, maxNonSuccess=10 ** 9, iprint=-1,
                   goal=c.get('goal', 'min'))
    orig = p.iterfcn

    def iterfcn(*a, **k):
        loc = sys._getframe(1).f_locals
        if 'P' in loc:                      # called from galileo.__solver__ (:89)
            P = loc['P']
            rec['x'].append(np.array(a[0], dtype=np.float64).copy())
            rec['f'].append(float(np.ravel(a[1])[0]))
            rec['nevals'].append(int(p.nEvals['f']))
            pop = np.array([ch.genes for ch in P.currentGeneration], dtype=np.float64)
            fit = np.array([float(np.ravel(ch.fitness)[0]) for ch in P.currentGeneration])
            rec['pop_sha'].append(sha(pop))
            rec['fit_sha'].append(sha(fit))
            rec['final_pop'], rec['final_fit'] = pop, fit
        return orig(*a, **k)

    p.iterfcn = iterfcn
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        extra = dict(maxTime=3) if c.get('doc') else {}
        r = p.solve('galileo', population=c['N'], crossoverRate=c.get('cr', 1.0), mutationRate=c.get('mr', 0.05), **extra)
    return r, rec


def main():
    os.makedirs(OUT, exist_ok=True)
    only = sys.argv[1:]
    for name, c in cases().items():
        if only and not any(name.startswith(o) for o in only):
            continue
        r, rec = run_reference(c)
        n = len(rec['f'])
        o = gorc.OracleGalileo(orc.OBJECTIVES[c['obj']], c['lb'], c['ub'], population=c['N'], crossover_rate=c.get('cr', 1.0),
                               mutation_rate=c.get('mr', 0.05), seed=SEED, stream='cpython', invert=(c.get('goal') == 'max'))
        pop_sha, fit_sha, nev = [], [], []
        for _ in range(n):
            o.step()
            pop_sha.append(sha(o.pop)); fit_sha.append(sha(o.fit)); nev.append(o.n_evals)
        # ---- pin the oracle to the reference, bit for bit ----
        assert np.array_equal(np.array(o.history['best_f']), np.array(rec['f'])), name
        assert np.array_equal(np.array(o.history['best_x']), np.array(rec['x'])), name
        assert pop_sha == rec['pop_sha'] and fit_sha == rec['fit_sha'], name
        assert np.array_equal(o.pop, rec['final_pop']) and np.array_equal(o.fit, rec['final_fit']), name
        # the reference's counter also holds the driver's own evaluations (x0, then one per iterfcn call): apart from
        # those the two counters move together (the end-to-end count is checked by tests/test_galileo_mirror.py)
        d = np.array(rec['nevals']) - np.array(nev)
        assert np.all(np.diff(d)[1:] == 1), (name, d)
        np.savez_compressed(
            os.path.join(OUT, name + '.npz'),
            obj=c['obj'], lb=c['lb'], ub=c['ub'], N=c['N'], iters=n, cr=c.get('cr', 1.0), mr=c.get('mr', 0.05), doc=bool(c.get('doc')),
            goal=c.get('goal', 'min'), seed=SEED,
            best_f=np.array(rec['f']), best_x=np.array(rec['x']), best_idx=np.array(o.history['best_idx'], dtype=np.int64),
            n_appended=np.array(o.history['n_appended'], dtype=np.int64), n_aliased=np.array(o.history['n_aliased'], dtype=np.int64),
            pop_sha=np.array(rec['pop_sha']), fit_sha=np.array(rec['fit_sha']),
            final_pop=rec['final_pop'], final_fit=rec['final_fit'],
            iter_f=np.array(r.iterValues.f, dtype=np.float64), xf=np.asarray(r.xf, dtype=np.float64), ff=np.float64(r.ff),
            istop=r.istop, n_evals=r.evals['f'])
        print('%-36s iters=%3d appended=%6d aliased=%6d ff=%r istop=%s  oracle==reference' % (
            name, n, int(np.sum(o.history['n_appended'])), int(np.sum(o.history['n_aliased'])), float(r.ff), r.istop))


if __name__ == '__main__':
    main()
