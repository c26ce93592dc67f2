"""Time the UNMODIFIED reference `GLP(f, ...).solve('de')` (baseline/_ref, loaded through oracle/ref_shim.py) on
ONE core of this box -- the reference is single-threaded by construction (BASELINE.md
section 3).  Run by bench.py's cpu_baseline leg as a subprocess; prints one JSON object.

    python baseline/time_reference.py OBJ D NP GENS LO HI [--numpy-rng]

--numpy-rng: the "friendly" second line of BASELINE.md section 3 -- de_oo.Rand / Randint are monkeypatched to
numpy.random.RandomState(seed) at run time (removes the per-element Python RNG loop; nothing on disk changes)."""
import json
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, os.path.join(ROOT, 'oracle'))
sys.path.insert(0, ROOT)


def main():
    obj, D, NP, gens, lo, hi = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]), float(sys.argv[5]), sys.argv[6]
    friendly = '--numpy-rng' in sys.argv
    if hasattr(os, 'sched_setaffinity'):
        os.sched_setaffinity(0, {sorted(os.sched_getaffinity(0))[0]})          # one core
    from ref_shim import load_reference
    ref_root = os.path.join(HERE, '_ref')
    if not os.path.isdir(os.path.join(ref_root, 'openopt')):
        print(json.dumps({'unavailable': 'baseline/_ref/openopt is missing (python baseline/install_ref.py)'}))
        return
    openopt, de_oo = load_reference(ref_root)
    import de_oracle as orc                                                   # numpy objective definitions (parity form)
    f = {'rosenbrock': orc.rosenbrock, 'rastrigin': orc.rastrigin, 'ackley': orc.ackley, 'griewank': orc.griewank,
         'sphere': orc.sphere}[obj]
    lb = lo * np.ones(D)
    ub = np.linspace(100.0, 900.0, D) if hi == 'lin' else float(hi) * np.ones(D)
    x0 = lb + 0.25 * (ub - lb)
    if friendly:
        rs = np.random.RandomState(150880)
        de_oo.Rand = lambda *shape: rs.rand(*shape)
        de_oo.Randint = lambda low, high=None, size=None: rs.randint(low, high, size)
        de_oo.Seed = lambda s: rs.seed(s)
    p = openopt.GLP(f, x0=x0, lb=lb, ub=ub, maxIter=gens + 1, maxFunEvals=10 ** 15, maxNonSuccess=10 ** 9, iprint=-1)
    t0 = time.perf_counter()
    r = p.solve('de', population=NP, differenceFactor=0.8, crossoverRate=0.5, seed=150880)
    dt = time.perf_counter() - t0
    evals = int(r.evals['f'])
    print(json.dumps({'evals_per_s': evals / dt, 'evals': evals, 'wall_s': dt, 'generations': gens, 'population': NP, 'D': D,
                      'objective': obj, 'cores': 1, 'nproc': os.cpu_count(), 'rng': 'numpy RandomState (patched at run time)'
                      if friendly else "stdlib random (the reference's live path)", 'ff': float(r.ff), 'istop': int(r.istop)}))


if __name__ == '__main__':
    main()
