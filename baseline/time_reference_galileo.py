"""Time the UNMODIFIED reference `GLP(f, ...).solve('galileo')` (baseline/_ref, loaded through
oracle/galileo_ref_shim.py, which restores the Python-2 behaviours the file relies on) on ONE core of this box.
Run by bench.py's cpu_baseline leg (--workload ga_*) as a subprocess; prints one JSON object.

    python baseline/time_reference_galileo.py OBJ D N ITERS LO HI GOAL

The metric is entries per second: 2*ceil(N/2) entries are selected, crossed, mutated, checked for duplicates and
ranked per iteration.  The reference's duplicate scan is O(N^2 * D) Python work per iteration
(galileo_oo.py:465-471), so only small populations finish; larger ones are not extrapolated."""
import json
import os
import sys
import time
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, os.path.join(ROOT, 'oracle'))
sys.path.insert(0, ROOT)


def main():
    obj, D, N, iters, lo, hi, goal = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]), float(sys.argv[5]), float(sys.argv[6]), sys.argv[7]
    if hasattr(os, 'sched_setaffinity'):
        os.sched_setaffinity(0, {sorted(os.sched_getaffinity(0))[0]})          # one core
    ref_root = os.path.join(HERE, '_ref')
    if not os.path.isdir(os.path.join(ref_root, 'openopt')):
        print(json.dumps({'unavailable': 'baseline/_ref/openopt is missing (python baseline/install_ref.py)'}))
        return
    from galileo_ref_shim import load_galileo
    openopt, _ = load_galileo(150880, ref_root)
    import de_oracle as orc
    f = orc.OBJECTIVES[obj]
    p = openopt.GLP(f, lb=lo * np.ones(D), ub=hi * np.ones(D), maxIter=iters, maxFunEvals=10 ** 15, maxNonSuccess=10 ** 9,
                    iprint=-1, goal=goal)
    t0 = time.perf_counter()
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        r = p.solve('galileo', population=N)
    dt = time.perf_counter() - t0
    done = len(r.iterValues.f) - 1
    print(json.dumps({'entries_per_s': 2 * ((N + 1) // 2) * done / dt, 'iterations': done, 'wall_s': dt, 'population': N, 'D': D,
                      'objective': obj, 'goal': goal, 'cores': 1, 'nproc': os.cpu_count(), 'ff': float(r.ff), 'istop': int(r.istop)}))


if __name__ == '__main__':
    main()
