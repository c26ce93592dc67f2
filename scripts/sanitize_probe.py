"""Small runs of every trial-kernel form for compute-sanitizer (memcheck / racecheck / synccheck):

    compute-sanitizer --tool racecheck python scripts/sanitize_probe.py

Shapes: wide rows (de_trial_tma_kernel), 120- and 100-column rows (de_trial_tma2_kernel), per-column bounds and a
neighbour-coupled objective (de_trial_kernel), each for a few generations, plus de_select_kernel's grid stage."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from openopt_b200 import _lib, objectives

cases = [('ackley', 1000, 600, False), ('ackley', 120, 700, False), ('rastrigin', 100, 900, False),
         ('griewank', 200, 500, True), ('rosenbrock', 37, 300, False), ('rastrigin', 100, 40000, False)]
if len(sys.argv) > 1:
    cases = [cases[int(a)] for a in sys.argv[1:]]
for obj, D, NP, asym in cases:
    o = objectives.get(obj)
    lb = -5.0 * np.ones(D)
    ub = np.linspace(2.0, 9.0, D) if asym else 5.0 * np.ones(D)
    with _lib.DeviceDE(o.device_id, lb, ub, x0=lb + 0.25 * (ub - lb), population=NP, obj_aux=o.aux(D), max_batch=8) as dev:
        dev.init()
        recs = dev.step(3)
        print('%-10s D=%4d NP=%6d %-22s best_f=%r accepted=%d' % (obj, D, NP, dev.trial_kernel_name(), recs[-1].best_f,
                                                                 recs[-1].n_accepted), flush=True)
