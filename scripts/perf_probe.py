"""Quick device-side timing probe (not the benchmark): ms/generation and GB/s for a few shapes.

    python scripts/perf_probe.py                      # default shape list
    python scripts/perf_probe.py ackley 128 1000000 20 [half_width]
"""
import sys
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from openopt_b200 import _lib, objectives

PEAK = 6553.0
shapes = [('rastrigin', 100, 10000, 200), ('rastrigin', 100, 1 << 20, 20), ('ackley', 1000, 131072, 10),
          ('ackley', 1000, 1000000, 10), ('griewank', 200, 262144, 20), ('sphere', 1000, 1000000, 10)]
hw = 5.12
if len(sys.argv) > 1:
    shapes = [(sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4]))]
    if len(sys.argv) > 5:
        hw = float(sys.argv[5])
for obj, D, NP, gens in shapes:
    o = objectives.get(obj)
    lb = -hw * np.ones(D)
    ub = hw * np.ones(D)
    with _lib.DeviceDE(o.device_id, lb, ub, x0=lb + 0.25 * (ub - lb), population=NP, obj_aux=o.aux(D), max_batch=64) as dev:
        dev.init()
        dev.step(3)
        c0 = dev.counters()
        recs = []
        left = gens
        while left > 0:
            recs += dev.step(min(left, 64))
            left -= min(left, 64)
        c1 = dev.counters()
        ms = (c1['kernel_ms'] - c0['kernel_ms']) / gens
        gb = NP * (40 * D + 16) / 1e9
        acc = np.mean([r.n_accepted for r in recs]) / NP
        print('%-10s D=%5d NP=%8d  %8.3f ms/gen  %8.1f GB/s  %5.1f%% of %g  evals/s=%.3e  acc=%.2f' % (
            obj, D, NP, ms, gb / ms * 1e3, 100 * gb / ms * 1e3 / PEAK, PEAK, NP / ms * 1e3, acc), flush=True)
