"""Population sweep on one GPU (BASELINE configs[1], [3], [4]): device ms/generation, evals/s and the
fraction of the measured HBM peak for Rastrigin 100-D at NP = 2^10 .. 2^24, plus C2 (NP = 10k) and
C4 (Griewank 200-D, asymmetric bounds, NP = 262144).  Writes one JSON object per line.

    python scripts/sweep.py [--max-log2 24] > gpurun_out/sweep.jsonl
"""
import argparse
import json
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from openopt_b200 import _lib, objectives

ap = argparse.ArgumentParser()
ap.add_argument('--max-log2', type=int, default=24)
ap.add_argument('--min-log2', type=int, default=10)
args = ap.parse_args()
try:
    PEAK = float(json.load(open(os.path.join(os.path.dirname(__file__), '..', 'MEASURED_PEAKS.json')))['hbm_gbs'])
except Exception:
    PEAK = 6650.0


def run(name, obj, D, NP, lb, ub, gens):
    o = objectives.get(obj)
    with _lib.DeviceDE(o.device_id, lb, ub, x0=lb + 0.25 * (ub - lb), population=NP, obj_aux=o.aux(D), max_batch=64) as dev:
        dev.init()
        dev.step(min(10, gens))
        c0 = dev.counters()
        left, recs = gens, []
        while left > 0:
            recs += dev.step(min(left, 64))
            left -= min(left, 64)
        c1 = dev.counters()
    ms = (c1['kernel_ms'] - c0['kernel_ms']) / gens
    gb = NP * (40 * D + 16) / 1e9
    line = dict(case=name, objective=obj, D=D, population=NP, generations=gens, ms_per_generation=ms,
                evals_per_s=NP / ms * 1e3, algorithmic_GBps=gb / ms * 1e3, frac_of_hbm_peak=gb / ms * 1e3 / PEAK,
                population_bytes=NP * D * 8, fits_l2=bool(2 * NP * D * 8 < 100e6),
                acceptance=float(np.mean([r.n_accepted for r in recs]) / NP))
    print(json.dumps(line), flush=True)


D = 100
run('c2', 'rastrigin', D, 10000, -5.12 * np.ones(D), 5.12 * np.ones(D), 200)
for lg in range(args.min_log2, args.max_log2 + 1, 2):
    NP = 1 << lg
    run('c5', 'rastrigin', D, NP, -5.12 * np.ones(D), 5.12 * np.ones(D), 200 if lg <= 16 else (40 if lg <= 20 else 12))
D = 200
run('c4', 'griewank', D, 262144, -600.0 * np.ones(D), np.linspace(100.0, 900.0, D), 40)
run('c4_reference_default_np', 'griewank', D, 2000, -600.0 * np.ones(D), np.linspace(100.0, 900.0, D), 200)
