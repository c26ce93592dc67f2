"""BASELINE configs[4]: population sweep 1k .. 16M at 100-D Rastrigin, on N GPUs (N = 1: this process; N > 1: launch
under torchrun, row shards) next to the CPU columns: the C port of the reference algorithm on every host core and the
UNMODIFIED reference (baseline/_ref, one core, populations it can finish; larger ones are flagged extrapolated --
its evals/s is ~independent of NP at fixed D).  One JSON object per population on stdout (rank 0).

    python scripts/sweep_c5.py [--min-log2 10] [--max-log2 24] [--no-cpu]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 scripts/sweep_c5.py --no-cpu
"""
import argparse
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
from openopt_b200 import _lib, objectives  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument('--max-log2', type=int, default=24)
ap.add_argument('--min-log2', type=int, default=10)
ap.add_argument('--no-cpu', action='store_true')
args = ap.parse_args()
world, rank, local = int(os.environ.get('WORLD_SIZE', '1')), int(os.environ.get('RANK', '0')), int(os.environ.get('LOCAL_RANK', '0'))
try:
    PEAK = float(json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))['hbm_gbs'])
except Exception:
    PEAK = 6650.0
D = 100
lb, ub = -5.12 * np.ones(D), 5.12 * np.ones(D)
x0 = lb + 0.25 * (ub - lb)
o = objectives.get('rastrigin')
extra = {}
if world > 1:
    import torch
    import torch.distributed as dist
    from openopt_b200 import distributed as ddist
    torch.cuda.set_device(local)
    dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    extra = dict(world_size=world, rank=rank, shard=_lib.SHARD_ROWS, nccl_id=ddist.shared_nccl_id())

ref_py = None
if rank == 0 and not args.no_cpu:
    sys.path.insert(0, os.path.join(ROOT, 'oracle'))
    import c_oracle
    threads = c_oracle.set_threads()
    r = subprocess.run([sys.executable, os.path.join(ROOT, 'baseline', 'time_reference.py'), 'rastrigin', str(D), '16384', '3', '-5.12', '5.12'],
                       capture_output=True, text=True)
    try:
        ref_py = json.loads(r.stdout.strip().splitlines()[-1])
    except Exception:
        ref_py = None

for lg in range(args.min_log2, args.max_log2 + 1, 2):
    NP = 1 << lg
    if world > 1 and NP < world:
        continue
    gens = 200 if lg <= 16 else (40 if lg <= 20 else 12)
    with _lib.DeviceDE(o.device_id, lb, ub, x0=x0, population=NP, device=local, max_batch=64, **extra) as dev:
        dev.init()
        dev.step(min(10, gens))
        if world > 1:
            torch.cuda.synchronize(); dist.barrier()
        c0 = dev.counters()
        left = gens
        while left > 0:
            dev.step(min(left, 64))
            left -= min(left, 64)
        c1 = dev.counters()
        kname = dev.trial_kernel_name()
    ms = (c1['kernel_ms'] - c0['kernel_ms']) / gens
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device='cuda')
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    if rank != 0:
        continue
    line = dict(workload='c5', objective='rastrigin', D=D, population=NP, n_gpus=world, generations=gens, ms_per_generation=ms,
                evals_per_s=NP / ms * 1e3, frac_of_hbm_peak_per_gpu=NP * (40 * D + 16) / world / 1e9 / ms * 1e3 / PEAK,
                trial_kernel=kname, sharding='row shards' if world > 1 else 'single GPU')
    if not args.no_cpu:
        rows = min(NP, 1 << 18)                      # the C port is timed on up to 256k rows (a few seconds)
        cg = max(2, min(200, int(2e8 // (rows * D))))
        c = c_oracle.COracleDE('rastrigin', lb, ub, x0=x0, population=rows, F=0.8, Cr=0.5, seed=150880, stream='philox')
        c.step(1)
        t0 = time.perf_counter()
        c.step(cg)
        dt = time.perf_counter() - t0
        c.close()
        line['cpu_port'] = dict(evals_per_s=rows * cg / dt, cores=threads, rows=rows, extrapolated=bool(rows < NP))
        if ref_py and 'evals_per_s' in ref_py:
            line['cpu_reference_python'] = dict(evals_per_s=ref_py['evals_per_s'], cores=1, rows=ref_py['population'],
                                                extrapolated=bool(ref_py['population'] != NP))
    print(json.dumps(line), flush=True)
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
