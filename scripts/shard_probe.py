"""Timing probe of a column-sharded population (run under torchrun): device ms/generation, and the
share of it spent in the trial kernel + exchange vs the selection kernel."""
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist
from openopt_b200 import _lib, objectives, distributed as ddist

local_rank = int(os.environ.get('LOCAL_RANK', '0'))
torch.cuda.set_device(local_rank)
backend = os.environ.get('PROBE_BACKEND', 'nccl')
if backend == 'nccl':
    dist.init_process_group('nccl', device_id=torch.device('cuda', local_rank))
else:
    dist.init_process_group(backend)
unsharded = bool(os.environ.get('PROBE_UNSHARDED'))
world, rank, _ = ddist.world()
obj, D, NP, gens = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
hw = float(sys.argv[5]) if len(sys.argv) > 5 else 32.768
o = objectives.get(obj)
lb, ub = -hw * np.ones(D), hw * np.ones(D)
kind = _lib.SHARD_ROWS if os.environ.get('PROBE_ROWS') else _lib.SHARD_COLS
extra = {} if unsharded else dict(world_size=world, rank=rank, shard=kind, nccl_id=ddist.shared_nccl_id())
with _lib.DeviceDE(o.device_id, lb, ub, x0=lb + 0.25 * (ub - lb), population=NP, obj_aux=o.aux(D), device=local_rank,
                   max_batch=64, **extra) as dev:
    mode = _lib.XCHG_NAMES[dev.exchange_mode()]
    dev.init()
    dev.step(3)
    torch.cuda.synchronize()
    dist.barrier()
    c0 = dev.counters()
    dev.step(gens)
    c1 = dev.counters()
    dev.set_profiling(True)
    dev.step(gens)
    c2 = dev.counters()
ms = (c1['kernel_ms'] - c0['kernel_ms']) / gens
dev_t = 'cuda' if backend == 'nccl' else 'cpu'
t = torch.tensor([ms, (c2['trial_kernel_ms'] - c1['trial_kernel_ms']) / gens, (c2['select_kernel_ms'] - c1['select_kernel_ms']) / gens,
                  (c2['exchange_wait_ms'] - c1['exchange_wait_ms']) / gens],
                 dtype=torch.float64, device=dev_t)
all_t = [torch.zeros_like(t) for _ in range(world)]
dist.all_gather(all_t, t)
if rank == 0:
    print('%s D=%d NP=%d N=%d [%s]%s pg=%s%s%s' % (obj, D, NP, world, mode, ' ROW SHARDS' if os.environ.get('PROBE_ROWS') else '', backend, ' UNSHARDED' if unsharded else '', ' LOCAL STORES ONLY' if os.environ.get('OODE_DEBUG_LOCAL_STORES') else ''))
    for r, v in enumerate(all_t):
        print('  rank %d: %.3f ms/gen   (profiled pass: trial+exchange %.3f of which publish+wait %.3f, select %.3f)'
              % (r, v[0].item(), v[1].item(), v[3].item(), v[2].item()))
    worst = max(v[0].item() for v in all_t)
    print('  max over ranks %.3f ms/gen = %.3e evals/s' % (worst, NP / worst * 1e3), flush=True)
dist.barrier()
dist.destroy_process_group()
