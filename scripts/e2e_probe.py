"""Where the wall time of a short `p.solve('de')` goes (run under torchrun for a sharded solve):
library phase timings (OODE_TIMING) + a cProfile of the host side on rank 0."""
import cProfile
import os
import pstats
import sys
import time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist
from openopt_b200 import GLP, objectives

world = int(os.environ.get('WORLD_SIZE', '1'))
local_rank = int(os.environ.get('LOCAL_RANK', '0'))
torch.cuda.set_device(local_rank)
if world > 1:
    dist.init_process_group('nccl', device_id=torch.device('cuda', local_rank))
D, NP, K = 1000, 1000000, int(sys.argv[1]) if len(sys.argv) > 1 else 20
lb, ub = -32.768 * np.ones(D), 32.768 * np.ones(D)


def solve():
    p = GLP(objectives.Ackley(), lb=lb, ub=ub, x0=lb + 0.25 * (ub - lb), maxIter=K + 1, maxFunEvals=10 ** 15,
            maxNonSuccess=10 ** 9, iprint=-1)
    return p.solve('de', population=NP, gensPerCall=min(K, 64))


solve()                                   # warm: communicator, block cache
for rep in range(2):
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    if rep == 1 and local_rank == 0:
        os.environ['OODE_TIMING'] = '1'
        pr = cProfile.Profile()
        t0 = time.perf_counter()
        pr.enable()
        solve()
        pr.disable()
        print('rank 0 wall %.1f ms' % (1e3 * (time.perf_counter() - t0)), file=sys.stderr)
        pstats.Stats(pr, stream=sys.stderr).sort_stats('cumulative').print_stats(18)
    else:
        t0 = time.perf_counter()
        solve()
        if local_rank == 0:
            print('rank 0 wall %.1f ms (no profiler)' % (1e3 * (time.perf_counter() - t0)), file=sys.stderr)
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
