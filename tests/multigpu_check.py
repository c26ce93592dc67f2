"""Run under torchrun (one process per GPU):

    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port 29511 \
        tests/multigpu_check.py [--big]

Checks that the column-sharded run over N GPUs is BIT-IDENTICAL to the single-GPU run of the same
problem (records, Best.x, final population and vals): draws are keyed by global (row, column) and
the leaf sums are folded in the same tree order, so the result must not depend on N.
With --big also times BASELINE configs[2] (Ackley 1000-D, population 1M) sharded over the N GPUs.
"""
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from openopt_b200 import GLP, _lib, objectives, distributed as ddist  # noqa: E402


def records(recs):
    return [(r.generation, r.trial_best_idx, r.trial_best_f, r.best_f, r.n_accepted, r.n_repaired, r.best_changed)
            for r in recs]


def main():
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    torch.cuda.set_device(local_rank)
    dist.init_process_group('nccl', device_id=torch.device('cuda', local_rank))
    world, rank, _ = ddist.world()
    ok = True
    # every case under both exchanges: leaf sums stored straight into the peers' memory by the trial
    # kernel (the default) and the NCCL all-gather in row chunks
    cases = (('ackley', 1000, 4096, False, 1), ('rastrigin', 1500, 1000, True, 3), ('griewank', 700, 777, True, 4),
             ('sphere', 2049, 300, False, 7), ('ackley', 1024, 70001, False, 2),
             ('rosenbrock', 1100, 500, True, 2))        # neighbour-coupled: the shards overlap by one column
    for xchg in (('p2p',) if '--p2p-only' in sys.argv else ('p2p', 'nccl')):
      for obj, D, NP, asym, chunks in cases:
        if len(_lib_leaves(D - 1 if obj == 'rosenbrock' else D)) < world:
            continue
        os.environ['OODE_EXCHANGE'] = xchg
        os.environ['OODE_CHUNKS'] = str(chunks)          # row chunks of the overlapped NCCL leaf-sum exchange
        lb = -5.0 * np.ones(D)
        ub = np.linspace(2.0, 9.0, D) if asym else 5.0 * np.ones(D)
        x0 = lb + 0.25 * (ub - lb)
        o = objectives.get(obj)
        gens = 7
        with _lib.DeviceDE(o.device_id, lb, ub, x0=x0, population=NP, obj_aux=o.aux(D), seed=99, device=local_rank,
                           max_batch=8) as one:
            r1 = [one.init()] + one.step(gens)
            bx1 = one.best_x()
            pop1, vals1 = one.population()
        nid = ddist.shared_nccl_id()
        with _lib.DeviceDE(o.device_id, lb, ub, x0=x0, population=NP, obj_aux=o.aux(D), seed=99, device=local_rank,
                           max_batch=8, world_size=world, rank=rank, shard=_lib.SHARD_COLS, nccl_id=nid) as sh:
            mode = _lib.XCHG_NAMES[sh.exchange_mode()]
            r2 = [sh.init()] + sh.step(gens)
            bx2 = sh.best_x()
            pop2 = np.zeros((NP, D))
            vals2 = np.zeros(NP)
            _lib._check(sh.lib.oode_get_population(sh.h, _lib._d(pop2), _lib._d(vals2)), 'get_population')
            col0, ncols, _, _ = _lib.shard_layout(D, o.device_id, world, rank)
        if obj == 'rosenbrock' and rank < world - 1:
            pop2[:, col0 + ncols - 1] = 0.0             # the column shared with the next shard: summed once below
        t = torch.from_numpy(pop2).cuda()
        dist.all_reduce(t)                              # every rank filled only its own columns
        pop2 = t.cpu().numpy()
        good = (records(r1) == records(r2) and np.array_equal(bx1, bx2) and np.array_equal(pop1, pop2)
                and np.array_equal(vals1, vals2, equal_nan=True))
        ok &= good
        if rank == 0:
            print('%-10s D=%5d NP=%6d  N=%d  [%s]  sharded == single GPU: %s' % (obj, D, NP, world, mode, good),
                  flush=True)
    os.environ.pop('OODE_CHUNKS', None)
    os.environ.pop('OODE_EXCHANGE', None)
    # row shards: rows too narrow to column-shard, neighbour-coupled objectives, non-default strategies
    row_cases = (('rastrigin', 100, 5000, False, {}), ('rosenbrock', 37, 1001, True, {}), ('griewank', 200, 777, True, {}),
                 ('ackley', 1000, 300, False, {}), ('sphere', 5, 64, True, {}),
                 ('rastrigin', 100, 3000, False, dict(base='best', direction='best', hndvi=3)),
                 ('rosenbrock', 130, 500, True, dict(base='best', factor='constant', hndvi=2)))
    for obj, D, NP, asym, strat in row_cases:
        lb = -5.0 * np.ones(D)
        ub = np.linspace(2.0, 9.0, D) if asym else 5.0 * np.ones(D)
        x0 = lb + 0.25 * (ub - lb)
        o = objectives.get(obj)
        gens = 9
        with _lib.DeviceDE(o.device_id, lb, ub, x0=x0, population=NP, obj_aux=o.aux(D), seed=7, device=local_rank,
                           max_batch=4, **strat) as one:
            r1 = [one.init()] + one.step(4) + one.step(4) + one.step(1)
            bx1 = one.best_x()
            pop1, vals1 = one.population()
            acc1, tv1 = one.accept_mask(), one.trial_vals()
        with _lib.DeviceDE(o.device_id, lb, ub, x0=x0, population=NP, obj_aux=o.aux(D), seed=7, device=local_rank,
                           max_batch=4, world_size=world, rank=rank, shard=_lib.SHARD_ROWS, nccl_id=ddist.shared_nccl_id(),
                           **strat) as sh:
            r2 = [sh.init()] + sh.step(4) + sh.step(4) + sh.step(1)
            bx2 = sh.best_x()
            pop2, vals2 = sh.population()
            acc2, tv2 = sh.accept_mask(), sh.trial_vals()
        good = (records(r1) == records(r2) and np.array_equal(bx1, bx2) and np.array_equal(pop1, pop2)
                and np.array_equal(vals1, vals2, equal_nan=True) and np.array_equal(acc1, acc2)
                and np.array_equal(tv1, tv2, equal_nan=True))
        t = torch.tensor([int(good)], device='cuda')
        dist.all_reduce(t, op=dist.ReduceOp.MIN)              # every rank's replica must match
        good = bool(t.item())
        ok &= good
        if rank == 0:
            print('%-10s D=%5d NP=%6d  N=%d  [row shards%s]  every replica == single GPU: %s'
                  % (obj, D, NP, world, ' ' + str(strat) if strat else '', good), flush=True)
    # the reference's own random stream (regenerated natively by every rank) on column and row shards
    for obj, D, NP, kind in (('ackley', 1000, 96, _lib.SHARD_COLS), ('rastrigin', 300, 64, _lib.SHARD_COLS),
                             ('rastrigin', 100, 200, _lib.SHARD_ROWS), ('rosenbrock', 37, 90, _lib.SHARD_ROWS)):
        if kind == _lib.SHARD_COLS and len(_lib_leaves(D)) < world:
            continue
        lb, ub = -5.0 * np.ones(D), np.linspace(2.0, 9.0, D)
        o = objectives.get(obj)
        kw = dict(x0=lb + 0.25 * (ub - lb), population=NP, obj_aux=o.aux(D), device=local_rank, max_batch=4, rng='cpython')
        with _lib.DeviceDE(o.device_id, lb, ub, **kw) as one:
            r1 = [one.init()] + one.step(4) + one.step(2)
            bx1, v1, a1 = one.best_x(), one.population_vals(), one.accept_mask()
        with _lib.DeviceDE(o.device_id, lb, ub, world_size=world, rank=rank, shard=kind, nccl_id=ddist.shared_nccl_id(), **kw) as sh:
            r2 = [sh.init()] + sh.step(4) + sh.step(2)
            bx2, v2, a2 = sh.best_x(), sh.population_vals(), sh.accept_mask()
        good = records(r1) == records(r2) and np.array_equal(bx1, bx2) and np.array_equal(v1, v2) and np.array_equal(a1, a2)
        t = torch.tensor([int(good)], device='cuda')
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        good = bool(t.item())
        ok &= good
        if rank == 0:
            print('%-10s D=%5d NP=%6d  N=%d  [CPython stream, %s shards]  sharded == single GPU: %s'
                  % (obj, D, NP, world, 'column' if kind == _lib.SHARD_COLS else 'row', good), flush=True)
    # the public API, sharded automatically because torch.distributed is initialised
    D, NP = 1000, 2000
    lb, ub = -32.768 * np.ones(D), 32.768 * np.ones(D)
    kw = dict(lb=lb, ub=ub, x0=lb + 0.25 * (ub - lb), maxIter=11, maxFunEvals=10 ** 12, maxNonSuccess=10 ** 9, iprint=-1)
    r_sh = GLP(objectives.Ackley(), **kw).solve('de', population=NP)
    r_one = GLP(objectives.Ackley(), **kw).solve('de', population=NP, shard='none')
    good = np.array_equal(r_sh.xf, r_one.xf) and r_sh.ff == r_one.ff and r_sh.iterValues.f == r_one.iterValues.f
    ok &= good
    if rank == 0:
        print('GLP.solve sharded == unsharded: %s (ff=%r)' % (good, r_sh.ff), flush=True)
    # narrow rows: shard='auto' falls back to row shards
    D, NP = 100, 4000
    lb, ub = -5.12 * np.ones(D), 5.12 * np.ones(D)
    kw = dict(lb=lb, ub=ub, x0=lb + 0.25 * (ub - lb), maxIter=21, maxFunEvals=10 ** 12, maxNonSuccess=10 ** 9, iprint=-1)
    r_sh = GLP(objectives.Rastrigin(), **kw).solve('de', population=NP)
    r_one = GLP(objectives.Rastrigin(), **kw).solve('de', population=NP, shard='none')
    good = np.array_equal(r_sh.xf, r_one.xf) and r_sh.ff == r_one.ff and r_sh.iterValues.f == r_one.iterValues.f
    ok &= good
    if rank == 0:
        print('GLP.solve row-sharded (auto, D=100) == unsharded: %s (ff=%r)' % (good, r_sh.ff), flush=True)

    if '--big' in sys.argv:
        D, NP = 1000, 1000000
        lb, ub = -32.768 * np.ones(D), 32.768 * np.ones(D)
        o = objectives.Ackley()
        nid = ddist.shared_nccl_id()
        with _lib.DeviceDE(o.device_id, lb, ub, x0=lb + 0.25 * (ub - lb), population=NP, device=local_rank, max_batch=32,
                           world_size=world, rank=rank, shard=_lib.SHARD_COLS, nccl_id=nid) as sh:
            sh.init()
            sh.step(3)
            torch.cuda.synchronize()
            dist.barrier()
            c0 = sh.counters()
            t0 = time.perf_counter()
            recs = sh.step(20)
            torch.cuda.synchronize()
            dist.barrier()
            wall = time.perf_counter() - t0
            ms = (sh.counters()['kernel_ms'] - c0['kernel_ms']) / 20
        t = torch.tensor([ms], device='cuda', dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        if rank == 0:
            print('C3 over %d GPUs: %.3f ms/generation (device, max over ranks), wall %.3f ms/gen, %.3e evals/s, acc=%.2f'
                  % (world, t.item(), 1e3 * wall / 20, NP / (t.item() * 1e-3), np.mean([r.n_accepted for r in recs]) / NP),
                  flush=True)
    dist.barrier()
    dist.destroy_process_group()
    if not ok:
        sys.exit(1)


def _lib_leaves(D):
    sys.path.insert(0, os.path.join(ROOT, 'oracle'))
    import de_oracle
    return de_oracle.pairwise_leaves(D)


if __name__ == '__main__':
    main()
