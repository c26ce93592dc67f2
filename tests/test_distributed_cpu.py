"""CPU tests of the multi-GPU host logic: the column-shard layout against the pairwise-sum tree,
and a world_size-2 gloo run of the plugin's distributed glue (device handle replaced by the
oracle): both ranks must agree on the NCCL id, pass the right shard description to the handle and
produce the same result as a single process."""
import os
import subprocess
import sys

import pytest

import de_oracle as orc
from conftest import ROOT
from openopt_b200 import _lib


@pytest.mark.parametrize('D', [129, 200, 1000, 1001, 4096, 10007])
@pytest.mark.parametrize('G', [1, 2, 4, 8])
def test_shard_layout_follows_the_pairwise_tree(D, G):
    leaves = orc.pairwise_leaves(D)
    if len(leaves) < G:
        with pytest.raises(_lib.LibraryError, match='fewer pairwise-sum leaves'):
            _lib.shard_layout(D, _lib.OBJ_ACKLEY, G, 0)
        return
    cover, nl = [], 0
    for r in range(G):
        col0, ncols, leaf0, nleaves = _lib.shard_layout(D, _lib.OBJ_ACKLEY, G, r)
        assert leaf0 == nl and nleaves >= 1
        assert col0 == leaves[leaf0][0] and col0 % 8 == 0
        assert ncols == sum(l for _, l in leaves[leaf0:leaf0 + nleaves])
        cover += list(range(col0, col0 + ncols))
        nl += nleaves
    assert cover == list(range(D)) and nl == len(leaves)


def test_rosenbrock_column_shards_overlap_by_one_column():
    """A neighbour-coupled objective (term j reads columns j and j+1) is sharded on the tree of its D-1 terms; every
    shard but the last also holds the first column of the next one."""
    leaves = orc.pairwise_leaves(999)
    cols = [_lib.shard_layout(1000, _lib.OBJ_ROSENBROCK, 4, r) for r in range(4)]
    assert [c[3] for c in cols] == [2, 2, 2, 3] and sum(c[3] for c in cols) == len(leaves)
    for r in range(3):
        assert cols[r][0] + cols[r][1] - 1 == cols[r + 1][0]          # last column == the next shard's first
    assert cols[3][0] + cols[3][1] == 1000
    assert _lib.shard_layout(1000, _lib.OBJ_ROSENBROCK, 1, 0)[1] == 1000


WORKER = r'''
import os, sys
sys.path.insert(0, %(root)r); sys.path.insert(0, os.path.join(%(root)r, 'oracle')); sys.path.insert(0, os.path.join(%(root)r, 'tests'))
import numpy as np, torch.distributed as dist
dist.init_process_group('gloo')
import helpers
from openopt_b200 import GLP, _lib, objectives, distributed as ddist
seen = {}
class Fake(helpers.OracleBackedDE):
    def __init__(self, *a, **kw):
        seen.update(world_size=kw.get('world_size'), rank=kw.get('rank'), shard=kw.get('shard'), nccl_id=kw.get('nccl_id'),
                    device=kw.get('device'))
        super().__init__(*a, **kw)
_lib.DeviceDE = Fake
_lib.nccl_unique_id = lambda: bytes(range(128))          # no GPU here: a stand-in id made on rank 0 only
lb, ub = -5.12 * np.ones(12), 5.12 * np.ones(12)
p = GLP(objectives.Rastrigin(), lb=lb, ub=ub, x0=lb + 1.0, maxIter=30, maxFunEvals=10 ** 9, maxNonSuccess=10 ** 9, iprint=-1)
r = p.solve('de', population=40, rng='cpython')
assert ddist.world()[0] == 2
# a 12-column row has one pairwise-sum leaf: it cannot be column-sharded over 2 ranks -> row shards
assert seen['world_size'] == 2 and seen['rank'] == dist.get_rank() and seen['shard'] == _lib.SHARD_ROWS
assert seen['nccl_id'] == bytes(range(128)) and seen['device'] == int(os.environ['LOCAL_RANK'])
out = [None, None]
dist.all_gather_object(out, (r.ff, r.xf.tolist(), r.evals['f'], r.istop))
assert out[0] == out[1]
p1 = GLP(objectives.Rastrigin(), lb=lb, ub=ub, x0=lb + 1.0, maxIter=30, maxFunEvals=10 ** 9, maxNonSuccess=10 ** 9, iprint=-1)
r1 = p1.solve('de', population=40, rng='cpython', shard='none')
assert seen['world_size'] is None and r1.ff == r.ff
# 300 columns = 3 leaves >= 2 ranks -> column shards (also for Rosenbrock: shards overlapping by one column);
# 'rows' forces row shards
lb3, ub3 = -5.12 * np.ones(300), 5.12 * np.ones(300)
kw3 = dict(lb=lb3, ub=ub3, x0=lb3 + 1.0, maxIter=4, maxFunEvals=10 ** 9, maxNonSuccess=10 ** 9, iprint=-1)
GLP(objectives.Rastrigin(), **kw3).solve('de', population=20, rng='cpython')
assert seen['shard'] == _lib.SHARD_COLS
GLP(objectives.Rastrigin(), **kw3).solve('de', population=20, rng='cpython', shard='rows')
assert seen['shard'] == _lib.SHARD_ROWS
GLP(objectives.Rosenbrock(), **kw3).solve('de', population=20, rng='cpython')
assert seen['shard'] == _lib.SHARD_COLS
# a stop only ONE rank asks for (user callback on rank 1) stops BOTH ranks on the same generation: a sharded handle is
# advanced by collective calls, so nobody may leave the loop alone (distributed.agree_stop)
cb = lambda q: bool(dist.get_rank() == 1 and q.iter >= 5)
p2 = GLP(objectives.Rastrigin(), lb=lb, ub=ub, x0=lb + 1.0, maxIter=30, maxFunEvals=10 ** 9, maxNonSuccess=10 ** 9, iprint=-1,
         callback=cb)
r2 = p2.solve('de', population=40, rng='cpython')
out = [None, None]
dist.all_gather_object(out, (len(r2.iterValues.f), r2.evals['f'], r2.istop))
assert out[0] == out[1], out
assert r2.istop == 80 and out[0][0] < 30, out
assert ('another rank' in r2.msg) == (dist.get_rank() == 0), r2.msg
# the sibling solver does not shard (its roulette couples every pick to the whole list): every rank runs its own full
# copy on its own device -- replicas only -- and the copies agree
gseen = {}
class FakeGA(helpers.OracleBackedGalileo):
    def __init__(self, *a, **kw):
        gseen.update(device=kw.get('device'))
        super().__init__(*a, **kw)
_lib.DeviceGalileo = FakeGA
pg = GLP(objectives.Rastrigin(), lb=lb, ub=ub, maxIter=20, maxFunEvals=10 ** 9, maxNonSuccess=10 ** 9, iprint=-1)
rg = pg.solve('galileo', population=16, rng='cpython')
assert gseen['device'] == int(os.environ['LOCAL_RANK'])
out = [None, None]
dist.all_gather_object(out, (rg.ff, rg.xf.tolist(), rg.evals['f'], rg.istop))
assert out[0] == out[1], out
sys.stdout.write('rank ' + str(dist.get_rank()) + ' ok\n')      # ONE write per rank: the two ranks share the pipe
sys.stdout.flush()
'''


def free_port():
    import socket
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


def test_plugin_distributed_glue_gloo_world2(tmp_path):
    script = tmp_path / 'worker.py'
    script.write_text(WORKER % {'root': ROOT})
    out = subprocess.run([sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', '2',
                          '--master-addr', '127.0.0.1', '--master-port', str(free_port()), str(script)],
                         capture_output=True, text=True, timeout=300, env=dict(os.environ, CUDA_VISIBLE_DEVICES=''))
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-3000:]
    assert 'rank 0 ok' in out.stdout and 'rank 1 ok' in out.stdout
