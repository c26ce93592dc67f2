"""One process per GPU: the glue between torch.distributed (plumbing only: rendezvous, shipping
the NCCL id) and a sharded oode handle.  Every rank runs the same p.solve('de') call; the library
takes identical selection decisions on every rank."""
import os


def world():
    """(world_size, rank, local_rank); (1, 0, 0) when torch.distributed is not initialised."""
    try:
        import torch.distributed as dist
    except Exception:
        return 1, 0, 0
    if not (dist.is_available() and dist.is_initialized()):
        return 1, 0, 0
    return dist.get_world_size(), dist.get_rank(), int(os.environ.get('LOCAL_RANK', dist.get_rank()))


def broadcast_bytes(payload, src=0):
    """Ship a bytes object from rank `src` to every rank (works on gloo and nccl groups)."""
    import torch.distributed as dist
    box = [payload if dist.get_rank() == src else None]
    dist.broadcast_object_list(box, src=src)
    return box[0]


_ID_CACHE = {}


def shared_nccl_id(fresh=False):
    """The 128-byte ncclUniqueId every rank passes to oode_create.  It is made once per process
    group and reused: liboode caches its NCCL communicators by id, so the second and later solves
    over the same ranks skip the NCCL bootstrap.  Collective on first use (every rank must call)."""
    import torch.distributed as dist
    from . import _lib
    key = (dist.get_world_size(), dist.get_rank())
    if fresh or key not in _ID_CACHE:
        _ID_CACHE[key] = broadcast_bytes(_lib.nccl_unique_id() if dist.get_rank() == 0 else None, src=0)
    return _ID_CACHE[key]


def agree_stop(istop):
    """Collective stop decision of a sharded solve (every rank calls it at the same point of the generation loop).

    A sharded population is advanced by collective library calls, so every rank must leave the loop on the SAME
    generation.  Criteria that depend only on the replicated records (maxIter, maxFunEvals, fEnough, ...) fire
    identically everywhere; wall-clock limits, user callbacks and rank-local errors do not.  Returns the istop every
    rank must adopt: 0 if nobody wants to stop, otherwise the stop code of the lowest rank that does."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return istop
    world = dist.get_world_size()
    dev = 'cuda' if dist.get_backend() == 'nccl' else 'cpu'
    mine = torch.tensor([float(istop) if istop else float('nan')], dtype=torch.float64, device=dev)
    gathered = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(gathered, mine)
    codes = [float(g.item()) for g in gathered]
    for c in codes:
        if c == c:                 # not NaN: this rank stops
            return int(c) if float(c).is_integer() else c
    return 0
