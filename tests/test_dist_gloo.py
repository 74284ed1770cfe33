"""world_size-2 test of the only cross-rank step on the path (the per-rank best exchange), gloo on CPU."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as tmp


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, ws, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=ws)
    from b200bo.dist import allreduce_best, shard_bounds
    vals = np.array([0.3, 2.0, np.nan, 2.0, 1.5, -1.0, 2.0])          # global candidate scores; ties at 1, 3, 6
    lo, hi = shard_bounds(vals.size, ws, rank)
    loc = vals[lo:hi]
    ok = ~np.isnan(loc)
    if ok.any():
        j = int(np.flatnonzero(ok & (loc == loc[ok].max()))[0])
        v, i = float(loc[j]), lo + j
    else:
        v, i = -np.inf, -1
    bv, bi = allreduce_best(v, i)
    # second round: one rank has nothing
    bv2, bi2 = allreduce_best(-np.inf if rank == 0 else 7.0, -1 if rank == 0 else 11)
    out[rank] = (bv, bi, bv2, bi2)
    dist.destroy_process_group()


def test_allreduce_best_world_size_2():
    import b200bo  # noqa: F401  (import before spawn so the shim is importable in the children)
    mgr = tmp.Manager()
    out = mgr.dict()
    tmp.spawn(_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    assert out[0] == out[1] == (2.0, 1, 7.0, 11)
