"""Multi-GPU: one process per GPU (torch.distributed plumbing), the M candidate columns sharded in contiguous
blocks, no data-path collective; ONE tiny exchange of the per-rank best (value, global index) at the end
(the reference's acquire_max loop, src/acquisition.jl:58-66, is a plain running max)."""
from __future__ import annotations

import numpy as np


def shard_bounds(M: int, world_size: int, rank: int):
    """contiguous block [lo, hi) of the candidate columns owned by `rank` (remainder spread over the low ranks)."""
    base, rem = divmod(int(M), int(world_size))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def select_best(values, indices):
    """deterministic arg-max over per-rank bests: max value, then lowest global index; index < 0 or NaN never wins
    (same rule as the kernel and as acquire_max's `f > maxf`)."""
    bv, bi = -np.inf, -1
    for v, i in zip(values, indices):
        i = int(i)
        if i < 0 or not (v == v):
            continue
        if v > bv or (v == bv and bi >= 0 and i < bi):
            bv, bi = float(v), i
    return bv, bi


def allreduce_best(value: float, index: int, device=None):
    """all-gather of (value, index) over the default process group (NCCL over NVLink on GPUs, gloo on CPU)."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return select_best([value], [index])
    ws = dist.get_world_size()
    mine = torch.tensor([value], dtype=torch.float64, device=device)
    mine_i = torch.tensor([index], dtype=torch.int64, device=device)
    vals = torch.empty(ws, dtype=torch.float64, device=device)
    idxs = torch.empty(ws, dtype=torch.int64, device=device)
    dist.all_gather_into_tensor(vals, mine)
    dist.all_gather_into_tensor(idxs, mine_i)
    return select_best(vals.cpu().numpy(), idxs.cpu().numpy())


def comm_attach(model, device=None):
    """One process per GPU: give the library its own NCCL communicator over the ranks of the default process group (rank 0 draws
    the ncclUniqueId, torch.distributed carries its 128 bytes).  Afterwards model.acquire() is a collective that returns the global
    best: the 272 B/rank all-gather and the deterministic merge happen inside libb200bo, on the handle's stream."""
    import ctypes as C
    import torch
    import torch.distributed as dist
    from . import _lib
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return False
    ws, rank = dist.get_world_size(), dist.get_rank()
    buf = C.create_string_buffer(128)
    if rank == 0:
        _lib.check(_lib.lib.b200bo_comm_unique_id(buf))
    t = torch.tensor(list(buf.raw), dtype=torch.uint8, device=device)
    dist.broadcast(t, src=0)
    model.comm_init_rank(ws, rank, bytes(t.cpu().numpy().tobytes()))
    return True


def acquire_sharded(model, kind: str, params, Xs: np.ndarray, seed: int = 0, device=None, **kw):
    """Each rank scores its block of the SAME candidate matrix; returns the global (best_value, best_index, best_x)."""
    import torch.distributed as dist
    ws = dist.get_world_size() if dist.is_initialized() else 1
    rank = dist.get_rank() if dist.is_initialized() else 0
    M = Xs.shape[1]
    lo, hi = shard_bounds(M, ws, rank)
    r = model.acquire(kind, params, Xs[:, lo:hi], seed=seed, idx_offset=lo, **kw)
    bv, bi = allreduce_best(r["best_value"], r["best_index"], device=device)
    bx = Xs[:, bi].copy() if bi >= 0 else None
    return bv, bi, bx, r
