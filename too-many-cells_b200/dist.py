"""Multi-GPU plumbing: one process per GPU, torch.distributed only bootstraps.

Rows (cells) are owned statically: rank r holds the contiguous row block
[m*r/world, m*(r+1)/world) of B for the whole tree build, so no matrix data ever moves
between GPUs; every tree node is a (possibly empty) local position range on every rank and
the ranks exchange only the feature-dimension vector y = B_S^T x plus a handful of scalars
per step, through the NCCL communicator that libtmc owns (`tmc_comm_init`).
"""
from __future__ import annotations

import ctypes as C
import os

from . import _lib


def row_block(m: int, rank: int, world: int):
    """[lo, hi) of the rows rank `rank` owns — the same arithmetic libtmc uses when it slices a whole matrix."""
    return m * rank // world, m * (rank + 1) // world


def exchange_unique_id(rank: int, world: int, device=None) -> bytes:
    """Rank 0 asks libtmc (NCCL) for a unique id and broadcasts the 128 bytes over torch.distributed."""
    import torch
    import torch.distributed as dist
    lib = _lib.load()
    buf = (C.c_uint8 * 128)()
    if rank == 0:
        _lib.check(lib.tmc_comm_unique_id(buf))
    t = torch.tensor(list(buf), dtype=torch.uint8)
    if device is not None:
        t = t.to(device)
    if world > 1:
        dist.broadcast(t, 0)
    return bytes(t.cpu().tolist())


def init_comm(rank=None, world=None, local_rank=None, device=None):
    """Create libtmc's NCCL communicator for this process (no-op for world == 1). Returns (rank, world)."""
    rank = int(os.environ.get("RANK", "0")) if rank is None else rank
    world = int(os.environ.get("WORLD_SIZE", "1")) if world is None else world
    local_rank = int(os.environ.get("LOCAL_RANK", "0")) if local_rank is None else local_rank
    if world <= 1:
        return rank, world
    uid = exchange_unique_id(rank, world, device)
    lib = _lib.load()
    _lib.check(lib.tmc_comm_init(C.create_string_buffer(uid, 128), rank, world, local_rank))
    return rank, world


def destroy_comm():
    _lib.load().tmc_comm_destroy()
