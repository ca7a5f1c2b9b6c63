"""Data-parallel plumbing: one process per GPU, gradients summed over NCCL (NVLink / NVSwitch).

The reference's only multi-GPU mechanism is nn.DataParallel (CVC-YOLOv3/train.py:193-195), whose
backward reduce-ADDS the replicas' gradients of per-replica MEAN losses onto GPU 0 -- so the
reduction here is a SUM (not a mean), BatchNorm statistics stay per replica (no SyncBN), and the
loss each rank reports is the loss of its own shard.
"""
from __future__ import annotations

import os

import torch
import torch.distributed as dist


def world_size() -> int:
    return dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1


def rank() -> int:
    return dist.get_rank() if dist.is_available() and dist.is_initialized() else 0


def init_from_env(backend: str | None = None) -> int:
    """Initialise torch.distributed from torchrun's environment (RANK/WORLD_SIZE/LOCAL_RANK/MASTER_*).
    Returns the local rank.  NVLink-only transport is requested for NCCL (no IB, no SHM)."""
    if "RANK" not in os.environ or int(os.environ.get("WORLD_SIZE", "1")) <= 1:
        return int(os.environ.get("LOCAL_RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if backend is None:
        backend = "nccl" if torch.cuda.is_available() else "gloo"
    if backend == "nccl":
        os.environ.setdefault("NCCL_P2P_LEVEL", "NVL")
        os.environ.setdefault("NCCL_IB_DISABLE", "1")
        torch.cuda.set_device(local)
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    if not dist.is_initialized():
        dist.init_process_group(backend=backend)
    return local


def allreduce_gradients(arena: torch.Tensor) -> None:
    """SUM the flat gradient arena over all ranks (no-op for a single process)."""
    if world_size() > 1:
        dist.all_reduce(arena, op=dist.ReduceOp.SUM)


def shard_batch(n: int, r: int | None = None, w: int | None = None):
    """[start, stop) of the contiguous shard of a global batch of n that rank r owns."""
    r = rank() if r is None else r
    w = world_size() if w is None else w
    base, rem = divmod(n, w)
    start = r * base + min(r, rem)
    return start, start + base + (1 if r < rem else 0)
