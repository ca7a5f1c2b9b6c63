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


# ---- unchanged single-GPU scripts under torchrun ("DataParallel-equivalent" mode) ------------------------------
# CVC-YOLOv3/train.py:193-196 wraps the model in nn.DataParallel whenever torch.cuda.device_count() > 1 and always
# uses device "cuda:0".  With B200CV_AUTO_DP=1 under torchrun every process is bound to ITS GPU before CUDA starts
# (CUDA_VISIBLE_DEVICES = the LOCAL_RANK-th visible device), so the script sees one GPU, does not wrap, and
# models.Darknet.forward takes this rank's DataParallel shard of the batch; gradients are summed over NCCL.
def auto_dp_enabled() -> bool:
    return os.environ.get("B200CV_AUTO_DP", "0") != "0" and int(os.environ.get("WORLD_SIZE", "1")) > 1


def bind_process_to_local_gpu() -> None:
    """Called when models / keypoint_net are imported.  No-op unless auto_dp_enabled()."""
    if not auto_dp_enabled() or os.environ.get("B200CV_AUTO_DP_BOUND"):
        return
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if torch.cuda.is_initialized():  # too late to mask devices: select ours instead
        torch.cuda.set_device(local)
        return
    vis = os.environ.get("CUDA_VISIBLE_DEVICES")
    ids = [v for v in vis.split(",") if v] if vis else None
    if ids is not None and local >= len(ids):
        raise RuntimeError(f"B200CV_AUTO_DP: LOCAL_RANK={local} but CUDA_VISIBLE_DEVICES={vis!r}")
    os.environ["CUDA_VISIBLE_DEVICES"] = ids[local] if ids is not None else str(local)
    os.environ["B200CV_AUTO_DP_BOUND"] = "1"  # DataLoader workers inherit the environment: bind once


def ensure_group() -> None:
    if auto_dp_enabled() and not dist.is_initialized():
        os.environ.setdefault("NCCL_P2P_LEVEL", "NVL")
        os.environ.setdefault("NCCL_IB_DISABLE", "1")
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group(backend="nccl" if torch.cuda.is_available() else "gloo")


def dp_chunk(n: int, r: int | None = None, w: int | None = None):
    """[start, stop) of the chunk nn.DataParallel's scatter (torch.chunk along the batch: ceil(n / w) per replica,
    trailing replicas may get less or nothing) hands replica r of w."""
    r = rank() if r is None else r
    w = world_size() if w is None else w
    size = -(-n // w)
    return min(n, r * size), min(n, (r + 1) * size)


def sum_over_replicas(local: torch.Tensor) -> torch.Tensor:
    """Value = sum over ranks (what nn.DataParallel's gather + `losses[0].sum()` give, train.py:70), gradient = this
    rank's own (its shard is the only part of the sum this process computed)."""
    if world_size() <= 1:
        return local
    tot = local.detach().clone()
    dist.all_reduce(tot, op=dist.ReduceOp.SUM)
    return local + (tot - local.detach())


def allreduce_gradients(arena: torch.Tensor) -> None:
    """SUM the flat gradient arena over all ranks (no-op for a single process)."""
    if world_size() > 1:
        dist.all_reduce(arena, op=dist.ReduceOp.SUM)


def allreduce_gradients_async(chunk: torch.Tensor):
    """Start the SUM all-reduce of one bucket of the gradient arena on the collective's own stream (it waits for the
    work already queued on the current stream, the current stream does NOT wait for it): backward of the earlier layers
    overlaps the transfer.  Returns a handle for finish_allreduces, or None for a single process."""
    if world_size() <= 1 or chunk.numel() == 0:
        return None
    return dist.all_reduce(chunk, op=dist.ReduceOp.SUM, async_op=True)


def finish_allreduces(handles) -> None:
    """Make the current stream wait for every bucket started with allreduce_gradients_async."""
    for h in handles:
        if h is not None:
            h.wait()


def grad_buckets() -> int:
    """Number of gradient buckets of a data-parallel step (B200CV_GRAD_BUCKETS, default 4; 1 = one all-reduce after
    backward).  A single process always uses one."""
    forced = os.environ.get("B200CV_GRAD_BUCKETS_FORCE")  # tests: segment the backward pass in a single process too
    if forced:
        return max(1, int(forced))
    return max(1, int(os.environ.get("B200CV_GRAD_BUCKETS", "4"))) if world_size() > 1 else 1


def shard_batch(n: int, r: int | None = None, w: int | None = None):
    """[start, stop) of the contiguous shard of a global batch of n that rank r owns."""
    r = rank() if r is None else r
    w = world_size() if w is None else w
    base, rem = divmod(n, w)
    start = r * base + min(r, rem)
    return start, start + base + (1 if r < rem else 0)
