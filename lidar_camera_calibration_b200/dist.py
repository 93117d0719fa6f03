"""Multi-GPU plumbing: one process per GPU, frames sharded in contiguous blocks, one exchange step.

Frames are independent units, so the data path needs no collective; the only exchange is the
all-gather of the per-frame corners (+ status) before the host-side PnP (ilcc2/test/calib_lidar_cam.cpp:116-165
consumes all frames' corners).  torch.distributed carries it: NCCL over NVLink on GPUs, gloo in CPU tests.
"""
from __future__ import annotations

import os


def init_from_env(backend: str | None = None):
    """Initialise torch.distributed from torchrun's environment. Returns (rank, world_size, local_rank)."""
    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local)
            dist.init_process_group(backend, device_id=torch.device("cuda", local))
        else:
            dist.init_process_group(backend)
    return rank, world, local


def shard_range(n_frames: int, world_size: int, rank: int):
    """Contiguous block of frames owned by `rank` (same arithmetic as ilcc2_shard_range in the C ABI)."""
    b = n_frames * rank // world_size
    e = n_frames * (rank + 1) // world_size
    return b, e - b


def allgather_corners(corners, status, world_size: int, out_corners=None, out_status=None):
    """corners (F_local, ncorn, 3) float32, status (F_local,) int32 -- equal F_local on every rank.
    Returns (all_corners (world*F_local, ncorn, 3), all_status (world*F_local,)) on every rank."""
    import torch
    import torch.distributed as dist
    if world_size == 1:
        return corners, status
    F = corners.shape[0]
    if out_corners is None:
        out_corners = torch.empty((world_size * F,) + tuple(corners.shape[1:]), dtype=corners.dtype, device=corners.device)
    if out_status is None:
        out_status = torch.empty((world_size * F,), dtype=status.dtype, device=status.device)
    dist.all_gather_into_tensor(out_corners, corners.contiguous())
    dist.all_gather_into_tensor(out_status, status.contiguous())
    return out_corners, out_status


def allgather_rows(x, world_size: int):
    """Concatenation over ranks (rank order) of a tensor with the same shape on every rank."""
    import torch
    import torch.distributed as dist
    if world_size == 1:
        return x
    out = torch.empty((world_size * x.shape[0],) + tuple(x.shape[1:]), dtype=x.dtype, device=x.device)
    dist.all_gather_into_tensor(out, x.contiguous())
    return out
