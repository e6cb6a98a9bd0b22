"""Batch-parallel inference across the GPUs of one box.

The reference has no multi-GPU inference (README.md:39 To-Do); images are independent units, so the
batch is split over ranks with replicated weights and NO collective inside the denoising loop; the only
exchange is one all-gather of the final packed latents at the sampler boundary (NCCL over NVLink;
`gloo` in the CPU tests).
"""
from __future__ import annotations

from typing import Tuple

import torch
import torch.distributed as dist


def shard_bounds(total: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous, balanced split: the first `total % world` ranks get one extra item."""
    base, rem = divmod(total, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_batch(t: torch.Tensor, rank: int = None, world: int = None) -> torch.Tensor:
    if rank is None:
        rank, world = (dist.get_rank(), dist.get_world_size()) if dist.is_initialized() else (0, 1)
    lo, hi = shard_bounds(t.shape[0], rank, world)
    return t[lo:hi]


def gather_latents(local: torch.Tensor, total: int) -> torch.Tensor:
    """All-gather of per-rank final latents [b_r, tokens, 64] into [total, tokens, 64] (rank order).
    Ragged shards (total % world != 0) are padded to the largest shard for the collective."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return local
    world = dist.get_world_size()
    sizes = [shard_bounds(total, r, world) for r in range(world)]
    mx = max(hi - lo for lo, hi in sizes)
    pad = local
    if local.shape[0] < mx:
        pad = torch.cat([local, local.new_zeros((mx - local.shape[0], *local.shape[1:]))], 0)
    out = local.new_empty((world * mx, *local.shape[1:]))
    dist.all_gather_into_tensor(out, pad.contiguous())
    return torch.cat([out[r * mx: r * mx + (hi - lo)] for r, (lo, hi) in enumerate(sizes)], 0)


def parallel_context(enabled: bool = True) -> Tuple[int, int]:
    """(rank, world) the pipelines shard a call over: the default process group when `torch.distributed` is initialised
    (and sharding was not switched off with `pipe.enable_batch_parallel(False)`), else (0, 1)."""
    if enabled and dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def init_from_env(backend: str = None):
    """Under `torchrun` (WORLD_SIZE > 1): bind this process to GPU LOCAL_RANK and initialise the default process group
    (NCCL on GPUs, gloo otherwise; MASTER_ADDR defaults to 127.0.0.1 — one box). Returns (rank, world, device string).
    A plain `python inference_*.py` run is (0, 1, 'cuda')."""
    import os
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world <= 1:
        return 0, 1, "cuda"
    local = int(os.environ.get("LOCAL_RANK", "0"))
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    os.environ.setdefault("MASTER_PORT", "29500")
    cuda = torch.cuda.is_available()
    if cuda:
        torch.cuda.set_device(local)
    if not dist.is_initialized():
        if cuda:
            dist.init_process_group(backend or "nccl", device_id=torch.device("cuda", local))
        else:
            dist.init_process_group(backend or "gloo")
    return dist.get_rank(), world, (f"cuda:{local}" if cuda else "cpu")
