"""Sample-parallel execution of Box2Video sampling: one process per GPU, clips sharded across
ranks, no data-path collective during sampling, one all-gather of the final latents.

The reference has no multi-GPU inference at all (every eval tool runs under
`if accelerator.is_main_process:`, /root/reference/tools/eval_video_controlnet.py:109); the unit
of work here is a whole clip (both CFG branches stay on one GPU, so the diffusers-0.27.2
`time_context` coupling of the two branches, SURVEY.md A.5, is preserved exactly)."""
from __future__ import annotations

import os
from typing import List, Sequence

import torch
import torch.distributed as dist


def shard_clips(n_clips: int, world_size: int, rank: int) -> List[int]:
    """Contiguous block partition of clip indices; the first (n_clips % world_size) ranks get one
    extra clip.  Deterministic, so every rank can reconstruct the global order."""
    if world_size <= 0 or not (0 <= rank < world_size):
        raise ValueError(f"bad rank/world_size {rank}/{world_size}")
    base, rem = divmod(n_clips, world_size)
    start = rank * base + min(rank, rem)
    return list(range(start, start + base + (1 if rank < rem else 0)))


def init_from_env(backend: str | None = None) -> tuple[int, int, int]:
    """Read RANK / WORLD_SIZE / LOCAL_RANK (torchrun) and initialise torch.distributed if needed."""
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        if backend == "nccl":
            torch.cuda.set_device(local)
        dist.init_process_group(backend=backend, rank=rank, world_size=world)
    return rank, world, local


def gather_latents(local: torch.Tensor, n_clips: int, world_size: int, rank: int) -> torch.Tensor:
    """All-gather per-rank final latents [n_local, T, 4, h, w] into the global [n_clips, ...] order.
    Ranks may hold different clip counts; shorter shards are padded for the collective."""
    if world_size == 1:
        return local
    counts = [len(shard_clips(n_clips, world_size, r)) for r in range(world_size)]
    cmax = max(counts)
    pad = local
    if local.shape[0] < cmax:
        pad = torch.cat([local, local.new_zeros((cmax - local.shape[0],) + tuple(local.shape[1:]))])
    bufs = [torch.empty_like(pad) for _ in range(world_size)]
    dist.all_gather(bufs, pad.contiguous())
    return torch.cat([b[:c] for b, c in zip(bufs, counts)])
