"""Sample-parallel execution of Box2Video sampling: one process per GPU, clips sharded across
ranks, no data-path collective during sampling, one all-gather of the final latents.  When there
are fewer clips than GPUs, `CfgPair` shards the two CFG branches of a clip over a GPU pair
(uncond on rank 2k, cond on rank 2k+1) with one small exchange of the model output per step.

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


# ---- CFG-branch sharding: one clip on a pair of GPUs ---------------------------------
def cfg_pair_of(rank: int) -> tuple[int, int]:
    """(pair index, CFG branch) of a rank: ranks 2k / 2k+1 run the uncond / cond half of pair k."""
    if rank < 0:
        raise ValueError(f"bad rank {rank}")
    return rank // 2, rank % 2


def time_context_rotation(branch: int, local_batch: int, sites: int) -> int:
    """Rows by which a branch-sharded process rotates the whole-batch context table.

    diffusers 0.27.2 builds `time_context` S-major (Appendix A.5): hidden row (b, s) of the batch of
    2B samples is paired with context (b*S + s) % 2B.  A process that owns global samples
    b0 = branch*B .. b0+B-1 sees local index (b_local*S + s), i.e. the global index minus b0*S."""
    return (branch * local_batch * sites) % (2 * local_batch)


class CfgPair:
    """Process-group plumbing of one GPU pair: `exchange(step)` all-gathers the two halves of the model
    output (fp32 [B*T*h*w, 4] per rank, 573 KB at 14x40x64) into `step.noise` over NVLink."""

    def __init__(self, rank: int, world_size: int):
        if world_size % 2:
            raise ValueError(f"CFG-branch sharding needs an even number of ranks, got {world_size}")
        self.pair, self.branch = cfg_pair_of(rank)
        self.n_pairs = world_size // 2
        self.group = None
        if world_size > 2:
            for k in range(self.n_pairs):  # every rank must create every group, in the same order
                grp = dist.new_group([2 * k, 2 * k + 1])
                if k == self.pair:
                    self.group = grp

    def exchange(self, step) -> None:
        dist.all_gather_into_tensor(step.noise, step.noise_local, group=self.group)
