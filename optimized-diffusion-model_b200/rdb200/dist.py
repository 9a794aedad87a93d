"""One process per GPU: the sampler shards by batch with NO collective inside the loop (samples are
independent; the Langevin step size uses the rank-local batch mean exactly like the reference's
per-rank snapshot sampling, run_train.py:124-127) and finishes with a single all-gather of the
samples (NCCL over NVLink on the GPU box; gloo in the CPU tests)."""
from __future__ import annotations

from typing import Tuple

import torch
import torch.distributed as dist


def shard_range(B: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous [lo, hi) slice of the global batch owned by `rank` (remainder spread over low ranks)."""
    base, rem = divmod(B, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def philox_seed_for_rank(seed: int, rank: int) -> int:
    """Distinct, reproducible Philox key per rank (SplitMix64 step of seed + rank)."""
    z = (seed + 0x9E3779B97F4A7C15 * (rank + 1)) & 0xFFFFFFFFFFFFFFFF
    z = ((z ^ (z >> 30)) * 0xBF58476D1CE4E5B9) & 0xFFFFFFFFFFFFFFFF
    z = ((z ^ (z >> 27)) * 0x94D049BB133111EB) & 0xFFFFFFFFFFFFFFFF
    return (z ^ (z >> 31)) & 0x3FFFFFFFFFFFFFFF


def all_gather_batch(local: torch.Tensor, B: int) -> torch.Tensor:
    """Concatenate the per-rank shards [B_r, ...] into [B, ...] on every rank (one collective)."""
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size() == 1:
        return local
    world = dist.get_world_size()
    sizes = [shard_range(B, r, world)[1] - shard_range(B, r, world)[0] for r in range(world)]
    if len(set(sizes)) == 1:
        out = torch.empty((B,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
        dist.all_gather_into_tensor(out, local.contiguous())
        return out
    parts = [torch.empty((s,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device) for s in sizes]
    dist.all_gather(parts, local.contiguous())
    return torch.cat(parts, dim=0)
