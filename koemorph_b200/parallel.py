"""Clip-sharded multi-GPU inference: one process per GPU, no data-path collective, ONE output gather.

Clips are independent (per-clip dB reference, per-token LayerNorm, no batch statistics: SURVEY.md section 8e),
so a batch or corpus is split into contiguous blocks of clip indices, every rank runs the same kernels on its
block with replicated weights, and the `[n_local, T_out, 52]` results are gathered with a single
`all_gather` (NCCL over NVLink/NVSwitch on GPUs; gloo in the CPU tests of this host logic).
"""
from __future__ import annotations

from typing import Callable, List, Optional, Tuple

import torch
import torch.distributed as dist


def shard_range(n_items: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous block [lo, hi) of rank `rank`: the first n % world ranks hold one extra item."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError(f"bad rank/world {rank}/{world}")
    base, extra = divmod(n_items, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_sizes(n_items: int, world: int) -> List[int]:
    return [shard_range(n_items, r, world)[1] - shard_range(n_items, r, world)[0] for r in range(world)]


def gather_outputs(local: torch.Tensor, n_total: int, group=None) -> torch.Tensor:
    """all_gather of the per-rank result blocks (dim 0 = clips) into the full `[n_total, ...]` tensor.

    Blocks may differ by one row; they are padded to the largest block for the collective."""
    if not (dist.is_available() and dist.is_initialized()):
        if local.shape[0] != n_total:
            raise ValueError("no process group: the local block must be the whole result")
        return local
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    sizes = shard_sizes(n_total, world)
    if local.shape[0] != sizes[rank]:
        raise ValueError(f"rank {rank} holds {local.shape[0]} rows, expected {sizes[rank]}")
    width = max(sizes)
    if local.shape[0] < width:
        pad = torch.zeros((width - local.shape[0],) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
        local = torch.cat([local, pad], dim=0)
    parts = [torch.empty_like(local) for _ in range(world)]
    dist.all_gather(parts, local.contiguous(), group=group)
    return torch.cat([p[:s] for p, s in zip(parts, sizes)], dim=0)


class ShardedInference:
    """Run `forward(audio_block, egemaps_block) -> (n_local, T_out, 52)` on this rank's clips and gather."""

    def __init__(self, forward: Callable[[torch.Tensor, torch.Tensor], torch.Tensor], group=None):
        self.forward, self.group = forward, group

    def rank_world(self) -> Tuple[int, int]:
        if dist.is_available() and dist.is_initialized():
            return dist.get_rank(self.group), dist.get_world_size(self.group)
        return 0, 1

    def __call__(self, audio: torch.Tensor, egemaps: torch.Tensor, gather: bool = True) -> torch.Tensor:
        """`audio` (N, L) / `egemaps` (N, 264) hold the WHOLE batch on every rank (or at least this rank's block
        at its global position); only rows [lo, hi) are read here."""
        rank, world = self.rank_world()
        n = audio.shape[0]
        lo, hi = shard_range(n, rank, world)
        local = self.forward(audio[lo:hi].contiguous(), egemaps[lo:hi].contiguous())
        return gather_outputs(local, n, self.group) if gather else local
