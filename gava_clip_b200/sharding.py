"""Data-parallel sharding helpers: one process per GPU, clips are independent.

Mirrors what the reference's latent multi-GPU code does: contiguous/strided shards per rank (video_dataset/dataloader.py:
107-120,156-159), SUM all-reduce of the [total, hit@1] counters (training/train.py:682, evaluation/evaluate.py:296) and,
optionally, a gather of the per-rank logits.  No collective touches the forward itself.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch
import torch.distributed as dist


def shard_bounds(n_items: int, rank: int, world_size: int) -> Tuple[int, int]:
    """[lo, hi) of the clips owned by `rank`: whole clips, sizes differ by at most one, order preserved."""
    if not (0 <= rank < world_size):
        raise ValueError(f"rank {rank} outside world of {world_size}")
    base, rem = divmod(n_items, world_size)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_batch(batch: torch.Tensor, rank: Optional[int] = None, world_size: Optional[int] = None) -> torch.Tensor:
    """The slice of a (B, ...) batch this rank processes (never splits a clip's T frames)."""
    rank = dist.get_rank() if rank is None else rank
    world_size = dist.get_world_size() if world_size is None else world_size
    lo, hi = shard_bounds(batch.shape[0], rank, world_size)
    return batch[lo:hi]


def gather_logits(local_logits: torch.Tensor, n_items: int) -> torch.Tensor:
    """All ranks receive the (n_items, n_cls) logits in the original clip order (shards may be ragged)."""
    world = dist.get_world_size()
    sizes = [shard_bounds(n_items, r, world) for r in range(world)]
    width = max(hi - lo for lo, hi in sizes)
    pad = torch.zeros(width, local_logits.shape[1], dtype=local_logits.dtype, device=local_logits.device)
    pad[: local_logits.shape[0]] = local_logits
    out = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(out, pad)
    return torch.cat([o[: hi - lo] for o, (lo, hi) in zip(out, sizes)], dim=0)


def reduce_hit_counts(logits: torch.Tensor, labels: torch.Tensor) -> Tuple[int, int]:
    """(total, hit@1) summed over ranks -- the reference's evaluate() counters (train.py:676-683)."""
    stats = torch.tensor([labels.numel(), int((logits.argmax(-1) == labels).sum())], dtype=torch.int64,
                         device=logits.device)
    dist.all_reduce(stats, op=dist.ReduceOp.SUM)
    return int(stats[0]), int(stats[1])
