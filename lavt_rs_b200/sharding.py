"""Clip sharding across the GPUs of one box (reference test_ytvos.py:112-137: contiguous slices of the sorted
video list, one process per GPU, no communication).  Inference has no data-path collective; the only exchange is
the max-over-ranks reduction of the step time used by bench.py."""
from __future__ import annotations

from typing import List, Tuple

import torch


def clip_slice(n_clips: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous [start, end) slice of clip indices owned by ``rank`` (sizes differ by at most one)."""
    if not (0 <= rank < world):
        raise ValueError(f"rank {rank} outside world {world}")
    base, rem = divmod(n_clips, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def all_slices(n_clips: int, world: int) -> List[Tuple[int, int]]:
    return [clip_slice(n_clips, r, world) for r in range(world)]


def max_over_ranks(value: float, device=None) -> float:
    """Slowest rank's value (device-side timings are reported as the max over ranks)."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device if device is not None else "cpu")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
