"""Data-parallel plumbing for the trainer and the sharded sampler (one process per GPU, ``torch.distributed``).

The reference has no multi-GPU code at all; this is the exchange step the data-parallel
training path needs: ONE gradient all-reduce per optimiser step over the flat gradient buffer, a scalar
all-reduce so every rank takes the same early-stopping decision, and a broadcast of the initial weights.
Sampling / log_prob shard by rows with no collective on the data path.  Works with ``nccl`` (GPU) and
``gloo`` (the CPU tests).
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch
import torch.distributed as dist


def is_dist() -> bool:
    return dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1


def rank_world() -> Tuple[int, int]:
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard_indices(idx: torch.Tensor, rank: Optional[int] = None, world: Optional[int] = None) -> torch.Tensor:
    """Rank ``r`` of ``W`` gets ``idx[r::W]`` truncated so every rank holds the same count (drop_last semantics)."""
    r, w = rank_world()
    rank = r if rank is None else rank
    world = w if world is None else world
    per = idx.numel() // world
    return idx[rank::world][:per]


def shard_rows(n: int, rank: Optional[int] = None, world: Optional[int] = None) -> Tuple[int, int]:
    """Contiguous row range [lo, hi) of rank ``r`` for ``n`` independent rows (sampling / log_prob)."""
    r, w = rank_world()
    rank = r if rank is None else rank
    world = w if world is None else world
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def allreduce_sum_(t: torch.Tensor) -> torch.Tensor:
    if is_dist():
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t


def broadcast_(t: torch.Tensor, src: int = 0) -> torch.Tensor:
    if is_dist():
        dist.broadcast(t, src=src)
    return t


def agree_sum(value: float, count: float, device) -> Tuple[float, float]:
    """Sum (value, count) over ranks -- used for the validation log-prob so all ranks stop together."""
    if not is_dist():
        return value, count
    t = torch.tensor([value, count], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t[0].item()), float(t[1].item())


def shared_seed(seed: Optional[int], device) -> int:
    """Every rank must draw the same train/validation split: rank 0's seed wins."""
    if seed is None:
        seed = int(torch.randint(0, 2 ** 31 - 1, (1,)).item())
    if is_dist():
        t = torch.tensor([seed], dtype=torch.int64, device=device)
        dist.broadcast(t, src=0)
        seed = int(t.item())
    return seed


def gather_rows(t: torch.Tensor) -> torch.Tensor:
    """All-gather variable-length row shards (end of a sharded sample / log_prob call)."""
    if not is_dist():
        return t
    w = dist.get_world_size()
    n = torch.tensor([t.shape[0]], dtype=torch.int64, device=t.device)
    sizes = [torch.zeros_like(n) for _ in range(w)]
    dist.all_gather(sizes, n)
    mx = int(max(int(s.item()) for s in sizes))
    pad = torch.zeros((mx,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
    pad[: t.shape[0]] = t
    outs = [torch.zeros_like(pad) for _ in range(w)]
    dist.all_gather(outs, pad)
    return torch.cat([o[: int(s.item())] for o, s in zip(outs, sizes)])
