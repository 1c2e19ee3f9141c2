"""Batch sharding of the hot path across the GPUs of one box.

Every op on the path is independent per cloud, so rank g owns clouds
``[g*B/G, (g+1)*B/G)`` of every frame of a window (a 3-frame window stays on one rank
because FlowModule couples frames, discriminator.py:309-320) and there is NO data-path
collective.  The only exchanges of a data-parallel GAN step are reductions: the loss
scalars / branch flag (train_step_final.py:117) and one flat gradient bucket per network.
Backend: NCCL over NVLink on GPUs; the same code runs over gloo on CPU tensors (tests).
"""
from __future__ import annotations

from typing import Iterable, List, Sequence, Tuple

import torch
import torch.distributed as dist


def shard_range(n_items: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous, balanced [lo, hi) slice of `n_items` for `rank` (first ranks get the remainder)."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError(f"bad rank/world: {rank}/{world}")
    base, rem = divmod(n_items, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_batch(tensors: Sequence[torch.Tensor], world: int, rank: int) -> List[torch.Tensor]:
    """Slice the leading (cloud) dimension of every tensor for this rank."""
    out = []
    for t in tensors:
        lo, hi = shard_range(t.shape[0], world, rank)
        out.append(t[lo:hi].contiguous())
    return out


def _world() -> int:
    return dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1


def reduce_mean_(t: torch.Tensor, weight: float = 1.0) -> torch.Tensor:
    """In-place weighted mean over ranks: sum_r(weight_r * t_r) / sum_r(weight_r).  With
    weight = local cloud count this reproduces the single-process batch mean of
    chamferdist's batch_reduction='mean' (loss.py:176-181) under uneven shards."""
    if _world() == 1:
        return t
    packed = torch.cat([t.reshape(-1).to(torch.float32) * weight,
                        torch.tensor([weight], dtype=torch.float32, device=t.device)])
    dist.all_reduce(packed, op=dist.ReduceOp.SUM)
    t.copy_((packed[:-1] / packed[-1]).reshape(t.shape))
    return t


def agree_any(flag: torch.Tensor) -> torch.Tensor:
    """Data-dependent branches must be taken by every rank alike (train_step_final.py:117,
    upsampling_network.py:147): logical OR over ranks of a 0/1 tensor."""
    if _world() > 1:
        f = flag.to(torch.float32)
        dist.all_reduce(f, op=dist.ReduceOp.MAX)
        return f.to(flag.dtype)
    return flag


def allreduce_buckets_(buckets: Iterable[torch.Tensor]) -> None:
    """Average one flat fp32 gradient bucket per network (G / tempo-D / spatial-D)."""
    w = _world()
    if w == 1:
        return
    for b in buckets:
        dist.all_reduce(b, op=dist.ReduceOp.SUM)
        b.div_(w)
