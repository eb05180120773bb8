"""Multi-GPU sampling = independent beatmaps spread over ranks, no collective on the data path
(SURVEY §8e).  The reference samples on one device only (sample.py:43); these helpers split the CFG
batch sample.py:87-108 builds — rows [0,n) conditional, [n,2n) unconditional — so that a beatmap's
two rows always land on the same rank (the guidance combine needs the pair, models.py:338-343), and
put the per-rank results back in order afterwards (result collection, off the timed path).
"""
from __future__ import annotations

import torch


def shard_range(n: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous, balanced [start, stop) of `n` beatmaps for `rank`; the first n % world ranks get one more."""
    if not 0 <= rank < world:
        raise ValueError(f"rank {rank} outside world of {world}")
    base, extra = divmod(n, world)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def shard_cfg_batch(tensors, rank: int, world: int):
    """Each tensor has 2n rows (cond | uncond). Returns this rank's (cond_r | uncond_r) rows of each."""
    out = []
    for t in tensors:
        if t.shape[0] % 2:
            raise ValueError("CFG batch must have an even number of rows")
        n = t.shape[0] // 2
        a, b = shard_range(n, rank, world)
        out.append(torch.cat([t[a:b], t[n + a:n + b]], 0))
    return out


def gather_cfg_samples(local: torch.Tensor, n: int, rank: int, world: int, group=None) -> torch.Tensor | None:
    """Reassemble per-rank sample tensors (2*n_r rows each) into the full (2n, ...) batch on rank 0."""
    import torch.distributed as dist
    if world == 1:
        return local
    parts = [None] * world
    dist.all_gather_object(parts, local.cpu(), group=group)
    if rank != 0:
        return None
    full = torch.empty((2 * n, *local.shape[1:]), dtype=local.dtype)
    for r, part in enumerate(parts):
        a, b = shard_range(n, r, world)
        full[a:b] = part[: b - a]
        full[n + a:n + b] = part[b - a:]
    return full
