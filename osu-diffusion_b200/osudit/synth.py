"""Seeded synthetic beatmap sequences of the shape the reference's data pipeline emits.

There is no network for datasets or checkpoints, so benchmarks and parity tests use
these.  Layout follows ``data_loading.py:32-39`` (19-feature datapoints: x, y, time ms,
16-way one-hot type), ``data_loading.py:146-151`` (distance to the previous datapoint,
first predecessor = playfield centre) and ``data_loading.py:172-187`` (x normalised by
the playfield, o = time, c = [sincos(distance) (128) | one-hot (16)]).  CPU tensors;
SURVEY.md §8(d) fixes the distributions.
"""
from __future__ import annotations

import math

import torch

CONTEXT_SIZE = 144  # 19 - 3 + 128 (sample.py:71)
NUM_CLASSES = 52670  # train.py:331


def _sincos(v: torch.Tensor, dim: int) -> torch.Tensor:
    half = dim // 2
    freqs = torch.exp(-math.log(10000.0) * torch.arange(half, dtype=torch.float32) / half)
    a = v.reshape(-1, 1).float() * freqs[None]
    return torch.cat([a.cos(), a.sin()], -1)


def beatmap_features(T: int, seed: int = 0, training: bool = False):
    """One synthetic beatmap: x (2,T) in [0,1], o (T,) ms, c (144,T)."""
    g = torch.Generator().manual_seed(seed)
    pos = torch.rand(2, T, generator=g) * torch.tensor([[512.0], [384.0]])
    time = torch.cumsum(torch.randint(50, 400, (T,), generator=g).float(), 0)
    kind = torch.randint(0, 16, (T,), generator=g)
    prev = torch.roll(pos, 1, 1)
    prev[0, 0], prev[1, 0] = 256.0, 192.0
    dist = torch.linalg.vector_norm(pos - prev, ord=2, dim=0)
    onehot = torch.nn.functional.one_hot(kind, 16).float().t()
    x = pos / torch.tensor([[512.0], [384.0]])
    o = time - time[0]
    if training:  # data_loading.py:198-200
        o = o + torch.rand((), generator=g) * 100000
    c = torch.cat([_sincos(dist, 128).t(), onehot], 0)
    return x, o, c


def sampling_batch(n: int, T: int, seed: int = 0, num_classes: int = NUM_CLASSES,
                   null_class: bool = False, distinct_maps: bool = True):
    """The CFG batch ``sample.py:87-108`` builds: z, o, c, y of 2n rows (cond | uncond)."""
    g = torch.Generator().manual_seed(seed + 7919)
    maps = [beatmap_features(T, seed + (i if distinct_maps else 0)) for i in range(n)]
    o = torch.stack([m[1] for m in maps])
    c = torch.stack([m[2] for m in maps])
    z = torch.randn(n, 2, T, generator=g)
    if null_class:
        y = torch.full((n,), num_classes, dtype=torch.long)
    else:
        y = torch.randint(0, num_classes, (n,), generator=g)
    y_null = torch.full((n,), num_classes, dtype=torch.long)
    return (torch.cat([z, z], 0), torch.cat([o, o], 0), torch.cat([c, c], 0),
            torch.cat([y, y_null], 0))


def training_batch(B: int, T: int, seed: int = 0, num_classes: int = NUM_CLASSES):
    """(x, o, c), y as ``train.py:243-247`` receives them from the DataLoader."""
    g = torch.Generator().manual_seed(seed + 104729)
    maps = [beatmap_features(T, seed * 1000003 + i, training=True) for i in range(B)]
    x = torch.stack([m[0] for m in maps])
    o = torch.stack([m[1] for m in maps])
    c = torch.stack([m[2] for m in maps])
    y = torch.randint(0, num_classes, (B,), generator=g)
    return (x, o, c), y


def band_mask(T: int, W: int = 128) -> torch.Tensor:
    """Closed form of the loop at ``sample.py:81-84``: True = blocked."""
    d = torch.arange(T)[None, :] - torch.arange(T)[:, None]  # key - query
    return ~((d >= -(W - 1)) & (d <= W))
