"""Device-side counterparts of the reference's feature-building functions (data_loading.py:146-203), SURVEY §8(f)2.

The reference builds `c (144, T)` per beatmap on the host (sin/cos embedding of the distances + one-hot types) and
ships it over PCIe with every batch (151 MB for 64 beatmaps x 2048).  Here the raw hit-object sequence (19 floats per
object) is uploaded instead and `libosudit.so` expands it on the GPU.
"""
from __future__ import annotations

import math

import torch

from . import _lib, ops

_freqs = {}


def _freqs64(device):
    key = str(device)
    if key not in _freqs:  # positional_embedding.py:39-44, evaluated on the host as the reference does
        f = torch.exp(-math.log(10000) * torch.arange(start=0, end=64, dtype=torch.float32) / 64)
        _freqs[key] = f.to(device)
    return _freqs[key]


def beatmap_features(seq, o_shift=None, want_x=True):
    """seq fp32 CUDA [B, R, T] (or [R, T]): rows x px, y px, time ms, one-hot type.  Returns (x, o, c) with
    x [B, 2, T] = pos / playfield (None when want_x is False), o [B, T] = time - time[0] (+ o_shift[b]),
    c [B, 128 + R - 3, T]."""
    single = seq.dim() == 2
    if single:
        seq = seq.unsqueeze(0)
    B, R, T = seq.shape
    dev = seq.device
    x = torch.empty(B, 2, T, device=dev) if want_x else None
    o = torch.empty(B, T, device=dev)
    c = torch.empty(B, 128 + R - 3, T, device=dev)
    lib = _lib.load()
    _lib.check(lib.osudit_beatmap_features(
        ops._chk(seq, torch.float32, "features.seq"), B, R, T, _freqs64(dev).data_ptr(),
        ops._chk(o_shift, torch.float32, "features.o_shift") if o_shift is not None else None,
        x.data_ptr() if want_x else None, o.data_ptr(), c.data_ptr(), ops._stream()), "osudit_beatmap_features")
    if single:
        return (x[0] if want_x else None), o[0], c[0]
    return x, o, c


def calc_distances(seq):
    """data_loading.py:146-151 on the device (small: plain strided arithmetic is enough here)."""
    prev = torch.roll(seq[:2, :], 1, 1)
    prev[0, 0], prev[1, 0] = 256.0, 192.0
    return torch.linalg.vector_norm(seq[:2, :] - prev, ord=2, dim=0)


def split_and_process_sequence_no_augment(seq):
    """Same name, argument and return value as data_loading.py:172-187 for a CUDA `seq (19, T)`:
    ((x (2,T), o (T,) absolute time, c (144,T)), T)."""
    x, _, c = beatmap_features(seq.float().contiguous())
    return (x, seq[2, :], c), seq.shape[1]
