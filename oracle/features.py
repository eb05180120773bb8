"""TEST INFRASTRUCTURE ONLY — CPU restatement of the reference's beatmap feature builder.

Reference: data_loading.py:146-151 (calc_distances), :172-187 (split_and_process_sequence_no_augment),
:195-203 (window_and_relative_time), sample.py:64-65 (relative time when sampling), positional_embedding.py:29-49
(timestep_embedding of the distances).  Pinned by tests/golden/features.npz, generated from the unmodified
reference by tests/golden/make_golden_features.py.
"""
from __future__ import annotations

import torch

from .dit import sincos

PLAYFIELD = (512.0, 384.0)


def calc_distances(seq: torch.Tensor) -> torch.Tensor:
    """Distance of every hit object to its predecessor; the first one is measured from the playfield centre."""
    prev = torch.roll(seq[:2, :], 1, 1)
    prev[0, 0] = PLAYFIELD[0] / 2
    prev[1, 0] = PLAYFIELD[1] / 2
    return torch.linalg.vector_norm(seq[:2, :] - prev, ord=2, dim=0)


def beatmap_features(seq: torch.Tensor, o_shift: float = 0.0):
    """seq (3 + n_types, T): rows x px, y px, time ms, one-hot object type.  Returns x (2,T) in playfield units,
    o (T,) = time - time[0] + o_shift, c (128 + n_types, T) = [cos | sin](dist * freqs) stacked on the one-hot."""
    d = calc_distances(seq)
    x = seq[:2, :] / torch.tensor(PLAYFIELD).unsqueeze(1)
    o = seq[2, :] - seq[2, 0] + o_shift
    c = torch.cat([sincos(d, 128).T, seq[3:, :]], 0)
    return x, o, c
