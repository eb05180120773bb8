"""Drop-in `diffusion.respace`: `space_timesteps` and `SpacedDiffusion`
(/root/reference/diffusion/respace.py:11-132).

`SpacedDiffusion` keeps a subset of the base process' timesteps, re-derives the betas of the
shortened chain from the base alphas_cumprod, and remembers the original index of every kept
step (`timestep_map`) so the denoiser is always called with ORIGINAL timesteps.  The reference
does that mapping in a `_WrappedModel` that rebuilds a tensor from a Python list on every call;
here the map lives on the device next to the coefficient tables
(`GaussianDiffusion._tables`) and the lookup happens inside `_model_output`.
"""
from __future__ import annotations

import numpy as np

from .gaussian_diffusion import GaussianDiffusion


def space_timesteps(num_timesteps, section_counts):
    """Which original steps to keep: per section, `count` steps at a fractional stride, each
    rounded (so "100" over 1000 steps is NOT uniform: 0,10,...,50,61,71,...); "ddimN" picks the
    integer stride that yields exactly N steps.  Returns a set."""
    if isinstance(section_counts, str):
        if section_counts.startswith("ddim"):
            want = int(section_counts[len("ddim"):])
            for stride in range(1, num_timesteps):
                if len(range(0, num_timesteps, stride)) == want:
                    return set(range(0, num_timesteps, stride))
            raise ValueError(f"cannot create exactly {num_timesteps} steps with an integer stride")
        section_counts = [int(v) for v in section_counts.split(",")]
    base, extra = divmod(num_timesteps, len(section_counts))
    kept, start = [], 0
    for i, count in enumerate(section_counts):
        size = base + (1 if i < extra else 0)
        if size < count:
            raise ValueError(f"cannot divide section of {size} steps into {count}")
        stride = 1 if count <= 1 else (size - 1) / (count - 1)
        kept += [start + round(j_stride) for j_stride in _cumulative(stride, count)]
        start += size
    return set(kept)


def _cumulative(stride, count):
    # repeated addition (not j*stride): reproduces the reference's float accumulation exactly
    cur, out = 0.0, []
    for _ in range(count):
        out.append(cur)
        cur += stride
    return out


class SpacedDiffusion(GaussianDiffusion):
    def __init__(self, use_timesteps, **kwargs):
        self.use_timesteps = set(use_timesteps)
        self.original_num_steps = len(kwargs["betas"])
        base_acp = np.cumprod(1.0 - np.array(kwargs["betas"], dtype=np.float64), axis=0)
        self.timestep_map = []
        betas, last = [], 1.0
        for i, acp in enumerate(base_acp):
            if i in self.use_timesteps:
                betas.append(1 - acp / last)
                last = acp
                self.timestep_map.append(i)
        kwargs["betas"] = np.array(betas)
        super().__init__(**kwargs)

    def _timestep_map(self):
        return self.timestep_map

    def _scale_timesteps(self, t):
        return t
