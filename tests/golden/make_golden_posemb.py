"""Golden outputs of the reference's positional_embedding.py:29-77 (UNMODIFIED reference, build container only):

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden_posemb.py   ->  posemb.npz
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.environ.get("OSU_DIFFUSION_REF", "/root/reference"))
import positional_embedding as ref  # noqa: E402  (reference)

g = torch.Generator().manual_seed(7)
t = torch.cat([torch.tensor([0.0, 1.0, 999.0]), torch.rand(5, generator=g) * 999])
o = torch.cumsum(torch.randint(50, 400, (2, 6), generator=g).float(), 1) + 12345.0
p = torch.rand(2, 6, 2, generator=g) * torch.tensor([512.0, 384.0])
out = {"t": t, "o": o, "p": p}
for dim in (256, 128, 9, 2):
    out[f"timestep_{dim}"] = ref.timestep_embedding(t, dim)
out["timestep_128_mp100"] = ref.timestep_embedding(t, 128, max_period=100)
out["offset_128"] = ref.offset_sequence_embedding(o / 10, 128)
out["position_128"] = ref.position_sequence_embedding(p, 128)
np.savez_compressed(os.path.join(HERE, "posemb.npz"), **{k: v.numpy() for k, v in out.items()})
print({k: tuple(v.shape) for k, v in out.items()})
