"""Golden outputs of the reference's diffusion/diffusion_utils.py (normal_kl :9-35, approx_standard_normal_cdf :38-43,
discretized_gaussian_log_likelihood :63-89) on a grid that reaches every branch — x at and beyond the +-0.999 edge
bins, |x - mean| / scale from 0 to ~1e4, log-variances from -20 to 6.  UNMODIFIED reference, build container only:

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden_utils.py   ->  diffusion_utils.npz
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.environ.get("OSU_DIFFUSION_REF", "/root/reference"))
from diffusion import diffusion_utils as du  # noqa: E402  (reference)

g = torch.Generator().manual_seed(3)
x = torch.cat([torch.tensor([-1.0, -0.9995, -0.999, -0.5, 0.0, 0.5, 0.999, 0.9995, 1.0, 2.0, -2.0]),
               torch.rand(117, generator=g) * 2.4 - 1.2])
mean = torch.cat([x[:11] + torch.tensor([0.0, 1e-3, -1e-3, 0.3, -4.0, 4.0, 0.0, 0.5, -0.5, 0.0, 0.1]),
                  torch.randn(117, generator=g)])
log_scale = torch.cat([torch.tensor([-10.0, -6.0, -3.0, -1.0, 0.0, 1.0, 3.0, -8.0, -5.0, -2.0, 0.5]),
                       torch.rand(117, generator=g) * 12 - 9])
m1, m2 = torch.randn(128, generator=g) * 2, torch.randn(128, generator=g) * 2
lv1 = torch.cat([torch.tensor([-20.0, -10.0, 0.0, 6.0]), torch.rand(124, generator=g) * 16 - 12])
lv2 = torch.cat([torch.tensor([6.0, -20.0, 0.0, -10.0]), torch.rand(124, generator=g) * 16 - 12])
z = torch.linspace(-12, 12, 97)
out = dict(x=x, mean=mean, log_scale=log_scale, m1=m1, m2=m2, lv1=lv1, lv2=lv2, z=z,
           loglik=du.discretized_gaussian_log_likelihood(x, means=mean, log_scales=log_scale),
           kl=du.normal_kl(m1, lv1, m2, lv2), cdf=du.approx_standard_normal_cdf(z))
np.savez_compressed(os.path.join(HERE, "diffusion_utils.npz"), **{k: v.numpy() for k, v in out.items()})
print({k: (float(v.min()), float(v.max())) for k, v in out.items() if k in ("loglik", "kl", "cdf")})
