"""Generate the golden fixtures in this directory from the UNMODIFIED reference.

Run in the build container only (``/root/reference`` does not exist on the GPU box):

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden.py

It imports ``models`` / ``diffusion`` from ``/root/reference`` (never copies them),
feeds them seeded inputs and writes small ``.npz`` / ``.json`` files that pin
``oracle/`` (tests/test_oracle_golden.py).  The reference has no tests or vectors
of its own, so these outputs of the reference itself are the pin.
"""
import importlib.util
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = os.environ.get("OSU_DIFFUSION_REF", "/root/reference")
sys.path.insert(0, REF)
sys.path.insert(1, ROOT)

import models as ref_models  # noqa: E402  (reference)
from diffusion import create_diffusion  # noqa: E402  (reference)
import diffusion.gaussian_diffusion as ref_gd  # noqa: E402

from oracle import dit as odit  # noqa: E402


def _load(name, path):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


synth = _load("synth", os.path.join(ROOT, "osu-diffusion_b200", "osudit", "synth.py"))


def npz(name, **arrs):
    out = {k: (v.detach().cpu().numpy() if isinstance(v, torch.Tensor) else np.asarray(v))
           for k, v in arrs.items()}
    np.savez_compressed(os.path.join(HERE, name), **out)
    print(name, os.path.getsize(os.path.join(HERE, name)) // 1024, "KiB")


def tiny_model():
    shape = odit.DiTShape(depth=2, hidden=32, heads=2, num_classes=10)
    sd = odit.init_state_dict(shape, seed=3, zero_init_std=0.2)
    for k in sd:  # non-zero biases so bias handling is exercised
        if k.endswith("bias") and sd[k].abs().sum() == 0:
            sd[k] = torch.randn(sd[k].shape, generator=torch.Generator().manual_seed(len(k))) * 0.05
    m = ref_models.DiT(hidden_size=32, depth=2, num_heads=2, num_classes=10, context_size=144,
                       class_dropout_prob=0.1)
    m.load_state_dict(sd, strict=True)
    return shape, sd, m.eval()


@torch.no_grad()
def golden_tiny_forward():
    shape, sd, m = tiny_model()
    T, n = 48, 2
    z, o, c, y = synth.sampling_batch(n, T, seed=5, num_classes=10)
    t = torch.tensor([999, 505, 10, 0])
    mask = synth.band_mask(T, W=8)
    # cross-check the closed-form mask against the reference's loop (sample.py:81-84)
    loop = torch.full((T, T), True)
    for i in range(T):
        loop[max(0, i - 8): min(T, i + 8), i] = False
    assert torch.equal(loop, mask)
    x = torch.randn(2 * n, 2, T, generator=torch.Generator().manual_seed(11))
    out_nomask = m(x, t, o=o, c=c, y=y)
    out_mask = m(x, t, o=o, c=c, y=y, attn_mask=mask)
    out_cfg = m.forward_with_cfg(x, t, o=o, c=c, y=y, cfg_scale=1.5, attn_mask=mask)
    npz("tiny_forward.npz", x=x, t=t, o=o, c=c, y=y, mask=mask, out_nomask=out_nomask,
        out_mask=out_mask, out_cfg=out_cfg, **{"sd." + k: v for k, v in sd.items()})


def golden_schedule():
    out = {}
    for tag, resp, sched in (("c100", "100", "squaredcos_cap_v2"), ("c250", "250", "squaredcos_cap_v2"),
                             ("c1000", "", "squaredcos_cap_v2"), ("l50", "50", "linear"),
                             ("c10_20", "10,20", "squaredcos_cap_v2")):
        d = create_diffusion(resp, noise_schedule=sched)
        out[tag + ".timestep_map"] = np.array(d.timestep_map)
        for name in ("betas", "alphas_cumprod", "sqrt_recip_alphas_cumprod",
                     "sqrt_recipm1_alphas_cumprod", "posterior_log_variance_clipped",
                     "posterior_mean_coef1", "posterior_mean_coef2", "sqrt_alphas_cumprod",
                     "sqrt_one_minus_alphas_cumprod", "posterior_variance"):
            out[tag + "." + name] = getattr(d, name)
    npz("schedule.npz", **out)


def golden_tiny_sampling():
    shape, sd, m = tiny_model()
    T, n = 48, 2
    z, o, c, y = synth.sampling_batch(n, T, seed=6, num_classes=10)
    mask = synth.band_mask(T, W=8)
    d = create_diffusion("10", noise_schedule="squaredcos_cap_v2")
    noises, real = [], torch.randn_like

    def rec(x):
        nz = real(x)
        noises.append(nz)
        return nz

    ref_gd.th.randn_like = rec
    try:
        torch.manual_seed(9)
        steps = list(d.p_sample_loop_progressive(
            m.forward_with_cfg, z.shape, z, clip_denoised=True,
            model_kwargs=dict(o=o, c=c, y=y, cfg_scale=1.5, attn_mask=mask), device="cpu"))
    finally:
        ref_gd.th.randn_like = real
    # noises[k] was drawn at respaced index K-1-k
    npz("tiny_sampling.npz", z=z, o=o, c=c, y=y, mask=mask,
        noises=torch.stack(noises[::-1]),  # index i -> noise used at respaced step i
        samples=torch.stack([s["sample"] for s in steps][::-1]),
        pred_xstart=torch.stack([s["pred_xstart"] for s in steps][::-1]))
    # in-paint callback variant (testing/test_toy.py:56-69): denoised_fn applied before the clamp
    keep = torch.zeros(1, 2, T, dtype=torch.bool)
    keep[..., : T // 2] = True
    target = z[:, :, :] * 0.1

    def in_paint(x0):
        return torch.where(keep, target, x0)

    torch.manual_seed(10)
    t = torch.full((2 * n,), 4)
    one = d.p_sample(m, z, t, clip_denoised=True, denoised_fn=in_paint,
                     model_kwargs=dict(o=o, c=c, y=y, attn_mask=None))
    torch.manual_seed(10)
    nz = torch.randn_like(z)
    npz("tiny_psample_inpaint.npz", z=z, t=t, keep=keep, target=target, noise=nz,
        sample=one["sample"], pred_xstart=one["pred_xstart"])


def golden_tiny_training():
    shape, sd, m = tiny_model()
    m.eval()  # no label drop: the drop is an RNG draw, tested separately
    T, B = 32, 4
    (x, o, c), y = synth.training_batch(B, T, seed=2, num_classes=10)
    g = torch.Generator().manual_seed(4)
    noise = torch.randn(B, 2, T, generator=g)
    t = torch.tensor([0, 1, 500, 999])
    out = {}
    for tag, l1 in (("l1", True), ("mse", False)):
        d = create_diffusion("", noise_schedule="squaredcos_cap_v2", use_l1=l1)
        m.zero_grad()
        terms = d.training_losses(m, x, t, dict(o=o, c=c, y=y), noise=noise)
        terms["loss"].mean().backward()
        for k, v in terms.items():
            out[f"{tag}.{k}"] = v.detach()
        out[f"{tag}.grad.final_w"] = m.final_layer.linear.weight.grad.clone()
        out[f"{tag}.grad.qkv0"] = m.blocks[0].attn.in_proj_weight.grad.clone()
        out[f"{tag}.grad.first_w"] = m.xoc_embedder.mlp[0].weight.grad.clone()
    npz("tiny_training.npz", x=x, o=o, c=c, y=y, t=t, noise=noise, **out)


@torch.no_grad()
def golden_registry():
    """Key names/shapes/order of every registry entry + a DiT-S forward on seeded weights."""
    layout = {}
    for name, ctor in ref_models.DiT_models.items():
        with torch.device("meta"):
            m = ctor(num_classes=52670, context_size=144)
        layout[name] = [[k, list(v.shape)] for k, v in m.state_dict().items()]
        layout[name + ".param_order"] = [k for k, _ in m.named_parameters()]
        layout[name + ".num_heads"] = m.num_heads
    with open(os.path.join(HERE, "registry_layout.json"), "w") as f:
        json.dump(layout, f)
    shape = odit.shape_of("DiT-S")
    sd = odit.init_state_dict(shape, seed=1)
    m = ref_models.DiT_models["DiT-S"](num_classes=52670, context_size=144)
    m.load_state_dict(sd, strict=True)
    m.eval()
    T, n = 256, 1
    z, o, c, y = synth.sampling_batch(n, T, seed=0)
    mask = synth.band_mask(T, 128)
    t = torch.tensor([505, 505])
    out = m.forward_with_cfg(z, t, o=o, c=c, y=y, cfg_scale=1.5, attn_mask=mask)
    npz("dit_s_forward.npz", out=out, t=t)


if __name__ == "__main__":
    torch.set_num_threads(8)
    golden_tiny_forward()
    golden_schedule()
    golden_tiny_sampling()
    golden_tiny_training()
    golden_registry()
