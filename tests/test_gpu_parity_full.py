"""Parity at the north star's own bars and at the BASELINE configurations' full sizes.

The CPU oracle is too slow beyond DiT-S / a few hundred datapoints, so the reference here is oracle/eager_cuda.py
(the reference's eager call sequence on stock torch CUDA kernels, fp32 with TF32 off; pinned to oracle/dit.py on the
CPU by tests/test_oracle_golden.py) running on the same GPU.

* fp32 mode, free-running 100 respaced steps on the damped fixture: EVERY final coordinate within 0.5 osu! px;
* bf16 mode, DiT-B, 256 datapoints under the band mask, free-running: the measured distribution (median / p99 / max),
  asserted with a 1.5x margin;
* full-depth DiT-XL (28 x 1152, head_dim 72) forward under the band mask and DiT-L (24 x 1024) seq-len-512 training
  gradients (BASELINE configs 4 and 5 per-sample shapes);
* BASELINE config 2 (128 rows x 2048 datapoints): one CFG denoising step at the first, a middle and the last
  respaced timestep.
"""
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import diffusion as odiff  # noqa: E402
from oracle import dit as odit  # noqa: E402
from oracle import eager_cuda  # noqa: E402
from osudit import synth  # noqa: E402

DEV = "cuda"


def rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


class _NoTF32:
    def __enter__(self):
        self.old = torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32
        torch.backends.cuda.matmul.allow_tf32 = torch.backends.cudnn.allow_tf32 = False

    def __exit__(self, *exc):
        torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = self.old


def _build(name=None, shape=None, seed=1, damp_x=None, std=0.02, **kw):
    import models
    shape = shape or odit.shape_of(name)
    sd = odit.init_state_dict(shape, seed=seed, zero_init_std=std, damp_x=damp_x)
    m = models.DiT(depth=shape.depth, hidden_size=shape.hidden, num_heads=shape.heads, num_classes=52670,
                   context_size=144, **kw)
    m.load_state_dict(sd, strict=True)
    return shape, sd, m.to(DEV).eval()


def _free_running(monkeypatch, name, T, W, precision, seed_noise=11):
    """100 respaced CFG steps, identical pre-drawn noise on both sides; returns per-datapoint distances in osu! px."""
    import diffusion.gaussian_diffusion as gd
    from diffusion import create_diffusion
    from osudit import graphs
    shape, sd, m = _build(name, damp_x=0.02)
    m.precision = precision
    n = 1
    z, o, c, y = [v.to(DEV) for v in synth.sampling_batch(n, T, seed=0)]
    mask = synth.band_mask(T, W).to(DEV) if W else None
    sdd = {k: v.to(DEV) for k, v in sd.items()}
    s = odiff.Schedule("100")
    d = create_diffusion("100", noise_schedule="squaredcos_cap_v2")
    g = torch.Generator().manual_seed(seed_noise)
    noises = [torch.randn(2 * n, 2, T, generator=g).to(DEV) for _ in range(100)]
    with torch.no_grad(), _NoTF32():
        x = z
        for i in reversed(range(100)):  # the reference loop, gaussian_diffusion.py:514-561
            t = torch.full((2 * n,), i, device=DEV)
            out = eager_cuda.forward_with_cfg(sdd, shape.heads, x, odiff.original_timesteps(s, t), o, c, y, 1.5, mask)
            x = odiff.p_sample(s, out, x, t, noises[i])["sample"]
        ref = x
        it = iter(noises[::-1])
        monkeypatch.setattr(graphs, "_ENABLED", False)  # a captured step draws its noise inside the graph
        monkeypatch.setattr(gd.th, "randn_like", lambda v: next(it))
        got = d.p_sample_loop(m.forward_with_cfg, z.shape, z, clip_denoised=True,
                              model_kwargs=dict(o=o, c=c, y=y, cfg_scale=1.5, attn_mask=mask), device=DEV)
    px = (got[:n] - ref[:n]).abs() * torch.tensor([512.0, 384.0], device=DEV)[None, :, None]
    return px.pow(2).sum(1).sqrt().flatten().cpu()


def test_fp32_mode_free_running_100_steps_every_coordinate_within_half_a_pixel(monkeypatch):
    """North star: "final x/y coordinates within 0.5 osu! pixels" — met strictly (max, not median) by the fp32 mode
    over a whole free-running 100-step CFG sampling on the damped fixture, band mask included."""
    dist = _free_running(monkeypatch, "DiT-S", 256, 128, "fp32")
    print(f"fp32 mode free-running px: median {dist.median():.2e} p99 {dist.quantile(0.99):.2e} max {dist.max():.2e}")
    assert float(dist.max()) < 0.5   # the north star's bar, for every coordinate
    assert float(dist.median()) < 0.01 and float(dist.max()) < 0.25  # measured: median 2e-3, p99 0.06, max 0.13 px


def test_bf16_free_running_distribution_dit_b_band_mask(monkeypatch):
    """The bf16 mode cannot promise 0.5 px for EVERY coordinate of a 100-step free-running trajectory (per-step eps
    error 1e-3, ~20 % of coordinates sitting on the clamp bounds: SURVEY A.8 measured max 1.7 px for the best bf16
    arithmetic); what it delivers on DiT-B / 256 datapoints / band mask — measured median 0.10, mean 0.27, p99 1.87,
    max 2.10 px, 16.4 % of datapoints beyond 0.5 px — is asserted at 1.5x those values."""
    dist = _free_running(monkeypatch, "DiT-B", 256, 128, "bf16")
    frac = float((dist > 0.5).float().mean())
    print(f"bf16 DiT-B T=256 band free-running px: median {dist.median():.3f} mean {dist.mean():.3f} "
          f"p99 {dist.quantile(0.99):.3f} max {dist.max():.3f} frac>0.5 {frac:.3f}")
    assert float(dist.median()) < 0.15
    assert float(dist.mean()) < 0.40
    assert float(dist.quantile(0.99)) < 2.8
    assert float(dist.max()) < 3.2
    assert frac < 0.25


@torch.no_grad()
def test_full_depth_dit_xl_forward_band_mask():
    """BASELINE config 4's model: DiT-XL, 28 blocks x 1152, 16 heads of 72, at 256 datapoints under the band mask,
    CFG rows; eps within the north star's 2e-3 of the fp32 eager reference."""
    shape, sd, m = _build("DiT-XL")
    T, n = 256, 2
    z, o, c, y = [v.to(DEV) for v in synth.sampling_batch(n, T, seed=0)]
    mask = synth.band_mask(T, 128).to(DEV)
    sdd = {k: v.to(DEV) for k, v in sd.items()}
    t = torch.tensor([505, 20, 505, 20], device=DEV)
    with _NoTF32():
        ref = eager_cuda.forward(sdd, shape.heads, z, t, o, c, y, mask)
    out = m(z, t, o=o, c=c, y=y, attn_mask=mask)
    e = rel(out[:, :2], ref[:, :2])
    print(f"DiT-XL full depth T=256 band: eps rel-L2 {e:.2e}, all {rel(out, ref):.2e}")
    assert e < 2e-3


def test_dit_l_seq_512_training_gradients():
    """BASELINE config 5's per-sample shape: DiT-L (24 x 1024, 16 heads), 512 datapoints, full attention; loss terms
    and every parameter gradient against fp32 autograd over the eager reference on the same GPU."""
    from diffusion import create_diffusion
    shape, sd, m = _build("DiT-L", std=0.05, class_dropout_prob=0.2)
    B, T = 2, 512
    (x, o, c), y = synth.training_batch(B, T, seed=3)
    g = torch.Generator().manual_seed(5)
    noise, t = torch.randn(B, 2, T, generator=g), torch.tensor([0, 640])
    x, o, c, y, noise, t = [v.to(DEV) for v in (x, o, c, y, noise, t)]
    sdg = {k: v.clone().to(DEV).requires_grad_(v.is_floating_point() and "playfield" not in k) for k, v in sd.items()}
    s = odiff.Schedule("")
    with _NoTF32():
        ref = odiff.training_losses(s, lambda x_t, tt: eager_cuda.forward(sdg, shape.heads, x_t, tt, o, c, y),
                                    x, t, noise, use_l1=True)
        ref["loss"].mean().backward()
    d = create_diffusion("", noise_schedule="squaredcos_cap_v2", use_l1=True)
    terms = d.training_losses(m, x, t, dict(o=o, c=c, y=y), noise=noise)
    torch.testing.assert_close(terms["l1"].detach(), ref["l1"].detach(), rtol=5e-3, atol=1e-4)
    terms["loss"].mean().backward()
    errs = sorted(((rel(p.grad, sdg[k].grad), k) for k, p in m.named_parameters() if p.requires_grad), reverse=True)
    print(f"DiT-L T=512: worst parameter-gradient rel-L2 {errs[0][0]:.2e} ({errs[0][1]}), median {errs[len(errs) // 2][0]:.2e}")
    assert errs[0][0] < 1e-2, errs[:3]


@pytest.mark.parametrize("step", [99, 60, 0])
@torch.no_grad()
def test_config2_size_denoising_step_parity(step):
    """BASELINE config 2 at full size (64 beatmaps x 2048 datapoints = 128 model rows, CFG 1.5, band W = 128): the raw
    eps and the x_{t-1} of one reverse step at the first, a middle and the last respaced timestep."""
    import diffusion.gaussian_diffusion as gd
    from diffusion import create_diffusion
    shape, sd, m = _build("DiT-B")
    n, T = 64, 2048
    z, o, c, y = [v.to(DEV) for v in synth.sampling_batch(n, T, seed=3, distinct_maps=False)]
    mask = synth.band_mask(T, 128).to(DEV)
    sdd = {k: v.to(DEV) for k, v in sd.items()}
    s = odiff.Schedule("100")
    d = create_diffusion("100", noise_schedule="squaredcos_cap_v2")
    t = torch.full((2 * n,), step, device=DEV, dtype=torch.long)
    t_orig = odiff.original_timesteps(s, t)
    noise = torch.randn_like(z)
    with _NoTF32():
        ref_out = eager_cuda.forward_with_cfg(sdd, shape.heads, z, t_orig, o, c, y, 1.5, mask)
    ref = odiff.p_sample(s, ref_out, z, t, noise)
    real = gd.th.randn_like
    gd.th.randn_like = lambda v: noise
    try:
        got = d.p_sample(m.forward_with_cfg, z, t, clip_denoised=True,
                         model_kwargs=dict(o=o, c=c, y=y, cfg_scale=1.5, attn_mask=mask))
    finally:
        gd.th.randn_like = real
    raw = m.forward_with_cfg(z, t_orig, o=o, c=c, y=y, cfg_scale=1.5, attn_mask=mask)
    e = rel(raw[:, :2], ref_out[:, :2])
    inside = (ref["pred_xstart"] > -0.999) & (ref["pred_xstart"] < 1.999)  # the clamp turns tiny eps errors into 0 / big
    dx = (got["sample"] - ref["sample"]).abs()
    amp = float(s.sqrt_recipm1_alphas_cumprod[step] * s.posterior_mean_coef1[step])
    tol = 4 * 2e-3 * amp * float(ref_out[:, :2].abs().max()) + 1e-5
    frac_bad = float((dx[inside] > tol).float().mean()) if bool(inside.any()) else 0.0
    print(f"config 2, step {step}: eps rel-L2 {e:.2e}; x_(t-1) beyond the amplified tolerance {tol:.2e}: {frac_bad:.4f}")
    assert e < 2e-3
    assert frac_bad < 0.01
