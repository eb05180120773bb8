"""time the reference's eager PyTorch CUDA call sequence (oracle/eager_cuda.py) next to
the native path on the same B200, same inputs, at the BASELINE config-2 shape.  The numbers are printed
(and written to gpurun_out/eager_baseline.json when that directory exists); the assertions are parity
of the two paths and that the native one is not slower."""
import json
import os

import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import diffusion as odiff  # noqa: E402
from oracle import dit as odit  # noqa: E402
from oracle import eager_cuda  # noqa: E402
from osudit import synth  # noqa: E402

DEV = "cuda"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _time(fn, warm=2, reps=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


@torch.no_grad()
def test_native_vs_eager_torch_cuda_config2_step():
    import models
    from diffusion import create_diffusion
    n, T = 64, 2048
    shape = odit.shape_of("DiT-B")
    sd = odit.init_state_dict(shape, seed=1, zero_init_std=0.02)
    m = models.DiT_models["DiT-B"](num_classes=52670, context_size=144)
    m.load_state_dict(sd, strict=True)
    m = m.to(DEV).eval()
    sdd = {k: v.to(DEV) for k, v in sd.items()}
    one = synth.sampling_batch(1, T, seed=0)
    z, o, c, y = [torch.cat([a[:1].repeat(n, *[1] * (a.dim() - 1)), a[1:].repeat(n, *[1] * (a.dim() - 1))]).to(DEV)
                  for a in one]
    mask = synth.band_mask(T, 128).to(DEV)
    s = odiff.Schedule("100")
    d = create_diffusion("100", noise_schedule="squaredcos_cap_v2")
    t = torch.full((2 * n,), 60, device=DEV, dtype=torch.long)
    t_orig = odiff.original_timesteps(s, t)
    noise = torch.randn_like(z)

    def eager_step():
        out = eager_cuda.forward_with_cfg(sdd, shape.heads, z, t_orig, o, c, y, 1.5, mask)
        return out, odiff.p_sample(s, out, z, t, noise)["sample"]

    def native_step():
        return d.p_sample(m.forward_with_cfg, z, t, clip_denoised=True,
                          model_kwargs=dict(o=o, c=c, y=y, cfg_scale=1.5, attn_mask=mask))["sample"]

    res = {}
    old = torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32
    try:
        torch.backends.cuda.matmul.allow_tf32 = True  # sample.py:25-26
        torch.backends.cudnn.allow_tf32 = True
        ref_tf32 = eager_step()[0]
        res["eager_fp32_tf32_ms"] = _time(eager_step)
        with torch.autocast("cuda", dtype=torch.bfloat16):
            ref_bf16 = eager_step()[0].float()
            res["eager_bf16_autocast_ms"] = _time(eager_step)
        torch.backends.cuda.matmul.allow_tf32 = False
        torch.backends.cudnn.allow_tf32 = False
        ref_fp32 = eager_step()[0]
        res["eager_fp32_ms"] = _time(eager_step, warm=1, reps=2)
    finally:
        torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = old
    got = m.forward_with_cfg(z, t_orig, o=o, c=c, y=y, cfg_scale=1.5, attn_mask=mask)
    res["native_ms"] = _time(native_step)

    def rel(a, b):
        return float((a[:, :2].double() - b[:, :2].double()).norm() / b[:, :2].double().norm())

    res["eps_rel_l2_native_vs_fp32"] = rel(got, ref_fp32)
    res["eps_rel_l2_tf32_vs_fp32"] = rel(ref_tf32, ref_fp32)
    res["eps_rel_l2_bf16_autocast_vs_fp32"] = rel(ref_bf16, ref_fp32)
    for k in ("eager_fp32_tf32_ms", "eager_bf16_autocast_ms", "eager_fp32_ms", "native_ms"):
        res[k.replace("_ms", "_beatmaps_per_s")] = round(n / (res[k] * 100 / 1e3), 3)  # 100 steps per beatmap
    res["config"] = f"DiT-B, {n} beatmaps x {T} datapoints (128 rows), one CFG denoising step, band W=128"
    print(json.dumps(res))
    out_dir = os.path.join(ROOT, "gpurun_out")
    if os.path.isdir(out_dir):
        with open(os.path.join(out_dir, "eager_baseline.json"), "w") as f:
            json.dump(res, f, indent=1)
    assert res["eps_rel_l2_native_vs_fp32"] < 2e-3
    assert res["native_ms"] < res["eager_bf16_autocast_ms"]
