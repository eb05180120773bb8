"""Every kernel family once at small shapes, for `compute-sanitizer --tool memcheck|racecheck python tools/sanitize_small.py`
(the reference has no race / memory checking).  CUDA graphs are off so that the sanitizer sees plain launches.

  part "sample": DiT-S, 2 beatmaps x 384 datapoints with the band mask (window attention kernel), one CFG denoising step,
                 then the same in fp32 mode and with a generic (non-band) mask on a ragged length
  part "gemm":   the CTA-pair GEMM in both tile widths and all four epilogues at the smallest M that selects it
  part "train":  DiT-S, 4 x 128 datapoints: forward, loss, native backward, fused AdamW + EMA
"""
import math
import os
import sys

os.environ["OSUDIT_CUDA_GRAPHS"] = "0"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "osu-diffusion_b200"))
import torch  # noqa: E402

import models  # noqa: E402
from diffusion import create_diffusion  # noqa: E402
from osudit import ops, synth  # noqa: E402
from osudit.optim import FusedAdamWEMA  # noqa: E402

dev = torch.device("cuda", 0)
parts = sys.argv[1:] or ["sample", "gemm", "train"]


def model(train=False):
    m = models.DiT_models["DiT-S"](num_classes=52670, context_size=144, class_dropout_prob=0.2 if train else 0.1)
    with torch.no_grad():
        for k, v in m.state_dict().items():
            if "adaLN_modulation" in k or k.startswith("final_layer.linear"):
                v.normal_(0, 0.02)
    return m.to(dev)


if "sample" in parts:
    m = model().eval()
    d = create_diffusion("100", noise_schedule="squaredcos_cap_v2")
    with torch.no_grad():
        for T, mask, prec in ((384, synth.band_mask(384, 128), "bf16"), (300, torch.rand(300, 300) < 0.3, "bf16"),
                              (256, synth.band_mask(256, 128), "fp32")):
            m.precision = prec
            mask = mask.clone()
            mask.fill_diagonal_(False)
            z, o, c, y = [v.to(dev) for v in synth.sampling_batch(2, T, seed=0)]
            out = d.p_sample(m.forward_with_cfg, z, torch.full((4,), 60, device=dev), clip_denoised=True,
                             model_kwargs=dict(o=o, c=c, y=y, cfg_scale=1.5, attn_mask=mask.to(dev)))["sample"]
            torch.cuda.synchronize()
            assert bool(torch.isfinite(out).all())
            print(f"sample T={T} {prec}: ok", flush=True)

if "gemm" in parts:
    M, K = 256 * 37, 128
    for N in (256, 192):
        a = torch.randn(M, K, device=dev).to(torch.bfloat16)
        w = (torch.randn(N, K, device=dev) / math.sqrt(K)).to(torch.bfloat16)
        bias = torch.randn(N, device=dev)
        out, aux = (torch.empty(M, N, device=dev, dtype=torch.bfloat16) for _ in range(2))
        ops.gemm([a], [w], bias, ops.EPI_BF16, out)
        ops.gemm([a], [w], bias, ops.EPI_BF16_GELU, out)
        ops.gemm_aux(a, w, bias, ops.EPI_BF16_GELU_SAVE, out, aux)
        ops.gemm_aux(a, w, None, ops.EPI_BF16_DGELU, out, aux)
        torch.cuda.synchronize()
        print(f"gemm pair-tile N={N}: ok", flush=True)

if "train" in parts:
    from copy import deepcopy
    m = model(train=True).train()
    ema = deepcopy(m).requires_grad_(False)
    d = create_diffusion("", noise_schedule="squaredcos_cap_v2", use_l1=True)
    opt = FusedAdamWEMA(m.parameters(), lr=1e-4, weight_decay=0)
    opt.attach_ema(ema, m)
    (x, o, c), y = synth.training_batch(4, 128, seed=0)
    x, o, c, y = [v.to(dev) for v in (x, o, c, y)]
    for _ in range(2):
        t = torch.randint(0, 1000, (4,), device=dev)
        loss = d.training_losses(m, x, t, dict(o=o, c=c, y=y))["loss"].mean()
        loss.backward()
        opt.step()
        opt.zero_grad(set_to_none=True)
    torch.cuda.synchronize()
    assert math.isfinite(float(loss))
    print("train: ok", flush=True)
