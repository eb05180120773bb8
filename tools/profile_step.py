"""A few denoising steps at the BASELINE config-2 shape (DiT-B, 64 beatmaps x 2048 datapoints,
CFG, band mask) — the command profiled by ncu (profiles/README.md).  Not a benchmark."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "osu-diffusion_b200"))
import torch
import bench
from diffusion import create_diffusion
from osudit import synth

n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
dev = torch.device("cuda", 0)
model = bench.build_native(dev)
d = create_diffusion("100", noise_schedule="squaredcos_cap_v2")
one = synth.sampling_batch(1, bench.SEQ, seed=0)
z, o, c, y = [torch.cat([t[:1].repeat(n, *[1] * (t.dim() - 1)), t[1:].repeat(n, *[1] * (t.dim() - 1))]).to(dev) for t in one]
mask = synth.band_mask(bench.SEQ, bench.BAND).to(dev)
x = z
with torch.no_grad():
    for i in range(steps):
        t = torch.full((2 * n,), 99 - i, device=dev, dtype=torch.long)
        x = d.p_sample(model.forward_with_cfg, x, t, clip_denoised=True,
                       model_kwargs=dict(o=o, c=c, y=y, cfg_scale=1.5, attn_mask=mask))["sample"]
torch.cuda.synchronize()
print("ok", float(x.abs().mean()))
