"""Diagnostic: per-step loss of the native training path with CUDA-graph replay vs eager launches vs a second
eager run (the run-to-run noise floor from atomic ordering), same init / data / AdamW."""
import copy
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "osu-diffusion_b200"))

import models  # noqa: E402
from diffusion import create_diffusion  # noqa: E402
from osudit import synth, train as otrain  # noqa: E402

dev = "cuda"
B, T, steps = 16, 128, 40
torch.manual_seed(0)
base = models.DiT_models["DiT-S"](num_classes=52670, context_size=144)
with torch.no_grad():
    for k, v in base.state_dict().items():
        if "adaLN_modulation" in k or k.startswith("final_layer.linear"):
            v.normal_(0, 0.05)
base = base.to(dev).eval()
d = create_diffusion("", noise_schedule="squaredcos_cap_v2", use_l1=True)
batches = []
for i in range(4):
    (x, o, c), y = synth.training_batch(B, T, seed=20 + i)
    batches.append([v.to(dev) for v in (x, o, c, y)])


def run(enabled):
    otrain._GRAPHS_ENABLED = enabled
    otrain._train_graphs.clear()
    m = copy.deepcopy(base)
    opt = torch.optim.AdamW(m.parameters(), lr=1e-4, weight_decay=0)
    g = torch.Generator().manual_seed(9)
    out = []
    for it in range(steps):
        x, o, c, y = batches[it % 4]
        t = torch.randint(0, 1000, (B,), generator=g).to(dev)
        noise = torch.randn(B, 2, T, generator=g).to(dev)
        loss = d.training_losses(m, x, t, dict(o=o, c=c, y=y), noise=noise)["loss"].mean()
        loss.backward()
        opt.step()
        opt.zero_grad(set_to_none=True)
        out.append(float(loss))
    return torch.tensor(out, dtype=torch.float64)


e1, e2, g1, g2 = run(False), run(False), run(True), run(True)
rel = lambda a, b: ((a - b).abs() / b.abs())  # noqa: E731
for name, v in (("eager vs eager", rel(e2, e1)), ("graph vs eager", rel(g1, e1)), ("graph vs graph", rel(g2, g1))):
    print(f"{name}: per-step rel loss diff at 0,1,2,3,5,10,20,39: " + " ".join(f"{float(v[i]):.1e}" for i in (0, 1, 2, 3, 5, 10, 20, 39))
          + f"  max {float(v.max()):.1e}")
