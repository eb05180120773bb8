"""Three DiT-B training steps at the BASELINE config-3 shape (batch 256 x 128 datapoints, fused optimizer) with CUDA
graphs off, so that ncu sees individual launches — the command behind profiles/r01c_* (not a benchmark)."""
import os, sys
os.environ["OSUDIT_CUDA_GRAPHS"] = "0"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "osu-diffusion_b200"))
import torch
from copy import deepcopy
import models
from diffusion import create_diffusion
from osudit import synth
from osudit.optim import FusedAdamWEMA

dev = torch.device("cuda", 0)
B = 256
model = models.DiT_models["DiT-B"](num_classes=52670, context_size=144, class_dropout_prob=0.2)
with torch.no_grad():
    for k, v in model.state_dict().items():
        if "adaLN_modulation" in k or k.startswith("final_layer.linear"):
            v.normal_(0, 0.02)
model = model.to(dev).train()
ema = deepcopy(model).requires_grad_(False)
d = create_diffusion("", noise_schedule="squaredcos_cap_v2", use_l1=True)
opt = FusedAdamWEMA(model.parameters(), lr=1e-4, weight_decay=0)
opt.attach_ema(ema, model)
scaler = torch.amp.GradScaler("cuda")
(x, o, c), y = synth.training_batch(B, 128, seed=0)
x, o, c, y = [t.to(dev) for t in (x, o, c, y)]
for _ in range(3):
    t = torch.randint(0, 1000, (B,), device=dev)
    with torch.autocast(device_type="cuda", dtype=torch.float16):
        loss = d.training_losses(model, x, t, dict(o=o, c=c, y=y))["loss"].mean()
    scaler.scale(loss).backward()
    scaler.step(opt); scaler.update(); opt.zero_grad(set_to_none=True)
torch.cuda.synchronize()
print("ok", float(loss))
