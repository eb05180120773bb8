"""Where a DiT-B training step (config 3) spends its time: per-phase wall time with syncs, and the
GPU-busy time from torch.profiler."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "osu-diffusion_b200"))
import torch
import bench_train, models
from copy import deepcopy
from diffusion import create_diffusion
from osudit import synth

dev = torch.device("cuda", 0)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
model = models.DiT_models["DiT-B"](num_classes=52670, context_size=144, class_dropout_prob=0.2)
with torch.no_grad():
    for k, v in model.state_dict().items():
        if "adaLN_modulation" in k or k.startswith("final_layer.linear"):
            v.normal_(0, 0.02)
model = model.to(dev).train()
ema = deepcopy(model).requires_grad_(False)
d = create_diffusion("", noise_schedule="squaredcos_cap_v2", use_l1=True)
FUSED = os.environ.get("FUSED", "0") == "1"
if FUSED:
    from osudit.optim import FusedAdamWEMA
    opt = FusedAdamWEMA(model.parameters(), lr=1e-4, weight_decay=0)
    opt.attach_ema(ema, model, decay=0.9999)
else:
    opt = torch.optim.AdamW(model.parameters(), lr=1e-4, weight_decay=0)
scaler = torch.amp.GradScaler("cuda")
(x, o, c), y = synth.training_batch(B, 128, seed=0)
x, o, c, y = [t.to(dev) for t in (x, o, c, y)]

def phases(sync):
    ts = [time.perf_counter()]
    def mark():
        if sync: torch.cuda.synchronize()
        ts.append(time.perf_counter())
    t = torch.randint(0, 1000, (B,), device=dev)
    with torch.autocast(device_type="cuda", dtype=torch.float16):
        loss = d.training_losses(model, x, t, dict(o=o, c=c, y=y))["loss"].mean()
    mark()
    scaler.scale(loss).backward(); mark()
    scaler.step(opt); scaler.update(); opt.zero_grad(set_to_none=True); mark()
    if not FUSED:
        bench_train.update_ema(ema, model)
    mark()
    return [b - a for a, b in zip(ts, ts[1:])]

for _ in range(3): phases(True)
acc = [0.0] * 4
for _ in range(5):
    for i, v in enumerate(phases(True)): acc[i] += v / 5
print("synced ms: forward+loss %.2f  backward %.2f  optimizer %.2f  ema %.2f" % tuple(1e3 * v for v in acc))
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(5): phases(False)
t_cpu = (time.perf_counter() - t0) / 5
torch.cuda.synchronize(); t_all = (time.perf_counter() - t0) / 5
print("async: CPU-side %.2f ms/step, wall %.2f ms/step" % (1e3 * t_cpu, 1e3 * t_all))
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    for _ in range(3): phases(False)
    torch.cuda.synchronize()
from torch.autograd import DeviceType
ka = [e for e in prof.key_averages() if e.device_type == DeviceType.CUDA]  # kernels / memcpys / memsets only
tot = sum(e.device_time_total for e in ka) / 3 / 1e3
print("GPU busy %.2f ms/step (sum of device activities)" % tot)
for e in sorted(ka, key=lambda e: -e.device_time_total)[:45]:
    print("  %-84s %8.3f ms  x%d" % (e.key[:84], e.device_time_total / 3 / 1e3, e.count // 3))

if os.environ.get("OSUDIT_CPROFILE"):
    import cProfile, pstats
    pr = cProfile.Profile()
    pr.enable()
    for _ in range(5): phases(False)
    torch.cuda.synchronize()
    pr.disable()
    st = pstats.Stats(pr); st.sort_stats("tottime").print_stats(22)
