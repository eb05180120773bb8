import sys, torch
sys.path.insert(0, "/root/repo/osu-diffusion_b200"); sys.path.insert(0, "/root/repo")
import models, bench
from osudit import synth
dev = "cuda"
m = bench.build_native(torch.device("cuda", 0))
n, T = 8, 2048
z, o, c, y = [t.to(dev) for t in synth.sampling_batch(n, T, seed=0)]
mask = synth.band_mask(T, 128).to(dev)
t = torch.full((2 * n,), 500, device=dev)
def run(): return m.forward_with_cfg(z, t, o=o, c=c, y=y, cfg_scale=1.5, attn_mask=mask)
res = {}
with torch.no_grad():
    for prec in ("bf16", "fp32"):
        m.precision = prec
        for _ in range(2): out = run()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5): out = run()
        e1.record(); torch.cuda.synchronize()
        res[prec] = (e0.elapsed_time(e1) / 5, out.clone())
print(f"DiT-B, {2*n} rows x {T}: bf16 {res['bf16'][0]:.2f} ms, fp32 mode {res['fp32'][0]:.2f} ms per forward "
      f"({res['fp32'][0]/res['bf16'][0]:.1f}x); eps rel-L2 between them {float((res['bf16'][1][:, :2]-res['fp32'][1][:, :2]).norm()/res['fp32'][1][:, :2].norm()):.2e}")
