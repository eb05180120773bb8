"""A/B timing of attention-kernel variants (alternate libosudit builds), isolated and inside a denoising step."""
import os, sys, subprocess
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if len(sys.argv) == 1:
    libs = {"current": ""}
    alt = os.path.join(ROOT, "osu-diffusion_b200", "alt")
    for f in sorted(os.listdir(alt)) if os.path.isdir(alt) else []:
        if f.endswith(".so"): libs[f] = os.path.join(alt, f)
    for name, path in libs.items():
        env = dict(os.environ)
        if path: env["OSUDIT_LIB"] = path
        out = subprocess.run([sys.executable, __file__, "run"], env=env, capture_output=True, text=True, timeout=400)
        print(f"{name}: {out.stdout.strip()} {out.stderr.strip()[-300:]}")
    sys.exit(0)
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "osu-diffusion_b200"))
import torch
from osudit import ops
B, T, H = 128, 2048, 12
qkv = torch.randn(B * T, 3 * H * 64, device="cuda").to(torch.bfloat16)
out = torch.empty(B * T, H * 64, device="cuda", dtype=torch.bfloat16)
def t_attn(n):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): ops.attn_band(qkv, out, B, T, H, 64, 127, 128, algo=ops.ATTN_TCGEN05)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
t_attn(5)
iso = t_attn(30)
# sustained (power-capped) regime: interleave with big GEMMs like the real step
a = torch.randn(262144, 768, device="cuda").to(torch.bfloat16); w = torch.randn(3072, 768, device="cuda").to(torch.bfloat16)
u = torch.empty(262144, 3072, device="cuda", dtype=torch.bfloat16)
evs = []
for i in range(40):
    ops.gemm([a], [w], None, ops.EPI_BF16_GELU, u); ops.gemm([a], [w], None, ops.EPI_BF16_GELU, u)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); ops.attn_band(qkv, out, B, T, H, 64, 127, 128, algo=ops.ATTN_TCGEN05); e1.record()
    evs.append((e0, e1))
torch.cuda.synchronize()
hot = sum(a_.elapsed_time(b_) for a_, b_ in evs[10:]) / 30
print(f"isolated {iso:.3f} ms, between GEMMs {hot:.3f} ms")
