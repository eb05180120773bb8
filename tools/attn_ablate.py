"""Timing of the window-attention kernel under its debug ablations (OSUDIT_ATTN_ABLATE)."""
import os, sys, subprocess
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if len(sys.argv) == 1:
    for a in ("0", "1", "4", "5", "6", "7"):
        env = dict(os.environ, OSUDIT_ATTN_ABLATE=a)
        out = subprocess.run([sys.executable, __file__, a], env=env, capture_output=True, text=True)
        print(f"ablate={a}:", out.stdout.strip(), out.stderr.strip()[-300:])
    sys.exit(0)
sys.path.insert(0, os.path.join(ROOT, "osu-diffusion_b200"))
import torch
from osudit import ops
B, T, H = 128, 2048, 12
qkv = torch.randn(B * T, 3 * H * 64, device="cuda").to(torch.bfloat16)
out = torch.empty(B * T, H * 64, device="cuda", dtype=torch.bfloat16)
for _ in range(3): ops.attn_band(qkv, out, B, T, H, 64, 127, 128, algo=ops.ATTN_TCGEN05)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20): ops.attn_band(qkv, out, B, T, H, 64, 127, 128, algo=ops.ATTN_TCGEN05)
e1.record(); torch.cuda.synchronize()
print(f"{e0.elapsed_time(e1) / 20:.3f} ms")
