"""Time one attention algorithm at the three BASELINE attention shapes (development aid for csrc/attn_*.cu variants
built into alternate libraries: OSUDIT_LIB=... python tools/attn_time.py stream)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "osu-diffusion_b200"))
import torch
from osudit import ops
algo = getattr(ops, "ATTN_" + (sys.argv[1] if len(sys.argv) > 1 else "stream").upper())
res = []
for B, T, H, wl, wr in [(128, 2048, 12, 127, 128), (256, 128, 12, -1, -1), (128, 512, 16, -1, -1)]:
    qkv = torch.randn(B * T, 3 * H * 64, device="cuda").to(torch.bfloat16)
    out = torch.empty(B * T, H * 64, device="cuda", dtype=torch.bfloat16)
    f = lambda: ops.attn_band(qkv, out, B, T, H, 64, wl, wr, None, algo)
    for _ in range(3):
        f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        f()
    e1.record(); torch.cuda.synchronize()
    res.append(e0.elapsed_time(e1) / 20)
print(os.environ.get("OSUDIT_LIB", "default").split("/")[-1], " ".join(f"{v:.3f}" for v in res), flush=True)
