"""One launch each of the streaming kernel and the window kernel on the sampling band (BASELINE config 2's attention
shape: 128 rows x 2048 datapoints x 12 heads, W = 128) after a warm-up launch, for `ncu --set full`.  Not a benchmark."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "osu-diffusion_b200"))
import torch
from osudit import ops
B, T, H, hd = 128, 2048, 12, 64
D = H * hd
qkv = torch.randn(B * T, 3 * D, device="cuda").to(torch.bfloat16)
out = torch.empty(B * T, D, device="cuda", dtype=torch.bfloat16)
algos = [getattr(ops, "ATTN_" + a.upper()) for a in (sys.argv[1:] or ["fa", "tcgen05"])]
for _ in range(2):
    for a in algos:
        ops.attn_band(qkv, out, B, T, H, hd, 127, 128, None, a)
torch.cuda.synchronize()
print("ok")
