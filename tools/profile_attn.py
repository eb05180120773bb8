"""One config-2-sized launch of each attention kernel (for ncu)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "osu-diffusion_b200"))
import torch
from osudit import ops
B, T, H = 128, 2048, 12
qkv = torch.randn(B * T, 3 * H * 64, device="cuda").to(torch.bfloat16)
out = torch.empty(B * T, H * 64, device="cuda", dtype=torch.bfloat16)
for _ in range(2):
    ops.attn_band(qkv, out, B, T, H, 64, 127, 128, algo=ops.ATTN_TCGEN05)
torch.cuda.synchronize()
