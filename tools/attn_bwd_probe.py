"""One forward + backward of the training attention at the config-3 shape (DiT-B, 256 x 128 datapoints) — the
command profiled by ncu for the single-CTA backward kernel; prints CUDA-event times."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "osu-diffusion_b200"))
import torch
from osudit import ops
B, T, H, hd = 256, 128, 12, 64
D = H * hd
torch.manual_seed(0)
qkv = torch.randn(B * T, 3 * D, device="cuda").to(torch.bfloat16)
dout = torch.randn(B * T, D, device="cuda").to(torch.bfloat16)
out = torch.empty(B * T, D, device="cuda", dtype=torch.bfloat16)
lse = torch.empty(B, H, T, device="cuda")
dqkv = torch.empty_like(qkv)
dbias = torch.zeros(3 * D, device="cuda")
def fwd(): ops.attn_band(qkv, out, B, T, H, hd, -1, -1, None, ops.ATTN_MMA_SYNC, lse=lse)
def bwd(): ops.attn_band_bwd(qkv, out, dout, lse, dqkv, B, T, H, hd, -1, -1, dbias=dbias)
for name, fn in (("forward", fwd), ("backward", bwd)):
    for _ in range(3): fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): fn()
    e1.record(); torch.cuda.synchronize()
    print(f"{name}: {e0.elapsed_time(e1) / 20 * 1e3:.1f} us")
