"""The attention kernels at the shapes the model runs them, for ncu (profiles/README.md): the streaming forward (attn_stream.cu) at
BASELINE config 3 (256 x 128, with log-sum-exp) and config 5 per GPU (128 x 512), the window kernel at config 2,
and the tcgen05 backward at config 3.  Not a benchmark."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "osu-diffusion_b200"))
import torch
from osudit import ops
dev, hd = "cuda", 64
bf = lambda t: t.to(torch.bfloat16)
for B, T, H, wl, wr, bwd in [(256, 128, 12, -1, -1, True), (128, 512, 16, -1, -1, False), (128, 2048, 12, 127, 128, False)]:
    D = H * hd
    qkv = bf(torch.randn(B * T, 3 * D, device=dev))
    out = torch.empty(B * T, D, device=dev, dtype=torch.bfloat16)
    lse = torch.empty(B, H, T, device=dev)
    for _ in range(3):
        ops.attn_band(qkv, out, B, T, H, hd, wl, wr, None, ops.ATTN_STREAM, lse=lse if bwd else None)
        if wl >= 0:
            ops.attn_band(qkv, out, B, T, H, hd, wl, wr, None, ops.ATTN_TCGEN05)
        if bwd:
            dout = bf(torch.randn(B * T, D, device=dev))
            ops.attn_band_bwd(qkv, out, dout, lse, torch.empty_like(qkv), B, T, H, hd, wl, wr, dbias=torch.zeros(3 * D, device=dev))
    torch.cuda.synchronize()
print("ok")
