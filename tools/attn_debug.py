"""Diagnose the tcgen05 window attention kernel on a GPU box."""
import os, sys, math
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "osu-diffusion_b200"))
import torch
from osudit import ops, synth

def ref(qkv, B, T, H, hd, mask):
    D = H * hd
    q, k, v = (z.reshape(B, T, H, hd).transpose(1, 2).float() for z in qkv.reshape(B, T, 3 * D).split(D, -1))
    s = q @ k.transpose(-1, -2) / math.sqrt(hd)
    if mask is not None:
        s = s.masked_fill(mask, float("-inf"))
    return (torch.softmax(s, -1) @ v).transpose(1, 2).reshape(B * T, D)

def run(B, T, H, W):
    torch.manual_seed(0)
    qkv = torch.randn(B * T, 3 * H * 64, device="cuda").to(torch.bfloat16)
    out = torch.full((B * T, H * 64), float("nan"), device="cuda", dtype=torch.bfloat16)
    if W: ops.attn_band(qkv, out, B, T, H, 64, W - 1, W, algo=ops.ATTN_TCGEN05)
    else: ops.attn_band(qkv, out, B, T, H, 64, algo=ops.ATTN_TCGEN05)
    torch.cuda.synchronize()
    r = ref(qkv, B, T, H, 64, synth.band_mask(T, W).cuda() if W else None)
    o = out.float(); nan = torch.isnan(o)
    err = (o - r).abs(); err[nan] = 9.0
    print(f"B={B} T={T} H={H} W={W}: max err {float(err.max()):.3e} nan {float(nan.float().mean()):.3f} "
          f"rel {float((torch.nan_to_num(o) - r).norm() / r.norm()):.3e}", flush=True)
    if float(err.max()) > 5e-2:
        e = err.reshape(B, T, H, 64)
        print("  err by q-tile of 128:", [f"{float(e[:, i:i+128].mean()):.2e}" for i in range(0, T, 128)][:8])
        print("  err by row%8:", [f"{float(e[:, i::8].mean()):.2e}" for i in range(8)])
        print("  err by dim/8:", [f"{float(e[..., i*8:(i+1)*8].mean()):.2e}" for i in range(8)])
        print("  out[0,:6]", o[0, :6].tolist(), "ref", r[0, :6].tolist(), flush=True)

for cfg in [(1, 128, 1, None), (1, 128, 1, 128), (1, 256, 1, 128), (2, 300, 2, 128), (1, 512, 1, 8), (2, 2048, 3, 128)]:
    run(*cfg)
# timing at config-2 scale
B, T, H = 128, 2048, 12
qkv = torch.randn(B * T, 3 * H * 64, device="cuda").to(torch.bfloat16)
out = torch.empty(B * T, H * 64, device="cuda", dtype=torch.bfloat16)
for algo, name in ((ops.ATTN_MMA_SYNC, "mma.sync"), (ops.ATTN_TCGEN05, "tcgen05")):
    for _ in range(3): ops.attn_band(qkv, out, B, T, H, 64, 127, 128, algo=algo)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): ops.attn_band(qkv, out, B, T, H, 64, 127, 128, algo=algo)
    e1.record(); torch.cuda.synchronize()
    print(f"{name}: {e0.elapsed_time(e1) / 10:.3f} ms per launch at B=128 T=2048 H=12", flush=True)
