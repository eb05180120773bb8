"""Development check of the streaming tcgen05 attention kernel (csrc/attn_fa.cu): parity against fp32 torch on the
shapes the model uses plus adversarial ones (rows whose scores keep growing: the O rescale path), and its time per
layer at BASELINE config 2 / config 3 next to the round-1 kernels.  Not a benchmark of record."""
import math
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "osu-diffusion_b200"))
import torch
from osudit import ops

DEV = "cuda"
ALGO = ops.ATTN_STREAM if "--stream" in sys.argv else ops.ATTN_FA


def reference(qkv, B, T, H, hd, wl, wr):
    D = H * hd
    q, k, v = [a.float().reshape(B, T, H, hd).transpose(1, 2) for a in qkv.float().split(D, dim=1)]
    s = (q @ k.transpose(2, 3)) / math.sqrt(hd)
    if wl >= 0:
        d = torch.arange(T, device=qkv.device)
        d = d[None, :] - d[:, None]
        s = s.masked_fill(~((d >= -wl) & (d <= wr)), float("-inf"))
    lse2 = torch.logsumexp(s, -1) / math.log(2.0)
    o = (torch.softmax(s, -1) @ v).transpose(1, 2).reshape(B * T, D)
    return o, lse2


def rel(a, b):
    return float((a.double() - b.double()).norm() / b.double().norm())


def check(B, T, H, wl, wr, scale=1.0, ramp=0.0, seed=0):
    hd = 64
    g = torch.Generator(device=DEV).manual_seed(seed)
    qkv = torch.randn(B * T, 3 * H * hd, device=DEV, generator=g) * scale
    if ramp:  # keys grow along the sequence: every row's maximum keeps moving, in jumps far beyond 2^24
        t = torch.arange(T, device=DEV).repeat(B).float()[:, None]
        qkv[:, :H * hd] = 1.0 + 0.05 * torch.randn(B * T, H * hd, device=DEV, generator=g)
        qkv[:, H * hd:2 * H * hd] = ramp * t / T + 0.05 * torch.randn(B * T, H * hd, device=DEV, generator=g)
    qkv = qkv.to(torch.bfloat16)
    out = torch.full((B * T, H * hd), float("nan"), device=DEV, dtype=torch.bfloat16)
    lse = torch.full((B, H, T), float("nan"), device=DEV)
    ops.attn_band(qkv, out, B, T, H, hd, wl, wr, None, ALGO, lse=lse)
    torch.cuda.synchronize()
    ref, lse_ref = reference(qkv, B, T, H, hd, wl, wr)
    e_o, e_l = rel(out.float(), ref), float((lse - lse_ref).abs().max())
    bad = int((~torch.isfinite(out.float())).sum())
    print(f"B={B} T={T} H={H} band=({wl},{wr}) scale={scale} ramp={ramp}: out rel-L2 {e_o:.2e}, lse max abs {e_l:.2e}, "
          f"non-finite {bad}", flush=True)
    return e_o < 1e-2 and e_l < 2e-2 and bad == 0


def timeit(fn, n=20):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


if __name__ == "__main__":
    ok = True
    for args in [(1, 128, 1, -1, -1), (2, 128, 3, -1, -1), (3, 100, 2, -1, -1), (1, 256, 2, 127, 128), (2, 300, 2, 127, 128),
                 (1, 2048, 2, 127, 128), (2, 512, 4, -1, -1), (1, 1000, 1, 63, 64), (1, 640, 2, 300, 10), (5, 128, 7, -1, -1),
                 (1, 1, 1, -1, -1), (1, 129, 1, 127, 128)]:
        ok &= check(*args)
    ok &= check(1, 512, 2, -1, -1, scale=6.0)            # wide score range
    ok &= check(1, 1024, 2, -1, -1, ramp=12.0)           # maxima grow slab after slab: O rescale path
    ok &= check(1, 1024, 2, 255, 256, ramp=40.0, seed=1)
    print("PARITY", "OK" if ok else "FAILED", flush=True)
    if "--time" in sys.argv:
        hd = 64
        for name, B, T, H, wl, wr in [("config 2 (sampling)", 128, 2048, 12, 127, 128), ("config 3 (training)", 256, 128, 12, -1, -1),
                                      ("config 5 per GPU", 128, 512, 16, -1, -1)]:
            qkv = torch.randn(B * T, 3 * H * hd, device=DEV).to(torch.bfloat16)
            out = torch.empty(B * T, H * hd, device=DEV, dtype=torch.bfloat16)
            lse = torch.empty(B, H, T, device=DEV)
            res = {}
            for tag, algo, l in [("stream", ops.ATTN_STREAM, None), ("stream+lse", ops.ATTN_STREAM, lse), ("fa", ops.ATTN_FA, None), ("fa+lse", ops.ATTN_FA, lse), ("mma.sync+lse", ops.ATTN_MMA_SYNC, lse)] + \
                    ([("window", ops.ATTN_TCGEN05, None)] if (wl >= 0 or T <= 256) else []):
                res[tag] = timeit(lambda: ops.attn_band(qkv, out, B, T, H, hd, wl, wr, None, algo, lse=l))
            pairs = sum(min(T - 1, j + (wr if wr >= 0 else T)) - max(0, j - (wl if wl >= 0 else T)) + 1 for j in range(T))
            fl = 4.0 * hd * pairs * H * B
            print(name, {k: f"{v:.3f} ms ({fl / v / 1e9:.0f} TFLOP/s)" for k, v in res.items()}, flush=True)
