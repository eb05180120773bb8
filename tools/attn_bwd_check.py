"""Development check of the tcgen05 attention backward (csrc/attn_bwd_tc.cu): parity against torch autograd on
training-window shapes, and its time per layer at BASELINE config 3 (run with OSUDIT_ATTN_BWD_TC=0 for the round-1
mma.sync kernel).  Not a benchmark of record."""
import math, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "osu-diffusion_b200"))
import torch
from osudit import ops
DEV = "cuda"
bf = lambda t: t.to(torch.bfloat16)
def rel(a, b): return float((a.double() - b.double()).norm() / b.double().norm().clamp_min(1e-6))

def check(B, T, H, wl, wr, seed=0):
    hd = 64; D = H * hd
    g = torch.Generator(device=DEV).manual_seed(seed + T)
    qkv = bf(torch.randn(B * T, 3 * D, device=DEV, generator=g))
    dout = bf(torch.randn(B * T, D, device=DEV, generator=g))
    x = qkv.float().requires_grad_()
    q, k, v = (z.reshape(B, T, H, hd).transpose(1, 2) for z in x.reshape(B, T, 3 * D).split(D, -1))
    s = q @ k.transpose(-1, -2) / math.sqrt(hd)
    if wl >= 0:
        d = torch.arange(T, device=DEV); d = d[None, :] - d[:, None]
        s = s.masked_fill(~((d >= -wl) & (d <= wr)), float("-inf"))
    ref = (torch.softmax(s, -1) @ v).transpose(1, 2).reshape(B * T, D)
    ref.backward(dout.float())
    out = torch.empty(B * T, D, device=DEV, dtype=torch.bfloat16)
    lse = torch.empty(B, H, T, device=DEV)
    ops.attn_band(qkv, out, B, T, H, hd, wl, wr, None, ops.ATTN_AUTO, lse=lse)
    dbias = torch.zeros(3 * D, device=DEV)
    dqkv = torch.full_like(qkv, float("nan"))
    ops.attn_band_bwd(qkv, out, dout, lse, dqkv, B, T, H, hd, wl, wr, dbias=dbias)
    torch.cuda.synchronize()
    g_ = x.grad
    errs = [rel(dqkv[:, i * D:(i + 1) * D].float(), g_[:, i * D:(i + 1) * D]) for i in range(3)]
    eb = rel(dbias, g_.sum(0))
    bad = int((~torch.isfinite(dqkv.float())).sum())
    print(f"B={B} T={T} H={H} band=({wl},{wr}): dq {errs[0]:.2e} dk {errs[1]:.2e} dv {errs[2]:.2e} dbias {eb:.2e} non-finite {bad}", flush=True)
    return max(errs) < 1.2e-2 and eb < 1.2e-2 and bad == 0

ok = True
for a in [(1, 128, 1, -1, -1), (2, 128, 3, -1, -1), (3, 100, 2, -1, -1), (2, 64, 1, -1, -1), (1, 33, 2, 7, 8), (3, 128, 2, 39, 40),
          (7, 128, 5, -1, -1), (1, 1, 1, -1, -1), (150, 128, 2, -1, -1)]:
    ok &= check(*a)
print("PARITY", "OK" if ok else "FAILED", flush=True)
B, T, H, hd = 256, 128, 12, 64
D = H * hd
qkv = bf(torch.randn(B * T, 3 * D, device=DEV)); dout = bf(torch.randn(B * T, D, device=DEV))
out = torch.empty(B * T, D, device=DEV, dtype=torch.bfloat16); lse = torch.empty(B, H, T, device=DEV)
ops.attn_band(qkv, out, B, T, H, hd, -1, -1, None, ops.ATTN_AUTO, lse=lse)
dqkv = torch.empty_like(qkv); dbias = torch.zeros(3 * D, device=DEV)
for with_bias in (False, True):
    f = lambda: ops.attn_band_bwd(qkv, out, dout, lse, dqkv, B, T, H, hd, -1, -1, dbias=dbias if with_bias else None)
    for _ in range(3): f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): f()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    print(f"config 3 backward (OSUDIT_ATTN_BWD_TC={os.environ.get('OSUDIT_ATTN_BWD_TC', '1')}, dbias={with_bias}): {ms:.3f} ms per layer "
          f"({5 * 2 * 128 * 128 * 64 * B * H / ms / 1e9:.0f} TFLOP/s)", flush=True)
