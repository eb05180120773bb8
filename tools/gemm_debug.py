"""Diagnose the tcgen05 GEMM on a GPU box: error maps by tile / row / column group."""
import os, sys, math
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "osu-diffusion_b200"))
import torch
from osudit import ops

def run(M, N, K, epi=ops.EPI_F32):
    torch.manual_seed(0)
    a = torch.randn(M, K, device="cuda").to(torch.bfloat16)
    w = (torch.randn(N, K, device="cuda") / math.sqrt(K)).to(torch.bfloat16)
    out = torch.full((M, N), float("nan"), device="cuda", dtype=torch.float32 if epi == 0 else torch.bfloat16)
    ops.gemm([a], [w], None, epi, out)
    torch.cuda.synchronize()
    ref = a.float() @ w.float().t()
    err = (out.float() - ref).abs()
    nan = torch.isnan(out.float())
    print(f"M={M} N={N} K={K} epi={epi}: max err {float(err[~nan].max()) if (~nan).any() else -1:.3e} "
          f"nan frac {float(nan.float().mean()):.3f} ref rms {float(ref.pow(2).mean().sqrt()):.3f}")
    if float(err[~nan].max() if (~nan).any() else 1) > 1e-2 or nan.any():
        e = torch.where(nan, torch.full_like(err, 99.0), err)
        print(" err by row%8      :", [f"{float(e[r::8].mean()):.2e}" for r in range(8)])
        print(" err by 32-row quad:", [f"{float(e[q*32:(q+1)*32].mean()):.2e}" for q in range(min(4, M // 32))])
        print(" err by 16-col grp :", [f"{float(e[:, g*16:(g+1)*16].mean()):.2e}" for g in range(min(16, N // 16))])
        print(" out[0,:8]", out[0, :8].float().tolist()); print(" ref[0,:8]", ref[0, :8].tolist())
        for k0 in range(0, K, 16):   # which K slices were accumulated?
            part = a[:, k0:k0+16].float() @ w[:, k0:k0+16].float().t()
            print(f"  corr with K[{k0}:{k0+16}] partial: {float((out.float()[~nan] * part[~nan]).sum() / part[~nan].pow(2).sum()):.3f}", end=";")
        print()

if __name__ == "__main__":
    for shape in [(128, 128, 64), (128, 128, 128), (128, 256, 64), (256, 128, 256), (128, 128, 528), (100, 136, 72)]:
        run(*shape)
    run(256, 256, 128, ops.EPI_BF16)
    run(256, 256, 128, ops.EPI_BF16_GELU)
