"""Gated-residual GEMM epilogue (csrc/gemm_2cta.cu, EPI_RESID) against the bf16-branch path at the config-2 block shapes:
out-projection and fc2 with the LayerNorm pass that follows each.  Development aid, not a benchmark of record."""
import math, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "osu-diffusion_b200"))
import torch
from osudit import ops
dev = "cuda"
B, T, D = 128, 2048, 768
M = B * T
def timeit(f, n=10):
    for _ in range(2): f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): f()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
x = torch.randn(M, D, device=dev)
mod = torch.randn(B, 6 * D, device=dev) * 0.1
h = torch.empty(M, D, device=dev, dtype=torch.bfloat16)
yb = torch.empty(M, D, device=dev, dtype=torch.bfloat16)
for K, name in ((768, "out-proj"), (3072, "fc2")):
    a = torch.randn(M, K, device=dev).to(torch.bfloat16)
    w = (torch.randn(D, K, device=dev) / math.sqrt(K)).to(torch.bfloat16)
    bias = torch.randn(D, device=dev)
    t_plain = timeit(lambda: ops.gemm([a], [w], bias, ops.EPI_BF16, yb))
    t_resid = timeit(lambda: ops.gemm_gated_residual(a, w, bias, mod, 2 * D, T, x))
    t_ln_fold = timeit(lambda: ops.ln_modulate(x, yb, mod, 2 * D, 3 * D, 4 * D, T, h))
    t_ln = timeit(lambda: ops.ln_modulate(x, None, mod, 0, 3 * D, 4 * D, T, h))
    fl = 2.0 * M * D * K
    print(f"{name}: gemm bf16 {t_plain:.3f} ms ({fl / t_plain / 1e9:.0f} TF/s) + LN fold {t_ln_fold:.3f} = {t_plain + t_ln_fold:.3f} | "
          f"gemm resid {t_resid:.3f} ms ({fl / t_resid / 1e9:.0f} TF/s) + LN {t_ln:.3f} = {t_resid + t_ln:.3f}", flush=True)
