"""1-CTA vs CTA-pair GEMM at the four block shapes of config 2 (correctness + TFLOP/s)."""
import os, sys, subprocess, math
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if len(sys.argv) == 1:
    for mode in ("0", "1"):
        env = dict(os.environ, OSUDIT_GEMM_2CTA=mode)
        out = subprocess.run([sys.executable, __file__, mode], env=env, capture_output=True, text=True, timeout=240)
        print(f"--- OSUDIT_GEMM_2CTA={mode}\n{out.stdout}{out.stderr[-600:]}")
    sys.exit(0)
sys.path.insert(0, os.path.join(ROOT, "osu-diffusion_b200"))
import torch
from osudit import ops
M = 262144
for (N, K, epi, name) in ((2304, 768, ops.EPI_BF16, "qkv"), (768, 768, ops.EPI_BF16, "out"),
                          (3072, 768, ops.EPI_BF16_GELU, "fc1"), (768, 3072, ops.EPI_BF16, "fc2")):
    torch.manual_seed(0)
    a = torch.randn(M, K, device="cuda").to(torch.bfloat16)
    w = (torch.randn(N, K, device="cuda") / math.sqrt(K)).to(torch.bfloat16)
    bias = torch.randn(N, device="cuda")
    out = torch.full((M, N), float("nan"), device="cuda", dtype=torch.bfloat16)
    ops.gemm([a], [w], bias, epi, out)
    torch.cuda.synchronize()
    idx = torch.randint(0, M, (2048,), device="cuda")
    ref = a[idx].float() @ w.float().t() + bias
    if epi == ops.EPI_BF16_GELU: ref = torch.nn.functional.gelu(ref, approximate="tanh")
    err = float((out[idx].float() - ref).norm() / ref.norm())
    tail = float((out[-300:].float() - ((a[-300:].float() @ w.float().t() + bias) if epi != ops.EPI_BF16_GELU else torch.nn.functional.gelu(a[-300:].float() @ w.float().t() + bias, approximate="tanh"))).abs().max())
    for _ in range(3): ops.gemm([a], [w], bias, epi, out)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): ops.gemm([a], [w], bias, epi, out)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    print(f"{name}: rel err {err:.2e} tail max abs {tail:.2e} nan {bool(torch.isnan(out.float()).any())}  {ms:.3f} ms  {2*M*N*K/ms/1e9:.0f} TFLOP/s", flush=True)
