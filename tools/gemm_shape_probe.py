import sys, math, torch
sys.path.insert(0, "/root/repo/osu-diffusion_b200")
from osudit import ops
M = 262144
SHAPES_XL = ((3456, 1152, ops.EPI_BF16, "XL qkv"), (1152, 1152, ops.EPI_BF16, "XL out-proj"),
             (4608, 1152, ops.EPI_BF16_GELU, "XL fc1+GELU"), (1152, 4608, ops.EPI_BF16, "XL fc2"),
             (1152, 384, ops.EPI_BF16, "S qkv"), (384, 1536, ops.EPI_BF16, "S fc2"))
for (N, K, epi, name) in SHAPES_XL if "xl" in sys.argv[1:] else ((3072, 768, ops.EPI_BF16, "fc1 shape, no GELU"), (3072, 768, ops.EPI_BF16_GELU, "fc1 shape, GELU"),
                          (2304, 768, ops.EPI_BF16, "qkv shape"), (2304, 768, ops.EPI_BF16_GELU, "qkv shape + GELU"),
                          (1536, 768, ops.EPI_BF16, "N=1536"), (4608, 768, ops.EPI_BF16, "N=4608")):
    a = torch.randn(M, K, device="cuda").to(torch.bfloat16)
    w = (torch.randn(N, K, device="cuda") / math.sqrt(K)).to(torch.bfloat16)
    bias = torch.randn(N, device="cuda")
    out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    if "once" in sys.argv[1:]:  # one launch per shape, for an ncu capture
        ops.gemm([a], [w], bias, epi, out); torch.cuda.synchronize(); continue
    for _ in range(3): ops.gemm([a], [w], bias, epi, out)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20): ops.gemm([a], [w], bias, epi, out)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 20
    print(f"{name}: {ms:.3f} ms  {2*M*N*K/ms/1e9:.0f} TFLOP/s", flush=True)
    del a, w, out
