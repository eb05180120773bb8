"""Timeline of the CTA-pair GEMM (needs a library built with -DOSUDIT_GEMM_TRACE, passed via OSUDIT_LIB): cycle
stamps of the MMA issuer and of epilogue thread 0 for a few consecutive tiles of CTA 0."""
import ctypes, math, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "osu-diffusion_b200"))
import numpy as np
import torch
from osudit import _lib, ops
N, K = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (2304, 768)
mode = sys.argv[3] if len(sys.argv) > 3 else "plain"
M = int(sys.argv[4]) if len(sys.argv) > 4 else 262144
epi = {"plain": ops.EPI_BF16, "gelu": ops.EPI_BF16_GELU, "gelu_save": ops.EPI_BF16_GELU_SAVE, "dgelu": ops.EPI_BF16_DGELU}[mode]
a = torch.randn(M, K, device="cuda").to(torch.bfloat16)
w = (torch.randn(N, K, device="cuda") / math.sqrt(K)).to(torch.bfloat16)
bias = torch.randn(N, device="cuda")
out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
aux = torch.randn(M, N, device="cuda").to(torch.bfloat16)
def run():
    if mode in ("plain", "gelu"): ops.gemm([a], [w], bias, epi, out)
    else: ops.gemm_aux(a, w, None if mode == "dgelu" else bias, epi, out, aux)
for _ in range(3): run()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): run()
e1.record(); torch.cuda.synchronize()
print(f"{mode} M={M}: {e0.elapsed_time(e1) / 10 * 1e3:.1f} us  {2 * M * N * K / (e0.elapsed_time(e1) / 10) / 1e9:.0f} TFLOP/s")
torch.cuda.synchronize()
buf = np.zeros(2 * 8 * 20, dtype=np.int64)
lib = _lib.load()
if not hasattr(lib, "osudit_debug_gemm_trace"):
    sys.exit(0)
lib.osudit_debug_gemm_trace.argtypes = [ctypes.c_void_p]
assert lib.osudit_debug_gemm_trace(buf.ctypes.data) == 0
tr = buf.reshape(2, 8, 20)
print(f"N={N} K={K} epi={epi}")
for i in range(1, 6):
    t0 = tr[0, i, 0]
    print(f"--- tile {4 + i}: mma period {tr[0, i, 0] - tr[0, i - 1, 0]}  acc free={tr[0, i, 1] - t0}  mmas issued={tr[0, i, 2] - t0}")
    e = tr[1, i]
    print(f"    epi: wait-top={e[0] - t0} acc ready={e[1] - t0} | " + " | ".join(
        f"c{c}: ld@{e[2 + 4 * c] - t0} buf@{e[3 + 4 * c] - t0} st@{e[4 + 4 * c] - t0} tma@{e[5 + 4 * c] - t0}" for c in range(4)))
    print(f"    c3 detail: after tcgen05 fence @{e[18] - t0}, after arrive @{e[19] - t0}")
