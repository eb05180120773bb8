"""Timeline of the double-buffered streaming attention kernel's hand-offs (csrc/attn_stream.cu built with
-DOSUDIT_ATTN_TRACE into an alternate library passed via OSUDIT_LIB): for consecutive slabs of CTA 0, the cycle stamps of
the MMA issuer and of softmax quadrant 0 (both halves), relative to the first stamp."""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "osu-diffusion_b200"))
import numpy as np
import torch
from osudit import _lib, ops
cfg = sys.argv[1] if len(sys.argv) > 1 else "2"
B, T, H, wl, wr = {"2": (128, 2048, 12, 127, 128), "3": (256, 128, 12, -1, -1), "5": (128, 512, 16, -1, -1)}[cfg]
hd = 64
D = H * hd
qkv = torch.randn(B * T, 3 * D, device="cuda").to(torch.bfloat16)
out = torch.empty(B * T, D, device="cuda", dtype=torch.bfloat16)
for _ in range(3):
    ops.attn_band(qkv, out, B, T, H, hd, wl, wr, None, ops.ATTN_STREAM)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20):
    ops.attn_band(qkv, out, B, T, H, hd, wl, wr, None, ops.ATTN_STREAM)
e1.record(); torch.cuda.synchronize()
print(f"config {cfg}: {e0.elapsed_time(e1) / 20 * 1e3:.1f} us per launch")
lib = _lib.load()
if not hasattr(lib, "osudit_debug_stream_trace"):
    sys.exit(0)
buf = np.zeros(3 * 48 * 6, dtype=np.int64)
lib.osudit_debug_stream_trace.argtypes = [ctypes.c_void_p]
assert lib.osudit_debug_stream_trace(buf.ctypes.data) == 0
tr = buf.reshape(3, 48, 6)
t0 = tr[tr > 0].min()
mn = ["top", "p_full ok", "done", "operands ok", "elected", "mmas out"]
sn = ["top", "s_full ok", "ld landed", "chunk0", "chunk1", "p_full arrived"]
for n in range(0, 24):
    print(f"slab {12 + n}  mma: " + " ".join(f"{mn[e]}={tr[0, n, e] - t0}" for e in (0, 1, 3, 4, 5, 2)))
    for h in range(2):
        print(f"      softmax half {h}: " + " ".join(f"{sn[e]}={tr[1 + h, n, e] - t0}" for e in range(6)))
