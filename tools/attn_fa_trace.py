"""Timeline of the streaming attention kernel's hand-offs (csrc/attn_fa.cu built with -DOSUDIT_ATTN_TRACE into an
alternate library passed via OSUDIT_LIB): for a few consecutive slab steps of CTA 0, the cycle stamps of each slot's
MMA issuer and of softmax quadrant 0, relative to the first stamp."""
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
    ops.attn_band(qkv, out, B, T, H, hd, wl, wr, None, ops.ATTN_FA)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20):
    ops.attn_band(qkv, out, B, T, H, hd, wl, wr, None, ops.ATTN_FA)
e1.record(); torch.cuda.synchronize()
print(f"config {cfg}: {e0.elapsed_time(e1) / 20 * 1e3:.1f} us per launch")
lib = _lib.load()
if not hasattr(lib, "osudit_debug_fa_trace"):
    sys.exit(0)
buf = np.zeros(2 * 2 * 32 * 8, dtype=np.int64)
lib.osudit_debug_fa_trace.argtypes = [ctypes.c_void_p]
assert lib.osudit_debug_fa_trace(buf.ctypes.data) == 0
tr = buf.reshape(2, 2, 32, 8)
t0 = tr[:, :, 0, :][tr[:, :, 0, :] > 0].min()
mn = ["top", "p_full ok", "S(j+1) issued", "PV(j) issued"]
sn = ["top", "s_full ok", "ld0 issued", "ld0 landed", "pre-exp", "exp done", "stored", "p_full arrived"]
for n in range(0, 14):
    for s in range(2):
        print(f"step {8 + n} slot {s}  mma: " + " ".join(f"{mn[e]}={tr[s, 0, n, e] - t0}" for e in range(4)))
        print(f"                softmax: " + " ".join(f"{sn[e]}={tr[s, 1, n, e] - t0}" for e in range(8)))
