"""Timeline of the window-attention kernel's hand-offs (needs a library built with -DOSUDIT_ATTN_TRACE, passed via
OSUDIT_LIB): prints, for a few consecutive tiles of CTA 0, the cycle stamps of the MMA warp and the two softmax
halves relative to the first one."""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "osu-diffusion_b200"))
import numpy as np
import torch
from osudit import _lib, ops
B, T, H, hd = 128, 2048, 12, 64
D = H * hd
qkv = torch.randn(B * T, 3 * D, device="cuda").to(torch.bfloat16)
out = torch.empty(B * T, D, device="cuda", dtype=torch.bfloat16)
for _ in range(3):
    ops.attn_band(qkv, out, B, T, H, hd, 127, 128, None, ops.ATTN_TCGEN05)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20):
    ops.attn_band(qkv, out, B, T, H, hd, 127, 128, None, ops.ATTN_TCGEN05)
e1.record(); torch.cuda.synchronize()
print(f"attention: {e0.elapsed_time(e1) / 20 * 1e3:.1f} us per launch")
if not hasattr(_lib.load(), "osudit_debug_attn_trace"):
    sys.exit(0)
buf = np.zeros(3 * 16 * 8, dtype=np.int64)
lib = _lib.load()
lib.osudit_debug_attn_trace.argtypes = [ctypes.c_void_p]
assert lib.osudit_debug_attn_trace(buf.ctypes.data) == 0
tr = buf.reshape(3, 16, 8)
t0 = tr[0, 0, 0]
names = [["mma:top", "qk_full", "slab0 issued", "v_full+o_free", "PV0 issued", "slab1 issued", "slab2 issued", "PV1 issued"],
         ["h0:top", "s01_done", "o_done0 ok", "chunks done", "", "", "", ""],
         ["h1:top", "s_done", "o_done1 ok", "chunks done", "xchg synced", "o_done both", "O read", "stored"]]
for i in range(2, 7):
    print(f"--- tile {8 + i} (period vs previous tile: {tr[0, i, 0] - tr[0, i - 1, 0]} cycles)")
    for role in range(3):
        print("   " + "  ".join(f"{names[role][e]}={tr[role, i, e] - tr[0, i, 0]}" for e in range(8) if names[role][e]))
if hasattr(lib, "osudit_debug_attn_trace_chunks"):
    b2 = np.zeros(16 * 16, dtype=np.int64)
    lib.osudit_debug_attn_trace_chunks.argtypes = [ctypes.c_void_p]
    assert lib.osudit_debug_attn_trace_chunks(b2.ctypes.data) == 0
    ch = b2.reshape(16, 16)
    for i in range(2, 5):
        e = ch[i]
        print(f"tile {8 + i}, half 0 / quadrant 0 chunks (wait->emit cycles | emit cycles): " + "  ".join(
            f"c{k}: +{e[2 * k] - (e[2 * k - 1] if k else e[0])} | {e[2 * k + 1] - e[2 * k]}" for k in range(6)))
