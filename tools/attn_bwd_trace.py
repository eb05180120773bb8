"""Timeline of the tcgen05 attention backward (csrc/attn_bwd_tc.cu built with -DOSUDIT_ATTN_TRACE, library passed via
OSUDIT_LIB): cycle stamps of the MMA issuer, softmax quadrant 0 and epilogue quadrant 0 for consecutive problems of CTA 0."""
import ctypes, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "osu-diffusion_b200"))
import numpy as np
import torch
from osudit import _lib, ops
B, T, H, hd = 256, 128, 12, 64
D = H * hd
bf = lambda t: t.to(torch.bfloat16)
qkv = bf(torch.randn(B * T, 3 * D, device="cuda")); dout = bf(torch.randn(B * T, D, device="cuda"))
out = torch.empty(B * T, D, device="cuda", dtype=torch.bfloat16); lse = torch.empty(B, H, T, device="cuda")
ops.attn_band(qkv, out, B, T, H, hd, -1, -1, None, ops.ATTN_AUTO, lse=lse)
dqkv = torch.empty_like(qkv)
for _ in range(3):
    ops.attn_band_bwd(qkv, out, dout, lse, dqkv, B, T, H, hd, -1, -1)
torch.cuda.synchronize()
lib = _lib.load()
buf = np.zeros(3 * 16 * 8, dtype=np.int64)
lib.osudit_debug_bwd_trace.argtypes = [ctypes.c_void_p]
assert lib.osudit_debug_bwd_trace(buf.ctypes.data) == 0
tr = buf.reshape(3, 16, 8)
t0 = tr[0, 0, 0]
names = [["top", "p_full ok", "S/dP(i+1) issued", "g_free ok", "grads issued"],
         ["top", "sdp_full ok", "c0 computed", "g_full(i-1) ok", "c1 computed", "c1 ok", "p_full arrived", "-"],
         ["top", "g_full ok", "g_free arrived", "stores read"]]
for i in range(0, 8):
    for role, tag in enumerate(("mma", "softmax", "epilogue")):
        print(f"problem {4 + i} {tag:8s}: " + "  ".join(f"{n}={tr[role, i, e] - t0}" for e, n in enumerate(names[role])))
