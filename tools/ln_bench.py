import os, sys
sys.path.insert(0, "/root/repo/osu-diffusion_b200")
import torch
from osudit import ops
rows, D, T, B = 262144, 768, 2048, 128
x = torch.randn(rows, D, device="cuda"); y = torch.randn(rows, D, device="cuda").to(torch.bfloat16)
mod = torch.randn(B, 6 * D, device="cuda") * 0.1; h = torch.empty(rows, D, device="cuda", dtype=torch.bfloat16)
for _ in range(3): ops.ln_modulate(x, y, mod, 0, D, 2 * D, T, h)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(50): ops.ln_modulate(x, y, mod, 0, D, 2 * D, T, h)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 50
print(f"ln_modulate: {ms*1e3:.1f} us  {12*D*rows/ms/1e6:.0f} GB/s")
