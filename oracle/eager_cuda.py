"""TEST INFRASTRUCTURE ONLY — the library calls the reference dispatches on a GPU, restated.

SURVEY §8(d) asks for "the reference PyTorch CUDA eager path on the same B200 (fp32+TF32 as
sample.py:25-26, and bf16 autocast) — that is the real bar".  /root/reference does not travel to the
GPU box, so this module restates the reference's *eager call sequence* with the same torch library
entry points its modules reach on CUDA: ``F.linear`` (nn.Linear), ``F.layer_norm`` (nn.LayerNorm,
eps 1e-6, no affine), ``F.scaled_dot_product_attention`` with the boolean mask (what
nn.MultiheadAttention's fast path calls, models.py:164-170), ``F.gelu(approximate="tanh")``,
``F.silu``, and elementwise torch ops for modulate / gate / CFG (models.py:12-13,151-175,327-343).
It shares weights (a reference-layout state dict) and semantics with oracle/dit.py and is checked
against it on CPU (tests/test_oracle_golden.py).  Nothing in the product imports this file; only
tests time it, next to the native path, to report the ratio.
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F

from . import dit

LN_EPS = 1e-6


def _sincos(values, dim):
    half = dim // 2
    k = torch.arange(half, dtype=torch.float32, device=values.device)
    freqs = torch.exp(-math.log(10000.0) * k / half)
    args = values.reshape(-1, 1).float() * freqs[None]
    return torch.cat([args.cos(), args.sin()], dim=-1)


def _modulate(x, shift, scale):
    return x * (1 + scale.unsqueeze(1)) + shift.unsqueeze(1)


def forward(sd, heads, x, t, o, c, y, attn_mask=None):
    """DiT.forward (models.py:306-325) in eval mode through stock torch CUDA kernels."""
    B, _, T = x.shape
    D = sd["t_embedder.mlp.2.bias"].shape[0]
    xt = x.transpose(1, 2)
    pos = xt * sd["xoc_embedder.playfield_size"]
    xin = torch.cat([_sincos(pos, dit.FREQ_DIM_SEQ).reshape(B, T, -1),
                     _sincos(o / 10, dit.FREQ_DIM_SEQ).reshape(B, T, -1), c.transpose(1, 2)], dim=-1)
    h = F.linear(xin, sd["xoc_embedder.mlp.0.weight"], sd["xoc_embedder.mlp.0.bias"])
    tf = _sincos(t, dit.FREQ_DIM_T)
    b = F.linear(F.silu(F.linear(tf, sd["t_embedder.mlp.0.weight"], sd["t_embedder.mlp.0.bias"])),
                 sd["t_embedder.mlp.2.weight"], sd["t_embedder.mlp.2.bias"])
    b = F.silu(b + sd["y_embedder.embedding_table.weight"][y])
    allowed = None if attn_mask is None else ~attn_mask
    depth = 1 + max(int(k.split(".")[1]) for k in sd if k.startswith("blocks."))
    hd = D // heads
    for i in range(depth):
        p = f"blocks.{i}."
        mod = F.linear(b, sd[p + "adaLN_modulation.1.weight"], sd[p + "adaLN_modulation.1.bias"])
        sh1, sc1, g1, sh2, sc2, g2 = mod.chunk(6, dim=1)
        m = _modulate(F.layer_norm(h, (D,), eps=LN_EPS), sh1, sc1)
        qkv = F.linear(m, sd[p + "attn.in_proj_weight"], sd[p + "attn.in_proj_bias"])
        q, k, v = (z.reshape(B, T, heads, hd).transpose(1, 2) for z in qkv.split(D, dim=-1))
        a = F.scaled_dot_product_attention(q, k, v, attn_mask=allowed).transpose(1, 2).reshape(B, T, D)
        h = h + g1.unsqueeze(1) * F.linear(a, sd[p + "attn.out_proj.weight"], sd[p + "attn.out_proj.bias"])
        m = _modulate(F.layer_norm(h, (D,), eps=LN_EPS), sh2, sc2)
        m = F.gelu(F.linear(m, sd[p + "mlp.fc1.weight"], sd[p + "mlp.fc1.bias"]), approximate="tanh")
        h = h + g2.unsqueeze(1) * F.linear(m, sd[p + "mlp.fc2.weight"], sd[p + "mlp.fc2.bias"])
    mod = F.linear(b, sd["final_layer.adaLN_modulation.1.weight"], sd["final_layer.adaLN_modulation.1.bias"])
    shift, scale = mod.chunk(2, dim=1)
    h = _modulate(F.layer_norm(h, (D,), eps=LN_EPS), shift, scale)
    return F.linear(h, sd["final_layer.linear.weight"], sd["final_layer.linear.bias"]).transpose(1, 2)


def forward_with_cfg(sd, heads, x, t, o, c, y, cfg_scale, attn_mask=None):
    """DiT.forward_with_cfg (models.py:327-343)."""
    half = x[: len(x) // 2]
    out = forward(sd, heads, torch.cat([half, half], 0), t, o, c, y, attn_mask)
    eps, rest = out[:, :2], out[:, 2:]
    cond, uncond = eps.split(len(eps) // 2, dim=0)
    g = uncond + cfg_scale * (cond - uncond)
    return torch.cat([torch.cat([g, g], 0), rest], dim=1)
