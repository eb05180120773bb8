"""Oracle: the DiT denoiser as a pure function of a reference ``state_dict``.

TEST INFRASTRUCTURE ONLY (see ``oracle/__init__.py``).

Restates, op by op, what ``/root/reference/models.py`` and
``/root/reference/positional_embedding.py`` compute, without ``nn.Module``s: every
function takes the weights as a flat ``dict[str, Tensor]`` with the reference's
checkpoint key names.  ``dtype=torch.float64`` gives a
higher-precision truth for error budgeting; fp32 is the parity target.
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import torch

# models.py:410-431 — the size registry (depth, hidden, heads).
SIZES = {
    "DiT-S": (12, 384, 6),
    "DiT-B": (12, 768, 12),
    "DiT-L": (24, 1024, 16),
    "DiT-XL": (28, 1152, 16),
}

LN_EPS = 1e-6  # models.py:129,136,187
FREQ_DIM_SEQ = 128  # models.py:213 (FirstLayer.frequency_embedding_size)
FREQ_DIM_T = 256  # models.py:26 (TimestepEmbedder.frequency_embedding_size)


@dataclass(frozen=True)
class DiTShape:
    depth: int
    hidden: int
    heads: int
    context_size: int = 144
    in_channels: int = 2
    num_classes: int = 52670
    mlp_ratio: float = 4.0

    @property
    def out_channels(self) -> int:  # models.py:262 (learn_sigma=True)
        return 2 * self.in_channels

    @property
    def first_in(self) -> int:  # models.py:216-218
        return self.in_channels * FREQ_DIM_SEQ + FREQ_DIM_SEQ + self.context_size


def shape_of(name: str, **kw) -> DiTShape:
    depth, hidden, heads = SIZES[name]
    return DiTShape(depth=depth, hidden=hidden, heads=heads, **kw)


# --------------------------------------------------------------------------- init


def init_state_dict(shape: DiTShape, seed: int = 1, zero_init_std: float = 0.02,
                    damp_x: float | None = None) -> dict:
    """Seeded weights with the reference's key names / shapes / order.

    Follows ``initialize_weights`` (models.py:275-304) — xavier-uniform Linears,
    N(0, 0.02) embedders — except that the tensors the reference zero-initialises
    (every ``adaLN_modulation.1`` and ``final_layer.linear``) are
    drawn from N(0, zero_init_std²), otherwise the model output is identically 0
    and any parity check is vacuous.  ``damp_x`` scales the x/y sin-cos columns
    of the first layer (the damped-feedback fixture of SURVEY F17).
    """
    g = torch.Generator().manual_seed(seed)
    D, H = shape.hidden, int(shape.hidden * shape.mlp_ratio)

    def xavier(out_f, in_f):
        bound = math.sqrt(6.0 / (in_f + out_f))
        return (torch.rand(out_f, in_f, generator=g) * 2 - 1) * bound

    def normal(*s, std=0.02):
        return torch.randn(*s, generator=g) * std

    sd = {}
    sd["xoc_embedder.playfield_size"] = torch.tensor((512.0, 384.0))
    sd["xoc_embedder.mlp.0.weight"] = normal(D, shape.first_in)
    sd["xoc_embedder.mlp.0.bias"] = torch.zeros(D)
    sd["t_embedder.mlp.0.weight"] = normal(D, FREQ_DIM_T)
    sd["t_embedder.mlp.0.bias"] = torch.zeros(D)
    sd["t_embedder.mlp.2.weight"] = normal(D, D)
    sd["t_embedder.mlp.2.bias"] = torch.zeros(D)
    sd["y_embedder.embedding_table.weight"] = normal(shape.num_classes + 1, D)
    for i in range(shape.depth):
        p = f"blocks.{i}."
        sd[p + "attn.in_proj_weight"] = xavier(3 * D, D)
        sd[p + "attn.in_proj_bias"] = torch.zeros(3 * D)
        sd[p + "attn.out_proj.weight"] = xavier(D, D)
        sd[p + "attn.out_proj.bias"] = torch.zeros(D)
        sd[p + "mlp.fc1.weight"] = xavier(H, D)
        sd[p + "mlp.fc1.bias"] = torch.zeros(H)
        sd[p + "mlp.fc2.weight"] = xavier(D, H)
        sd[p + "mlp.fc2.bias"] = torch.zeros(D)
        sd[p + "adaLN_modulation.1.weight"] = normal(6 * D, D, std=zero_init_std)
        sd[p + "adaLN_modulation.1.bias"] = normal(6 * D, std=zero_init_std)
    sd["final_layer.linear.weight"] = normal(shape.out_channels, D, std=zero_init_std)
    sd["final_layer.linear.bias"] = normal(shape.out_channels, std=zero_init_std)
    sd["final_layer.adaLN_modulation.1.weight"] = normal(2 * D, D, std=zero_init_std)
    sd["final_layer.adaLN_modulation.1.bias"] = normal(2 * D, std=zero_init_std)
    if damp_x is not None:
        sd["xoc_embedder.mlp.0.weight"][:, : shape.in_channels * FREQ_DIM_SEQ] *= damp_x
    return sd


def state_dict_layout(shape: DiTShape) -> list:
    """[(key, shape)] in the reference's registration order, without allocating."""
    D, H = shape.hidden, int(shape.hidden * shape.mlp_ratio)
    out = [("xoc_embedder.playfield_size", [2]),
           ("xoc_embedder.mlp.0.weight", [D, shape.first_in]), ("xoc_embedder.mlp.0.bias", [D]),
           ("t_embedder.mlp.0.weight", [D, FREQ_DIM_T]), ("t_embedder.mlp.0.bias", [D]),
           ("t_embedder.mlp.2.weight", [D, D]), ("t_embedder.mlp.2.bias", [D]),
           ("y_embedder.embedding_table.weight", [shape.num_classes + 1, D])]
    for i in range(shape.depth):
        p = f"blocks.{i}."
        out += [(p + "attn.in_proj_weight", [3 * D, D]), (p + "attn.in_proj_bias", [3 * D]),
                (p + "attn.out_proj.weight", [D, D]), (p + "attn.out_proj.bias", [D]),
                (p + "mlp.fc1.weight", [H, D]), (p + "mlp.fc1.bias", [H]),
                (p + "mlp.fc2.weight", [D, H]), (p + "mlp.fc2.bias", [D]),
                (p + "adaLN_modulation.1.weight", [6 * D, D]), (p + "adaLN_modulation.1.bias", [6 * D])]
    out += [("final_layer.linear.weight", [shape.out_channels, D]),
            ("final_layer.linear.bias", [shape.out_channels]),
            ("final_layer.adaLN_modulation.1.weight", [2 * D, D]),
            ("final_layer.adaLN_modulation.1.bias", [2 * D])]
    return out


def rerandomise_zero_init(sd: dict, seed: int = 1, std: float = 0.02) -> dict:
    """In place: redraw the tensors ``initialize_weights`` zeroes (models.py:295-304)."""
    g = torch.Generator().manual_seed(seed)
    for k, v in sd.items():
        if "adaLN_modulation" in k or k.startswith("final_layer.linear"):
            v.copy_(torch.randn(v.shape, generator=g) * std)
    return sd


# ---------------------------------------------------------------------- embeddings


def sincos(values: torch.Tensor, dim: int, max_period: float = 10000.0) -> torch.Tensor:
    """``timestep_embedding`` (positional_embedding.py:29-49): [cos | sin], cos first.

    The product ``float32(v) * freqs`` is formed in fp32 exactly as the reference
    does, whatever dtype the caller asked for downstream.
    """
    half = dim // 2
    k = torch.arange(half, dtype=torch.float32)
    freqs = torch.exp(-math.log(max_period) * k / half)
    args = values.reshape(-1, 1).to(torch.float32) * freqs[None]
    return torch.cat([args.cos(), args.sin()], dim=-1)


def first_layer_input(sd, x_btc, o_bt, c_btc):
    """models.py:227-233: [sincos(x*512) | sincos(y*384) | sincos(o/10) | c]."""
    B, T, _ = x_btc.shape
    pos = x_btc * sd["xoc_embedder.playfield_size"].to(x_btc.dtype)
    x_freq = sincos(pos, FREQ_DIM_SEQ).reshape(B, T, -1)  # positional_embedding.py:65-77
    o_freq = sincos(o_bt / 10, FREQ_DIM_SEQ).reshape(B, T, -1)  # :52-62
    return torch.cat([x_freq, o_freq, c_btc.to(torch.float32)], dim=-1)


def _linear(x, w, b):
    return x @ w.t() + b


def _silu(x):
    return x * torch.sigmoid(x)


def _gelu_tanh(x):  # nn.GELU(approximate="tanh"), models.py:138
    return 0.5 * x * (1.0 + torch.tanh(math.sqrt(2.0 / math.pi) * (x + 0.044715 * x ** 3)))


def _layer_norm(x):  # nn.LayerNorm(elementwise_affine=False, eps=1e-6)
    mu = x.mean(-1, keepdim=True)
    var = ((x - mu) ** 2).mean(-1, keepdim=True)
    return (x - mu) / torch.sqrt(var + LN_EPS)


def _modulate(x, shift, scale):  # models.py:12-13
    return x * (1 + scale[:, None]) + shift[:, None]


def conditioning(sd, t, y, dtype=torch.float32):
    """models.py:318-320: b = t_embedder(t) + y_embedder(y)  (eval mode: no label drop)."""
    tf = sincos(t, FREQ_DIM_T).to(dtype)
    h = _linear(tf, sd["t_embedder.mlp.0.weight"], sd["t_embedder.mlp.0.bias"])
    h = _linear(_silu(h), sd["t_embedder.mlp.2.weight"], sd["t_embedder.mlp.2.bias"])
    return h + sd["y_embedder.embedding_table.weight"][y]


def attention(h, w_in, b_in, w_out, b_out, heads, attn_mask):
    """nn.MultiheadAttention(batch_first) self-attention (models.py:130-135,164-170).

    Packed in_proj rows are [Wq; Wk; Wv]; heads are contiguous slices; a boolean
    ``attn_mask`` (T,T) is True where attention is NOT allowed (sample.py:82-84).
    """
    B, T, D = h.shape
    hd = D // heads
    qkv = _linear(h, w_in, b_in)
    q, k, v = (z.reshape(B, T, heads, hd).transpose(1, 2) for z in qkv.split(D, dim=-1))
    s = (q @ k.transpose(-1, -2)) / math.sqrt(hd)
    if attn_mask is not None:
        s = s.masked_fill(attn_mask, float("-inf"))
    p = torch.softmax(s, dim=-1)
    a = (p @ v).transpose(1, 2).reshape(B, T, D)
    return _linear(a, w_out, b_out)


def block(sd, i, x, b, heads, attn_mask):
    """DiTBlock.forward (models.py:151-175)."""
    p = f"blocks.{i}."
    mod = _linear(_silu(b), sd[p + "adaLN_modulation.1.weight"], sd[p + "adaLN_modulation.1.bias"])
    sh1, sc1, g1, sh2, sc2, g2 = mod.chunk(6, dim=1)
    h = _modulate(_layer_norm(x), sh1, sc1)
    x = x + g1[:, None] * attention(
        h, sd[p + "attn.in_proj_weight"], sd[p + "attn.in_proj_bias"],
        sd[p + "attn.out_proj.weight"], sd[p + "attn.out_proj.bias"], heads, attn_mask)
    h = _modulate(_layer_norm(x), sh2, sc2)
    h = _gelu_tanh(_linear(h, sd[p + "mlp.fc1.weight"], sd[p + "mlp.fc1.bias"]))
    h = _linear(h, sd[p + "mlp.fc2.weight"], sd[p + "mlp.fc2.bias"])
    return x + g2[:, None] * h


def final_layer(sd, x, b):
    """FinalLayer.forward (models.py:192-196)."""
    mod = _linear(_silu(b), sd["final_layer.adaLN_modulation.1.weight"],
                  sd["final_layer.adaLN_modulation.1.bias"])
    shift, scale = mod.chunk(2, dim=1)
    h = _modulate(_layer_norm(x), shift, scale)
    return _linear(h, sd["final_layer.linear.weight"], sd["final_layer.linear.bias"])


def forward(sd, heads, x, t, o, c, y, attn_mask=None, dtype=torch.float32, taps=None):
    """DiT.forward (models.py:306-325), eval mode. x (B,2,T) t (B,) o (B,T) c (B,E,T) y (B,).

    ``taps``: optional dict that receives intermediate tensors (for per-op parity tests).
    """
    if dtype != torch.float32:
        sd = {k: v.to(dtype) for k, v in sd.items()}
    xin = first_layer_input(sd, x.transpose(1, 2), o, c.transpose(1, 2)).to(dtype)
    h = _linear(xin, sd["xoc_embedder.mlp.0.weight"], sd["xoc_embedder.mlp.0.bias"])
    b = conditioning(sd, t, y, dtype)
    depth = 1 + max(int(k.split(".")[1]) for k in sd if k.startswith("blocks."))
    if taps is not None:
        taps["first_in"], taps["h0"], taps["cond"] = xin, h, b
    for i in range(depth):
        h = block(sd, i, h, b, heads, attn_mask)
        if taps is not None:
            taps[f"h{i + 1}"] = h
    out = final_layer(sd, h, b)
    return out.transpose(1, 2)


def forward_with_cfg(sd, heads, x, t, o, c, y, cfg_scale, attn_mask=None, dtype=torch.float32,
                     in_channels=2):
    """DiT.forward_with_cfg (models.py:327-343): only the first half of x is read."""
    half = x[: len(x) // 2]
    out = forward(sd, heads, torch.cat([half, half], 0), t, o, c, y, attn_mask, dtype)
    eps, rest = out[:, :in_channels], out[:, in_channels:]
    cond, uncond = eps.split(len(eps) // 2, dim=0)
    g = uncond + cfg_scale * (cond - uncond)
    return torch.cat([torch.cat([g, g], 0), rest], dim=1)


def band_mask(T: int, W: int = 128) -> torch.Tensor:
    """sample.py:81-84 in closed form: True (= blocked) unless -(W-1) <= key-query <= W."""
    q = torch.arange(T)[:, None]
    k = torch.arange(T)[None, :]
    d = k - q
    return ~((d >= -(W - 1)) & (d <= W))
