"""Host-side schedule of one DiT forward over the native kernels.

Mirrors DiT.forward (reference models.py:306-325) as a fixed sequence of libosudit launches on the
current stream.  Owns (a) the packed copies of the module's fp32 parameters — bf16 for the large
GEMMs, split-bf16 (hi, lo) for the precision-critical small ones — rebuilt when a parameter's
version counter changes, and (b) per-shape activation workspaces, so that steady-state forwards
allocate nothing and reuse the same TMA descriptors.
"""
from __future__ import annotations

import math
import os

import torch

from . import ops

FREQ_SEQ = 128  # FirstLayer.frequency_embedding_size (models.py:213)
FREQ_T = 256  # TimestepEmbedder.frequency_embedding_size (models.py:26)


def _freqs(half: int, device) -> torch.Tensor:
    # positional_embedding.py:39-44, evaluated by torch on the host exactly as the reference does
    f = torch.exp(-math.log(10000) * torch.arange(start=0, end=half, dtype=torch.float32) / half)
    return f.to(device)


class MaskSpec:
    """How the attention kernel should treat `attn_mask`: band (w_left, w_right), none, or generic."""

    __slots__ = ("w_left", "w_right", "generic")

    def __init__(self, w_left=-1, w_right=-1, generic=None):
        self.w_left, self.w_right, self.generic = w_left, w_right, generic


_mask_cache: dict = {}  # key -> (MaskSpec, the mask tensor): the entry keeps the tensor (and so its address) alive
_MASK_CACHE_ENTRIES = 16
_RESID_EPILOGUE = os.environ.get("OSUDIT_GEMM_RESID", "1") != "0"


def classify_mask(attn_mask, T: int) -> MaskSpec:
    """sample.py:81-84 builds a (T,T) bool band (True = blocked).  Recognise the closed form once
    per mask tensor so attention can skip whole key tiles; anything else is applied element-wise.
    The classification reads the mask back once (a host sync); later calls with the same tensor are sync-free."""
    if attn_mask is None:
        return MaskSpec()
    if attn_mask.dtype != torch.bool or attn_mask.shape != (T, T):
        raise ValueError("attn_mask must be a (T, T) bool tensor (True = not allowed) or None")
    key = (attn_mask.data_ptr(), attn_mask._version, T, str(attn_mask.device))
    hit = _mask_cache.get(key)
    if hit is not None and hit[1].untyped_storage().data_ptr() == attn_mask.untyped_storage().data_ptr():
        return hit[0]
    allowed = ~attn_mask
    wr = int(allowed[0].sum().item()) - 1
    wl = int(allowed[:, 0].sum().item()) - 1
    d = torch.arange(T, device=attn_mask.device)
    d = d[None, :] - d[:, None]
    if wl >= 0 and wr >= 0 and torch.equal(allowed, (d >= -wl) & (d <= wr)):
        spec = MaskSpec(wl, wr)
    else:
        spec = MaskSpec(-1, -1, attn_mask.to(torch.uint8).contiguous())
    while len(_mask_cache) >= _MASK_CACHE_ENTRIES:  # FIFO; a captured graph keeps its own reference to its spec
        _mask_cache.pop(next(iter(_mask_cache)))
    _mask_cache[key] = (spec, attn_mask)  # holding the tensor means its address cannot be recycled under this key
    return spec


def check_inputs(model, x, t, o, c, y, x_rows=None):
    """Shape checks the reference gets for free from its matmuls / embedding lookups (models.py:227-235,306-325):
    the native kernels take raw pointers, so a mismatch would read or write out of bounds instead of raising."""
    if o.dim() != 2:
        raise ValueError(f"DiT.forward: `o` must be (N, T), got {tuple(o.shape)}")
    B, T = o.shape
    if tuple(x.shape) != (B, model.in_channels, T):  # with CFG only the first x_rows rows are read, all B are passed
        raise ValueError(f"DiT.forward: `x` must be ({B}, {model.in_channels}, {T}), got {tuple(x.shape)}")
    if x_rows is not None and not 0 < x_rows <= B:
        raise ValueError(f"DiT.forward: x_rows={x_rows} outside (0, {B}]")
    if tuple(c.shape) != (B, model.context_size, T):
        raise ValueError(f"DiT.forward: `c` must be ({B}, {model.context_size}, {T}), got {tuple(c.shape)}")
    if tuple(t.shape) != (B,) or tuple(y.shape) != (B,):
        raise ValueError(f"DiT.forward: `t` and `y` must be ({B},), got {tuple(t.shape)} and {tuple(y.shape)}")
    if B * T == 0:
        raise ValueError("DiT.forward: empty batch")


_label_cache: dict = {}  # (ptr, version, n, rows) -> the tensor whose range was already checked on the host


def check_label_range(y, table_rows: int, sync: bool):
    """A label >= table rows is an IndexError in the reference (embedding lookup, models.py:73).  Inference checks it
    on the host once per label tensor (`sync`); every call also queues the device-side assert in front of the gather
    (ops.check_labels), which is what protects the training path without a per-step host synchronisation."""
    if sync:
        key = (y.data_ptr(), y._version, y.numel(), table_rows)
        hit = _label_cache.get(key)
        if hit is None or hit.untyped_storage().data_ptr() != y.untyped_storage().data_ptr():
            lo, hi = int(y.min()), int(y.max())
            if lo < 0 or hi >= table_rows:
                raise IndexError(f"class label out of range: [{lo}, {hi}] outside the embedding table [0, {table_rows})")
            while len(_label_cache) >= 16:
                _label_cache.pop(next(iter(_label_cache)))
            _label_cache[key] = y
    ops.check_labels(y, table_rows)


class PackedWeights:
    """GEMM-ready copies of a DiT module's parameters."""

    def __init__(self, model):
        self.versions = None
        self.refresh(model)

    @staticmethod
    def _signature(model):
        return tuple((p.data_ptr(), p._version) for p in model.parameters())

    def refresh(self, model):
        sig = self._signature(model)
        if sig == self.versions:
            return
        bf = lambda p: p.detach().to(torch.bfloat16).contiguous()  # noqa: E731
        f32 = lambda p: p.detach().float().contiguous()  # noqa: E731
        self.first_w = ops.split_bf16(model.xoc_embedder.mlp[0].weight)
        self.first_b = f32(model.xoc_embedder.mlp[0].bias)
        self.pf = [float(v) for v in model.xoc_embedder.playfield_size.detach().cpu()]
        self.t0_w = ops.split_bf16(model.t_embedder.mlp[0].weight)
        self.t0_b = f32(model.t_embedder.mlp[0].bias)
        self.t2_w = ops.split_bf16(model.t_embedder.mlp[2].weight)
        self.t2_b = f32(model.t_embedder.mlp[2].bias)
        self.table = f32(model.y_embedder.embedding_table.weight)
        mods_w = [blk.adaLN_modulation[1].weight for blk in model.blocks] + \
                 [model.final_layer.adaLN_modulation[1].weight]
        mods_b = [blk.adaLN_modulation[1].bias for blk in model.blocks] + \
                 [model.final_layer.adaLN_modulation[1].bias]
        self.mod_w = ops.split_bf16(torch.cat([w.detach() for w in mods_w], 0))
        self.mod_b = torch.cat([b.detach() for b in mods_b], 0).float().contiguous()
        self.blocks = []
        for blk in model.blocks:
            self.blocks.append(dict(
                qkv_w=bf(blk.attn.in_proj_weight), qkv_b=f32(blk.attn.in_proj_bias),
                out_w=bf(blk.attn.out_proj.weight), out_b=f32(blk.attn.out_proj.bias),
                fc1_w=bf(blk.mlp.fc1.weight), fc1_b=f32(blk.mlp.fc1.bias),
                fc2_w=bf(blk.mlp.fc2.weight), fc2_b=f32(blk.mlp.fc2.bias)))
        self.final_w = f32(model.final_layer.linear.weight)
        self.final_b = f32(model.final_layer.linear.bias)
        self.versions = sig


def _gemm3(a_hi, a_lo, w, bias, out):
    """~fp32-accurate product on the bf16 tensor cores: hi*Whi + lo*Whi + hi*Wlo."""
    w_hi, w_lo = w
    return ops.gemm([a_hi, a_lo, a_hi], [w_hi, w_hi, w_lo], bias, ops.EPI_F32, out)


class DiTEngine:
    def __init__(self, model):
        self.model = model
        self.D = model.hidden_size
        self.H = model.num_heads
        self.depth = len(model.blocks)
        self.E = model.context_size
        self.hidden_mlp = model.blocks[0].mlp.fc1.weight.shape[0]
        self.weights = None
        self._fp32 = None
        self._ws = {}
        self._freqs = {}
        # False while a CUDA graph is being warmed up / captured: a replayed graph must rewrite every column of the
        # first-layer operand itself, because another caller may have used the shared workspace in between
        self.reuse_oc_columns = True

    # ------------------------------------------------------------------ helpers
    def packed(self) -> PackedWeights:
        if self.weights is None:
            self.weights = PackedWeights(self.model)
        else:
            self.weights.refresh(self.model)
        return self.weights

    def freqs(self, half, device):
        key = (half, str(device))
        if key not in self._freqs:
            self._freqs[key] = _freqs(half, device)
        return self._freqs[key]

    def workspace(self, B: int, T: int, device):
        key = (B, T, str(device))
        ws = self._ws.get(key)
        if ws is None:
            if len(self._ws) >= 4:
                self._ws.clear()
            rows, D = B * T, self.D
            e = lambda *s, dt=torch.bfloat16: torch.empty(*s, dtype=dt, device=device)  # noqa: E731
            kin = 3 * FREQ_SEQ + self.E
            ws = dict(
                a_hi=e(rows, kin), a_lo=e(rows, kin), x=e(rows, D, dt=torch.float32), h=e(rows, D),
                qkv=e(rows, 3 * D), att=e(rows, D), y=e(rows, D), u=e(rows, self.hidden_mlp),
                tf_hi=e(B, FREQ_T), tf_lo=e(B, FREQ_T), t1=e(B, D, dt=torch.float32),
                s_hi=e(B, D), s_lo=e(B, D), temb=e(B, D, dt=torch.float32),
                c_hi=e(B, D), c_lo=e(B, D),
                mod=e(B, (6 * self.depth + 2) * D, dt=torch.float32),
                out=e(B, 4, T, dt=torch.float32))
            self._ws[key] = ws
        return ws

    def keepalive(self, B: int, T: int, device, attn_mask):
        """Everything a captured forward of this shape points into (activation workspace, packed weights, frequency
        tables, the mask's classification): a CUDA graph holds the returned objects so that evicting a workspace or
        re-packing the weights can never leave it replaying into freed memory."""
        precision = getattr(self.model, "precision", "bf16")
        if precision == "fp32":
            sched = self._fp32
            return (sched._ws.get((B, T, str(device))), dict(vars(sched.weights)) if sched.weights else None,
                    dict(self._freqs), classify_mask(attn_mask, T))
        return (self._ws.get((B, T, str(device))), dict(vars(self.weights)) if self.weights else None,
                dict(self._freqs), classify_mask(attn_mask, T))

    # ------------------------------------------------------------------ forward
    def conditioning(self, ws, t, y, w):
        """mod[B, depth*6D + 2D] = every adaLN Linear applied to SiLU(t_embedder(t) + y_embedder(y))
        (models.py:318-320,152-159,193), one batched GEMM for all blocks."""
        dev = t.device
        ops.timestep_features(t, self.freqs(FREQ_T // 2, dev), ws["tf_hi"], ws["tf_lo"])
        _gemm3(ws["tf_hi"], ws["tf_lo"], w.t0_w, w.t0_b, ws["t1"])
        ops.silu_split(ws["t1"], ws["s_hi"], ws["s_lo"])
        _gemm3(ws["s_hi"], ws["s_lo"], w.t2_w, w.t2_b, ws["temb"])
        check_label_range(y, w.table.shape[0], sync=not torch.cuda.is_current_stream_capturing())
        ops.silu_split(ws["temb"], ws["c_hi"], ws["c_lo"], table=w.table, y=y)
        _gemm3(ws["c_hi"], ws["c_lo"], w.mod_w, w.mod_b, ws["mod"])
        return ws["mod"]

    def conditioning_steps(self, t_steps, y, max_bytes=2 << 30):
        """adaLN modulation of SEVERAL denoising steps in one pass (the timesteps of a sampling loop are known up
        front, gaussian_diffusion.py:514-561): t_steps int64 [K, B] -> list of K tensors fp32 [B, depth*6D + 2D].
        One GEMM over K*B rows instead of K rounds of six small launches; computed in chunks of at most `max_bytes`."""
        K, B = t_steps.shape
        dev = t_steps.device
        w = self.packed()
        width = (6 * self.depth + 2) * self.D
        per = max(1, min(K, max_bytes // (B * width * 4)))
        out = []
        e = lambda *s, dt=torch.bfloat16: torch.empty(*s, dtype=dt, device=dev)  # noqa: E731
        for k0 in range(0, K, per):
            k1 = min(K, k0 + per)
            rows = (k1 - k0) * B
            D = self.D
            ws = dict(tf_hi=e(rows, FREQ_T), tf_lo=e(rows, FREQ_T), t1=e(rows, D, dt=torch.float32), s_hi=e(rows, D),
                      s_lo=e(rows, D), temb=e(rows, D, dt=torch.float32), c_hi=e(rows, D), c_lo=e(rows, D),
                      mod=e(rows, width, dt=torch.float32))
            mod = self.conditioning(ws, t_steps[k0:k1].reshape(-1).contiguous(), y.repeat(k1 - k0), w)
            out.extend(mod[i * B:(i + 1) * B] for i in range(k1 - k0))
        return out

    def forward(self, x, t, o, c, y, attn_mask=None, x_rows=None, mod=None):
        """Returns the raw model output fp32 [B, 4, T] (a workspace tensor, overwritten by the next
        call with the same shape).  x: [x_rows, 2, T] with x_rows in {B, B/2}."""
        precision = getattr(self.model, "precision", "bf16")
        if precision == "fp32":
            if self._fp32 is None:
                from .fp32 import Fp32Schedule
                self._fp32 = Fp32Schedule(self)
            return self._fp32.forward(x, t, o, c, y, attn_mask, x_rows)
        if precision != "bf16":
            raise ValueError(f"DiT.precision must be 'bf16' or 'fp32', got {precision!r}")
        B, T = o.shape
        D, H = self.D, self.H
        x_rows = B if x_rows is None else x_rows
        w = self.packed()
        ws = self.workspace(B, T, o.device)
        spec = classify_mask(attn_mask, T)

        # First-layer operand [x sincos (256) | o sincos (128) | c (E)]: within a sampling loop only the x columns
        # change from step to step (models.py:227-233), so when o and c are the tensors of the previous call on this
        # workspace only those 256 columns are rewritten (no re-read of c, 151 MB at config 2)
        oc_sig = (o.data_ptr(), o._version, c.data_ptr(), c._version, w.pf[0], w.pf[1])
        if self.reuse_oc_columns and ws.get("oc_sig") == oc_sig and not torch.cuda.is_current_stream_capturing():
            ops.embed_x(x, self.freqs(FREQ_SEQ // 2, o.device), w.pf[0], w.pf[1], B, T, self.E, x_rows,
                        ws["a_hi"], ws["a_lo"])
        else:
            ops.embed_xoc(x, o, c, self.freqs(FREQ_SEQ // 2, o.device), w.pf[0], w.pf[1], x_rows,
                          ws["a_hi"], ws["a_lo"])
            ws["oc_sig"] = oc_sig if self.reuse_oc_columns else None
            ws["oc_keep"] = (o, c)  # the signature is only meaningful while these addresses cannot be recycled
        _gemm3(ws["a_hi"], ws["a_lo"], w.first_w, w.first_b, ws["x"])
        if mod is None:
            mod = self.conditioning(ws, t, y, w)

        xres, h, qkv, att, yb, u = ws["x"], ws["h"], ws["qkv"], ws["att"], ws["y"], ws["u"]
        # The gated residual updates (models.py:164-174) run in the epilogue of the out-projection / fc2 GEMMs when the
        # CTA-pair kernel takes the shape (an fp32 TMA reduce-add into the residual stream): the LayerNorm kernels,
        # HBM-bound, then only read x (6 D bytes per token instead of 12 D).  OSUDIT_GEMM_RESID=0 keeps the bf16 branch.
        fused = _RESID_EPILOGUE and ops.gemm_gated_residual_applicable(B * T, D, T)
        for i, bw in enumerate(w.blocks):
            base = 6 * D * i
            if i == 0 or fused:
                ops.ln_modulate(xres, None, mod, 0, base, base + D, T, h)
            else:  # fold the previous block's gated MLP residual into this LayerNorm pass
                ops.ln_modulate(xres, yb, mod, base - D, base, base + D, T, h)
            ops.gemm([h], [bw["qkv_w"]], bw["qkv_b"], ops.EPI_BF16, qkv)
            ops.attn_band(qkv, att, B, T, H, D // H, spec.w_left, spec.w_right, spec.generic)
            if fused:
                ops.gemm_gated_residual(att, bw["out_w"], bw["out_b"], mod, base + 2 * D, T, xres)
                ops.ln_modulate(xres, None, mod, 0, base + 3 * D, base + 4 * D, T, h)
            else:
                ops.gemm([att], [bw["out_w"]], bw["out_b"], ops.EPI_BF16, yb)
                ops.ln_modulate(xres, yb, mod, base + 2 * D, base + 3 * D, base + 4 * D, T, h)
            ops.gemm([h], [bw["fc1_w"]], bw["fc1_b"], ops.EPI_BF16_GELU, u)
            if fused:
                ops.gemm_gated_residual(u, bw["fc2_w"], bw["fc2_b"], mod, base + 5 * D, T, xres)
            else:
                ops.gemm([u], [bw["fc2_w"]], bw["fc2_b"], ops.EPI_BF16, yb)
        fbase = 6 * D * self.depth
        ops.final_layer(xres, None if fused else yb, mod, fbase - D, fbase, fbase + D, T, w.final_w, w.final_b, ws["out"])
        return ws["out"]
