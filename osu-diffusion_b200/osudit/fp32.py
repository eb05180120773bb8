"""fp32 mode of the DiT forward (north star: eps within 1e-5 relative L2 of the fp32 reference).

Selected with `model.precision = "fp32"` (or OSUDIT_PRECISION=fp32 at construction).  Same schedule as
`engine.DiTEngine.forward` (reference models.py:306-325), but every activation stays fp32 and every
GEMM runs as six bf16 tensor-core products of three-way operand splits accumulated in one fp32 TMEM
accumulator (csrc/fp32_mode.cu explains the layout); attention runs in fp32 on the CUDA cores.  This
mode exists to validate checkpoints / kernels against the reference, it is not the benchmarked path
(an order of magnitude slower than the bf16 mode).  Inference only.
"""
from __future__ import annotations

import ctypes

import torch

from . import _lib, ops
from .engine import FREQ_SEQ, FREQ_T, check_label_range, classify_mask

_PTR3 = ctypes.c_void_p * 3
_LONG3 = ctypes.c_int64 * 3


# ------------------------------------------------------------------ C-ABI wrappers
def split3(a, out3, act=0, table=None, y=None):
    """out3 bf16 [rows, 3K] = [hi | mid | lo] of act(a (+ table[y])); act 0 none, 1 GELU(tanh), 2 SiLU."""
    rows, K = a.shape
    lib = _lib.load()
    _lib.check(lib.osudit_split3_bf16(
        ops._chk(a, torch.float32, "split3.in"), rows, K, act,
        ops._chk(table, torch.float32, "split3.table") if table is not None else None,
        ops._chk(y, torch.int64, "split3.y") if y is not None else None,
        ops._chk(out3, torch.bfloat16, "split3.out"), ops._stream()), "osudit_split3_bf16")
    return out3


def pack_weight6(w):
    """nn.Linear weight fp32 [N, K] -> bf16 [N, 6K] = [hi hi hi mid mid lo]."""
    N, K = w.shape
    w3 = split3(w.detach().float().contiguous(), torch.empty(N, 3 * K, dtype=torch.bfloat16, device=w.device))
    hi, mid, lo = w3[:, :K], w3[:, K:2 * K], w3[:, 2 * K:]
    return torch.cat([hi, hi, hi, mid, mid, lo], dim=1).contiguous()


# 64-wide k-blocks per tensor-core accumulation chain: 8 (32 MMAs) keeps a GEMM within 6e-7 of fp64 (2: 1.5e-7) at a
# quarter of the fp32 reduce-add traffic, which is what the fp32 mode's run time consists of
KB_PER_SPLIT = 8


def gemm_f32(a3, w6, bias, out, kb_per_split=None):
    """out fp32 [M, N] = A W^T + bias with A given as split3 [M, 3K] and W as packed [N, 6K]."""
    kb_per_split = KB_PER_SPLIT if kb_per_split is None else kb_per_split
    M, K3 = a3.shape
    K = K3 // 3
    N = w6.shape[0]
    if w6.shape[1] != 6 * K or tuple(out.shape) != (M, N):
        raise _lib.OsuditError(f"gemm_f32: shape mismatch {tuple(a3.shape)} x {tuple(w6.shape)} -> {tuple(out.shape)}")
    pa, pw = ops._chk(a3, torch.bfloat16, "gemm_f32.a"), ops._chk(w6, torch.bfloat16, "gemm_f32.w")
    ap, bp, lda, ldb, ks = _PTR3(), _PTR3(), _LONG3(), _LONG3(), _LONG3()
    for s, (width, woff) in enumerate(((3 * K, 0), (2 * K, 3 * K), (K, 5 * K))):
        ap[s], bp[s] = pa, pw + 2 * woff  # (hi+mid+lo)Whi, (hi+mid)Wmid, hi Wlo
        lda[s], ldb[s], ks[s] = 3 * K, 6 * K, width
    lib = _lib.load()
    bias_p = ops._chk(bias, torch.float32, "gemm_f32.bias") if bias is not None else None
    out_p = ops._chk(out, torch.float32, "gemm_f32.out")
    if kb_per_split > 0:  # short tensor-core accumulation chains + fp32 round-to-nearest reduce-adds
        _lib.check(lib.osudit_gemm_bf16_splitk(3, ap, lda, bp, ldb, ks, M, N, bias_p, kb_per_split, out_p,
                                               out.stride(0), ops._stream()), "osudit_gemm_bf16_splitk")
    else:
        _lib.check(lib.osudit_gemm_bf16(3, ap, lda, bp, ldb, ks, M, N, bias_p, ops.EPI_F32, out_p, out.stride(0),
                                        ops._stream()), "osudit_gemm_bf16")
    return out


def ln_modulate(x, branch, mod, gate_col, shift_col, scale_col, T, h3):
    rows, D = x.shape
    lib = _lib.load()
    bp = ops._chk(branch, torch.float32, "ln32.branch") if branch is not None else None
    gp = ops._off(mod, gate_col) if branch is not None else None
    _lib.check(lib.osudit_ln_modulate_f32(ops._chk(x, torch.float32, "ln32.x"), bp, gp, ops._off(mod, shift_col),
                                          ops._off(mod, scale_col), mod.stride(0), rows, T, D,
                                          ops._chk(h3, torch.bfloat16, "ln32.h3"), ops._stream()),
               "osudit_ln_modulate_f32")
    return h3


def final_layer(x, branch, mod, gate_col, shift_col, scale_col, T, w, bias, out):
    rows, D = x.shape
    lib = _lib.load()
    bp = ops._chk(branch, torch.float32, "final32.branch") if branch is not None else None
    gp = ops._off(mod, gate_col) if branch is not None else None
    _lib.check(lib.osudit_final_layer_f32(ops._chk(x, torch.float32, "final32.x"), bp, gp,
                                          ops._off(mod, shift_col), ops._off(mod, scale_col), mod.stride(0), rows,
                                          T, D, ops._chk(w, torch.float32, "final32.w"),
                                          ops._chk(bias, torch.float32, "final32.bias"), w.shape[0],
                                          ops._chk(out, torch.float32, "final32.out"), ops._stream()),
               "osudit_final_layer_f32")
    return out


def attn_band(qkv, out, B, T, H, head_dim, w_left=-1, w_right=-1, mask=None):
    lib = _lib.load()
    _lib.check(lib.osudit_attn_band_f32(ops._chk(qkv, torch.float32, "attn32.qkv"),
                                        ops._chk(out, torch.float32, "attn32.out"), B, T, H, head_dim, w_left,
                                        w_right, ops._chk(mask, torch.uint8, "attn32.mask") if mask is not None else None,
                                        ops._stream()), "osudit_attn_band_f32")
    return out


def embed_xoc(x, o, c, freqs64, pf_x, pf_y, xrows, a):
    B, T = o.shape
    lib = _lib.load()
    _lib.check(lib.osudit_embed_xoc_f32(ops._chk(x, torch.float32, "embed32.x"), ops._chk(o, torch.float32, "embed32.o"),
                                        ops._chk(c, torch.float32, "embed32.c"),
                                        ops._chk(freqs64, torch.float32, "embed32.freqs"), pf_x, pf_y, B, xrows, T,
                                        c.shape[1], ops._chk(a, torch.float32, "embed32.a"), ops._stream()),
               "osudit_embed_xoc_f32")
    return a


def timestep_features(t, freqs128, out):
    lib = _lib.load()
    _lib.check(lib.osudit_timestep_features_f32(ops._chk(t, torch.int64, "tfeat32.t"),
                                                ops._chk(freqs128, torch.float32, "tfeat32.freqs"), t.shape[0],
                                                ops._chk(out, torch.float32, "tfeat32.out"), ops._stream()),
               "osudit_timestep_features_f32")
    return out


# ------------------------------------------------------------------ packed weights + schedule
class PackedWeightsF32:
    def __init__(self, model):
        self.versions = None
        self.refresh(model)

    def refresh(self, model):
        sig = tuple((p.data_ptr(), p._version) for p in model.parameters())
        if sig == self.versions:
            return
        f32 = lambda p: p.detach().float().contiguous()  # noqa: E731
        self.first_w = pack_weight6(model.xoc_embedder.mlp[0].weight)
        self.first_b = f32(model.xoc_embedder.mlp[0].bias)
        self.pf = [float(v) for v in model.xoc_embedder.playfield_size.detach().cpu()]
        self.t0_w, self.t0_b = pack_weight6(model.t_embedder.mlp[0].weight), f32(model.t_embedder.mlp[0].bias)
        self.t2_w, self.t2_b = pack_weight6(model.t_embedder.mlp[2].weight), f32(model.t_embedder.mlp[2].bias)
        self.table = f32(model.y_embedder.embedding_table.weight)
        mods = [blk.adaLN_modulation[1] for blk in model.blocks] + [model.final_layer.adaLN_modulation[1]]
        self.mod_w = pack_weight6(torch.cat([m.weight.detach() for m in mods], 0))
        self.mod_b = torch.cat([m.bias.detach() for m in mods], 0).float().contiguous()
        self.blocks = [dict(
            qkv_w=pack_weight6(b.attn.in_proj_weight), qkv_b=f32(b.attn.in_proj_bias),
            out_w=pack_weight6(b.attn.out_proj.weight), out_b=f32(b.attn.out_proj.bias),
            fc1_w=pack_weight6(b.mlp.fc1.weight), fc1_b=f32(b.mlp.fc1.bias),
            fc2_w=pack_weight6(b.mlp.fc2.weight), fc2_b=f32(b.mlp.fc2.bias)) for b in model.blocks]
        self.final_w = f32(model.final_layer.linear.weight)
        self.final_b = f32(model.final_layer.linear.bias)
        self.versions = sig


class Fp32Schedule:
    """Owned by DiTEngine; `forward` has the signature and return convention of DiTEngine.forward."""

    def __init__(self, engine):
        self.eng = engine
        self.weights = None
        self._ws = {}

    def packed(self):
        if self.weights is None:
            self.weights = PackedWeightsF32(self.eng.model)
        else:
            self.weights.refresh(self.eng.model)
        return self.weights

    def workspace(self, B, T, device):
        key = (B, T, str(device))
        ws = self._ws.get(key)
        if ws is None:
            self._ws.clear()
            eng = self.eng
            rows, D, Hm = B * T, eng.D, eng.hidden_mlp
            kin = 3 * FREQ_SEQ + eng.E
            f = lambda *s: torch.empty(*s, dtype=torch.float32, device=device)  # noqa: E731
            b = lambda *s: torch.empty(*s, dtype=torch.bfloat16, device=device)  # noqa: E731
            ws = dict(a=f(rows, kin), a3=b(rows, 3 * kin), x=f(rows, D), h3=b(rows, 3 * D), qkv=f(rows, 3 * D),
                      att=f(rows, D), att3=b(rows, 3 * D), y=f(rows, D), pre=f(rows, Hm), u3=b(rows, 3 * Hm),
                      tf=f(B, FREQ_T), tf3=b(B, 3 * FREQ_T), t1=f(B, D), s3=b(B, 3 * D), temb=f(B, D),
                      c3=b(B, 3 * D), mod=f(B, (6 * eng.depth + 2) * D), out=f(B, 4, T))
            self._ws[key] = ws
        return ws

    def forward(self, x, t, o, c, y, attn_mask=None, x_rows=None):
        eng = self.eng
        B, T = o.shape
        D, H = eng.D, eng.H
        x_rows = B if x_rows is None else x_rows
        w = self.packed()
        ws = self.workspace(B, T, o.device)
        spec = classify_mask(attn_mask, T)
        dev = o.device

        embed_xoc(x, o, c, eng.freqs(FREQ_SEQ // 2, dev), w.pf[0], w.pf[1], x_rows, ws["a"])
        gemm_f32(split3(ws["a"], ws["a3"]), w.first_w, w.first_b, ws["x"])
        # conditioning (models.py:318-320,152-159,193)
        timestep_features(t, eng.freqs(FREQ_T // 2, dev), ws["tf"])
        gemm_f32(split3(ws["tf"], ws["tf3"]), w.t0_w, w.t0_b, ws["t1"])
        gemm_f32(split3(ws["t1"], ws["s3"], act=2), w.t2_w, w.t2_b, ws["temb"])
        check_label_range(y, w.table.shape[0], sync=not torch.cuda.is_current_stream_capturing())
        gemm_f32(split3(ws["temb"], ws["c3"], act=2, table=w.table, y=y), w.mod_w, w.mod_b, ws["mod"])
        mod = ws["mod"]

        xres, h3, qkv, att, att3, yb, pre, u3 = (ws[k] for k in ("x", "h3", "qkv", "att", "att3", "y", "pre", "u3"))
        for i, bw in enumerate(w.blocks):
            base = 6 * D * i
            if i == 0:
                ln_modulate(xres, None, mod, 0, base, base + D, T, h3)
            else:
                ln_modulate(xres, yb, mod, base - D, base, base + D, T, h3)
            gemm_f32(h3, bw["qkv_w"], bw["qkv_b"], qkv)
            attn_band(qkv, att, B, T, H, D // H, spec.w_left, spec.w_right, spec.generic)
            gemm_f32(split3(att, att3), bw["out_w"], bw["out_b"], yb)
            ln_modulate(xres, yb, mod, base + 2 * D, base + 3 * D, base + 4 * D, T, h3)
            gemm_f32(h3, bw["fc1_w"], bw["fc1_b"], pre)
            gemm_f32(split3(pre, u3, act=1), bw["fc2_w"], bw["fc2_b"], yb)
        fbase = 6 * D * eng.depth
        final_layer(xres, yb, mod, fbase - D, fbase, fbase + D, T, w.final_w, w.final_b, ws["out"])
        return ws["out"]
