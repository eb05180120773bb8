"""Tensor-level wrappers over the C ABI: validate, pass raw pointers + the current stream.

PyTorch is used for device memory and streams only; all arithmetic happens in libosudit.so.
"""
from __future__ import annotations

import ctypes

import torch

from . import _lib

EPI_F32, EPI_BF16, EPI_BF16_GELU = 0, 1, 2
_PTR3 = ctypes.c_void_p * 3
_LONG3 = ctypes.c_int64 * 3


launch_count = 0  # native kernel launches issued through this module (bench.py reports it)


def _stream() -> int:
    global launch_count
    launch_count += 1
    return torch.cuda.current_stream().cuda_stream


def _chk(t: torch.Tensor, dtype, name: str) -> int:
    if not t.is_cuda:
        raise _lib.OsuditError(f"{name}: expected a CUDA tensor (this path has no CPU fallback)")
    if t.dtype != dtype:
        raise _lib.OsuditError(f"{name}: expected {dtype}, got {t.dtype}")
    if not t.is_contiguous():
        raise _lib.OsuditError(f"{name}: expected a contiguous tensor")
    return t.data_ptr()


def gemm(a_segs, b_segs, bias, epilogue: int, out: torch.Tensor) -> torch.Tensor:
    """out[M,N] = sum_s a_segs[s][M,K_s] @ b_segs[s][N,K_s].T + bias (bf16 in, fp32 accumulate)."""
    n = len(a_segs)
    M, N = out.shape
    ap, bp, lda, ldb, ks = _PTR3(), _PTR3(), _LONG3(), _LONG3(), _LONG3()
    for s in range(n):
        a, b = a_segs[s], b_segs[s]
        if a.shape[0] != M or b.shape[0] != N or a.shape[1] != b.shape[1]:
            raise _lib.OsuditError(f"gemm: shape mismatch {tuple(a.shape)} x {tuple(b.shape)} -> {tuple(out.shape)}")
        ap[s] = _chk(a, torch.bfloat16, "gemm.a")
        bp[s] = _chk(b, torch.bfloat16, "gemm.b")
        lda[s], ldb[s], ks[s] = a.stride(0), b.stride(0), a.shape[1]
    want = torch.float32 if epilogue == EPI_F32 else torch.bfloat16
    op = _chk(out, want, "gemm.out")
    bptr = _chk(bias, torch.float32, "gemm.bias") if bias is not None else None
    lib = _lib.load()
    _lib.check(lib.osudit_gemm_bf16(n, ap, lda, bp, ldb, ks, M, N, bptr, epilogue, op, out.stride(0),
                                    _stream()), "osudit_gemm_bf16")
    return out


EPI_BF16_GELU_SAVE, EPI_BF16_DGELU = 3, 4


def gemm_aux(a, w, bias, epilogue: int, out, aux):
    """Single-segment bf16 GEMM out = f(a @ w.T + bias) with a second [M, N] bf16 tensor in the epilogue:
    GELU_SAVE writes gelu(pre) to `out` and gelu'(pre) to `aux`; DGELU multiplies the product by `aux`."""
    M, N = out.shape
    if a.shape[0] != M or w.shape[0] != N or a.shape[1] != w.shape[1] or tuple(aux.shape) != (M, N):
        raise _lib.OsuditError(f"gemm_aux: shape mismatch {tuple(a.shape)} x {tuple(w.shape)} -> {tuple(out.shape)}")
    lib = _lib.load()
    _lib.check(lib.osudit_gemm_bf16_aux(_chk(a, torch.bfloat16, "gemm_aux.a"), a.stride(0),
                                        _chk(w, torch.bfloat16, "gemm_aux.w"), w.stride(0), a.shape[1], M, N,
                                        _chk(bias, torch.float32, "gemm_aux.bias") if bias is not None else None,
                                        epilogue, _chk(out, torch.bfloat16, "gemm_aux.out"), out.stride(0),
                                        _chk(aux, torch.bfloat16, "gemm_aux.aux"), aux.stride(0), _stream()),
               "osudit_gemm_bf16_aux")
    return out


def gemm_gated_residual_applicable(M: int, N: int, rows_per_batch: int) -> bool:
    return bool(_lib.load().osudit_gemm_gated_residual_applicable(M, N, rows_per_batch))


def gemm_gated_residual(a, w, bias, mod, gate_col: int, rows_per_batch: int, x):
    """x fp32 [M, N] += mod[row // rows_per_batch, gate_col : gate_col + N] * (a @ w.T + bias): the block's gated
    residual update (models.py:164-174) in the epilogue of the GEMM that produces the branch."""
    M, N = x.shape
    if a.shape[0] != M or w.shape[0] != N or a.shape[1] != w.shape[1]:
        raise _lib.OsuditError(f"gemm_gated_residual: shape mismatch {tuple(a.shape)} x {tuple(w.shape)} -> {tuple(x.shape)}")
    _chk(mod, torch.float32, "gemm_gated_residual.mod")
    lib = _lib.load()
    _lib.check(lib.osudit_gemm_gated_residual(_chk(a, torch.bfloat16, "gemm_gated_residual.a"), a.stride(0),
                                              _chk(w, torch.bfloat16, "gemm_gated_residual.w"), w.stride(0), a.shape[1],
                                              M, N, _chk(bias, torch.float32, "gemm_gated_residual.bias") if bias is not None else None,
                                              _off(mod, gate_col), mod.stride(0), rows_per_batch,
                                              _chk(x, torch.float32, "gemm_gated_residual.x"), x.stride(0), _stream()),
               "osudit_gemm_gated_residual")
    return x


ATTN_AUTO, ATTN_MMA_SYNC, ATTN_TCGEN05, ATTN_FA, ATTN_STREAM = 0, 1, 2, 3, 4


def attn_band(qkv, out, B, T, H, head_dim, w_left=-1, w_right=-1, mask=None, algo=ATTN_AUTO, lse=None):
    mp = _chk(mask, torch.uint8, "attn.mask") if mask is not None else None
    lib = _lib.load()
    _lib.check(lib.osudit_attn_band(_chk(qkv, torch.bfloat16, "attn.qkv"),
                                    _chk(out, torch.bfloat16, "attn.out"), B, T, H, head_dim,
                                    w_left, w_right, mp, algo,
                                    _chk(lse, torch.float32, "attn.lse") if lse is not None else None,
                                    _stream()), "osudit_attn_band")
    return out


def _off(t: torch.Tensor, col: int) -> int:
    return t.data_ptr() + col * t.element_size()


def ln_modulate(x, branch, mod, gate_col, shift_col, scale_col, T, h, x_out=None):
    """x fp32 [rows,D] (+= gate*branch in place), h bf16 = LN(x)*(1+scale)+shift; the three chunks
    live at columns gate_col/shift_col/scale_col of `mod` fp32 [B, mod_ld]."""
    rows, D = x.shape
    _chk(mod, torch.float32, "ln.mod")
    lib = _lib.load()
    bp = _chk(branch, torch.bfloat16, "ln.branch") if branch is not None else None
    gp = _off(mod, gate_col) if branch is not None else None
    _lib.check(lib.osudit_ln_modulate(_chk(x, torch.float32, "ln.x"), bp, gp, _off(mod, shift_col),
                                      _off(mod, scale_col), mod.stride(0), rows, T, D,
                                      _chk(h, torch.bfloat16, "ln.h"),
                                      _chk(x_out, torch.float32, "ln.x_out") if x_out is not None else None,
                                      _stream()), "osudit_ln_modulate")
    return h


def final_layer(x, branch, mod, gate_col, shift_col, scale_col, T, w, bias, out):
    rows, D = x.shape
    _chk(mod, torch.float32, "final.mod")
    lib = _lib.load()
    bp = _chk(branch, torch.bfloat16, "final.branch") if branch is not None else None
    gp = _off(mod, gate_col) if branch is not None else None
    _lib.check(lib.osudit_final_layer(_chk(x, torch.float32, "final.x"), bp, gp, _off(mod, shift_col),
                                      _off(mod, scale_col), mod.stride(0), rows, T, D,
                                      _chk(w, torch.float32, "final.w"),
                                      _chk(bias, torch.float32, "final.bias"), w.shape[0],
                                      _chk(out, torch.float32, "final.out"), _stream()),
               "osudit_final_layer")
    return out


def embed_xoc(x, o, c, freqs64, pf_x, pf_y, xrows, a_hi, a_lo):
    B, T = o.shape
    E = c.shape[1]
    lib = _lib.load()
    _lib.check(lib.osudit_embed_xoc(_chk(x, torch.float32, "embed.x"), _chk(o, torch.float32, "embed.o"),
                                    _chk(c, torch.float32, "embed.c"),
                                    _chk(freqs64, torch.float32, "embed.freqs"), pf_x, pf_y, B, xrows,
                                    T, E, _chk(a_hi, torch.bfloat16, "embed.hi"),
                                    _chk(a_lo, torch.bfloat16, "embed.lo"), _stream()), "osudit_embed_xoc")


def embed_x(x, freqs64, pf_x, pf_y, B, T, E, xrows, a_hi, a_lo):
    """Rewrite only the 256 x columns of a_hi / a_lo [B*T, 384 + E] (the o / c columns keep a previous embed_xoc)."""
    lib = _lib.load()
    _lib.check(lib.osudit_embed_x(_chk(x, torch.float32, "embed.x"), _chk(freqs64, torch.float32, "embed.freqs"),
                                  pf_x, pf_y, B, xrows, T, E, _chk(a_hi, torch.bfloat16, "embed.hi"),
                                  _chk(a_lo, torch.bfloat16, "embed.lo"), _stream()), "osudit_embed_x")


def timestep_features(t, freqs128, hi, lo):
    lib = _lib.load()
    _lib.check(lib.osudit_timestep_features(_chk(t, torch.int64, "tfeat.t"),
                                            _chk(freqs128, torch.float32, "tfeat.freqs"), t.shape[0],
                                            _chk(hi, torch.bfloat16, "tfeat.hi"),
                                            _chk(lo, torch.bfloat16, "tfeat.lo"), _stream()),
               "osudit_timestep_features")


def silu_split(a, hi, lo, a_index=None, table=None, y=None):
    rows, D = hi.shape
    lib = _lib.load()
    _lib.check(lib.osudit_silu_split(
        _chk(a, torch.float32, "silu.a"),
        _chk(a_index, torch.int32, "silu.index") if a_index is not None else None,
        _chk(table, torch.float32, "silu.table") if table is not None else None,
        _chk(y, torch.int64, "silu.y") if y is not None else None,
        rows, D, _chk(hi, torch.bfloat16, "silu.hi"), _chk(lo, torch.bfloat16, "silu.lo"), _stream()),
        "osudit_silu_split")


def check_labels(y, table_rows: int):
    """Device-side assert that every label indexes the embedding table (the launch traps otherwise)."""
    lib = _lib.load()
    _lib.check(lib.osudit_check_labels(_chk(y, torch.int64, "labels.y"), y.numel(), int(table_rows), _stream()),
               "osudit_check_labels")


def split_bf16(a, need_lo=True):
    a = a.detach().contiguous()
    hi = torch.empty(a.shape, dtype=torch.bfloat16, device=a.device)
    lo = torch.empty_like(hi) if need_lo else None
    lib = _lib.load()
    _lib.check(lib.osudit_split_bf16(_chk(a, torch.float32, "split.a"), a.numel(), hi.data_ptr(),
                                     lo.data_ptr() if need_lo else None, _stream()), "osudit_split_bf16")
    return hi, lo


def diffusion_step(model_out, x, noise, t, coef_table, cfg_half, cfg_scale, clip, phase, sample,
                   pred_xstart, x0_in=None, mean=None, log_variance=None):
    B, _, T = x.shape
    lib = _lib.load()
    _lib.check(lib.osudit_diffusion_step(
        _chk(model_out, torch.float32, "step.model_out"), _chk(x, torch.float32, "step.x"),
        _chk(noise, torch.float32, "step.noise") if noise is not None else None,
        _chk(x0_in, torch.float32, "step.x0_in") if x0_in is not None else None,
        _chk(t, torch.int64, "step.t"), _chk(coef_table, torch.float32, "step.coef"), B, T, cfg_half,
        float(cfg_scale), int(bool(clip)), phase,
        _chk(sample, torch.float32, "step.sample") if sample is not None else None,
        _chk(pred_xstart, torch.float32, "step.pred_xstart"),
        _chk(mean, torch.float32, "step.mean") if mean is not None else None,
        _chk(log_variance, torch.float32, "step.log_variance") if log_variance is not None else None,
        _stream()), "osudit_diffusion_step")


def cfg_combine(model_out, cfg_scale, out):
    B, _, T = model_out.shape
    lib = _lib.load()
    _lib.check(lib.osudit_cfg_combine(_chk(model_out, torch.float32, "cfg.in"), B, T, float(cfg_scale),
                                      _chk(out, torch.float32, "cfg.out"), _stream()), "osudit_cfg_combine")
    return out


def q_sample(x0, noise, t, sqrt_acp, sqrt_1m_acp, out):
    B = x0.shape[0]
    lib = _lib.load()
    _lib.check(lib.osudit_q_sample(_chk(x0, torch.float32, "q.x0"), _chk(noise, torch.float32, "q.noise"),
                                   _chk(t, torch.int64, "q.t"), _chk(sqrt_acp, torch.float32, "q.a"),
                                   _chk(sqrt_1m_acp, torch.float32, "q.b"), B, x0.numel() // B,
                                   _chk(out, torch.float32, "q.out"), _stream()), "osudit_q_sample")
    return out


# ------------------------------------------------------------------------------ backward ops
def _pad8(n: int) -> int:
    return (n + 7) // 8 * 8


def transpose(a: torch.Tensor) -> torch.Tensor:
    """[R, C] (bf16 or fp32) -> bf16 [C, pad8(R)] with zero padding: a K-major GEMM operand whose
    K dimension is the (padded) row count of `a`."""
    R, C = a.shape
    ld = _pad8(R)
    out = torch.empty(C, ld, dtype=torch.bfloat16, device=a.device)
    if ld != R:
        out[:, R:].zero_()
    lib = _lib.load()
    is_f32 = a.dtype == torch.float32
    _lib.check(lib.osudit_transpose_bf16(_chk(a, torch.float32 if is_f32 else torch.bfloat16, "transpose.in"),
                                         out.data_ptr(), R, C, ld, int(is_f32), _stream()), "osudit_transpose_bf16")
    return out


def gemm_wgrad(dy, x, out):
    """out[M,N] (fp32) += dy[rows,M]^T @ x[rows,N] (bf16, token-major; no transposes)."""
    rows, M = dy.shape
    N = x.shape[1]
    if x.shape[0] != rows or tuple(out.shape) != (M, N):
        raise _lib.OsuditError(f"gemm_wgrad: shape mismatch {tuple(dy.shape)} {tuple(x.shape)} -> {tuple(out.shape)}")
    lib = _lib.load()
    _lib.check(lib.osudit_gemm_wgrad(_chk(dy, torch.bfloat16, "wgrad.dy"), dy.stride(0),
                                     _chk(x, torch.bfloat16, "wgrad.x"), x.stride(0), rows, M, N,
                                     _chk(out, torch.float32, "wgrad.out"), out.stride(0), _stream()),
               "osudit_gemm_wgrad")
    return out


def gelu(pre, out, dy=None):
    lib = _lib.load()
    _lib.check(lib.osudit_gelu(_chk(pre, torch.bfloat16, "gelu.pre"),
                               _chk(dy, torch.bfloat16, "gelu.dy") if dy is not None else None,
                               _chk(out, torch.bfloat16, "gelu.out"), pre.numel(), int(dy is not None), _stream()),
               "osudit_gelu")
    return out


def colsum(a, out):
    """out[N] (fp32) += column sums of a[rows, N]."""
    rows, N = a.shape
    lib = _lib.load()
    is_f32 = a.dtype == torch.float32
    _lib.check(lib.osudit_colsum(_chk(a, torch.float32 if is_f32 else torch.bfloat16, "colsum.in"), int(is_f32),
                                 rows, N, _chk(out, torch.float32, "colsum.out"), _stream()), "osudit_colsum")
    return out


def gelu_bwd(pre, dy, out, dbias=None):
    """out = dy * gelu'(pre) (may alias dy); dbias (fp32 [N], accumulated) += column sums of out."""
    rows, N = pre.shape
    lib = _lib.load()
    _lib.check(lib.osudit_gelu_bwd(_chk(pre, torch.bfloat16, "gelu_bwd.pre"), _chk(dy, torch.bfloat16, "gelu_bwd.dy"),
                                   _chk(out, torch.bfloat16, "gelu_bwd.out"), rows, N,
                                   _chk(dbias, torch.float32, "gelu_bwd.dbias") if dbias is not None else None,
                                   _stream()), "osudit_gelu_bwd")
    return out


def gate_residual_bwd(dx, y, mod, dmod, gate_col, B, T, dy, dbias=None):
    D = dx.shape[1]
    lib = _lib.load()
    _lib.check(lib.osudit_gate_residual_bwd(_chk(dx, torch.float32, "grb.dx"), _chk(y, torch.bfloat16, "grb.y"),
                                            _off(mod, gate_col), _off(dmod, gate_col), mod.stride(0), B, T, D,
                                            _chk(dy, torch.bfloat16, "grb.dy"),
                                            _chk(dbias, torch.float32, "grb.dbias") if dbias is not None else None,
                                            _stream()),
               "osudit_gate_residual_bwd")
    return dy


def ln_modulate_bwd(x, dh, mod, dmod, shift_col, scale_col, B, T, dx, accumulate):
    D = x.shape[1]
    lib = _lib.load()
    _lib.check(lib.osudit_ln_modulate_bwd(_chk(x, torch.float32, "lnb.x"), _chk(dh, torch.bfloat16, "lnb.dh"),
                                          _off(mod, scale_col), _off(dmod, shift_col), _off(dmod, scale_col),
                                          mod.stride(0), B, T, D, _chk(dx, torch.float32, "lnb.dx"),
                                          int(accumulate), _stream()), "osudit_ln_modulate_bwd")
    return dx


def ln_gate_bwd(x, dh, mod, dmod, shift_col, scale_col, B, T, dx, accumulate, y=None, gate_col=None, dy=None,
                dbias=None):
    """ln_modulate_bwd, then (when y is given) gate_residual_bwd on the updated dx, in one pass."""
    D = x.shape[1]
    lib = _lib.load()
    has = y is not None
    _lib.check(lib.osudit_ln_gate_bwd(
        _chk(x, torch.float32, "lgb.x"), _chk(dh, torch.bfloat16, "lgb.dh"), _off(mod, scale_col),
        _off(dmod, shift_col), _off(dmod, scale_col), mod.stride(0), B, T, D, _chk(dx, torch.float32, "lgb.dx"),
        int(accumulate), _chk(y, torch.bfloat16, "lgb.y") if has else None,
        _off(mod, gate_col) if has else None, _off(dmod, gate_col) if has else None,
        _chk(dy, torch.bfloat16, "lgb.dy") if has else None,
        _chk(dbias, torch.float32, "lgb.dbias") if (has and dbias is not None) else None, _stream()),
        "osudit_ln_gate_bwd")
    return dy if has else dx


def final_layer_bwd(x, dout, mod, dmod, shift_col, scale_col, B, T, w, dw, dbias, dx):
    D = x.shape[1]
    lib = _lib.load()
    _lib.check(lib.osudit_final_layer_bwd(_chk(x, torch.float32, "fb.x"), _chk(dout, torch.float32, "fb.dout"),
                                          _off(mod, shift_col), _off(mod, scale_col), _off(dmod, shift_col),
                                          _off(dmod, scale_col), mod.stride(0), B, T, D,
                                          _chk(w, torch.float32, "fb.w"), _chk(dw, torch.float32, "fb.dw"),
                                          _chk(dbias, torch.float32, "fb.dbias"), _chk(dx, torch.float32, "fb.dx"),
                                          _stream()), "osudit_final_layer_bwd")


def silu_bwd(a, ds, dcond, table=None, y=None, dtable=None):
    rows, D = a.shape
    lib = _lib.load()
    _lib.check(lib.osudit_silu_bwd(_chk(a, torch.float32, "sb.a"),
                                   _chk(table, torch.float32, "sb.table") if table is not None else None,
                                   _chk(y, torch.int64, "sb.y") if y is not None else None,
                                   _chk(ds, torch.float32, "sb.ds"), rows, D, _chk(dcond, torch.float32, "sb.dcond"),
                                   _chk(dtable, torch.float32, "sb.dtable") if dtable is not None else None,
                                   _stream()), "osudit_silu_bwd")
    return dcond


def attn_band_bwd(qkv, out, dout, lse, dqkv, B, T, H, head_dim, w_left=-1, w_right=-1, dbias=None):
    """dbias (fp32 [3*H*head_dim], accumulated) receives the in_proj_bias gradient when given."""
    delta = torch.empty(B, H, T, dtype=torch.float32, device=qkv.device)
    lib = _lib.load()
    _lib.check(lib.osudit_attn_band_bwd(_chk(qkv, torch.bfloat16, "ab.qkv"), _chk(out, torch.bfloat16, "ab.out"),
                                        _chk(dout, torch.bfloat16, "ab.dout"), _chk(lse, torch.float32, "ab.lse"),
                                        delta.data_ptr(), _chk(dqkv, torch.bfloat16, "ab.dqkv"), B, T, H, head_dim,
                                        w_left, w_right,
                                        _chk(dbias, torch.float32, "ab.dbias") if dbias is not None else None,
                                        _stream()), "osudit_attn_band_bwd")
    return dqkv


def diffusion_loss(model_out, x0, x_t, noise, t, coef_table, use_l1, term_main, term_vb, dmodel_out):
    B, _, T = x0.shape
    lib = _lib.load()
    _lib.check(lib.osudit_diffusion_loss(_chk(model_out, torch.float32, "loss.out"), _chk(x0, torch.float32, "loss.x0"),
                                         _chk(x_t, torch.float32, "loss.x_t"), _chk(noise, torch.float32, "loss.noise"),
                                         _chk(t, torch.int64, "loss.t"), _chk(coef_table, torch.float32, "loss.coef"),
                                         B, T, int(bool(use_l1)), _chk(term_main, torch.float32, "loss.main"),
                                         _chk(term_vb, torch.float32, "loss.vb"),
                                         _chk(dmodel_out, torch.float32, "loss.dout"), _stream()),
               "osudit_diffusion_loss")


def scale_rows(a, g, out):
    B = a.shape[0]
    lib = _lib.load()
    _lib.check(lib.osudit_scale_rows(_chk(a, torch.float32, "sr.in"), _chk(g, torch.float32, "sr.g"), B,
                                     a.numel() // B, _chk(out, torch.float32, "sr.out"), _stream()),
               "osudit_scale_rows")
    return out
