"""ctypes binding of libosudit.so (the C ABI declared in include/osudit.h).

This is the whole host<->native boundary: plain pointers, sizes and a stream handle.  There is
no CPU fallback — if the library is missing or a call fails, an exception is raised.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import c_char_p, c_float, c_int, c_int64, c_void_p

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("OSUDIT_LIB") or os.path.join(os.path.dirname(_HERE), "libosudit.so")

_P = c_void_p
_I = c_int
_L = c_int64
_F = c_float

# name -> argtypes (restype is always int unless listed in _RESTYPES)
SIGNATURES = {
    "osudit_version": [],
    "osudit_last_error": [],
    "osudit_set_sm_limit": [_I],
    "osudit_gemm_bf16": [_I, _P, _P, _P, _P, _P, _L, _L, _P, _I, _P, _L, _P],
    "osudit_attn_band": [_P, _P, _I, _I, _I, _I, _I, _I, _P, _I, _P, _P],
    "osudit_gemm_wgrad": [_P, _L, _P, _L, _L, _L, _L, _P, _L, _P],
    "osudit_attn_band_bwd": [_P, _P, _P, _P, _P, _P, _I, _I, _I, _I, _I, _I, _P, _P],
    "osudit_transpose_bf16": [_P, _P, _L, _L, _L, _I, _P],
    "osudit_repack_weights": [_P, _I, _I, _P],
    "osudit_gelu": [_P, _P, _P, _L, _I, _P],
    "osudit_colsum": [_P, _I, _L, _I, _P, _P],
    "osudit_gate_residual_bwd": [_P, _P, _P, _P, _L, _I, _I, _I, _P, _P, _P],
    "osudit_gelu_bwd": [_P, _P, _P, _L, _I, _P, _P],
    "osudit_ln_modulate_bwd": [_P, _P, _P, _P, _P, _L, _I, _I, _I, _P, _I, _P],
    "osudit_ln_gate_bwd": [_P, _P, _P, _P, _P, _L, _I, _I, _I, _P, _I, _P, _P, _P, _P, _P, _P],
    "osudit_final_layer_bwd": [_P, _P, _P, _P, _P, _P, _L, _I, _I, _I, _P, _P, _P, _P, _P],
    "osudit_silu_bwd": [_P, _P, _P, _P, _L, _I, _P, _P, _P],
    "osudit_diffusion_loss": [_P, _P, _P, _P, _P, _P, _I, _I, _I, _P, _P, _P, _P],
    "osudit_scale_rows": [_P, _P, _I, _L, _P, _P],
    "osudit_ln_modulate": [_P, _P, _P, _P, _P, _L, _L, _I, _I, _P, _P, _P],
    "osudit_final_layer": [_P, _P, _P, _P, _P, _L, _L, _I, _I, _P, _P, _I, _P, _P],
    "osudit_embed_xoc": [_P, _P, _P, _P, _F, _F, _I, _I, _I, _I, _P, _P, _P],
    "osudit_embed_x": [_P, _P, _F, _F, _I, _I, _I, _I, _P, _P, _P],
    "osudit_timestep_features": [_P, _P, _I, _P, _P, _P],
    "osudit_silu_split": [_P, _P, _P, _P, _L, _I, _P, _P, _P],
    "osudit_split_bf16": [_P, _L, _P, _P, _P],
    "osudit_check_labels": [_P, _L, _L, _P],
    "osudit_diffusion_step": [_P, _P, _P, _P, _P, _P, _I, _I, _I, _F, _I, _I, _P, _P, _P, _P, _P],
    "osudit_cfg_combine": [_P, _I, _I, _F, _P, _P],
    "osudit_q_sample": [_P, _P, _P, _P, _P, _I, _L, _P, _P],
    "osudit_gemm_bf16_aux": [_P, _L, _P, _L, _L, _L, _L, _P, _I, _P, _L, _P, _L, _P],
    "osudit_gemm_gated_residual": [_P, _L, _P, _L, _L, _L, _L, _P, _P, _L, _L, _P, _L, _P],
    "osudit_gemm_gated_residual_applicable": [_L, _L, _L],
    "osudit_beatmap_features": [_P, _I, _I, _I, _P, _P, _P, _P, _P, _P],
    "osudit_opt_chunk_elems": [],
    "osudit_adamw_ema_step": [_P, _P, _I, _F, _F, _F, _F, _F, _F, _P, _P, _P, _P],
    "osudit_gemm_bf16_splitk": [_I, _P, _P, _P, _P, _P, _L, _L, _P, _I, _P, _L, _P],
    "osudit_split3_bf16": [_P, _L, _I, _I, _P, _P, _P, _P],
    "osudit_ln_modulate_f32": [_P, _P, _P, _P, _P, _L, _L, _I, _I, _P, _P],
    "osudit_final_layer_f32": [_P, _P, _P, _P, _P, _L, _L, _I, _I, _P, _P, _I, _P, _P],
    "osudit_attn_band_f32": [_P, _P, _I, _I, _I, _I, _I, _I, _P, _P],
    "osudit_embed_xoc_f32": [_P, _P, _P, _P, _F, _F, _I, _I, _I, _I, _P, _P],
    "osudit_timestep_features_f32": [_P, _P, _I, _P, _P],
}
_RESTYPES = {"osudit_last_error": c_char_p}

_lib = None


class OsuditError(RuntimeError):
    pass


def load() -> ctypes.CDLL:
    """Load the library once; raise (never fall back) when it is absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise OsuditError(
            f"{LIB_PATH} not found: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "or `make -C osu-diffusion_b200/csrc`. There is no CPU/PyTorch fallback for this path.")
    import torch  # noqa: F401  (loads the CUDA runtime the library links against)
    lib = ctypes.CDLL(LIB_PATH)
    for name, argtypes in SIGNATURES.items():
        fn = getattr(lib, name)
        fn.argtypes = argtypes
        fn.restype = _RESTYPES.get(name, c_int)
    _lib = lib
    return lib


def check(rc: int, what: str) -> None:
    if rc != 0:
        msg = load().osudit_last_error()
        raise OsuditError(f"{what} failed ({rc}): {msg.decode() if msg else '?'}")
