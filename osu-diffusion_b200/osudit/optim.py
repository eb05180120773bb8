"""Fused AdamW + EMA optimizer step — opt-in replacement for the tail of train.py's step.

Reference (train.py:154,258-261,36-45):

    opt = torch.optim.AdamW(model.parameters(), lr=1e-4, weight_decay=0)
    ...
    scaler.step(opt); scaler.update(); update_ema(ema, model.module)

With this module:

    opt = FusedAdamWEMA(model.parameters(), lr=1e-4, weight_decay=0)
    opt.attach_ema(ema, model.module, decay=0.9999)     # update_ema(...) is then dropped from the loop
    ...
    scaler.step(opt); scaler.update()

One launch updates every parameter, both Adam moments and the EMA copy (40 B per parameter instead of ~600
small launches), reads the gradients still multiplied by GradScaler's scale and skips the step on the device
when they contain inf/nan (`_step_supports_amp_scaling`), so `scaler.step` no longer synchronises the host.
The per-parameter state keeps torch.optim.AdamW's keys (`step`, `exp_avg`, `exp_avg_sq`), so `state_dict()` /
`load_state_dict()` interoperate with checkpoints written by the reference (`train.py:287-293`).
"""
from __future__ import annotations

import ctypes

import numpy as np
import torch

from . import _lib, ops


class _Seg(ctypes.Structure):
    _fields_ = [("p", ctypes.c_void_p), ("g", ctypes.c_void_p), ("m", ctypes.c_void_p), ("v", ctypes.c_void_p),
                ("ema", ctypes.c_void_p), ("n", ctypes.c_longlong)]


_SEG_DTYPE = np.dtype([("p", "<u8"), ("g", "<u8"), ("m", "<u8"), ("v", "<u8"), ("ema", "<u8"), ("n", "<i8")])
assert _SEG_DTYPE.itemsize == ctypes.sizeof(_Seg) == 48


class FusedAdamWEMA(torch.optim.Optimizer):
    _step_supports_amp_scaling = True  # GradScaler hands us grad_scale / found_inf instead of syncing

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2):
        if lr < 0 or eps < 0 or not 0 <= betas[0] < 1 or not 0 <= betas[1] < 1 or weight_decay < 0:
            raise ValueError("invalid AdamW hyper-parameter")
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))
        self._ema = {}  # id(param) -> ema tensor
        self._ema_decay = 0.0
        self._tables = {}  # per param group: cached chunk table + pinned / device segment tables

    # ------------------------------------------------------------------ EMA (train.py:36-45)
    def attach_ema(self, ema_model, model, decay=0.9999):
        """From now on step() also does `update_ema(ema_model, model, decay)` for every parameter it updates."""
        ema_params = dict(ema_model.named_parameters())
        self._ema = {}
        for name, p in model.named_parameters():
            e = ema_params[name]
            if e.shape != p.shape or e.dtype != torch.float32 or not e.is_contiguous():
                raise ValueError(f"EMA copy of {name} does not match the parameter")
            self._ema[id(p)] = e
        self._ema_decay = float(decay)
        self._tables.clear()

    # ------------------------------------------------------------------ step
    def _group_tables(self, gi, params):
        key = tuple(id(p) for p in params)
        tb = self._tables.get(gi)
        if tb is None or tb["key"] != key:
            E = _lib.load().osudit_opt_chunk_elems()
            counts = [(p.numel() + E - 1) // E for p in params]
            seg_idx = np.repeat(np.arange(len(params), dtype=np.int32), counts)
            offs = np.concatenate([np.arange(c, dtype=np.int32) for c in counts])
            chunks = torch.from_numpy(np.stack([seg_idx, offs], 1).copy()).to(params[0].device)
            # the segment table changes every step (fresh .grad tensors); it is staged through a small ring of
            # pinned buffers so that the host can run ahead of the device without overwriting a pending copy
            nbytes = len(params) * _SEG_DTYPE.itemsize
            ring = []
            for _ in range(4):
                host = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
                ring.append(dict(host=host, view=host.numpy().view(_SEG_DTYPE), event=None,
                                 dev=torch.empty(nbytes, dtype=torch.uint8, device=params[0].device)))
            tb = dict(key=key, chunks=chunks, nchunks=int(chunks.shape[0]), ring=ring, turn=0)
            self._tables[gi] = tb
        return tb

    @torch.no_grad()
    def step(self, closure=None):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        grad_scale = getattr(self, "grad_scale", None)
        found_inf = getattr(self, "found_inf", None)
        lib = _lib.load()
        for gi, group in enumerate(self.param_groups):
            params = [p for p in group["params"] if p.grad is not None]
            if not params:
                continue
            dev = params[0].device
            for p in params:
                if p.dtype != torch.float32 or p.grad.dtype != torch.float32 or not p.is_contiguous() \
                        or not p.grad.is_contiguous() or p.grad.is_sparse or not p.is_cuda:
                    raise RuntimeError("FusedAdamWEMA needs contiguous fp32 CUDA parameters and dense fp32 gradients")
                st = self.state[p]
                if not st:
                    st["step"] = torch.zeros((), dtype=torch.float32, device=dev)
                    st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                    st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
            # one device-side step counter per group drives the kernel; every parameter's `step` entry aliases it while
            # training and is cloned per parameter by state_dict() (torch.optim.AdamW keeps one tensor per parameter and
            # its foreach path would otherwise advance a shared tensor once per parameter after a load)
            step_t = self.state[params[0]]["step"]
            if not (torch.is_tensor(step_t) and step_t.is_cuda and step_t.dtype == torch.float32):
                step_t = torch.as_tensor(float(step_t), dtype=torch.float32, device=dev)
            for p in params:
                self.state[p]["step"] = step_t
            tb = self._group_tables(gi, params)
            slot = tb["ring"][tb["turn"] % len(tb["ring"])]
            tb["turn"] += 1
            if slot["event"] is not None:
                slot["event"].synchronize()  # the copy issued from this buffer four steps ago
            v = slot["view"]
            for i, p in enumerate(params):
                st = self.state[p]
                e = self._ema.get(id(p))
                v[i] = (p.data_ptr(), p.grad.data_ptr(), st["exp_avg"].data_ptr(), st["exp_avg_sq"].data_ptr(),
                        e.data_ptr() if e is not None else 0, p.numel())
            slot["dev"].copy_(slot["host"], non_blocking=True)
            slot["event"] = torch.cuda.Event()
            slot["event"].record()
            b1, b2 = group["betas"]
            _lib.check(lib.osudit_adamw_ema_step(
                slot["dev"].data_ptr(), tb["chunks"].data_ptr(), tb["nchunks"], float(group["lr"]), float(b1), float(b2),
                float(group["eps"]), float(group["weight_decay"]), self._ema_decay, step_t.data_ptr(),
                grad_scale.data_ptr() if grad_scale is not None else None,
                found_inf.data_ptr() if found_inf is not None else None, ops._stream()), "osudit_adamw_ema_step")
            # the kernel wrote through raw pointers: tell autograd's version counters, which every packed-weight cache
            # (engine.PackedWeights, fp32.PackedWeightsF32, train.TrainWeights, graphs.StepGraph) keys on
            touched = list(params)
            touched += [self._ema[id(p)] for p in params if id(p) in self._ema]
            torch.autograd.graph.increment_version(touched)
        return loss

    def state_dict(self):
        """torch.optim.AdamW's layout, with one private `step` tensor per parameter (the live state shares one per
        group), so the checkpoint loads into the reference's optimizer and back (train.py:287-293)."""
        sd = super().state_dict()  # its per-parameter dicts ARE the live ones: build new dicts, never edit those
        sd["state"] = {k: ({**st, "step": st["step"].clone()} if torch.is_tensor(st.get("step")) else dict(st))
                       for k, st in sd["state"].items()}
        return sd
