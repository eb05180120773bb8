"""CUDA-graph replay of one native denoising step.

A step at the headline batch (128 rows x 2048 datapoints) keeps the GPU busy for ~60 ms, so its ~100
launches are free; a single beatmap (what `sample.py` usually runs: 2 rows) is launch-bound instead.
The step = `randn_like` + DiT forward (all libosudit launches into the engine's persistent
workspace) + fused CFG/diffusion update is captured once per (model, shapes, conditioning tensors,
mask, guidance scale, clip flag) and replayed with only `x` and `t` refreshed.  Results are cloned
out of the graph's static buffers because `p_sample_loop_progressive` hands them to the caller.

The captured work is exactly the eager path's; parity with it is bit-exact (tests/test_gpu_model.py).
Set OSUDIT_CUDA_GRAPHS=0 to disable.
"""
from __future__ import annotations

import os

import torch

from . import ops

_ENABLED = os.environ.get("OSUDIT_CUDA_GRAPHS", "1") != "0"
_MAX_ROWS_TOKENS = int(os.environ.get("OSUDIT_GRAPH_MAX_ROW_TOKENS", 1 << 17))  # above: GPU-bound, replay only pins memory
_cache: dict = {}


def eligible(x) -> bool:
    return _ENABLED and x.is_cuda and x.shape[0] * x.shape[-1] <= _MAX_ROWS_TOKENS and \
        not torch.cuda.is_current_stream_capturing()


class StepGraph:
    def __init__(self, diffusion, module, uses_cfg, x, t, kw, clip):
        dev = x.device
        self.tb = diffusion._tables(dev)
        self.module, self.uses_cfg, self.clip = module, uses_cfg, bool(clip)
        self.o, self.c, self.y, self.mask = kw["o"], kw["c"], kw["y"], kw.get("attn_mask")
        self.scale = float(kw.get("cfg_scale", 0.0)) if uses_cfg else 0.0
        self.x = x.clone()
        self.t = t.clone()
        self.t_orig = self.tb["tmap"][self.t]
        self.sample = torch.empty_like(self.x)
        self.x0 = torch.empty_like(self.x)
        self.param_sig = self._sig()
        self.diffusion = diffusion  # keeps id(diffusion) in the cache key unique while cached
        rng = torch.cuda.get_rng_state(dev)  # the warm-up draws must not shift the caller's noise stream
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):  # warm-up: packs weights, sizes workspaces, sets func attributes
            for _ in range(2):
                self._body()
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.set_rng_state(rng, dev)
        # the captured launches hold raw pointers into the engine's workspace, packed weights and mask copy: keep those
        # objects alive for as long as this graph can be replayed, whatever the engine's own caches evict
        self.keep = module.engine().keepalive(self.x.shape[0], self.x.shape[-1], dev, self.mask)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self._body()

    def _sig(self):
        return tuple((p.data_ptr(), p._version) for p in self.module.parameters())

    def _body(self):
        B = self.x.shape[0]
        noise = torch.randn_like(self.x)  # the reference's draw (gaussian_diffusion.py:454)
        half = B // 2 if self.uses_cfg else 0
        eng = self.module.engine()
        keep, eng.reuse_oc_columns = eng.reuse_oc_columns, False  # a replay must rewrite the whole first-layer operand
        try:
            raw = self.module._raw_forward(self.x, self.t_orig, self.o, self.c, self.y, self.mask,
                                           x_rows=half if self.uses_cfg else None)
        finally:
            eng.reuse_oc_columns = keep
        ops.diffusion_step(raw, self.x, noise, self.t, self.tb["step"], half, self.scale, self.clip, 0,
                           self.sample, self.x0)

    def run(self, x, t):
        self.x.copy_(x)
        self.t.copy_(t)
        torch.index_select(self.tb["tmap"], 0, self.t, out=self.t_orig)
        if self.keep[0] is not None:
            self.keep[0]["oc_sig"] = None  # the replay rewrites the shared workspace's o / c columns with ITS o / c
        self.graph.replay()
        return self.sample.clone(), self.x0.clone()


def _key(diffusion, module, uses_cfg, x, kw, clip):
    def ident(v):
        return None if v is None else (v.data_ptr(), v._version, tuple(v.shape))
    return (id(diffusion), id(module), uses_cfg, tuple(x.shape), str(x.device), ident(kw["o"]), ident(kw["c"]),
            ident(kw["y"]), ident(kw.get("attn_mask")), float(kw.get("cfg_scale", 0.0)), bool(clip),
            module.training, getattr(module, "precision", "bf16"))


def release_graphs():
    """Drop every captured sampling step (each one pins its static buffers and the model it was captured for)."""
    _cache.clear()


def step(diffusion, module, uses_cfg, x, t, kw, clip):
    """(sample, pred_xstart) of one reverse step through a cached graph."""
    key = _key(diffusion, module, uses_cfg, x, kw, clip)
    g = _cache.get(key)
    if g is not None and g.param_sig != g._sig():  # weights changed (load_state_dict, training): re-capture
        g = None
    if g is None:
        if len(_cache) >= 4:
            _cache.pop(next(iter(_cache)))
        g = StepGraph(diffusion, module, uses_cfg, x, t, kw, clip)
        _cache[key] = g
    return g.run(x, t)
