"""Training-time forward (activations kept) and backward of the DiT.

The reference gets its gradients from PyTorch autograd over `DiT.forward` (train.py:249-257).  Here the
forward is one fixed schedule of libosudit launches and the backward another, cut into depth + 2 pieces
(final layer, blocks, embedders) that hang off a chain of identity autograd nodes: autograd, DDP and the
optimizer see the same leaf parameters and receive fp32 gradients for them group by group, nothing else
runs in torch.

Data-gradient GEMMs use transposed weight copies, weight-gradient GEMMs use transposed activations
(`ops.transpose` pads the token dimension to a multiple of 8 with zeros); both then run on the same
tcgen05 GEMM as the forward.  Gradients flow in bf16 between GEMMs and in fp32 along the residual
stream and into every parameter.
"""
from __future__ import annotations

import os
import weakref

import torch

from . import ops
from .engine import FREQ_SEQ, FREQ_T, check_label_range, classify_mask


_SEG_DTYPE = None


def _seg_dtype():
    global _SEG_DTYPE
    if _SEG_DTYPE is None:
        import numpy as np
        _SEG_DTYPE = np.dtype([("src", "<u8"), ("copy", "<u8"), ("trans", "<u8"), ("hi", "<u8"), ("lo", "<u8"),
                               ("ld_trans", "<i8"), ("rows", "<i4"), ("cols", "<i4"), ("tile0", "<i4"), ("tiles_x", "<i4")])
        assert _SEG_DTYPE.itemsize == 64  # struct RepackSeg, csrc/backward.cu
    return _SEG_DTYPE


class TrainWeights:
    """bf16 / split-bf16 / transposed copies of the parameters, rebuilt when a version changes.

    The destinations are allocated once per parameter set and ALL of them are refreshed by one launch
    (`osudit_repack_weights`, a table of 64 x 64 tiles over every weight matrix): same-layout bf16 for the forward
    GEMMs, transposed bf16 (zero-padded to a multiple of 8 columns) for the data-gradient GEMMs, hi / lo splits for the
    precision-critical small GEMMs; the adaLN Linears of all blocks land directly in their slices of one stacked
    matrix (no torch.cat)."""

    def __init__(self):
        self.sig = None
        self.ptr_sig = None

    def _build(self, model):
        import numpy as np
        dev = next(model.parameters()).device
        bf = lambda *s: torch.empty(*s, dtype=torch.bfloat16, device=dev)  # noqa: E731
        zbf = lambda *s: torch.zeros(*s, dtype=torch.bfloat16, device=dev)  # noqa: E731
        segs = []

        def add(p, copy=None, trans=None, trans_off=0, hi=None, lo=None):
            w = p.detach()
            if w.dtype != torch.float32 or not w.is_contiguous() or not w.is_cuda:
                raise RuntimeError("the native training path needs contiguous fp32 CUDA parameters")
            rows, cols = w.shape
            segs.append((w.data_ptr(), copy.data_ptr() if copy is not None else 0,
                         trans.data_ptr() + 2 * trans_off if trans is not None else 0,
                         hi.data_ptr() if hi is not None else 0, lo.data_ptr() if lo is not None else 0,
                         trans.stride(0) if trans is not None else 0, rows, cols))

        def split_of(p):
            hi, lo = bf(*p.shape), bf(*p.shape)
            return hi, lo

        def trans_of(p):  # [out, in] -> [in, pad8(out)], the pad columns stay zero
            return zbf(p.shape[1], ops._pad8(p.shape[0]))

        first, t0, t2 = model.xoc_embedder.mlp[0].weight, model.t_embedder.mlp[0].weight, model.t_embedder.mlp[2].weight
        self.first_w, self.t0_w, self.t2_w = split_of(first), split_of(t0), split_of(t2)
        self.t2_wt = trans_of(t2)
        add(first, hi=self.first_w[0], lo=self.first_w[1])
        add(t0, hi=self.t0_w[0], lo=self.t0_w[1])
        add(t2, hi=self.t2_w[0], lo=self.t2_w[1], trans=self.t2_wt)
        mods = [b.adaLN_modulation[1] for b in model.blocks] + [model.final_layer.adaLN_modulation[1]]
        total = sum(m.weight.shape[0] for m in mods)
        D = mods[0].weight.shape[1]
        self.mod_w = (bf(total, D), bf(total, D))
        self.mod_wt = zbf(D, ops._pad8(total))
        r = 0
        for m in mods:
            n = m.weight.shape[0]
            add(m.weight, hi=self.mod_w[0][r:r + n], lo=self.mod_w[1][r:r + n], trans=self.mod_wt, trans_off=r)
            r += n
        self._mod_biases = [m.bias for m in mods]
        self.blocks = []
        for blk in model.blocks:
            d = {}
            for name, p in (("qkv", blk.attn.in_proj_weight), ("out", blk.attn.out_proj.weight),
                            ("fc1", blk.mlp.fc1.weight), ("fc2", blk.mlp.fc2.weight)):
                d[name + "_w"], d[name + "_wt"] = bf(*p.shape), trans_of(p)
                add(p, copy=d[name + "_w"], trans=d[name + "_wt"])
            self.blocks.append(d)
        tab = np.zeros(len(segs), dtype=_seg_dtype())
        tile0 = 0
        for i, (src, copy, trans, hi, lo, ldt, rows, cols) in enumerate(segs):
            tx = (cols + 63) // 64
            tab[i] = (src, copy, trans, hi, lo, ldt, rows, cols, tile0, tx)
            tile0 += tx * ((rows + 63) // 64)
        self._table = torch.from_numpy(tab.view(np.uint8).copy()).to(dev)
        self._nseg, self._tiles = len(segs), tile0

    def refresh(self, model):
        pfp = model.xoc_embedder.playfield_size  # frozen: read it back (a host sync) only when it changes
        pf_sig = (pfp.data_ptr(), pfp._version)
        if pf_sig != getattr(self, "pf_sig", None):
            self.pf = [float(v) for v in pfp.detach().cpu()]
            self.pf_sig = pf_sig
        sig = tuple((p.data_ptr(), p._version) for p in model.parameters())
        if sig == self.sig:
            return self
        ptr_sig = tuple(s[0] for s in sig)
        if ptr_sig != self.ptr_sig:  # first use, or the parameters moved (.to(), load into new storage): new table
            if torch.cuda.is_current_stream_capturing():
                raise RuntimeError("TrainWeights: the re-pack table must exist before a CUDA graph is captured")
            self._build(model)
            self.ptr_sig = ptr_sig
        lib = ops._lib.load()
        ops._lib.check(lib.osudit_repack_weights(self._table.data_ptr(), self._nseg, self._tiles, ops._stream()),
                       "osudit_repack_weights")
        self.mod_b = torch.cat([b.detach() for b in self._mod_biases], 0).float().contiguous()
        self.sig = sig
        return self


def _gemm3(a_hi, a_lo, w, bias, out):
    w_hi, w_lo = w
    return ops.gemm([a_hi, a_lo, a_hi], [w_hi, w_hi, w_lo], bias, ops.EPI_F32, out)


def _e(*shape, dtype=torch.bfloat16, device=None):
    return torch.empty(*shape, dtype=dtype, device=device)


def forward_train(model, tw: TrainWeights, x, t, o, c, y, attn_mask):
    """DiT.forward (models.py:306-325) keeping what the backward needs. Returns (out [B,4,T], saved)."""
    B, T = o.shape
    D, H, depth = model.hidden_size, model.num_heads, len(model.blocks)
    hd, E, dev = D // H, model.context_size, o.device
    if hd not in (64, 72):
        raise NotImplementedError("the native attention covers head_dim 64 (DiT-S/B/L) and 72 (DiT-XL)")
    rows = B * T
    spec = classify_mask(attn_mask, T)
    if spec.generic is not None:
        raise NotImplementedError("training with a generic attention mask is not built (None or a band only)")
    eng = model.engine()
    f32 = lambda p: p.detach().float().contiguous()  # noqa: E731
    S = dict(B=B, T=T, spec=spec, y=y)

    kin = 3 * FREQ_SEQ + E
    a_hi, a_lo = _e(rows, kin, device=dev), _e(rows, kin, device=dev)
    ops.embed_xoc(x, o, c, eng.freqs(FREQ_SEQ // 2, dev), tw.pf[0], tw.pf[1], B, a_hi, a_lo)
    xa = _e(rows, D, dtype=torch.float32, device=dev)
    _gemm3(a_hi, a_lo, tw.first_w, f32(model.xoc_embedder.mlp[0].bias), xa)
    S["a_hi"] = a_hi

    # conditioning: mod = Linear_all(SiLU(t_emb + y_emb))
    tf_hi, tf_lo = _e(B, FREQ_T, device=dev), _e(B, FREQ_T, device=dev)
    ops.timestep_features(t, eng.freqs(FREQ_T // 2, dev), tf_hi, tf_lo)
    h1t = _e(B, D, dtype=torch.float32, device=dev)
    _gemm3(tf_hi, tf_lo, tw.t0_w, f32(model.t_embedder.mlp[0].bias), h1t)
    s1_hi, s1_lo = _e(B, D, device=dev), _e(B, D, device=dev)
    ops.silu_split(h1t, s1_hi, s1_lo)
    temb = _e(B, D, dtype=torch.float32, device=dev)
    _gemm3(s1_hi, s1_lo, tw.t2_w, f32(model.t_embedder.mlp[2].bias), temb)
    table = f32(model.y_embedder.embedding_table.weight)
    c_hi, c_lo = _e(B, D, device=dev), _e(B, D, device=dev)
    check_label_range(y, table.shape[0], sync=False)  # device-side assert only: no host sync on the training path
    ops.silu_split(temb, c_hi, c_lo, table=table, y=y)
    mod = _e(B, (6 * depth + 2) * D, dtype=torch.float32, device=dev)
    _gemm3(c_hi, c_lo, tw.mod_w, tw.mod_b, mod)
    S.update(tf_hi=tf_hi, h1t=h1t, s1_hi=s1_hi, temb=temb, table=table, c_hi=c_hi, mod=mod)

    saved_blocks = []
    y2 = None
    for i, (blk, bw) in enumerate(zip(model.blocks, tw.blocks)):
        base = 6 * D * i
        h1 = _e(rows, D, device=dev)
        if i == 0:
            ops.ln_modulate(xa, None, mod, 0, base, base + D, T, h1)
        else:
            xa_new = _e(rows, D, dtype=torch.float32, device=dev)
            ops.ln_modulate(xb, y2, mod, base - D, base, base + D, T, h1, x_out=xa_new)
            xa = xa_new
        qkv = _e(rows, 3 * D, device=dev)
        ops.gemm([h1], [bw["qkv_w"]], f32(blk.attn.in_proj_bias), ops.EPI_BF16, qkv)
        att = _e(rows, D, device=dev)
        lse = _e(B, H, T, dtype=torch.float32, device=dev)
        ops.attn_band(qkv, att, B, T, H, hd, spec.w_left, spec.w_right, None, ops.ATTN_AUTO, lse=lse)
        y1 = _e(rows, D, device=dev)
        ops.gemm([att], [bw["out_w"]], f32(blk.attn.out_proj.bias), ops.EPI_BF16, y1)
        xb = _e(rows, D, dtype=torch.float32, device=dev)
        h2 = _e(rows, D, device=dev)
        ops.ln_modulate(xa, y1, mod, base + 2 * D, base + 3 * D, base + 4 * D, T, h2, x_out=xb)
        pre = _e(rows, bw["fc1_w"].shape[0], device=dev)  # receives gelu'(fc1 output): what the backward needs of it
        u = ops.gemm_aux(h2, bw["fc1_w"], f32(blk.mlp.fc1.bias), ops.EPI_BF16_GELU_SAVE, torch.empty_like(pre), pre)
        y2 = _e(rows, D, device=dev)
        ops.gemm([u], [bw["fc2_w"]], f32(blk.mlp.fc2.bias), ops.EPI_BF16, y2)
        saved_blocks.append(dict(xa=xa, h1=h1, qkv=qkv, att=att, lse=lse, y1=y1, xb=xb, h2=h2, pre=pre, u=u, y2=y2))
    fbase = 6 * D * depth
    xf = _e(rows, D, dtype=torch.float32, device=dev)
    ops.ln_modulate(xb, y2, mod, fbase - D, fbase, fbase + D, T, _e(rows, D, device=dev), x_out=xf)
    out = _e(B, 4, T, dtype=torch.float32, device=dev)
    ops.final_layer(xf, None, mod, 0, fbase, fbase + D, T, f32(model.final_layer.linear.weight),
                    f32(model.final_layer.linear.bias), out)
    S.update(blocks=saved_blocks, xf=xf)
    return out, S


class _Zeros:
    """Zero-initialised fp32 gradient buffers of one backward piece, carved out of ONE allocation (one memset instead
    of a fill launch per parameter): the weight-gradient GEMM and the bias column sums accumulate into them.  The
    total is learnt on the piece's first run (which still allocates per tensor)."""
    totals: dict = {}

    def __init__(self, key, dev):
        self.key, self.dev, self.off, self.used = key, dev, 0, 0
        n = _Zeros.totals.get(key, 0)
        self.buf = torch.zeros(n, dtype=torch.float32, device=dev) if n else None

    def __call__(self, *shape):
        n = 1
        for v in shape:
            n *= int(v)
        step = (n + 63) // 64 * 64  # 256-byte aligned slices (TMA reduce-add needs 16)
        self.used += step
        if self.buf is None or self.off + n > self.buf.numel():
            return torch.zeros(*shape, dtype=torch.float32, device=self.dev)
        out = self.buf[self.off:self.off + n].view(*shape)
        self.off += step
        return out

    def close(self):
        _Zeros.totals[self.key] = max(self.used, _Zeros.totals.get(self.key, 0))


def _wgrad(dy, x, z):
    """dW[out, in] = dY[rows, out]^T . X[rows, in], read token-major (no transposes), fp32, into a zeroed buffer of `z`."""
    return ops.gemm_wgrad(dy, x, z(dy.shape[1], x.shape[1]))


def _wgrad_small(dy_f32, x_bf16, z):
    """Same for the per-sample (conditioning) matrices whose row count is the batch size: the fp32
    gradient is rounded to bf16 first."""
    dy_bf, _ = ops.split_bf16(dy_f32, need_lo=False)
    return _wgrad(dy_bf, x_bf16, z)


# ------------------------------------------------------------------------------ backward, in pieces
# The backward runs as depth + 2 pieces — final layer, blocks from last to first, embedders ("head") — each of
# which leaves the gradients of ITS parameters final.  One autograd node per piece (below) hands them to autograd
# as soon as they exist, in the reverse registration order DistributedDataParallel builds its buckets in, so the
# gradient all-reduce of a bucket overlaps the backward of the blocks before it (train.py:152,257).
class _Bwd:
    """Buffers shared by the pieces of one backward pass."""
    __slots__ = ("dx", "dmod", "dy_buf", "dh_buf", "dy2")


def _adaln_grads(S, dmod_cols, lin, z):
    """Gradients of one adaLN Linear (mod = SiLU(cond) W^T + b) from its finished columns of dmod."""
    dm = dmod_cols.contiguous()
    dm_bf, _ = ops.split_bf16(dm, need_lo=False)
    gw = _wgrad(dm_bf, S["c_hi"], z)
    gb = ops.colsum(dm, z(dm.shape[1]))
    return {lin.weight: gw, lin.bias: gb}


def bwd_final(model, tw: TrainWeights, S, dout):
    """Final layer + the gate of the last block's MLP branch.  Returns (shared buffers, {param: grad})."""
    B, T = S["B"], S["T"]
    D, depth = model.hidden_size, len(model.blocks)
    rows, dev, mod = B * T, dout.device, S["mod"]
    z32 = _Zeros(("final", D, depth, B), dev)
    bs, grads = _Bwd(), {}
    bs.dmod = z32(B, mod.shape[1])
    bs.dx = _e(rows, D, dtype=torch.float32, device=dev)
    fbase = 6 * D * depth
    fl = model.final_layer
    grads[fl.linear.weight], grads[fl.linear.bias] = z32(4, D), z32(4)
    ops.final_layer_bwd(S["xf"], dout.contiguous(), mod, bs.dmod, fbase, fbase + D, B, T,
                        fl.linear.weight.detach().float().contiguous(), grads[fl.linear.weight],
                        grads[fl.linear.bias], bs.dx)
    grads.update(_adaln_grads(S, bs.dmod[:, fbase:fbase + 2 * D], fl.adaLN_modulation[1], z32))
    # Bias gradients are the column sums of the branch gradients and are accumulated by the kernels that
    # produce those.  Along the residual stream each LayerNorm backward is fused with the gated-residual
    # backward that follows it (ops.ln_gate_bwd); only the very first gate (last block's MLP) stands alone.
    bs.dy_buf, bs.dh_buf = _e(rows, D, device=dev), _e(rows, D, device=dev)
    last = model.blocks[depth - 1]
    grads[last.mlp.fc2.bias] = z32(D)  # handed out by the last block's piece
    bs.dy2 = ops.gate_residual_bwd(bs.dx, S["blocks"][depth - 1]["y2"], mod, bs.dmod, 6 * D * (depth - 1) + 5 * D,
                                   B, T, bs.dy_buf, dbias=grads[last.mlp.fc2.bias])
    z32.close()
    return bs, grads


def bwd_block(model, tw: TrainWeights, S, bs, i, carry):
    """Backward of block i.  `carry` holds fc2.bias of this block (accumulated by the previous piece); the
    returned dict covers this block's parameters, and the new carry the fc2.bias of block i-1."""
    B, T, spec = S["B"], S["T"], S["spec"]
    D, H = model.hidden_size, model.num_heads
    rows, dev, mod, dmod, dx = B * T, bs.dx.device, S["mod"], bs.dmod, bs.dx
    z32 = _Zeros(("block", D, model.blocks[0].mlp.fc1.weight.shape[0], i > 0), dev)
    blk, bw, sv = model.blocks[i], tw.blocks[i], S["blocks"][i]
    base = 6 * D * i
    hidden = sv["pre"].shape[1]
    grads = {blk.mlp.fc2.bias: carry}
    dy2 = bs.dy2
    # ---- MLP branch: x_out = xb + gate_mlp * y2 (dy2 = gate_mlp * dx is already in dy_buf)
    grads[blk.mlp.fc2.weight] = _wgrad(dy2, sv["u"], z32)
    # d pre = (dy2 W2) * gelu'(pre): the saved derivative is applied in the data-gradient GEMM's epilogue
    dpre = ops.gemm_aux(dy2, bw["fc2_wt"], None, ops.EPI_BF16_DGELU, _e(rows, hidden, device=dev), sv["pre"])
    grads[blk.mlp.fc1.bias] = ops.colsum(dpre, z32(hidden))
    grads[blk.mlp.fc1.weight] = _wgrad(dpre, sv["h2"], z32)
    dh2 = ops.gemm([dpre], [bw["fc1_wt"]], None, ops.EPI_BF16, bs.dh_buf)
    # ---- LN2 backward into dx, then the attention branch's gate: xb = xa + gate_msa * y1
    grads[blk.attn.out_proj.bias] = z32(D)
    dy1 = ops.ln_gate_bwd(sv["xb"], dh2, mod, dmod, base + 3 * D, base + 4 * D, B, T, dx, True,
                          y=sv["y1"], gate_col=base + 2 * D, dy=bs.dy_buf, dbias=grads[blk.attn.out_proj.bias])
    grads[blk.attn.out_proj.weight] = _wgrad(dy1, sv["att"], z32)
    datt = ops.gemm([dy1], [bw["out_wt"]], None, ops.EPI_BF16, bs.dh_buf)
    grads[blk.attn.in_proj_bias] = z32(3 * D)
    dqkv = ops.attn_band_bwd(sv["qkv"], sv["att"], datt, sv["lse"], _e(rows, 3 * D, device=dev), B, T, H,
                             D // H, spec.w_left, spec.w_right, dbias=grads[blk.attn.in_proj_bias])
    grads[blk.attn.in_proj_weight] = _wgrad(dqkv, sv["h1"], z32)
    dh1 = ops.gemm([dqkv], [bw["qkv_wt"]], None, ops.EPI_BF16, bs.dh_buf)
    # ---- LN1 backward into dx, then the previous block's MLP gate
    new_carry = None
    if i > 0:
        new_carry = z32(D)
        bs.dy2 = ops.ln_gate_bwd(sv["xa"], dh1, mod, dmod, base, base + D, B, T, dx, True,
                                 y=S["blocks"][i - 1]["y2"], gate_col=base - D, dy=bs.dy_buf, dbias=new_carry)
    else:
        ops.ln_gate_bwd(sv["xa"], dh1, mod, dmod, base, base + D, B, T, dx, True)
    # every column of this block's slice of dmod is final now (its MLP gate was accumulated one piece earlier)
    grads.update(_adaln_grads(S, dmod[:, base:base + 6 * D], blk.adaLN_modulation[1], z32))
    z32.close()
    return grads, new_carry


def bwd_head(model, tw: TrainWeights, S, bs):
    """First layer (no gradient to the inputs) and the conditioning path: s = SiLU(temb + table[y]),
    temb = SiLU(tf W0^T + b0) W2^T + b2."""
    B = S["B"]
    D = model.hidden_size
    dev, dx, dmod = bs.dx.device, bs.dx, bs.dmod
    z32 = _Zeros(("head", D, model.y_embedder.embedding_table.weight.shape[0], model.context_size), dev)
    grads = {}
    first = model.xoc_embedder.mlp[0]
    grads[first.weight] = _wgrad_small(dx, S["a_hi"], z32)
    grads[first.bias] = ops.colsum(dx, z32(D))
    dmod_bf, _ = ops.split_bf16(dmod, need_lo=False)
    ds = ops.gemm([dmod_bf], [tw.mod_wt], None, ops.EPI_F32, _e(B, D, dtype=torch.float32, device=dev))
    table_p = model.y_embedder.embedding_table.weight
    grads[table_p] = z32(*table_p.shape)
    dcond = ops.silu_bwd(S["temb"], ds, torch.empty_like(ds), table=S["table"], y=S["y"], dtable=grads[table_p])
    t0, t2 = model.t_embedder.mlp[0], model.t_embedder.mlp[2]
    dcond_bf, _ = ops.split_bf16(dcond, need_lo=False)
    grads[t2.weight] = _wgrad(dcond_bf, S["s1_hi"], z32)
    grads[t2.bias] = ops.colsum(dcond, z32(D))
    ds1 = ops.gemm([dcond_bf], [tw.t2_wt], None, ops.EPI_F32, torch.empty_like(ds))
    dh1t = ops.silu_bwd(S["h1t"], ds1, torch.empty_like(ds1))
    grads[t0.weight] = _wgrad_small(dh1t, S["tf_hi"], z32)
    grads[t0.bias] = ops.colsum(dh1t, z32(D))
    z32.close()
    return grads


def param_groups(model):
    """(head, [block 0..L-1], final): the parameters each backward piece is responsible for."""
    head = [p for m in (model.xoc_embedder, model.t_embedder, model.y_embedder) for p in m.parameters()]
    return head, [list(b.parameters()) for b in model.blocks], list(model.final_layer.parameters())


def _pick(grads, params):
    return [grads.get(p) if p.requires_grad else None for p in params]


def backward_train(model, tw: TrainWeights, S, dout):
    """All pieces in order: gradients of every trainable parameter in `model.parameters()` order."""
    bs, grads = bwd_final(model, tw, S, dout)
    carry = grads.pop(model.blocks[-1].mlp.fc2.bias)
    for i in reversed(range(len(model.blocks))):
        g, carry = bwd_block(model, tw, S, bs, i, carry)
        grads.update(g)
    grads.update(bwd_head(model, tw, S, bs))
    return _pick(grads, model.parameters())


class TrainGraph:
    """CUDA-graph replay of one training step's model work for a fixed (model, batch shape, mask).

    A config-3 step (DiT-B, 256 x 128 datapoints) is ~370 launches of 10-300 us each plus ~300 small
    allocations: eager, the host needs longer to enqueue them than the GPU needs to run them.  The graphs are
    captured once — (weight re-pack + forward_train), then one per backward piece, all sharing one memory pool so
    the saved activations stay where the backward graphs expect them — and replayed every step with only the inputs
    and the incoming output gradient copied into static buffers.  The weight copies are rebuilt INSIDE the forward
    graph from the live fp32 parameters (same addresses every step), so optimizer updates need no re-capture.
    The returned gradient tensors are the graphs' static buffers; autograd / DDP copy out of them (they never
    take ownership because this object also references them).
    """

    def __init__(self, model, x, t, o, c, y, attn_mask):
        dev = o.device
        self.model, self.mask = model, attn_mask
        self.inputs = [v.clone() for v in (x, t, o, c, y)]
        self.tw = TrainWeights()
        self.sig = self.signature(model)
        self.pending = None  # weakref to the call whose backward has not finished yet
        head, blocks, final = param_groups(model)
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side), torch.no_grad():  # eager warm-up: host-side caches, function attributes
            self.tw.refresh(model)
            out, S = forward_train(model, self.tw, *self.inputs, attn_mask)
            backward_train(model, self.tw, S, torch.zeros_like(out))
            del out, S
        torch.cuda.current_stream(dev).wait_stream(side)
        self.g_fwd = torch.cuda.CUDAGraph()
        with torch.no_grad(), torch.cuda.graph(self.g_fwd):
            self.tw.sig = None
            self.tw.refresh(model)
            self.out, self.S = forward_train(model, self.tw, *self.inputs, attn_mask)
        pool = self.g_fwd.pool()
        self.dout = torch.zeros_like(self.out)
        self.g_final = torch.cuda.CUDAGraph()
        with torch.no_grad(), torch.cuda.graph(self.g_final, pool=pool):
            self.bs, g = bwd_final(model, self.tw, self.S, self.dout)
            carry = g.pop(model.blocks[-1].mlp.fc2.bias)
            self.grads_final = _pick(g, final)
        self.g_block, self.grads_block = {}, {}
        for i in reversed(range(len(model.blocks))):
            self.g_block[i] = torch.cuda.CUDAGraph()
            with torch.no_grad(), torch.cuda.graph(self.g_block[i], pool=pool):
                g, carry = bwd_block(model, self.tw, self.S, self.bs, i, carry)
                self.grads_block[i] = _pick(g, blocks[i])
        self.g_head = torch.cuda.CUDAGraph()
        with torch.no_grad(), torch.cuda.graph(self.g_head, pool=pool):
            self.grads_head = _pick(bwd_head(model, self.tw, self.S, self.bs), head)
        # Gradients are handed to autograd as fresh views of the static buffers: AccumulateGrad adopts a tensor nobody
        # else references instead of cloning it (131 device copies, 681 MB per DiT-B step).  `p.grad` then aliases a
        # buffer the next replay overwrites, which is only correct if it is gone by then (train.py:260 sets it to None
        # every step); _detach_stale_grads() clones whatever is still there before a replay (gradient accumulation,
        # zero_grad(set_to_none=False)).
        self._grad_ptrs = {g.data_ptr() for gs in [self.grads_final, self.grads_head, *self.grads_block.values()]
                           for g in gs if g is not None}

    def _detach_stale_grads(self):
        for p in self.model.parameters():
            g = p.grad
            if g is not None and g.data_ptr() in self._grad_ptrs:
                p.grad = g.clone()

    @staticmethod
    def _fresh(grads):
        return [None if g is None else g.view_as(g) for g in grads]

    @staticmethod
    def signature(model):
        return tuple((p.data_ptr(), p.requires_grad) for p in model.parameters())

    def busy(self):
        if self.pending is None:
            return False
        if self.pending() is None:  # that forward's graph was dropped without a backward
            self.pending = None
            return False
        return True

    def run_forward(self, call, x, t, o, c, y):
        self._detach_stale_grads()
        for dst, src in zip(self.inputs, (x, t, o, c, y)):
            dst.copy_(src)
        self.g_fwd.replay()
        self.pending = weakref.ref(call)
        return self.out.clone()

    def run_final(self, dout):
        self.dout.copy_(dout)
        self.g_final.replay()
        return self._fresh(self.grads_final)

    def run_block(self, i):
        self.g_block[i].replay()
        return self._fresh(self.grads_block[i])

    def run_head(self):
        self.g_head.replay()
        self.pending = None
        return self._fresh(self.grads_head)


_GRAPHS_ENABLED = os.environ.get("OSUDIT_CUDA_GRAPHS", "1") != "0"
_train_graphs: dict = {}


def release_graphs():
    """Drop every captured training graph (and the activations / gradient buffers its memory pool pins)."""
    _train_graphs.clear()


def _train_graph_for(model, x, t, o, c, y, attn_mask):
    """The cached TrainGraph for this call, a new one, or None (disabled / saved activations still in use)."""
    if not _GRAPHS_ENABLED or torch.cuda.is_current_stream_capturing():
        return None
    spec = classify_mask(attn_mask, o.shape[1])
    key = (id(model), tuple(o.shape), str(o.device), spec.w_left, spec.w_right, spec.generic is not None)
    g = _train_graphs.get(key)
    if g is not None and g.sig != TrainGraph.signature(model):
        g = None  # parameters moved or were (un)frozen: capture again
    if g is None:
        _train_graphs.pop(key, None)
        while len(_train_graphs) >= 2:  # each graph pins its activations: keep at most two shapes alive
            _train_graphs.pop(next(iter(_train_graphs)))
        g = TrainGraph(model, x, t, o, c, y, attn_mask)
        _train_graphs[key] = g
        return g
    return None if g.busy() else g


class _Call:
    """What the autograd nodes of one forward call share."""
    __slots__ = ("model", "tw", "graph", "S", "bs", "carry", "__weakref__")


class _HeadFn(torch.autograd.Function):
    """Runs the whole forward; its backward is the LAST piece (first layer + conditioning path)."""

    @staticmethod
    def forward(ctx, call, x, t, o, c, y, attn_mask, *params):
        model = call.model
        ctx.call = call
        call.graph = _train_graph_for(model, x, t, o, c, y, attn_mask)
        if call.graph is not None:
            return call.graph.run_forward(call, x, t, o, c, y)
        call.tw.refresh(model)
        out, call.S = forward_train(model, call.tw, x, t, o, c, y, attn_mask)
        return out

    @staticmethod
    def backward(ctx, g):
        call = ctx.call
        head = param_groups(call.model)[0]
        with torch.no_grad():
            if call.graph is not None:
                grads = call.graph.run_head()
            else:
                grads = _pick(bwd_head(call.model, call.tw, call.S, call.bs), head)
                call.S = call.bs = None
        return (None,) * 7 + tuple(grads)


class _BlockFn(torch.autograd.Function):
    """Identity in the forward; the backward of block i, returning that block's parameter gradients."""

    @staticmethod
    def forward(ctx, tok, call, i, *params):
        ctx.call, ctx.i = call, i
        return tok.view_as(tok)

    @staticmethod
    def backward(ctx, g):
        call, i = ctx.call, ctx.i
        with torch.no_grad():
            if call.graph is not None:
                grads = call.graph.run_block(i)
            else:
                gd, call.carry = bwd_block(call.model, call.tw, call.S, call.bs, i, call.carry)
                grads = _pick(gd, list(call.model.blocks[i].parameters()))
        return (g, None, None) + tuple(grads)


class _FinalFn(torch.autograd.Function):
    """Identity in the forward; the FIRST backward piece (final layer), which receives d loss / d out."""

    @staticmethod
    def forward(ctx, tok, call, *params):
        ctx.call = call
        return tok.view_as(tok)

    @staticmethod
    def backward(ctx, g):
        call = ctx.call
        model = call.model
        with torch.no_grad():
            if call.graph is not None:
                grads = call.graph.run_final(g.float())
            else:
                call.bs, gd = bwd_final(model, call.tw, call.S, g.float())
                call.carry = gd.pop(model.blocks[-1].mlp.fc2.bias)
                grads = _pick(gd, list(model.final_layer.parameters()))
        return (g, None) + tuple(grads)


def _reserve_sms_for_nccl():
    """Data-parallel training (train.py:152): NCCL's all-reduce kernels run beside the backward, but a persistent
    GEMM grid that fills all 148 SMs leaves them no room and the step time doubles (measured, osudit/ddp.py).  A
    training forward in a multi-rank process therefore caps the persistent grids when nobody has (16 SMs stay free;
    OSUDIT_DDP_RESERVE_SMS overrides, 0 disables) — also for an unmodified train.py that never calls osudit.ddp.wrap."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1):
        return
    reserve = int(os.environ.get("OSUDIT_DDP_RESERVE_SMS", "16"))
    lib = ops._lib.load()
    if reserve > 0 and lib.osudit_set_sm_limit(-1) == 0:
        sms = torch.cuda.get_device_properties(torch.cuda.current_device()).multi_processor_count
        lib.osudit_set_sm_limit(max(sms - reserve, sms // 2))


def dit_forward_train(model, tw, x, t, o, c, y, attn_mask):
    """out = DiT(x, t, o, c, y) as a chain of autograd nodes over one native forward: head -> block 0 -> ... ->
    block L-1 -> final.  Autograd walks it backwards, so parameter gradients are released group by group."""
    _reserve_sms_for_nccl()
    call = _Call()
    call.model, call.tw, call.graph, call.S, call.bs, call.carry = model, tw, None, None, None, None
    head, blocks, final = param_groups(model)
    tok = _HeadFn.apply(call, x, t, o, c, y, attn_mask, *head)
    for i, ps in enumerate(blocks):
        tok = _BlockFn.apply(tok, call, i, *ps)
    return _FinalFn.apply(tok, call, *final)


class LossFunction(torch.autograd.Function):
    """training_losses' arithmetic (gaussian_diffusion.py:822-870): values + d loss / d model_out."""

    @staticmethod
    def forward(ctx, model_out, x0, x_t, noise, t, coef, use_l1):
        B = x0.shape[0]
        main = torch.empty(B, dtype=torch.float32, device=x0.device)
        vb = torch.empty_like(main)
        dunit = torch.empty_like(model_out)
        ops.diffusion_loss(model_out.contiguous(), x0, x_t, noise, t, coef, use_l1, main, vb, dunit)
        ctx.save_for_backward(dunit)
        loss = main + vb  # B-element add; everything per datapoint happened in the kernel
        ctx.mark_non_differentiable(main, vb)
        return loss, main, vb

    @staticmethod
    def backward(ctx, g_loss, g_main, g_vb):
        (dunit,) = ctx.saved_tensors
        g = g_loss.contiguous().float()
        return ops.scale_rows(dunit, g, torch.empty_like(dunit)), None, None, None, None, None, None
