"""Drop-in `models` module: same `DiT_models` registry, constructor kwargs, parameter tree,
`state_dict` layout and `forward` / `forward_with_cfg` signatures as the reference
(/root/reference/models.py:243-343,410-431), with the arithmetic done by libosudit.so.

Only the parameter *containers* are torch modules (so `load_state_dict`, `deepcopy`, `.to()`,
DDP wrapping and optimizers see exactly the reference's names, shapes and registration order —
SURVEY.md F10); none of their `forward`s is on the product path.  There is no CPU fallback:
calling the model on CPU tensors raises.
"""
from __future__ import annotations

import os

import torch
import torch.nn as nn

from osudit import ops
from osudit.engine import DiTEngine, check_inputs


class _Holder(nn.Module):
    """A module that only owns parameters / sub-modules."""

    def forward(self, *a, **k):  # pragma: no cover
        raise RuntimeError("parameter container; the computation runs in libosudit.so")


class _Attention(_Holder):
    # same own-parameter names and order as nn.MultiheadAttention (models.py:130-135)
    def __init__(self, hidden, heads):
        super().__init__()
        self.num_heads = heads
        self.in_proj_weight = nn.Parameter(torch.empty(3 * hidden, hidden))
        self.in_proj_bias = nn.Parameter(torch.zeros(3 * hidden))
        self.out_proj = nn.Linear(hidden, hidden, bias=True)
        # drawn here, after out_proj's own constructor draws, as nn.MultiheadAttention does: the global generator
        # is consumed in the reference's order, so a seeded construction yields the reference's initial weights
        nn.init.xavier_uniform_(self.in_proj_weight)


class _Mlp(_Holder):  # models.py:83-119
    def __init__(self, hidden, mlp_hidden):
        super().__init__()
        self.fc1 = nn.Linear(hidden, mlp_hidden, bias=True)
        self.fc2 = nn.Linear(mlp_hidden, hidden, bias=True)


class _Block(_Holder):  # models.py:122-149
    def __init__(self, hidden, heads, mlp_ratio):
        super().__init__()
        self.attn = _Attention(hidden, heads)
        self.mlp = _Mlp(hidden, int(hidden * mlp_ratio))
        self.adaLN_modulation = nn.Sequential(nn.SiLU(), nn.Linear(hidden, 6 * hidden, bias=True))


class _FinalLayer(_Holder):  # models.py:178-190
    def __init__(self, hidden, out_channels):
        super().__init__()
        self.linear = nn.Linear(hidden, out_channels, bias=True)
        self.adaLN_modulation = nn.Sequential(nn.SiLU(), nn.Linear(hidden, 2 * hidden, bias=True))


class _FirstLayer(_Holder):  # models.py:199-225
    def __init__(self, hidden, context_size, in_channels, frequency_embedding_size=128):
        super().__init__()
        self.frequency_embedding_size = frequency_embedding_size
        self.mlp = nn.Sequential(nn.Linear(
            in_channels * frequency_embedding_size + frequency_embedding_size + context_size,
            hidden, bias=True))
        self.playfield_size = nn.Parameter(torch.tensor((512, 384), dtype=torch.float32),
                                           requires_grad=False)


class _TimestepEmbedder(_Holder):  # models.py:21-33
    def __init__(self, hidden, frequency_embedding_size=256):
        super().__init__()
        self.frequency_embedding_size = frequency_embedding_size
        self.mlp = nn.Sequential(nn.Linear(frequency_embedding_size, hidden, bias=True), nn.SiLU(),
                                 nn.Linear(hidden, hidden, bias=True))


class _LabelEmbedder(_Holder):  # models.py:41-54
    def __init__(self, num_classes, hidden, dropout_prob):
        super().__init__()
        self.embedding_table = nn.Embedding(num_classes + (dropout_prob > 0), hidden)
        self.num_classes = num_classes
        self.dropout_prob = dropout_prob

    def token_drop(self, labels, force_drop_ids=None):  # models.py:56-67
        if force_drop_ids is None:
            drop = torch.rand(labels.shape[0], device=labels.device) < self.dropout_prob
        else:
            drop = force_drop_ids == 1
        return torch.where(drop, self.num_classes, labels)


class DiT(nn.Module):
    """adaLN-Zero DiT over beatmap datapoint sequences (reference models.py:238-343)."""

    def __init__(self, in_channels=2, context_size=142, hidden_size=1152, depth=28, num_heads=16,
                 mlp_ratio=4.0, class_dropout_prob=0.1, num_classes=1000, learn_sigma=True):
        super().__init__()
        if in_channels != 2 or not learn_sigma:
            raise NotImplementedError("the native path covers in_channels=2, learn_sigma=True "
                                      "(the only configuration the reference scripts build)")
        kin = in_channels * 128 + 128 + context_size  # first-layer input width (models.py:216-219)
        if kin % 8 != 0:
            raise ValueError(f"context_size={context_size} gives a first-layer width of {kin}; the native GEMM needs a "
                             "multiple of 8 (16-byte bf16 rows for TMA); the reference scripts use 144 (sample.py:71)")
        self.learn_sigma = learn_sigma
        self.in_channels = in_channels
        self.context_size = context_size
        self.out_channels = in_channels * 2
        self.num_heads = num_heads
        self.hidden_size = hidden_size

        self.xoc_embedder = _FirstLayer(hidden_size, context_size, in_channels)
        self.t_embedder = _TimestepEmbedder(hidden_size)
        self.y_embedder = _LabelEmbedder(num_classes, hidden_size, class_dropout_prob)
        self.blocks = nn.ModuleList([_Block(hidden_size, num_heads, mlp_ratio) for _ in range(depth)])
        self.final_layer = _FinalLayer(hidden_size, self.out_channels)
        self.initialize_weights()
        self._engine = None
        self._train_weights = None
        # "bf16" (default: bf16 tensor-core operands, fp32 accumulate / residual / statistics; eps within 2e-3
        # of the fp32 reference) or "fp32" (osudit/fp32.py: eps within 1e-5, inference only, ~17x slower)
        self.precision = os.environ.get("OSUDIT_PRECISION", "bf16")

    def initialize_weights(self):
        """Same distributions as models.py:275-304 (xavier-uniform Linears with zero bias,
        N(0, 0.02) embedders, zeros for every adaLN modulation and the output projection), drawn in the same
        order: together with the constructors above, `torch.manual_seed(s); DiT_models[name](...)` gives
        bit-identical weights to the reference's (tests/golden/init_digest.json).  The packed in-projection keeps
        the xavier draw of its constructor, as in the reference (it is not an nn.Linear)."""
        for m in self.modules():
            if isinstance(m, nn.Linear):
                nn.init.xavier_uniform_(m.weight)
                if m.bias is not None:
                    nn.init.zeros_(m.bias)
        nn.init.normal_(self.xoc_embedder.mlp[0].weight, std=0.02)
        nn.init.normal_(self.y_embedder.embedding_table.weight, std=0.02)
        nn.init.normal_(self.t_embedder.mlp[0].weight, std=0.02)
        nn.init.normal_(self.t_embedder.mlp[2].weight, std=0.02)
        for lin in [b.adaLN_modulation[-1] for b in self.blocks] + \
                   [self.final_layer.adaLN_modulation[-1], self.final_layer.linear]:
            nn.init.zeros_(lin.weight)
            nn.init.zeros_(lin.bias)

    # ------------------------------------------------------------------ native path
    def __deepcopy__(self, memo):  # EMA copies (train.py:147) must not share the engine's buffers
        import copy
        cls = self.__class__
        new = cls.__new__(cls)
        memo[id(self)] = new
        for k, v in self.__dict__.items():
            new.__dict__[k] = None if k in ("_engine", "_train_weights") else copy.deepcopy(v, memo)
        return new

    def engine(self) -> DiTEngine:
        if self._engine is None:
            self._engine = DiTEngine(self)
        return self._engine

    def _labels(self, y):
        if self.training and self.y_embedder.dropout_prob > 0:  # models.py:69-72
            y = self.y_embedder.token_drop(y)
        return y

    def _raw_forward(self, x, t, o, c, y, attn_mask, x_rows=None, mod=None):
        for name, v in (("x", x), ("t", t), ("o", o), ("c", c), ("y", y)):
            if not v.is_cuda:
                raise RuntimeError(f"DiT.forward: `{name}` is on {v.device}; the native path runs on "
                                   "CUDA only and has no CPU fallback")
        check_inputs(self, x, t, o, c, y, x_rows)
        return self.engine().forward(x.float().contiguous(), t.long().contiguous(),
                                     o.float().contiguous(), c.float().contiguous(),
                                     self._labels(y.long()).contiguous(), attn_mask, x_rows, mod)

    def _needs_grad(self):
        return torch.is_grad_enabled() and any(p.requires_grad for p in self.parameters())

    def forward(self, x, t, o, c, y, attn_mask=None):
        """x (N,2,T), t (N,), o (N,T) ms, c (N,E,T), y (N,) -> (N,4,T); models.py:306-325.

        With autograd on (train.py:249-257) the call becomes one autograd node whose backward is the
        native schedule in osudit/train.py; under no_grad it is the inference schedule."""
        if self._needs_grad():
            from osudit.train import TrainWeights, dit_forward_train
            if self.precision != "bf16":
                raise NotImplementedError("precision='fp32' is an inference mode; train with precision='bf16'")
            for name, v in (("x", x), ("t", t), ("o", o), ("c", c), ("y", y)):
                if not v.is_cuda:
                    raise RuntimeError(f"DiT.forward: `{name}` is on {v.device}; the native path runs "
                                       "on CUDA only and has no CPU fallback")
            check_inputs(self, x, t, o, c, y)
            if self._train_weights is None:
                self._train_weights = TrainWeights()
            tw = self._train_weights  # re-packed inside the head node (eagerly, or as part of the CUDA graph)
            return dit_forward_train(self, tw, x.detach().float().contiguous(), t.long().contiguous(),
                                     o.float().contiguous(), c.float().contiguous(),
                                     self._labels(y.long()).contiguous(), attn_mask)
        return self._raw_forward(x, t, o, c, y, attn_mask).clone()

    def forward_with_cfg(self, x, t, o, c, y, cfg_scale, attn_mask=None):
        """models.py:327-343: the first half of x feeds both the conditional and unconditional
        rows; eps channels are guided, variance channels pass through."""
        raw = self._raw_forward(x, t, o, c, y, attn_mask, x_rows=len(x) // 2)
        return ops.cfg_combine(raw, cfg_scale, torch.empty_like(raw))


def DiT_XL(**kwargs):
    return DiT(depth=28, hidden_size=1152, num_heads=16, **kwargs)


def DiT_L(**kwargs):
    return DiT(depth=24, hidden_size=1024, num_heads=16, **kwargs)


def DiT_B(**kwargs):
    return DiT(depth=12, hidden_size=768, num_heads=12, **kwargs)


def DiT_S(**kwargs):
    return DiT(depth=12, hidden_size=384, num_heads=6, **kwargs)


DiT_models = {"DiT-XL": DiT_XL, "DiT-L": DiT_L, "DiT-B": DiT_B, "DiT-S": DiT_S}
