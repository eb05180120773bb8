"""Drop-in `positional_embedding`: the three embeddings the reference's data pipeline imports
(/root/reference/positional_embedding.py:29-77).

On the model's hot path these features are produced inside libosudit.so (csrc/embed.cu), straight
into the GEMM operand layout.  The functions here exist for the host-side callers that need the
same values as tensors — `data_loading.py:161` builds the distance context with
`timestep_embedding` — i.e. data preparation, not the denoising path.
"""
import math

import torch


def timestep_embedding(t, dim, max_period=10000):
    """[cos(t f_k) | sin(t f_k)], f_k = exp(-ln(max_period) k / (dim/2)); (N,) -> (N, dim)."""
    half = dim // 2
    k = torch.arange(start=0, end=half, dtype=torch.float32, device=t.device)
    freqs = torch.exp(-math.log(max_period) * k / half)
    args = t[:, None].float() * freqs[None]
    emb = torch.cat([torch.cos(args), torch.sin(args)], dim=-1)
    if dim % 2:
        emb = torch.cat([emb, torch.zeros_like(emb[:, :1])], dim=-1)
    return emb


def offset_sequence_embedding(t, dim, max_period=10000):
    """(N, T) time offsets -> (N, T, dim)."""
    n, length = t.shape
    return timestep_embedding(t.flatten(), dim, max_period).reshape(n, length, dim)


def position_sequence_embedding(t, dim, max_period=10000):
    """(N, T, D) positions -> (N, T, D * dim)."""
    n, length, d = t.shape
    return timestep_embedding(t.flatten(), dim, max_period).reshape(n, length, d * dim)
