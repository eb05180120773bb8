"""Data-parallel training around the native DiT: the reference wraps the model in stock
`DistributedDataParallel` (train.py:152) and lets its NCCL all-reduce run under the backward (train.py:257).

That keeps working unchanged with this repo's `models.DiT` (the autograd chain in osudit/train.py hands DDP the
parameter gradients block by block, in bucket order).  `wrap` is the same call with the settings that matter at the
small per-GPU batches of a strong-scaling run (measured on 2 and 8 B200s, DiT-B, 32 sequences per GPU, fused optimizer;
`gpurun_out` logs summarised in DESIGN.md §7):

* an SM cap for the persistent kernels (`osudit_set_sm_limit`, 16 SMs left free): a persistent GEMM grid otherwise fills
  all 148 SMs with 225 KB of shared memory each, NCCL's kernels cannot co-reside and every all-reduce queues behind a
  full-machine grid (whose successor then runs as two waves): 15.8 ms per step without the cap, 8.7 ms with it;
* 128 MB buckets instead of 25 MB: a 12 MB all-reduce over NVSwitch is latency-bound (28 of them per step), six large
  ones run at link bandwidth: 9.1 -> 8.7 ms at N=2;
* `gradient_as_bucket_view=True`: `p.grad` aliases the bucket, one copy of every gradient less per step;
* bf16 gradient buckets (`bf16_compress_hook`) are available (`bf16_buckets=True` / OSUDIT_DDP_BF16=1) but OFF by
  default: NVLink 5 moves the fp32 buckets faster than the two cast passes the hook adds (8.61 ms with fp32 buckets vs
  8.93 ms with bf16 ones at N=8, 8.23 vs 8.68 at N=2).
"""
from __future__ import annotations

import os

import torch

from . import _lib


def set_sm_limit(n: int) -> int:
    """Cap persistent grids at `n` CTAs (0 = one per SM); returns the previous cap."""
    return _lib.load().osudit_set_sm_limit(int(n))


def wrap(model, device_ids=None, bf16_buckets: bool | None = None, reserve_sms: int | None = None,
         bucket_cap_mb: int | None = None):
    """`DistributedDataParallel(model, device_ids=...)` as train.py:152, plus the opt-ins described above.

    `reserve_sms`: SMs left to NCCL while this process trains (default: $OSUDIT_DDP_RESERVE_SMS or 16); 0 disables."""
    import torch.distributed as dist
    from torch.nn.parallel import DistributedDataParallel as DDP

    if not dist.is_initialized():
        raise RuntimeError("osudit.ddp.wrap: init_process_group first (train.py:104)")
    if bucket_cap_mb is None:
        bucket_cap_mb = int(os.environ.get("OSUDIT_DDP_BUCKET_MB", "128"))
    if bf16_buckets is None:
        bf16_buckets = os.environ.get("OSUDIT_DDP_BF16", "0") == "1"
    net = DDP(model, device_ids=device_ids, gradient_as_bucket_view=True, bucket_cap_mb=bucket_cap_mb)
    if bf16_buckets and dist.get_backend() == "nccl":
        from torch.distributed.algorithms.ddp_comm_hooks import default_hooks
        net.register_comm_hook(None, default_hooks.bf16_compress_hook)
    if reserve_sms is None:
        reserve_sms = int(os.environ.get("OSUDIT_DDP_RESERVE_SMS", "16"))
    if reserve_sms > 0 and dist.get_world_size() > 1 and torch.cuda.is_available():
        sms = torch.cuda.get_device_properties(torch.cuda.current_device()).multi_processor_count
        set_sm_limit(max(sms - reserve_sms, sms // 2))
    return net
