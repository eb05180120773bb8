#!/usr/bin/env python
"""Secondary metric of BASELINE.json: training sequences/sec, DiT-B, seq-len 128, global batch 256
(BASELINE config 3), data-parallel over the GPUs of one node.

  python bench_train.py [--steps K] [--warmup W] [--model DiT-B] [--global-batch 256]
  python -m torch.distributed.run --nproc-per-node N ... bench_train.py --gpus N

Restates the step body of train.py:243-261 around the drop-in modules: label dropout, t ~ U{0..999},
`diffusion.training_losses` under fp16 autocast, GradScaler, AdamW(lr 1e-4, wd 0), EMA update, with
DistributedDataParallel (NCCL gradient all-reduce) when WORLD_SIZE > 1.  `--impl reference` times the
same step on the CPU oracle port (fp32 autograd, all host threads) on a bounded batch.
Prints one JSON line on rank 0.  (bench.py remains the headline benchmark.)
"""
import argparse
import json
import os
import sys
import time
from copy import deepcopy

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (os.path.join(ROOT, "osu-diffusion_b200"), ROOT):
    if p not in sys.path:
        sys.path.insert(0, p)

import torch  # noqa: E402

SEQ = 128


def flops_per_seq(model_name):
    D, depth, H = {"DiT-S": (384, 12, 6), "DiT-B": (768, 12, 12), "DiT-L": (1024, 24, 16), "DiT-XL": (1152, 28, 16)}[model_name]
    gemm_tok = 24 * D * D * depth + 2 * 528 * D + 2 * D * 4
    attn = SEQ * SEQ * 4 * D * depth
    return 3 * (SEQ * gemm_tok + attn)


@torch.no_grad()
def update_ema(ema, model, decay=0.9999):  # train.py:36-45
    for pe, pm in zip(ema.parameters(), model.parameters()):
        pe.mul_(decay).add_(pm.detach(), alpha=1 - decay)


def run_native(args, rank, world, local_rank):
    import models
    from diffusion import create_diffusion
    from osudit import ops, synth
    device = torch.device("cuda", local_rank)
    torch.cuda.set_device(device)
    dist = None
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
        dist.init_process_group("nccl", device_id=device)
    torch.manual_seed(0)
    model = models.DiT_models[args.model](num_classes=52670, context_size=144, class_dropout_prob=0.2)
    g = torch.Generator().manual_seed(1)
    with torch.no_grad():  # non-zero adaLN / output layers so every gradient is exercised
        for k, v in model.state_dict().items():
            if "adaLN_modulation" in k or k.startswith("final_layer.linear"):
                v.copy_(torch.randn(v.shape, generator=g) * 0.02)
    model = model.to(device)
    ema = deepcopy(model).requires_grad_(False)
    net = model
    if world > 1:
        if args.fused_optimizer:  # the documented opt-ins: bf16 buckets, bucket views, SMs reserved for NCCL
            from osudit import ddp
            net = ddp.wrap(model, device_ids=[local_rank])
        else:
            net = torch.nn.parallel.DistributedDataParallel(model, device_ids=[local_rank])  # train.py:152
    diffusion = create_diffusion("", noise_schedule="squaredcos_cap_v2", use_l1=True)
    if args.fused_optimizer:  # opt-in: AdamW + EMA + unscale + inf-skip in one launch (osudit/optim.py)
        from osudit.optim import FusedAdamWEMA
        opt = FusedAdamWEMA(net.parameters(), lr=1e-4, weight_decay=0)
        opt.attach_ema(ema, model, decay=0.9999)
    else:
        opt = torch.optim.AdamW(net.parameters(), lr=1e-4, weight_decay=0)
    scaler = torch.amp.GradScaler("cuda")
    B = args.global_batch // world
    (x, o, c), y = synth.training_batch(B, SEQ, seed=rank)
    host = [t.pin_memory() for t in (x, o, c, y)]
    model.train()

    def step():
        xd, od, cd, yd = [t.to(device, non_blocking=True) for t in host]
        t = torch.randint(0, diffusion.num_timesteps, (B,), device=device)
        with torch.autocast(device_type="cuda", dtype=torch.float16):
            loss = diffusion.training_losses(net, xd, t, dict(o=od, c=cd, y=yd))["loss"].mean()
        scaler.scale(loss).backward()
        scaler.step(opt)
        scaler.update()
        opt.zero_grad(set_to_none=True)
        if not args.fused_optimizer:
            update_ema(ema, model)
        return loss

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step()
    barrier()
    l0 = ops.launch_count
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        loss = step()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    if dist is not None:
        tt = torch.tensor([ms], device=device)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms = float(tt.item())
    barrier()
    launches = ops.launch_count - l0
    final_loss = float(loss)
    if dist is not None:
        dist.destroy_process_group()
    if rank != 0:
        return None
    value = args.global_batch * args.steps / (ms / 1e3)
    return {"metric": f"train seq/s {args.model} seq-len {SEQ}", "value": round(value, 1), "unit": "seq/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(ms / args.steps, 3),
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": f"{args.model} training, seq-len {SEQ}, global batch {args.global_batch}, L1+VB loss, "
                                   "AdamW 1e-4, EMA, fp16-autocast context + GradScaler as train.py; DDP NCCL all-reduce",
                       "optimizer": "osudit FusedAdamWEMA (opt-in)" if args.fused_optimizer
                       else "torch.optim.AdamW + update_ema loop (train.py unchanged)",
                       "parallelism": f"dp{world}"},
            "model_tflops": round(flops_per_seq(args.model) * value / world / 1e12, 1),
            "gpu_launches": launches, "final_loss": round(final_loss, 4)}


def run_reference(args, rank):
    if rank != 0:
        return None
    from oracle import diffusion as odiff, dit as odit
    from osudit import synth
    torch.set_num_threads(os.cpu_count() or 1)
    shape = odit.shape_of(args.model)
    sd = odit.init_state_dict(shape, seed=1)
    params = {k: v.requires_grad_(v.is_floating_point() and "playfield" not in k) for k, v in sd.items()}
    opt = torch.optim.AdamW([p for p in params.values() if p.requires_grad], lr=1e-4, weight_decay=0)
    s = odiff.Schedule("")
    B = 8
    (x, o, c), y = synth.training_batch(B, SEQ, seed=0)

    def step():
        t = torch.randint(0, 1000, (B,))
        noise = torch.randn_like(x)
        loss = odiff.training_losses(s, lambda xt, tt: odit.forward(params, shape.heads, xt, tt, o, c, y),
                                     x, t, noise, use_l1=True)["loss"].mean()
        loss.backward()
        opt.step()
        opt.zero_grad(set_to_none=True)

    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = (time.perf_counter() - t0) / args.steps
    value = B / dt
    return {"impl": "reference", "metric": f"train seq/s {args.model} seq-len {SEQ}", "value": round(value, 2),
            "unit": "seq/s", "n_gpus": 1, "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(dt * 1e3, 1),
            "higher_is_better": True, "dtype": "f32", "data": "synthetic",
            "cpu_baseline": {"value": round(value, 2), "unit": "seq/s", "cores": torch.get_num_threads(), "kind": "port",
                             "sample": f"oracle port, fp32 autograd + AdamW, batch {B} per step"}}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--model", default="DiT-B")
    ap.add_argument("--global-batch", type=int, default=256)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--seq-len", type=int, default=128, help="datapoints per training window (config 5: 512)")
    ap.add_argument("--fused-optimizer", action="store_true",
                    help="use osudit.optim.FusedAdamWEMA instead of torch.optim.AdamW + the update_ema loop")
    args = ap.parse_args()
    global SEQ
    SEQ = args.seq_len
    # stdout carries exactly ONE JSON line: keep the real stdout aside and point fd 1 at stderr, so that anything a
    # library writes to C stdout (NCCL prints its version banner there when NCCL_DEBUG is set) cannot precede it
    sys.stdout.flush()
    out_fd = os.dup(1)
    os.dup2(2, 1)
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    line = run_reference(args, rank) if args.impl == "reference" else \
        run_native(args, rank, world, int(os.environ.get("LOCAL_RANK", 0)))
    if rank == 0 and line is not None:
        os.write(out_fd, (json.dumps(line) + "\n").encode())


if __name__ == "__main__":
    main()
