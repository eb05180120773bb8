#!/usr/bin/env python
"""Headline benchmark: beatmaps/sec of DiT-B 100-step CFG sampling (BASELINE.json config 2).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl native|reference]

One "step" = one pass of the hot path over one batch: the full 100-step classifier-free-guidance
sampling of 64 synthetic 2048-datapoint beatmaps (128 model rows) through the public API
`create_diffusion("100", "squaredcos_cap_v2").p_sample_loop(model.forward_with_cfg, ...)`.
N>1 (launched by torchrun) shards independent beatmaps over ranks with no collective on the data
path (weak scaling: 64 beatmaps per GPU); timing is CUDA events, max over ranks.

The same JSON line carries BASELINE.json's second metric under "train": sequences/sec of DiT-B seq-len-128
training at global batch 256 (config 3), data-parallel over the N ranks with the NCCL gradient all-reduce of
train.py:152,257 (strong scaling: 256 / N sequences per GPU), timed in the same process after the sampling arm.

`--impl reference` times the reference's own CPU path on a bounded sample of the same workload: the UNMODIFIED
reference modules when `baseline/_ref/` (written by `__graft_entry__.build()` from /root/reference; git-ignored,
travels to the GPU box) or $OSU_DIFFUSION_REF holds them (`kind: "reference"`), else the oracle port (`kind: "port"`).
Prints exactly one JSON line on rank 0.
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "osu-diffusion_b200")
for p in (PKG, ROOT):
    if p not in sys.path:
        sys.path.insert(0, p)

import torch  # noqa: E402

MODEL, N_BEATMAPS, SEQ, STEPS_DIFF, CFG, BAND = "DiT-B", 64, 2048, 100, 1.5, 128
METRIC = "beatmaps/sec DiT-B 100-step CFG sampling"
UNIT = "beatmaps/s"
TRAIN_MODEL, TRAIN_SEQ, TRAIN_GLOBAL_BATCH = "DiT-B", 128, 256  # BASELINE config 3


def ncu_gemm_traffic():
    """(GB per launch, source) of the dominant GEMM's DRAM traffic: mean over the `gemm2_kernel` rows of the newest
    committed `--set full` summary under profiles/ (dram bytes_read + bytes_write columns), None when absent."""
    import glob
    import re
    for path in sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_summary.md")), reverse=True):
        rd = wr = None
        vals = []
        for line in open(path):
            cells = [c.strip() for c in line.strip().strip("|").split("|")]
            if "kernel" in cells[0] and ("dram bytes_read [Gbyte]" in line or "dram rd [Gbyte]" in line):
                rd = next(i for i, c in enumerate(cells) if c.startswith(("dram bytes_read", "dram rd")))
                wr = next(i for i, c in enumerate(cells) if c.startswith(("dram bytes_write", "dram wr")))
                if "[Gbyte]" not in cells[wr]:  # mixed units in this table: not usable as is
                    rd = wr = None
                continue
            if rd is not None and re.match(r"`(g2::)?gemm2_kernel<", cells[0]) and len(cells) > max(rd, wr):
                try:
                    vals.append(float(cells[rd]) + float(cells[wr]))
                except ValueError:
                    pass
        if vals:
            return round(sum(vals) / len(vals), 3), f"{os.path.relpath(path, ROOT)} ({len(vals)} gemm2_kernel launches)"
    return None, "no ncu --set full summary under profiles/"


def workload_config(n_gpus):
    return {"workload": f"{MODEL} sampling, {SEQ}-datapoint synthetic beatmaps, {STEPS_DIFF} respaced "
                        f"steps (squaredcos_cap_v2), CFG {CFG} with style class, batch {N_BEATMAPS} "
                        f"beatmaps ({2 * N_BEATMAPS} model rows) per GPU, band mask W={BAND}",
            "beatmaps_per_gpu": N_BEATMAPS, "datapoints": SEQ, "diffusion_steps": STEPS_DIFF,
            "cfg_scale": CFG, "parallelism": f"independent beatmaps x{n_gpus} (no collective)",
            "weights": "seeded random init, zero-init layers redrawn N(0,0.02^2) (SURVEY F4)",
            "l2_policy": "inputs and activations (>5 GB per step) exceed the 126 MB L2; no flush needed"}


SIZES = {"DiT-S": (384, 12, 6), "DiT-B": (768, 12, 12), "DiT-L": (1024, 24, 16), "DiT-XL": (1152, 28, 16)}


def flops_per_beatmap():
    """Algorithmic flop count, attention over the +-128 band only."""
    D, depth, H = SIZES[MODEL]
    hd = D // H
    gemm_tok = 24 * D * D * depth + 2 * 528 * D + 2 * D * 4
    pairs = sum(min(SEQ - 1, j + BAND) - max(0, j - (BAND - 1)) + 1 for j in range(SEQ))
    attn_row = pairs * 4 * hd * H * depth
    return 2 * STEPS_DIFF * (SEQ * gemm_tok + attn_row)


# ------------------------------------------------------------------------------ clocks
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.path = index, None, None

    def start(self):
        try:
            f = tempfile.NamedTemporaryFile("w", suffix=".csv", delete=False)
            self.path = f.name
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.index), "-lms", "200"], stdout=f, stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, power = [], 0, set(), 0.0
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in open(self.path):
            c = [v.strip() for v in line.split(",")]
            if len(c) < 9:
                continue
            try:
                sm.append(float(c[1])); mx = max(mx, float(c[2])); power = max(power, float(c[3]))
            except ValueError:
                continue
            for nm, v in zip(names, c[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        os.unlink(self.path)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx or None,
                "power_w_max": power, "samples": len(sm), "reasons": sorted(reasons)}


# --------------------------------------------------------------------------- native arm
def build_native(device, seed=1):
    import models
    torch.manual_seed(seed)
    m = models.DiT_models[MODEL](num_classes=52670, context_size=144)
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():  # redraw the zero-initialised tensors, otherwise the output is identically 0
        for k, v in m.state_dict().items():
            if "adaLN_modulation" in k or k.startswith("final_layer.linear"):
                v.copy_(torch.randn(v.shape, generator=g) * 0.02)
    return m.to(device).eval()


def run_native(args, rank, world, local_rank):
    from diffusion import create_diffusion
    from osudit import ops, synth

    device = torch.device("cuda", local_rank)
    torch.cuda.set_device(device)
    dist = None
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")  # stdout carries exactly one JSON line
        dist.init_process_group("nccl", device_id=device)
    model = build_native(device)
    diffusion = create_diffusion(str(STEPS_DIFF), noise_schedule="squaredcos_cap_v2")
    z, o, c, y = synth.sampling_batch(N_BEATMAPS, SEQ, seed=1000 * rank)  # each rank: its own beatmaps
    mask = synth.band_mask(SEQ, BAND)
    host = [t.pin_memory() for t in (z, o, c, y)]
    mask_d = mask.to(device)
    zd, od, cd, yd = [t.to(device) for t in (z, o, c, y)]

    def sample_resident():
        return diffusion.p_sample_loop(model.forward_with_cfg, zd.shape, zd, clip_denoised=True,
                                       model_kwargs=dict(o=od, c=cd, y=yd, cfg_scale=CFG, attn_mask=mask_d),
                                       device=device)

    out_host = torch.empty(N_BEATMAPS, 2, SEQ).pin_memory()

    def sample_e2e():
        zz, oo, cc, yy = [t.to(device, non_blocking=True) for t in host]
        s = diffusion.p_sample_loop(model.forward_with_cfg, zz.shape, zz, clip_denoised=True,
                                    model_kwargs=dict(o=oo, c=cc, y=yy, cfg_scale=CFG, attn_mask=mask_d),
                                    device=device)
        out_host.copy_(s.chunk(2, dim=0)[0], non_blocking=True)  # sample.py:111 keeps the cond half
        return s

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, k):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(k):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        if dist is not None:
            t = torch.tensor([ms], device=device)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        barrier()
        return ms

    with torch.no_grad():
        for _ in range(args.warmup):
            sample_resident()
        clocks = ClockSampler(local_rank)
        if rank == 0:
            clocks.start()
        l0 = ops.launch_count
        ms = timed(sample_resident, args.steps)
        launches = ops.launch_count - l0
        clk = clocks.stop() if rank == 0 else None
        if args.no_e2e:
            ms_e2e = None
        else:
            sample_e2e()
            ms_e2e = timed(sample_e2e, args.steps)
        roof = kernel_roofline(model, diffusion, zd, od, cd, yd, mask_d, device) if rank == 0 else None

    total = N_BEATMAPS * world * args.steps
    value = total / (ms / 1e3)
    train = None
    if not args.no_train:
        model._engine = None  # drop the sampling workspaces (5.6 GB) before the training arm allocates its own
        del model, zd, od, cd, yd
        torch.cuda.empty_cache()
        train = run_train(rank, world, local_rank, dist, device)
    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return None
    line = {
        "metric": METRIC, "value": round(value, 4), "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": round(ms / args.steps, 3), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": workload_config(world), "clocks": clk,
        "e2e": None if ms_e2e is None else {
            "value": round(total / (ms_e2e / 1e3), 4), "unit": UNIT,
            "h2d_bytes_per_step": sum(t.numel() * t.element_size() for t in host),
            "d2h_bytes_per_step": out_host.numel() * out_host.element_size()},
        "gpu_launches": launches,
        "model_tflops": round(flops_per_beatmap() * value / world / 1e12, 1),
        "roofline": roof,
        "train": train,
    }
    line["cpu_baseline"] = cpu_baseline_sample() if (world == 1 and not args.no_cpu_baseline) else None
    if dist is not None:
        dist.destroy_process_group()
    return line



# ------------------------------------------------------------------------- training arm
def train_flops_per_seq():
    """fwd + bwd (3x forward) algorithmic flops of one training sequence, full attention over the window."""
    D, depth = 768, 12
    gemm_tok = 24 * D * D * depth + 2 * 528 * D + 2 * D * 4
    return 3 * (TRAIN_SEQ * gemm_tok + TRAIN_SEQ * TRAIN_SEQ * 4 * D * depth)


def run_train(rank, world, local_rank, dist, device, steps=20, warmup=8):
    """BASELINE config 3 through the drop-in API: train.py:243-261's step (label dropout, t ~ U{0..999},
    `training_losses` under fp16 autocast, GradScaler, AdamW 1e-4, EMA 0.9999) on synthetic windows, DDP over the
    ranks.  Two variants are timed: "stock" = train.py's own optimizer / EMA loop / DistributedDataParallel call,
    and the headline = the documented opt-ins (osudit.optim.FusedAdamWEMA, osudit.ddp.wrap: 128 MB buckets as bucket
    views, SMs reserved for NCCL).  `allreduce_ms_exposed` = step time minus the same step under `no_sync()`."""
    import contextlib
    from copy import deepcopy
    import models
    from diffusion import create_diffusion
    from osudit import ddp, ops, synth, train as otrain
    from osudit.optim import FusedAdamWEMA

    B = TRAIN_GLOBAL_BATCH // world
    diffusion = create_diffusion("", noise_schedule="squaredcos_cap_v2", use_l1=True)
    (x, o, c), y = synth.training_batch(B, TRAIN_SEQ, seed=rank)
    host = [t.pin_memory() for t in (x, o, c, y)]

    def build(fused):
        torch.manual_seed(0)
        m = models.DiT_models[TRAIN_MODEL](num_classes=52670, context_size=144, class_dropout_prob=0.2)
        g = torch.Generator().manual_seed(1)
        with torch.no_grad():  # non-zero adaLN / output layers so every gradient is exercised
            for k, v in m.state_dict().items():
                if "adaLN_modulation" in k or k.startswith("final_layer.linear"):
                    v.copy_(torch.randn(v.shape, generator=g) * 0.02)
        m = m.to(device).train()
        ema = deepcopy(m).requires_grad_(False)
        net = m
        if world > 1:
            if fused:
                net = ddp.wrap(m, device_ids=[local_rank])
            else:
                net = torch.nn.parallel.DistributedDataParallel(m, device_ids=[local_rank])  # train.py:152
        if fused:
            opt = FusedAdamWEMA(net.parameters(), lr=1e-4, weight_decay=0)
            opt.attach_ema(ema, m, decay=0.9999)
        else:
            opt = torch.optim.AdamW(net.parameters(), lr=1e-4, weight_decay=0)
        return m, ema, net, opt

    def measure(fused):
        m, ema, net, opt = build(fused)
        scaler = torch.amp.GradScaler("cuda")

        def step(sync=True):
            xd, od, cd, yd = [t.to(device, non_blocking=True) for t in host]
            t = torch.randint(0, diffusion.num_timesteps, (B,), device=device)
            ctx = contextlib.nullcontext() if (sync or world == 1) else net.no_sync()
            with ctx:
                with torch.autocast(device_type="cuda", dtype=torch.float16):
                    loss = diffusion.training_losses(net, xd, t, dict(o=od, c=cd, y=yd))["loss"].mean()
                scaler.scale(loss).backward()
            scaler.step(opt)
            scaler.update()
            opt.zero_grad(set_to_none=True)
            if not fused:
                with torch.no_grad():  # update_ema, train.py:36-45
                    for pe, pm in zip(ema.parameters(), m.parameters()):
                        pe.mul_(0.9999).add_(pm.detach(), alpha=1 - 0.9999)
            return loss

        def timed(k, sync=True):
            if dist is not None:
                dist.barrier()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(k):
                loss = step(sync)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1)
            if dist is not None:
                tt = torch.tensor([ms], device=device)
                dist.all_reduce(tt, op=dist.ReduceOp.MAX)
                ms = float(tt.item())
            return ms / k, float(loss.detach())

        for _ in range(warmup):
            step()
        l0 = ops.launch_count
        ms, loss = timed(steps)
        launches = ops.launch_count - l0
        ms_nosync = timed(max(steps // 2, 4), sync=False)[0] if world > 1 else ms
        nparam = sum(p.numel() for p in m.parameters() if p.requires_grad)
        del net, opt, ema, m
        otrain.release_graphs()
        ddp.set_sm_limit(0)
        torch.cuda.empty_cache()
        return ms, ms_nosync, loss, launches, nparam

    ms_f, ns_f, loss_f, launches, nparam = measure(True)
    ms_s, ns_s, loss_s, _, _ = measure(False)
    pk_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    tf_peak = json.load(open(pk_path))["bf16_tflops_sustained"] if os.path.exists(pk_path) else 1400.0
    value = TRAIN_GLOBAL_BATCH / (ms_f / 1e3)
    tfl = train_flops_per_seq() * value / world / 1e12
    bytes_ar = nparam * 4
    return {
        "metric": f"train seq/s {TRAIN_MODEL} seq-len {TRAIN_SEQ}", "value": round(value, 1), "unit": "seq/s",
        "n_gpus": world, "steps": steps, "warmup": warmup, "ms_per_step": round(ms_f, 3), "scaling": "strong",
        "global_batch": TRAIN_GLOBAL_BATCH, "per_gpu_batch": B, "dtype": "bf16", "data": "synthetic",
        "final_loss": round(loss_f, 4), "gpu_launches": launches,
        "config": {"workload": f"{TRAIN_MODEL} training, seq-len {TRAIN_SEQ}, global batch {TRAIN_GLOBAL_BATCH}, L1+VB loss, "
                               "AdamW 1e-4, EMA, fp16-autocast context + GradScaler as train.py:243-261",
                   "optimizer": "osudit.optim.FusedAdamWEMA (opt-in: AdamW + EMA + unscale in one launch)",
                   "parallelism": f"dp{world}: DistributedDataParallel over NCCL via osudit.ddp.wrap (fp32 gradient buckets of "
                                  "128 MB as bucket views, 16 SMs reserved for NCCL)" if world > 1 else "dp1 (no collective)"},
        "allreduce_ms_exposed": round(ms_f - ns_f, 3) if world > 1 else 0.0,
        "collective": None if world == 1 else {
            "op": "NCCL all-reduce of the parameter gradients (train.py:152,257), the only collective on the path",
            "bytes_per_step": bytes_ar, "dtype": "fp32 buckets",
            "floor_ms_at_725GBs_busbw": round(bytes_ar * 2 * (world - 1) / world / 725e9 * 1e3, 3)},
        "roofline": {"bound": "tensor", "achieved": round(tfl, 1), "peak": tf_peak, "unit": "TFLOP/s (model-level, per GPU)",
                     "frac": round(tfl / tf_peak, 4)},
        "stock": {"value": round(TRAIN_GLOBAL_BATCH / (ms_s / 1e3), 1), "ms_per_step": round(ms_s, 3),
                  "allreduce_ms_exposed": round(ms_s - ns_s, 3) if world > 1 else 0.0, "final_loss": round(loss_s, 4),
                  "what": "train.py unchanged: torch.optim.AdamW + update_ema loop + stock DistributedDataParallel "
                          "(fp32 buckets)"},
    }


def kernel_roofline(model, diffusion, zd, od, cd, yd, mask_d, device):
    """Per-launch CUDA-event timing of the dominant kernel (the tcgen05 GEMM, all QKV / out-proj /
    fc1 / fc2 launches of one denoising step) and of the HBM-bound LayerNorm kernel."""
    from osudit import ops
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        pk = json.load(open(peaks_path))
        tf_peak, hbm_peak, src = pk["bf16_tflops_sustained"], pk["hbm_gbs"], "MEASURED_PEAKS.json (sustained bf16, copy GB/s)"
    else:
        tf_peak, hbm_peak, src = 1400.0, 6650.0, "fallback (B200_PROFILING.md)"
    rec = {"gemm": [], "ln": [], "attn": []}
    real_gemm, real_ln, real_attn, real_resid = ops.gemm, ops.ln_modulate, ops.attn_band, ops.gemm_gated_residual

    def wrap(kind, fn, work):
        def inner(*a, **k):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            r = fn(*a, **k)
            e1.record()
            rec[kind].append((e0, e1, work(*a, **k)))
            return r
        return inner

    def gemm_work(a_segs, b_segs, bias, epi, out):
        return 2.0 * out.shape[0] * out.shape[1] * sum(a.shape[1] for a in a_segs), tuple(out.shape) + (a_segs[0].shape[1],)

    def resid_work(a, w, bias, mod, gate_col, rows_per_batch, x):  # out-projection / fc2 with the residual epilogue
        return 2.0 * x.shape[0] * x.shape[1] * a.shape[1], tuple(x.shape) + (a.shape[1], "resid")

    def ln_work(x, branch, mod, g, sh, sc, T, h):
        return x.numel() * (12.0 if branch is not None else 6.0), None

    def attn_work(qkv, out, B, T, H, hd, wl=-1, wr=-1, mask=None):
        pairs = sum(min(T - 1, j + wr) - max(0, j - wl) + 1 for j in range(T)) if wl >= 0 else T * T
        return 4.0 * hd * pairs * H * B, None

    ops.gemm = wrap("gemm", real_gemm, gemm_work)
    ops.gemm_gated_residual = wrap("gemm", real_resid, resid_work)
    ops.ln_modulate = wrap("ln", real_ln, ln_work)
    ops.attn_band = wrap("attn", real_attn, attn_work)
    from osudit import graphs
    graphs_on, graphs._ENABLED = graphs._ENABLED, False  # per-launch events need the individual launches
    try:
        t = torch.full((zd.shape[0],), diffusion.num_timesteps // 2, device=device, dtype=torch.long)
        for _ in range(3):
            diffusion.p_sample(model.forward_with_cfg, zd, t, clip_denoised=True,
                               model_kwargs=dict(o=od, c=cd, y=yd, cfg_scale=CFG, attn_mask=mask_d))
        torch.cuda.synchronize()
    finally:
        ops.gemm, ops.ln_modulate, ops.attn_band, ops.gemm_gated_residual = real_gemm, real_ln, real_attn, real_resid
        graphs._ENABLED = graphs_on

    def agg(items, pred=lambda tag: True):
        items = [(e0.elapsed_time(e1), w) for e0, e1, (w, tag) in items if pred(tag)]
        ms = sum(m for m, _ in items)
        return sum(w for _, w in items), ms, len(items)

    big = lambda tag: tag is not None and tag[0] >= 65536 and tag[2] >= 768  # noqa: E731  (block GEMMs)
    fl, ms, n = agg(rec["gemm"], big)
    by_shape = {}
    resid = [0.0, 0.0, 0.0]  # flops, ms, algorithmic bytes of the launches that also carry the residual update
    plain = [0.0, 0.0]
    for e0, e1, (w, tag) in rec["gemm"]:
        if big(tag):
            d = by_shape.setdefault("x".join(map(str, tag[:3])), [0.0, 0.0])
            t_ms = e0.elapsed_time(e1)
            d[0] += w; d[1] += t_ms
            if len(tag) > 3:  # A and W read in bf16, the fp32 residual stream read and written
                resid[0] += w; resid[1] += t_ms
                resid[2] += 2.0 * (tag[0] * tag[2] + tag[1] * tag[2]) + 8.0 * tag[0] * tag[1]
            else:
                plain[0] += w; plain[1] += t_ms
    lb, lms, ln_n = agg(rec["ln"])
    af, ams, an = agg(rec["attn"])
    step_ms = sum(e0.elapsed_time(e1) for k in rec for e0, e1, _ in rec[k])
    ach = fl / (ms / 1e3) / 1e12
    traffic, traffic_src = ncu_gemm_traffic()
    # algorithmic bytes of one launch: A read + W read + out written, bf16 (mean over the timed block GEMMs)
    shapes = [tag for _, _, (w, tag) in rec["gemm"] if big(tag)]
    algo_gb = sum(2.0 * (t[0] * t[2] + t[1] * t[2]) + (8.0 if len(t) > 3 else 2.0) * t[0] * t[1] for t in shapes) / max(len(shapes), 1) / 1e9
    return {"bound": "tensor", "kernel": "g2::gemm2_kernel (cta_group::2 tcgen05 GEMM: QKV, fc1+GELU, and out-proj / fc2 "
            "with the gated-residual fp32 reduce-add epilogue)",
            "achieved": round(ach, 1), "peak": tf_peak, "unit": "TFLOP/s", "frac": round(ach / tf_peak, 4),
            "traffic": traffic, "traffic_unit": f"GB per launch (ncu dram read+write, mean over {traffic_src}); "
            f"algorithmic {round(algo_gb, 2)}", "peak_source": src,
            "launches_timed": n,
            "avg_launch_ms": round(ms / max(n, 1), 4), "share_of_step": round(ms / step_ms, 3),
            "by_epilogue": {
                "bf16 / GELU (QKV, fc1)": {"achieved": round(plain[0] / max(plain[1], 1e-9) / 1e9, 1), "unit": "TFLOP/s",
                                            "frac": round(plain[0] / max(plain[1], 1e-9) / 1e9 / tf_peak, 4)},
                "gated residual (out-proj, fc2): x += gate*(acc+bias) by fp32 TMA reduce-add, 6 D bytes per token moved here "
                "from the LayerNorm kernel": {"achieved": round(resid[0] / max(resid[1], 1e-9) / 1e9, 1), "unit": "TFLOP/s",
                                              "hbm_GBs": round(resid[2] / max(resid[1], 1e-9) / 1e6, 1),
                                              "hbm_frac": round(resid[2] / max(resid[1], 1e-9) / 1e6 / hbm_peak, 4)}},
            "per_shape_tflops": {k: round(v[0] / (v[1] / 1e3) / 1e12, 1) for k, v in by_shape.items()},
            "hbm_kernel": {"kernel": "ln_modulate_kernel", "bound": "hbm",
                           "achieved": round(lb / (lms / 1e3) / 1e9, 1), "peak": hbm_peak, "unit": "GB/s",
                           "frac": round(lb / (lms / 1e3) / 1e9 / hbm_peak, 4), "launches_timed": ln_n,
                           "share_of_step": round(lms / step_ms, 3)},
            "attention_kernel": {"kernel": "attn_st::attn_stream_kernel (tcgen05, P in TMEM)" if SIZES[MODEL][0] // SIZES[MODEL][2] == 64
                                 else "attn_band_kernel<72> (mma.sync)", "achieved": round(af / (ams / 1e3) / 1e12, 1),
                                 "unit": "TFLOP/s (banded algorithmic flops)", "share_of_step": round(ams / step_ms, 3)}}


# ------------------------------------------------------------------------ CPU reference arm
def reference_root():
    """Directory holding the UNMODIFIED reference modules (models.py, positional_embedding.py, diffusion/), or None."""
    for cand in (os.environ.get("OSU_DIFFUSION_REF"), os.path.join(ROOT, "baseline", "_ref"), "/root/reference"):
        if cand and os.path.exists(os.path.join(cand, "models.py")) and \
                os.path.exists(os.path.join(cand, "diffusion", "gaussian_diffusion.py")):
            return cand
    return None


def _load_synth():
    """osudit/synth.py by file path: the reference arm must not put this repo's drop-in `models` / `diffusion` on
    sys.path next to the reference's modules of the same names."""
    import importlib.util
    spec = importlib.util.spec_from_file_location("osudit_synth", os.path.join(PKG, "osudit", "synth.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


class ReferenceCPU:
    """sample.py:69-108,174-182 and train.py:243-261 restated around the unmodified reference modules on the host
    cores (fp32, all threads): the scripts themselves need `slider` / matplotlib / CUDA and cannot be launched."""
    kind = "reference"

    def __init__(self, root):
        for q in (PKG, ROOT):
            while q in sys.path:
                sys.path.remove(q)
        sys.path.insert(0, root)
        os.environ.setdefault("PYTHONDONTWRITEBYTECODE", "1")
        sys.dont_write_bytecode = True
        import models as ref_models  # noqa: the reference's own module
        from diffusion import create_diffusion as ref_create
        self.models, self.create = ref_models, ref_create
        self.synth = _load_synth()
        self.where = root

    def _model(self, name, dropout):
        torch.manual_seed(1)
        m = self.models.DiT_models[name](num_classes=52670, context_size=144, class_dropout_prob=dropout)
        g = torch.Generator().manual_seed(1)
        with torch.no_grad():
            for k, v in m.state_dict().items():
                if "adaLN_modulation" in k or k.startswith("final_layer.linear"):
                    v.copy_(torch.randn(v.shape, generator=g) * 0.02)
        return m

    def denoise_step_seconds(self, n, T, repeats):
        if getattr(self, "_sampler", None) is None:  # built once: the timed region is sample.py's loop body only
            self._sampler = (self._model(MODEL, 0.1).eval(), self.create(str(STEPS_DIFF), noise_schedule="squaredcos_cap_v2"))
        m, d = self._sampler
        z, o, c, y = self.synth.sampling_batch(n, T, seed=0)
        mask = self.synth.band_mask(T, BAND)
        x, times = z, []
        with torch.no_grad():
            for r in range(repeats):
                t = torch.full((2 * n,), STEPS_DIFF - 1 - r)
                t0 = time.perf_counter()
                x = d.p_sample(m.forward_with_cfg, x, t, clip_denoised=True,
                               model_kwargs=dict(o=o, c=c, y=y, cfg_scale=CFG, attn_mask=mask))["sample"]
                times.append(time.perf_counter() - t0)
        return times

    def train_seq_per_s(self, B=8, steps=2):
        m = self._model(TRAIN_MODEL, 0.2).train()
        opt = torch.optim.AdamW(m.parameters(), lr=1e-4, weight_decay=0)
        d = self.create("", noise_schedule="squaredcos_cap_v2", use_l1=True)
        (x, o, c), y = self.synth.training_batch(B, TRAIN_SEQ, seed=0)

        def step():
            t = torch.randint(0, d.num_timesteps, (B,))
            loss = d.training_losses(m, x, t, dict(o=o, c=c, y=y))["loss"].mean()
            opt.zero_grad()
            loss.backward()
            opt.step()

        step()
        t0 = time.perf_counter()
        for _ in range(steps):
            step()
        return B * steps / (time.perf_counter() - t0), B


class PortCPU:
    """Fallback when the reference modules are not on this machine: the oracle's torch-CPU restatement."""
    kind = "port"
    where = "oracle/"

    def denoise_step_seconds(self, n, T, repeats):
        from oracle import diffusion as odiff
        from oracle import dit as odit
        from osudit import synth
        shape = odit.shape_of(MODEL)
        if getattr(self, "_sd", None) is None:
            self._sd = odit.init_state_dict(shape, seed=1)
        sd = self._sd
        z, o, c, y = synth.sampling_batch(n, T, seed=0)
        mask = synth.band_mask(T, BAND)
        s = odiff.Schedule(str(STEPS_DIFF))
        g = torch.Generator().manual_seed(0)
        x, times = z, []
        with torch.no_grad():
            for r in range(repeats):
                i = STEPS_DIFF - 1 - r
                t = torch.full((2 * n,), i)
                t0 = time.perf_counter()
                out = odit.forward_with_cfg(sd, shape.heads, x, odiff.original_timesteps(s, t), o, c, y, CFG, mask)
                x = odiff.p_sample(s, out, x, t, torch.randn(x.shape, generator=g))["sample"]
                times.append(time.perf_counter() - t0)
        return times

    def train_seq_per_s(self, B=8, steps=2):
        from oracle import diffusion as odiff, dit as odit
        from osudit import synth
        shape = odit.shape_of(TRAIN_MODEL)
        sd = odit.init_state_dict(shape, seed=1)
        params = {k: v.requires_grad_(v.is_floating_point() and "playfield" not in k) for k, v in sd.items()}
        opt = torch.optim.AdamW([p for p in params.values() if p.requires_grad], lr=1e-4, weight_decay=0)
        s = odiff.Schedule("")
        (x, o, c), y = synth.training_batch(B, TRAIN_SEQ, seed=0)

        def step():
            t = torch.randint(0, 1000, (B,))
            loss = odiff.training_losses(s, lambda xt, tt: odit.forward(params, shape.heads, xt, tt, o, c, y),
                                         x, t, torch.randn_like(x), use_l1=True)["loss"].mean()
            opt.zero_grad()
            loss.backward()
            opt.step()

        step()
        t0 = time.perf_counter()
        for _ in range(steps):
            step()
        return B * steps / (time.perf_counter() - t0), B


def cpu_arm():
    root = None if os.environ.get("OSUDIT_BENCH_CPU_ARM") == "port" else reference_root()
    return ReferenceCPU(root) if root else PortCPU()


def cpu_baseline_sample():
    """The reference arm on a bounded sample, in its own process (the reference's module names collide with the
    drop-in's): 1 warm-up + 1 timed denoising step of one beatmap; returns its `cpu_baseline` object."""
    try:
        out = subprocess.run([sys.executable, os.path.abspath(__file__), "--impl", "reference", "--steps", "1",
                              "--warmup", "1"], capture_output=True, text=True, timeout=900, cwd=ROOT,
                             env={k: v for k, v in os.environ.items() if k not in ("RANK", "WORLD_SIZE", "LOCAL_RANK")})
        line = json.loads(out.stdout.strip().splitlines()[-1])
        cb = line["cpu_baseline"]
        cb["train"] = line.get("train")
        return cb
    except Exception as e:  # the baseline is a reported number, never a reason to lose the GPU measurement
        return {"value": None, "unit": UNIT, "cores": os.cpu_count(), "kind": "unavailable", "sample": f"failed: {e}"}


def run_reference(args, rank, world):
    if rank != 0:
        return None
    torch.set_num_threads(os.cpu_count() or 1)
    arm = cpu_arm()
    arm.denoise_step_seconds(1, SEQ, max(args.warmup, 1))
    ts = arm.denoise_step_seconds(1, SEQ, args.steps)  # each bench step = one denoising step of one beatmap
    per_step = per_sample = sum(ts) / len(ts)
    value = 1.0 / (per_step * STEPS_DIFF)
    train_v, train_b = arm.train_seq_per_s()
    what = "the unmodified reference modules (" + arm.where + ")" if arm.kind == "reference" else \
        "oracle/ (torch fp32 CPU restatement of the reference)"
    sample = (f"{what}: sample.py's p_sample(model.forward_with_cfg, ...) call on 1 beatmap (2 CFG rows) x {SEQ} "
              f"datapoints, fp32, {torch.get_num_threads()} threads; each bench step = 1 of the {STEPS_DIFF} denoising steps "
              f"({args.steps} timed after {max(args.warmup, 1)} warm-up), extrapolated linearly to the full sampling")
    return {"impl": "reference", "metric": METRIC, "value": round(value, 6), "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(per_sample * 1e3, 1),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(world),
            "cpu_baseline": {"value": round(value, 6), "unit": UNIT, "cores": torch.get_num_threads(),
                             "kind": arm.kind, "sample": sample},
            "e2e": {"value": round(value, 6), "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "train": {"metric": f"train seq/s {TRAIN_MODEL} seq-len {TRAIN_SEQ}", "value": round(train_v, 2),
                      "unit": "seq/s", "kind": arm.kind,
                      "sample": f"train.py:243-261's step (training_losses -> backward -> AdamW) in fp32 on the host "
                                f"cores, batch {train_b}, 2 steps timed after 1 warm-up"},
            "gpu_launches": 0}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--no-train", action="store_true", help="skip the training arm (the \"train\" record)")
    ap.add_argument("--model", default=MODEL, choices=sorted(SIZES),
                    help="sampling model (default: the headline DiT-B; DiT-XL + --diffusion-steps 1000 = BASELINE config 4)")
    ap.add_argument("--diffusion-steps", type=int, default=STEPS_DIFF)
    ap.add_argument("--no-cpu-baseline", action="store_true", help="skip the CPU reference sample (non-default workloads)")
    ap.add_argument("--no-e2e", action="store_true", help="skip the host-buffer leg (long non-default workloads only)")
    args = ap.parse_args()
    globals().update(MODEL=args.model, STEPS_DIFF=args.diffusion_steps)
    if (args.model, args.diffusion_steps) != ("DiT-B", 100):
        globals().update(METRIC=f"beatmaps/sec {args.model} {args.diffusion_steps}-step CFG sampling")
    # stdout carries exactly ONE JSON line: keep the real stdout aside and point fd 1 at stderr, so that anything a
    # library writes to C stdout (NCCL prints its version banner there when NCCL_DEBUG is set) cannot precede it
    sys.stdout.flush()
    out_fd = os.dup(1)
    os.dup2(2, 1)
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    local_rank = int(os.environ.get("LOCAL_RANK", 0))
    if args.impl == "reference":
        line = run_reference(args, rank, world)
    else:
        if not torch.cuda.is_available():
            raise SystemExit("bench.py: no CUDA device; the native arm has no CPU fallback")
        line = run_native(args, rank, world, local_rank)
    if rank == 0 and line is not None:
        os.write(out_fd, (json.dumps(line) + "\n").encode())


if __name__ == "__main__":
    main()
