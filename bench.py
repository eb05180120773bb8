#!/usr/bin/env python
"""Headline benchmark: beatmaps/sec of DiT-B 100-step CFG sampling (BASELINE.json config 2).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl native|reference]

One "step" = one pass of the hot path over one batch: the full 100-step classifier-free-guidance
sampling of 64 synthetic 2048-datapoint beatmaps (128 model rows) through the public API
`create_diffusion("100", "squaredcos_cap_v2").p_sample_loop(model.forward_with_cfg, ...)`.
N>1 (launched by torchrun) shards independent beatmaps over ranks with no collective on the data
path (weak scaling: 64 beatmaps per GPU); timing is CUDA events, max over ranks.

`--impl reference` times the reference algorithm's CPU port (oracle/, torch fp32 on all host
threads; the reference itself is Python and absent on the GPU box) on a bounded sample of the same
workload.  Prints exactly one JSON line on rank 0.
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
NCU_GEMM_TRAFFIC_GB = 1.57  # ncu --set full, profiles/r01d_summary.md: (1.560 + 0.765 + 1.963 + 2.011) / 4
UNIT = "beatmaps/s"


def workload_config(n_gpus):
    return {"workload": f"{MODEL} sampling, {SEQ}-datapoint synthetic beatmaps, {STEPS_DIFF} respaced "
                        f"steps (squaredcos_cap_v2), CFG {CFG} with style class, batch {N_BEATMAPS} "
                        f"beatmaps ({2 * N_BEATMAPS} model rows) per GPU, band mask W={BAND}",
            "beatmaps_per_gpu": N_BEATMAPS, "datapoints": SEQ, "diffusion_steps": STEPS_DIFF,
            "cfg_scale": CFG, "parallelism": f"independent beatmaps x{n_gpus} (no collective)",
            "weights": "seeded random init, zero-init layers redrawn N(0,0.02^2) (SURVEY F4)",
            "l2_policy": "inputs and activations (>5 GB per step) exceed the 126 MB L2; no flush needed"}


def flops_per_beatmap():
    """Algorithmic flop count, attention over the +-128 band only."""
    D, depth, H, hd = 768, 12, 12, 64
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
        sample_e2e()
        ms_e2e = timed(sample_e2e, args.steps)
        roof = kernel_roofline(model, diffusion, zd, od, cd, yd, mask_d, device) if rank == 0 else None

    total = N_BEATMAPS * world * args.steps
    value = total / (ms / 1e3)
    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return None
    line = {
        "metric": METRIC, "value": round(value, 4), "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": round(ms / args.steps, 3), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": workload_config(world), "clocks": clk,
        "e2e": {"value": round(total / (ms_e2e / 1e3), 4), "unit": UNIT,
                "h2d_bytes_per_step": sum(t.numel() * t.element_size() for t in host),
                "d2h_bytes_per_step": out_host.numel() * out_host.element_size()},
        "gpu_launches": launches,
        "model_tflops": round(flops_per_beatmap() * value / world / 1e12, 1),
        "roofline": roof,
    }
    line["cpu_baseline"] = cpu_baseline_sample(sample_steps=1) if world == 1 else None
    if dist is not None:
        dist.destroy_process_group()
    return line


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
    real_gemm, real_ln, real_attn = ops.gemm, ops.ln_modulate, ops.attn_band

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

    def ln_work(x, branch, mod, g, sh, sc, T, h):
        return x.numel() * (12.0 if branch is not None else 6.0), None

    def attn_work(qkv, out, B, T, H, hd, wl=-1, wr=-1, mask=None):
        pairs = sum(min(T - 1, j + wr) - max(0, j - wl) + 1 for j in range(T)) if wl >= 0 else T * T
        return 4.0 * hd * pairs * H * B, None

    ops.gemm = wrap("gemm", real_gemm, gemm_work)
    ops.ln_modulate = wrap("ln", real_ln, ln_work)
    ops.attn_band = wrap("attn", real_attn, attn_work)
    try:
        t = torch.full((zd.shape[0],), 50, device=device, dtype=torch.long)
        for _ in range(3):
            diffusion.p_sample(model.forward_with_cfg, zd, t, clip_denoised=True,
                               model_kwargs=dict(o=od, c=cd, y=yd, cfg_scale=CFG, attn_mask=mask_d))
        torch.cuda.synchronize()
    finally:
        ops.gemm, ops.ln_modulate, ops.attn_band = real_gemm, real_ln, real_attn

    def agg(items, pred=lambda tag: True):
        items = [(e0.elapsed_time(e1), w) for e0, e1, (w, tag) in items if pred(tag)]
        ms = sum(m for m, _ in items)
        return sum(w for _, w in items), ms, len(items)

    big = lambda tag: tag is not None and tag[0] >= 65536 and tag[2] >= 768  # noqa: E731  (block GEMMs)
    fl, ms, n = agg(rec["gemm"], big)
    by_shape = {}
    for e0, e1, (w, tag) in rec["gemm"]:
        if big(tag):
            d = by_shape.setdefault("x".join(map(str, tag)), [0.0, 0.0])
            d[0] += w; d[1] += e0.elapsed_time(e1)
    lb, lms, ln_n = agg(rec["ln"])
    af, ams, an = agg(rec["attn"])
    step_ms = sum(e0.elapsed_time(e1) for k in rec for e0, e1, _ in rec[k])
    ach = fl / (ms / 1e3) / 1e12
    return {"bound": "tensor", "kernel": "g2::gemm2_kernel (cta_group::2 tcgen05 GEMM: QKV, out-proj, fc1+GELU, fc2 "
            "launches)",
            "achieved": round(ach, 1), "peak": tf_peak, "unit": "TFLOP/s", "frac": round(ach / tf_peak, 4),
            "traffic": NCU_GEMM_TRAFFIC_GB, "traffic_unit": "GB per launch (dram read+write, mean of the QKV/out-proj/"
            "fc1/fc2 launches in profiles/r01d_summary.md; algorithmic 1.56)", "peak_source": src, "launches_timed": n,
            "avg_launch_ms": round(ms / max(n, 1), 4), "share_of_step": round(ms / step_ms, 3),
            "per_shape_tflops": {k: round(v[0] / (v[1] / 1e3) / 1e12, 1) for k, v in by_shape.items()},
            "hbm_kernel": {"kernel": "ln_modulate_kernel", "bound": "hbm",
                           "achieved": round(lb / (lms / 1e3) / 1e9, 1), "peak": hbm_peak, "unit": "GB/s",
                           "frac": round(lb / (lms / 1e3) / 1e9 / hbm_peak, 4), "launches_timed": ln_n,
                           "share_of_step": round(lms / step_ms, 3)},
            "attention_kernel": {"kernel": "attn_tc::attn_window_kernel", "achieved": round(af / (ams / 1e3) / 1e12, 1),
                                 "unit": "TFLOP/s (banded algorithmic flops)", "share_of_step": round(ams / step_ms, 3)}}


# ------------------------------------------------------------------------ CPU reference arm
def cpu_denoise_step_seconds(n, T, repeats):
    """Time `repeats` CFG denoising steps of the oracle port (fp32, all host threads)."""
    from oracle import diffusion as odiff
    from oracle import dit as odit
    from osudit import synth
    shape = odit.shape_of(MODEL)
    sd = odit.init_state_dict(shape, seed=1)
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


def cpu_baseline_sample(sample_steps=1):
    torch.set_num_threads(os.cpu_count() or 1)
    cpu_denoise_step_seconds(1, SEQ, 1)  # warm-up (allocator, thread pool)
    ts = cpu_denoise_step_seconds(1, SEQ, sample_steps)
    per_step = sum(ts) / len(ts)
    return {"value": round(1.0 / (per_step * STEPS_DIFF), 6), "unit": UNIT, "cores": torch.get_num_threads(),
            "kind": "port",
            "sample": f"oracle/ (torch fp32 CPU restatement of the reference) on 1 beatmap (2 CFG rows) x {SEQ} "
                      f"datapoints, {sample_steps} of {STEPS_DIFF} denoising steps timed after 1 warm-up step, "
                      f"extrapolated linearly to {STEPS_DIFF} steps"}


def run_reference(args, rank, world):
    if rank != 0:
        return None
    torch.set_num_threads(os.cpu_count() or 1)
    for _ in range(args.warmup):
        cpu_denoise_step_seconds(1, SEQ, 1)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cpu_denoise_step_seconds(1, SEQ, 1)
    per_sample = (time.perf_counter() - t0) / args.steps  # one denoising step of one beatmap
    value = 1.0 / (per_sample * STEPS_DIFF)
    sample = (f"each bench step = 1 of {STEPS_DIFF} denoising steps of 1 beatmap (2 CFG rows) x {SEQ} datapoints "
              f"on the oracle port, extrapolated linearly to the full {STEPS_DIFF}-step sampling")
    return {"impl": "reference", "metric": METRIC, "value": round(value, 6), "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(per_sample * 1e3, 1),
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(world),
            "cpu_baseline": {"value": round(value, 6), "unit": UNIT, "cores": torch.get_num_threads(),
                             "kind": "port", "sample": sample},
            "e2e": {"value": round(value, 6), "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    args = ap.parse_args()
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
