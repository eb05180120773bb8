"""BASELINE config 4 data point: DiT-XL sampling, 64 beatmaps x 2048 datapoints per GPU (the per-GPU share of 512
beatmaps over 8 GPUs, no collective), CFG 1.5, band W=128.  A full 1000-step pass takes minutes per GPU, so this
times K denoising steps of the 1000-step schedule (t spread over the schedule) and reports ms per denoising step and
the 1000-step extrapolation, next to the algorithmic roofline of SURVEY §8(d).  Not the headline bench."""
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "osu-diffusion_b200"))

import models  # noqa: E402
from diffusion import create_diffusion  # noqa: E402
from osudit import synth  # noqa: E402


@torch.no_grad()
def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "DiT-XL"
    K = int(sys.argv[2]) if len(sys.argv) > 2 else 10
    n, T, W, total_steps = 64, 2048, 128, 1000
    dev = torch.device("cuda", 0)
    m = models.DiT_models[name](num_classes=52670, context_size=144)
    g = torch.Generator().manual_seed(1)
    for k, v in m.state_dict().items():
        if "adaLN_modulation" in k or k.startswith("final_layer.linear"):
            v.copy_(torch.randn(v.shape, generator=g) * 0.02)
    m = m.to(dev).eval()
    d = create_diffusion(str(total_steps), noise_schedule="squaredcos_cap_v2")
    z, o, c, y = [t.to(dev) for t in synth.sampling_batch(n, T, seed=0)]
    kw = dict(o=o, c=c, y=y, cfg_scale=1.5, attn_mask=synth.band_mask(T, W).to(dev))
    ts = [int(v) for v in torch.linspace(total_steps - 1, 0, K)]

    def run():
        x = z
        for i in ts:
            x = d.p_sample(m.forward_with_cfg, x, torch.full((2 * n,), i, device=dev), model_kwargs=kw)["sample"]
        return x

    run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(2):
        run()
    e1.record()
    torch.cuda.synchronize()
    ms_step = e0.elapsed_time(e1) / (2 * K)
    if os.environ.get("OSUDIT_KERNEL_TABLE"):  # per-kernel device time of one pass, to stderr
        from torch.autograd import DeviceType
        from torch.profiler import ProfilerActivity, profile
        with profile(activities=[ProfilerActivity.CUDA]) as prof:
            run()
            torch.cuda.synchronize()
        ka = [e for e in prof.key_averages() if e.device_type == DeviceType.CUDA]
        tot = sum(e.device_time_total for e in ka) / K / 1e3
        print("GPU busy %.2f ms per denoising step" % tot, file=sys.stderr)
        for e in sorted(ka, key=lambda e: -e.device_time_total)[:16]:
            print("  %-90s %8.3f ms  x%d  (%.1f%%)" % (e.key[:90], e.device_time_total / K / 1e3, e.count // K,
                                                      e.device_time_total / K / 10 / tot), file=sys.stderr)
    D, depth, H = m.hidden_size, len(m.blocks), m.num_heads
    pairs = sum(min(T - 1, j + W) - max(0, j - (W - 1)) + 1 for j in range(T))
    flop_step = 2 * n * (T * (24 * D * D * depth + 2 * 528 * D + 8 * D) + pairs * 4 * D * depth)
    peak = 1406.1e12
    print(json.dumps({
        "workload": f"{name} sampling, {n} beatmaps x {T} datapoints per GPU, CFG 1.5, band W={W}, {K} of "
                    f"{total_steps} denoising steps timed",
        "ms_per_denoising_step": round(ms_step, 2),
        "beatmaps_per_s_per_gpu_at_1000_steps": round(n / (ms_step * total_steps / 1e3), 4),
        "model_tflops": round(flop_step / (ms_step / 1e3) / 1e12, 1),
        "roofline_beatmaps_per_s_per_gpu": round(n / (flop_step * total_steps / peak), 4),
        "fraction_of_roofline": round(flop_step / (ms_step / 1e3) / peak, 3)}))


if __name__ == "__main__":
    main()
