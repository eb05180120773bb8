"""Single-beatmap sampling latency (what the reference's sample.py runs): eager launches vs the
CUDA-graph replay of osudit/graphs.py.  Prints one JSON line per (model, T, beatmaps)."""
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "osu-diffusion_b200"))

import models  # noqa: E402
from diffusion import create_diffusion  # noqa: E402
from osudit import graphs, synth  # noqa: E402


@torch.no_grad()
def main():
    dev = "cuda"
    for name, T, n in (("DiT-B", 2048, 1), ("DiT-B", 1024, 1), ("DiT-B", 2048, 8), ("DiT-S", 2048, 1),
                       ("DiT-XL", 2048, 1)):
        m = models.DiT_models[name](num_classes=52670, context_size=144).to(dev).eval()
        for p in m.parameters():
            if p.dim() > 1 and float(p.abs().max()) == 0:
                torch.nn.init.normal_(p, std=0.02)
        z, o, c, y = synth.sampling_batch(n, T, seed=0)
        kw = dict(o=o.to(dev), c=c.to(dev), y=y.to(dev), cfg_scale=1.5, attn_mask=synth.band_mask(T, 128).to(dev))
        d = create_diffusion("100", noise_schedule="squaredcos_cap_v2")
        res = {}
        for mode in ("eager", "graph"):
            graphs._ENABLED = mode == "graph"
            graphs._cache.clear()
            for rep in range(3):  # rep 0 warms up (and captures)
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                out = d.p_sample_loop(m.forward_with_cfg, z.shape, z.to(dev), model_kwargs=kw, device=dev)
                out.cpu()
                res[mode] = time.perf_counter() - t0
        print(json.dumps(dict(model=name, T=T, beatmaps=n, eager_s=round(res["eager"], 4),
                              graph_s=round(res["graph"], 4),
                              speedup=round(res["eager"] / res["graph"], 2))), flush=True)
        del m


if __name__ == "__main__":
    main()
