"""Golden digests of the reference's SEEDED initial weights (models.py:243-304), UNMODIFIED reference, build container:

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden_init.py   ->  init_digest.json

For each case: `torch.manual_seed(seed); DiT_models[name](**kwargs)`, then per state-dict entry a CRC32 of its fp32
bytes, and the next `torch.rand(4)` of the global generator (the construction must leave the generator where the
reference leaves it, so that whatever is drawn next — label dropout, noise — matches too).
"""
import json
import os
import sys
import zlib

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.environ.get("OSU_DIFFUSION_REF", "/root/reference"))
import models  # noqa: E402  (reference)

CASES = [("DiT-S", 0, dict(num_classes=100, context_size=144)),
         ("DiT-S", 3, dict(num_classes=37, context_size=136, class_dropout_prob=0.0)),
         ("DiT-B", 1, dict(num_classes=10, context_size=144, class_dropout_prob=0.2))]
out = []
for name, seed, kw in CASES:
    torch.manual_seed(seed)
    m = models.DiT_models[name](**kw)
    nxt = torch.rand(4).tolist()
    out.append({"name": name, "seed": seed, "kwargs": kw, "next_rand": nxt,
                "crc": {k: "%08x" % zlib.crc32(v.detach().contiguous().numpy().tobytes())
                        for k, v in m.state_dict().items()}})
json.dump(out, open(os.path.join(HERE, "init_digest.json"), "w"), separators=(",", ":"))
print([len(c["crc"]) for c in out])
