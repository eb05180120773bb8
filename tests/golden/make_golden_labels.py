"""Golden for the reference's label dropout (models.py:56-67), from the UNMODIFIED reference, build container only:

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden_labels.py   ->  labels.json
"""
import json
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.environ.get("OSU_DIFFUSION_REF", "/root/reference"))
from models import LabelEmbedder  # noqa: E402  (reference)

labels = torch.randint(0, 52670, (64,), generator=torch.Generator().manual_seed(11))
force = (torch.arange(64) % 3 == 0).long()
out = {"labels": labels.tolist(), "force": force.tolist(), "cases": []}
for p in (0.1, 0.2, 0.5):
    emb = LabelEmbedder(52670, 8, p)
    torch.manual_seed(5)
    a = emb.token_drop(labels)
    b = emb.token_drop(labels)                      # the next draw of the same generator
    c = emb.token_drop(labels, force_drop_ids=force)
    out["cases"].append({"p": p, "table_rows": emb.embedding_table.weight.shape[0],
                         "first": a.tolist(), "second": b.tolist(), "forced": c.tolist()})
out["table_rows_without_dropout"] = LabelEmbedder(52670, 8, 0.0).embedding_table.weight.shape[0]
json.dump(out, open(os.path.join(HERE, "labels.json"), "w"), separators=(",", ":"))
print([sum(v == 52670 for v in c["first"]) for c in out["cases"]], out["table_rows_without_dropout"])
