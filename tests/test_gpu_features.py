"""Device-side beatmap feature builder against the reference's own outputs
(tests/golden/features.npz, from data_loading.py with a stubbed `slider`) and the CPU oracle."""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import features as ofeat  # noqa: E402
from osudit import data  # noqa: E402

DEV = "cuda"
GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "features.npz")


def test_features_match_reference_golden():
    g = {k: torch.from_numpy(np.asarray(v)) for k, v in np.load(GOLD).items()}
    seq = g["seq"].to(DEV)
    (x, o_abs, c), n = data.split_and_process_sequence_no_augment(seq)
    assert n == seq.shape[1] and c.shape == (144, n)
    assert torch.equal(x.cpu(), g["x"]) and torch.equal(o_abs.cpu(), g["o"])
    assert torch.equal(c[128:].cpu(), g["c"][128:])                    # one-hot rows: exact
    assert float((c[:128].cpu() - g["c"][:128]).abs().max()) < 2e-4    # sin/cos of up to ~600 rad
    _, o_rel, _ = data.beatmap_features(seq, want_x=False)
    assert torch.equal(o_rel.cpu(), g["o_sampling"])                   # sample.py:65
    s, e = [int(v) for v in g["win"]]
    win = seq[:, s:e].contiguous()
    shift = torch.full((1,), float(g["shift"]), device=DEV)
    xw, ow, cw = data.beatmap_features(win.unsqueeze(0), o_shift=shift)
    assert torch.equal(xw[0].cpu(), g["xw"]) and torch.equal(ow[0].cpu(), g["ow"])
    # a window's first distance is measured from the playfield centre, the reference slices the full-map c instead
    assert float((cw[0, :128, 1:].cpu() - g["cw"][:128, 1:]).abs().max()) < 2e-4
    assert float((data.calc_distances(seq).cpu() - g["dist"]).abs().max()) < 1e-4


def test_features_batch_matches_oracle():
    gen = torch.Generator().manual_seed(3)
    B, T = 5, 333
    seq = torch.zeros(B, 19, T)
    seq[:, 0] = torch.rand(B, T, generator=gen) * 512
    seq[:, 1] = torch.rand(B, T, generator=gen) * 384
    seq[:, 2] = torch.cumsum(torch.randint(20, 900, (B, T), generator=gen).float(), 1) + 777.0
    seq[:, 3:] = torch.nn.functional.one_hot(torch.randint(0, 16, (B, T), generator=gen), 16).float().transpose(1, 2)
    shift = torch.rand(B, generator=gen) * 1e5
    x, o, c = data.beatmap_features(seq.to(DEV), o_shift=shift.to(DEV))
    for b in range(B):
        xr, orr, cr = ofeat.beatmap_features(seq[b], float(shift[b]))
        assert torch.equal(x[b].cpu(), xr)
        assert float((o[b].cpu() - orr).abs().max()) <= 1e-2  # fp32 ulp at ~3e5 ms
        assert float((c[b].cpu() - cr).abs().max()) < 2e-4
