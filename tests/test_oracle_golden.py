"""Pin ``oracle/`` against outputs of the unmodified reference (tests/golden/*.npz,
written by tests/golden/make_golden.py) and against SURVEY.md §A.4's schedule KATs."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import diffusion as odiff
from oracle import dit as odit
from osudit import synth


def _load(golden_dir, name):
    z = np.load(os.path.join(golden_dir, name))
    return {k: torch.from_numpy(z[k]) for k in z.files}


def _rel(a, b):
    return float((a - b).norm() / b.norm())


def _tiny_sd(g):
    return {k[3:]: v for k, v in g.items() if k.startswith("sd.")}


def test_tiny_forward_matches_reference(golden_dir):
    g = _load(golden_dir, "tiny_forward.npz")
    sd = _tiny_sd(g)
    args = (g["x"], g["t"], g["o"], g["c"], g["y"])
    assert _rel(odit.forward(sd, 2, *args), g["out_nomask"]) < 2e-6
    assert _rel(odit.forward(sd, 2, *args, attn_mask=g["mask"]), g["out_mask"]) < 2e-6
    cfg = odit.forward_with_cfg(sd, 2, *args, cfg_scale=1.5, attn_mask=g["mask"])
    assert _rel(cfg, g["out_cfg"]) < 2e-6
    # fp64 oracle agrees with the fp32 reference to fp32 round-off
    f64 = odit.forward(sd, 2, *args, attn_mask=g["mask"], dtype=torch.float64)
    assert _rel(f64.float(), g["out_mask"]) < 5e-6


def test_mask_changes_the_answer(golden_dir):
    g = _load(golden_dir, "tiny_forward.npz")
    assert _rel(g["out_nomask"], g["out_mask"]) > 1e-3


def test_dit_s_forward_on_seeded_weights(golden_dir):
    """Weights regenerated from the seed here; the reference output was stored."""
    g = _load(golden_dir, "dit_s_forward.npz")
    shape = odit.shape_of("DiT-S")
    sd = odit.init_state_dict(shape, seed=1)
    z, o, c, y = synth.sampling_batch(1, 256, seed=0)
    out = odit.forward_with_cfg(sd, shape.heads, z, g["t"], o, c, y, 1.5, synth.band_mask(256, 128))
    assert out.abs().max() > 1e-3  # not the zero-init trap
    assert _rel(out, g["out"]) < 5e-6


def test_registry_layout(golden_dir):
    with open(os.path.join(golden_dir, "registry_layout.json")) as f:
        layout = json.load(f)
    for name in odit.SIZES:
        shape = odit.shape_of(name)
        assert layout[name + ".num_heads"] == shape.heads
        assert [[k, s] for k, s in odit.state_dict_layout(shape)] == layout[name]
        assert layout[name][7] == ["y_embedder.embedding_table.weight", [52671, shape.hidden]]
        assert layout[name + ".param_order"] == [k for k, _ in layout[name]]
    sd = odit.init_state_dict(odit.shape_of("DiT-S"), seed=0)
    assert [[k, list(v.shape)] for k, v in sd.items()] == layout["DiT-S"]


@pytest.mark.parametrize("tag,resp,sched", [
    ("c100", "100", "squaredcos_cap_v2"), ("c250", "250", "squaredcos_cap_v2"),
    ("c1000", "", "squaredcos_cap_v2"), ("l50", "50", "linear"), ("c10_20", "10,20", "squaredcos_cap_v2")])
def test_schedule_tables_bit_exact(golden_dir, tag, resp, sched):
    z = np.load(os.path.join(golden_dir, "schedule.npz"))
    s = odiff.Schedule(resp, sched)
    assert list(z[tag + ".timestep_map"]) == s.timestep_map
    for name in ("betas", "alphas_cumprod", "sqrt_recip_alphas_cumprod", "sqrt_recipm1_alphas_cumprod",
                 "posterior_log_variance_clipped", "posterior_mean_coef1", "posterior_mean_coef2",
                 "sqrt_alphas_cumprod", "sqrt_one_minus_alphas_cumprod", "posterior_variance"):
        np.testing.assert_array_equal(z[f"{tag}.{name}"], getattr(s, name), err_msg=name)


def test_schedule_known_answers():
    """SURVEY.md §A.4 (computed from the reference)."""
    s = odiff.Schedule("100")
    assert s.timestep_map[:12] == [0, 10, 20, 30, 40, 50, 61, 71, 81, 91, 101, 111]
    assert s.timestep_map[-3:] == [979, 989, 999] and len(s.timestep_map) == 100
    np.testing.assert_allclose(s.betas[[0, 1, 50, 99]],
                               [4.1284224822e-05, 6.7984006851e-04, 3.4171701337e-02, 9.9998999920e-01], rtol=1e-9)
    np.testing.assert_allclose(s.alphas_cumprod[[0, 50, 99]],
                               [9.9995871578e-01, 4.8449452062e-01, 2.4287669070e-09], rtol=1e-9)
    np.testing.assert_allclose(s.posterior_log_variance_clipped[[0, 1, 99]],
                               [-10.153945111, -10.153945111, -2.5288514518e-04], rtol=1e-9)
    np.testing.assert_allclose(s.sqrt_recip_alphas_cumprod[99], 20291.169634711496, rtol=1e-12)
    np.testing.assert_allclose(s.sqrt_recipm1_alphas_cumprod[99], 20291.169610070236, rtol=1e-12)
    assert odiff.Schedule("250").timestep_map[:4] == [0, 4, 8, 12]
    full = odiff.Schedule("")
    assert full.num_timesteps == 1000 and full.betas[999] == 0.999


def test_sampling_loop_matches_reference(golden_dir):
    g = _load(golden_dir, "tiny_sampling.npz")
    sd = _tiny_sd(_load(golden_dir, "tiny_forward.npz"))
    s = odiff.Schedule("10")
    rec = []

    def model_fn(x, t):
        return odit.forward_with_cfg(sd, 2, x, t, g["o"], g["c"], g["y"], 1.5, g["mask"])

    # The free-running map is chaotic on random weights (fp32 round-off order
    # alone moves the end point), so the gate is teacher-forced: the reference's own x_in
    # at every step.  The free-running loop must still run and stay in the clamp range.
    final = odiff.p_sample_loop(s, model_fn, g["z"], g["noises"], record=rec)
    assert len(rec) == 10 and torch.isfinite(final).all()
    for i in reversed(range(10)):
        x_in = g["z"] if i == 9 else g["samples"][i + 1]
        t = torch.full((4,), i)
        out = odiff.p_sample(s, model_fn(x_in, odiff.original_timesteps(s, t)), x_in, t, g["noises"][i])
        assert _rel(out["sample"], g["samples"][i]) < 1e-4, i
        assert _rel(out["pred_xstart"], g["pred_xstart"][i]) < 1e-4, i


def test_inpaint_callback_before_clamp(golden_dir):
    g = _load(golden_dir, "tiny_psample_inpaint.npz")
    f = _load(golden_dir, "tiny_sampling.npz")
    sd = _tiny_sd(_load(golden_dir, "tiny_forward.npz"))
    s = odiff.Schedule("10")
    keep, target = g["keep"].bool(), g["target"]
    out_m = odit.forward(sd, 2, g["z"], odiff.original_timesteps(s, g["t"]), f["o"], f["c"], f["y"])
    out = odiff.p_sample(s, out_m, g["z"], g["t"], g["noise"],
                         denoised_fn=lambda x0: torch.where(keep, target, x0))
    assert _rel(out["sample"], g["sample"]) < 1e-5
    assert _rel(out["pred_xstart"], g["pred_xstart"]) < 1e-5


@pytest.mark.parametrize("tag,l1", [("l1", True), ("mse", False)])
def test_training_losses_and_grads(golden_dir, tag, l1):
    g = _load(golden_dir, "tiny_training.npz")
    sd = {k: v.clone().requires_grad_(v.is_floating_point() and "playfield" not in k)
          for k, v in _tiny_sd(_load(golden_dir, "tiny_forward.npz")).items()}
    s = odiff.Schedule("")

    def model_fn(x, t):
        return odit.forward(sd, 2, x, t, g["o"], g["c"], g["y"])

    terms = odiff.training_losses(s, model_fn, g["x"], g["t"], g["noise"], use_l1=l1)
    for k in ("loss", "vb", tag):
        torch.testing.assert_close(terms[k], g[f"{tag}.{k}"], rtol=2e-5, atol=1e-6)
    terms["loss"].mean().backward()
    assert _rel(sd["final_layer.linear.weight"].grad, g[f"{tag}.grad.final_w"]) < 1e-4
    assert _rel(sd["blocks.0.attn.in_proj_weight"].grad, g[f"{tag}.grad.qkv0"]) < 1e-4
    assert _rel(sd["xoc_embedder.mlp.0.weight"].grad, g[f"{tag}.grad.first_w"]) < 1e-4


def test_eager_library_restatement_matches_oracle():
    """oracle/eager_cuda.py (the stock-torch call sequence timed on the GPU as 'the real bar') computes the
    same function as the pinned oracle, with and without the band mask and CFG."""
    from oracle import eager_cuda
    shape = odit.shape_of("DiT-S")
    sd = odit.init_state_dict(shape, seed=3, zero_init_std=0.02)
    T = 96
    z, o, c, y = synth.sampling_batch(1, T, seed=1)
    t = torch.tensor([700, 31])
    mask = synth.band_mask(T, 16)
    for mk in (None, mask):
        a = eager_cuda.forward(sd, shape.heads, z, t, o, c, y, mk)
        b = odit.forward(sd, shape.heads, z, t, o, c, y, mk)
        assert _rel(a, b) < 1e-5
    a = eager_cuda.forward_with_cfg(sd, shape.heads, z, t, o, c, y, 1.5, mask)
    b = odit.forward_with_cfg(sd, shape.heads, z, t, o, c, y, 1.5, mask)
    assert _rel(a, b) < 1e-5


def test_feature_builder_oracle_matches_reference_golden(golden_dir):
    """oracle/features.py against data_loading.py's own outputs (tests/golden/make_golden_features.py): bit-exact."""
    from oracle import features as ofeat
    g = _load(golden_dir, "features.npz")
    x, o, c = ofeat.beatmap_features(g["seq"])
    assert torch.equal(x, g["x"]) and torch.equal(o, g["o_sampling"]) and torch.equal(c, g["c"])
    assert torch.equal(ofeat.calc_distances(g["seq"].clone()), g["dist"])
    s, e = [int(v) for v in g["win"]]
    _, ow, _ = ofeat.beatmap_features(g["seq"][:, s:e], float(g["shift"]))
    assert torch.equal(ow, g["ow"])
    # synthetic inputs used by the benches are built the same way
    xs, os_, cs = synth.beatmap_features(64, seed=5)
    assert cs.shape == (144, 64) and float(os_[0]) == 0.0 and float(xs.max()) <= 1.0


def test_diffusion_utils_restatement_covers_every_branch(golden_dir):
    """oracle._normal_kl / _std_normal_cdf / _disc_gauss_loglik against the reference's diffusion_utils on a grid that
    reaches the edge bins (x < -0.999, x > 0.999), the 1e-12 clamp and extreme log-variances
    (tests/golden/make_golden_utils.py)."""
    z = {k: torch.from_numpy(v) for k, v in np.load(os.path.join(golden_dir, "diffusion_utils.npz")).items()}
    torch.testing.assert_close(odiff._disc_gauss_loglik(z["x"], z["mean"], z["log_scale"]), z["loglik"],
                               rtol=2e-6, atol=1e-6)
    torch.testing.assert_close(odiff._normal_kl(z["m1"], z["lv1"], z["m2"], z["lv2"]), z["kl"], rtol=2e-6, atol=1e-6)
    torch.testing.assert_close(odiff._std_normal_cdf(z["z"]), z["cdf"], rtol=2e-6, atol=1e-7)
    assert float(z["loglik"].min()) < -27 and float(z["loglik"].max()) == 0.0  # clamp and saturated bins are in the grid
