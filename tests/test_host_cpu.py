"""CPU-side checks: the C-ABI library loads and exports what include/osudit.h declares, the
drop-in modules keep the reference's surface (names, shapes, order, schedules), and the product
path refuses to run without CUDA instead of falling back."""
import ctypes
import json
import os
import re

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    from osudit import _lib
    with open(os.path.join(ROOT, "include", "osudit.h")) as f:
        declared = set(re.findall(r"\b(osudit_\w+)\s*\(", f.read()))
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), name
    assert _lib.load().osudit_version() == 2


def test_no_reference_or_oracle_imports_in_product():
    """The product path must not route through the oracle or read /root/reference."""
    pkg = os.path.join(ROOT, "osu-diffusion_b200")
    for dirpath, _, files in os.walk(pkg):
        for fn in files:
            if fn.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, fn)).read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", src, re.M), fn
                assert "sys.path" not in src or "reference" not in src, fn


@pytest.mark.parametrize("name", ["DiT-S", "DiT-B", "DiT-L", "DiT-XL"])
def test_parameter_tree_matches_reference(golden_dir, name):
    import models
    with open(os.path.join(golden_dir, "registry_layout.json")) as f:
        layout = json.load(f)
    with torch.device("meta"):
        m = models.DiT_models[name](num_classes=52670, context_size=144)
    assert [[k, list(v.shape)] for k, v in m.state_dict().items()] == layout[name]
    assert [k for k, _ in m.named_parameters()] == layout[name + ".param_order"]
    assert m.num_heads == layout[name + ".num_heads"]
    assert [k for k, p in m.named_parameters() if not p.requires_grad] == ["xoc_embedder.playfield_size"]


def test_constructor_kwargs_and_init_statistics():
    import models
    torch.manual_seed(0)
    m = models.DiT(hidden_size=128, depth=2, num_heads=2, num_classes=10, context_size=144,
                   class_dropout_prob=0.2)
    sd = m.state_dict()
    assert sd["y_embedder.embedding_table.weight"].shape == (11, 128)
    for k, v in sd.items():
        if "adaLN_modulation" in k or k.startswith("final_layer.linear"):
            assert float(v.abs().max()) == 0.0, k  # zero-init (reference models.py:295-304)
        elif k.endswith("bias"):
            assert float(v.abs().max()) == 0.0, k
    assert abs(float(sd["xoc_embedder.mlp.0.weight"].std()) - 0.02) < 2e-3
    bound = (6.0 / (128 + 384)) ** 0.5
    assert float(sd["blocks.0.attn.in_proj_weight"].abs().max()) <= bound
    m2 = models.DiT(hidden_size=128, depth=1, num_heads=2, num_classes=10, class_dropout_prob=0.0, context_size=144)
    assert m2.state_dict()["y_embedder.embedding_table.weight"].shape == (10, 128)
    with pytest.raises(ValueError, match="multiple of 8"):  # the constructor default (142) gives a 526-wide first layer,
        models.DiT(hidden_size=128, depth=1, num_heads=2, num_classes=10)  # which no TMA row can hold: rejected up front
    import copy
    m3 = copy.deepcopy(m)  # EMA copy (train.py:147)
    assert m3._engine is None and list(m3.state_dict()) == list(sd)


def test_cpu_call_fails_loudly():
    import models
    from osudit import synth
    m = models.DiT(hidden_size=128, depth=1, num_heads=2, num_classes=10, context_size=144).eval()
    z, o, c, y = synth.sampling_batch(1, 32, seed=0, num_classes=10)
    with pytest.raises(RuntimeError, match="no CPU fallback"), torch.no_grad():
        m(z, torch.tensor([0, 0]), o=o, c=c, y=y)
    with pytest.raises(RuntimeError, match="no CPU fallback"), torch.no_grad():
        m.forward_with_cfg(z, torch.tensor([0, 0]), o=o, c=c, y=y, cfg_scale=2.0)


@pytest.mark.parametrize("tag,resp,sched", [
    ("c100", "100", "squaredcos_cap_v2"), ("c250", "250", "squaredcos_cap_v2"),
    ("c1000", "", "squaredcos_cap_v2"), ("l50", "50", "linear"), ("c10_20", "10,20", "squaredcos_cap_v2")])
def test_create_diffusion_tables_bit_exact(golden_dir, tag, resp, sched):
    from diffusion import create_diffusion
    z = np.load(os.path.join(golden_dir, "schedule.npz"))
    d = create_diffusion(resp, noise_schedule=sched)
    assert list(z[tag + ".timestep_map"]) == d.timestep_map
    assert d.num_timesteps == len(d.timestep_map)
    for name in ("betas", "alphas_cumprod", "sqrt_recip_alphas_cumprod", "sqrt_recipm1_alphas_cumprod",
                 "posterior_log_variance_clipped", "posterior_mean_coef1", "posterior_mean_coef2",
                 "sqrt_alphas_cumprod", "sqrt_one_minus_alphas_cumprod", "posterior_variance"):
        np.testing.assert_array_equal(z[f"{tag}.{name}"], getattr(d, name), err_msg=name)


def test_create_diffusion_surface():
    from diffusion import create_diffusion, gaussian_diffusion as gd, space_timesteps
    d = create_diffusion("", noise_schedule="squaredcos_cap_v2", use_l1=True)
    assert d.loss_type == gd.LossType.L1 and d.model_var_type == gd.ModelVarType.LEARNED_RANGE
    assert d.model_mean_type == gd.ModelMeanType.EPSILON and d.original_num_steps == 1000
    assert create_diffusion("").loss_type == gd.LossType.MSE
    assert sorted(space_timesteps(1000, "ddim50"))[:3] == [0, 20, 40]
    assert sorted(space_timesteps(300, [10, 15, 20]))[:3] == [0, 11, 22]
    with pytest.raises(ValueError):
        space_timesteps(10, "20")
    for name in ("p_sample", "p_sample_loop", "p_sample_loop_progressive", "training_losses",
                 "q_sample", "p_mean_variance"):
        assert callable(getattr(d, name))


def test_mask_classification_closed_form():
    from osudit import synth
    from osudit.engine import classify_mask
    T = 300
    loop = torch.full((T, T), True)
    for i in range(T):  # the reference's loop, sample.py:81-84
        loop[max(0, i - 128): min(T, i + 128), i] = False
    spec = classify_mask(loop, T)
    assert (spec.w_left, spec.w_right, spec.generic) == (127, 128, None)
    assert torch.equal(loop, synth.band_mask(T, 128))
    none = classify_mask(None, T)
    assert (none.w_left, none.w_right, none.generic) == (-1, -1, None)
    weird = loop.clone()
    weird[0, T - 1] = False
    g = classify_mask(weird, T)
    assert g.generic is not None and g.generic.dtype == torch.uint8
    full = classify_mask(torch.zeros(T, T, dtype=torch.bool), T)
    assert full.generic is None and full.w_left == T - 1 and full.w_right == T - 1
    with pytest.raises(ValueError):
        classify_mask(torch.zeros(T, T), T)


def test_synthetic_inputs_shape_and_layout():
    from osudit import synth
    z, o, c, y = synth.sampling_batch(3, 64, seed=1)
    assert z.shape == (6, 2, 64) and o.shape == (6, 64) and c.shape == (6, 144, 64) and y.shape == (6,)
    assert torch.equal(z[:3], z[3:]) and torch.equal(c[:3], c[3:]) and bool((y[3:] == 52670).all())
    assert bool((o[:, 0] == 0).all()) and bool((o[:, 1:] > o[:, :-1]).all())
    assert bool((c[:, 128:].sum(1) == 1).all())  # one-hot datapoint type
    (x, o2, c2), y2 = synth.training_batch(4, 128, seed=0)
    assert x.shape == (4, 2, 128) and float(x.min()) >= 0 and float(x.max()) <= 1 and y2.shape == (4,)


def _spacing_digest(fn, n, spec):
    import zlib
    try:
        kept = sorted(fn(n, spec))
    except Exception as e:
        return type(e).__name__
    return "%d:%08x" % (len(kept), zlib.crc32(",".join(map(str, kept)).encode()))


def test_space_timesteps_sweep_matches_reference(golden_dir):
    """Every single count and every "ddimN" over 1000 / 250 / 37 steps plus multi-section lists (2 613 specs), kept
    steps bit-exact and the same exception type where the reference raises (tests/golden/make_golden_spacing.py)."""
    from diffusion import space_timesteps
    from oracle.diffusion import spaced_steps
    want = {k: v for k, v in json.load(open(os.path.join(golden_dir, "spacing.json"))).items()
            if not k.startswith("betas|")}
    assert len(want) == 2613
    bad = []
    for key, digest in want.items():
        n, spec = key.split("|")
        if _spacing_digest(space_timesteps, int(n), spec) != digest:
            bad.append(("product", key))
        if not spec.startswith("ddim") and _spacing_digest(spaced_steps, int(n), spec) != digest:
            bad.append(("oracle", key))
    assert not bad, bad[:10]


def test_named_beta_schedules_bit_exact(golden_dir):
    """`get_named_beta_schedule` for both schedules at eight lengths: float64 bytes identical to the reference's, and
    the same exception for an unknown name (gaussian_diffusion.py:112-134); the oracle's two builders likewise."""
    import zlib
    from diffusion import get_named_beta_schedule
    from oracle import diffusion as odiff
    want = {k: v for k, v in json.load(open(os.path.join(golden_dir, "spacing.json"))).items() if k.startswith("betas|")}
    assert len(want) == 24

    def digest(fn, *a):
        try:
            b = np.asarray(fn(*a))
        except Exception as e:
            return type(e).__name__
        return "%d:%08x" % (len(b), zlib.crc32(b.astype("<f8").tobytes()))

    for key, d in want.items():
        _, name, n = key.split("|")
        assert digest(get_named_beta_schedule, name, int(n)) == d, key
        if name == "linear":
            assert digest(odiff.linear_betas, int(n)) == d, ("oracle", key)
        elif name == "squaredcos_cap_v2":
            assert digest(odiff.cosine_betas, int(n)) == d, ("oracle", key)


def test_positional_embedding_module_matches_reference(golden_dir):
    """The host-side drop-in `positional_embedding` (data preparation callers, data_loading.py:161) against outputs
    of the reference's module on the same inputs: even / odd / tiny dims, a non-default max_period, sequence forms."""
    import positional_embedding as pe
    z = np.load(os.path.join(golden_dir, "posemb.npz"))
    t, o, p = (torch.from_numpy(z[k]) for k in ("t", "o", "p"))
    for dim in (256, 128, 9, 2):
        np.testing.assert_array_equal(pe.timestep_embedding(t, dim).numpy(), z[f"timestep_{dim}"])
    np.testing.assert_array_equal(pe.timestep_embedding(t, 128, max_period=100).numpy(), z["timestep_128_mp100"])
    np.testing.assert_array_equal(pe.offset_sequence_embedding(o / 10, 128).numpy(), z["offset_128"])
    np.testing.assert_array_equal(pe.position_sequence_embedding(p, 128).numpy(), z["position_128"])


def test_label_dropout_matches_reference(golden_dir):
    """models.py:56-67: the null class replaces a label where `rand(B) < p` on the global generator (same draws as the
    reference under the same seed) or where `force_drop_ids == 1`; the table has one extra row only when p > 0."""
    import models
    g = json.load(open(os.path.join(golden_dir, "labels.json")))
    labels, force = torch.tensor(g["labels"]), torch.tensor(g["force"])
    for case in g["cases"]:
        emb = models._LabelEmbedder(52670, 8, case["p"])
        assert emb.embedding_table.weight.shape[0] == case["table_rows"]
        torch.manual_seed(5)
        assert emb.token_drop(labels).tolist() == case["first"]
        assert emb.token_drop(labels).tolist() == case["second"]
        assert emb.token_drop(labels, force_drop_ids=force).tolist() == case["forced"]
    assert models._LabelEmbedder(52670, 8, 0.0).embedding_table.weight.shape[0] == g["table_rows_without_dropout"]


def test_seeded_construction_gives_the_reference_weights(golden_dir):
    """`torch.manual_seed(s); DiT_models[name](**kw)` consumes the global generator in the reference's order
    (models.py:243-304 and the constructors it calls): every initial tensor is bit-identical to the reference's,
    and so is the generator state afterwards (tests/golden/make_golden_init.py)."""
    import zlib
    import models
    for case in json.load(open(os.path.join(golden_dir, "init_digest.json"))):
        torch.manual_seed(case["seed"])
        m = models.DiT_models[case["name"]](**case["kwargs"])
        assert torch.rand(4).tolist() == case["next_rand"]
        sd = m.state_dict()
        assert list(sd) == list(case["crc"])
        bad = [k for k, v in sd.items() if "%08x" % zlib.crc32(v.contiguous().numpy().tobytes()) != case["crc"][k]]
        assert not bad, (case["name"], bad[:6])


@pytest.mark.parametrize("force_port", [False, True])
def test_bench_reference_arm_prints_the_contract_line(force_port):
    """`bench.py --impl reference` (the CPU arm the driver times next to the GPU arm) needs no GPU and prints one
    JSON line with the GPU arm's metric / unit / config keys plus `impl`, `cpu_baseline` and a zero-copy `e2e`.  It runs
    the unmodified reference modules when they are installed (baseline/_ref, /root/reference), else the oracle port."""
    import subprocess
    import sys
    env = dict(os.environ, OSUDIT_BENCH_CPU_ARM="port") if force_port else dict(os.environ)
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1",
                        "--warmup", "0"], capture_output=True, text=True, timeout=600, cwd=ROOT, env=env)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [ln for ln in r.stdout.splitlines() if ln.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"].startswith("beatmaps/sec") and d["unit"] == "beatmaps/s"
    assert d["higher_is_better"] is True and d["n_gpus"] == 1 and d["steps"] == 1 and d["value"] > 0
    assert d["e2e"] == {"value": d["value"], "unit": d["unit"], "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    have_ref = any(os.path.exists(os.path.join(p, "models.py")) for p in
                   (os.path.join(ROOT, "baseline", "_ref"), "/root/reference"))
    want = "port" if (force_port or not have_ref) else "reference"
    assert d["cpu_baseline"]["kind"] == want and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["sample"]
    assert d["train"]["unit"] == "seq/s" and d["train"]["value"] > 0 and d["train"]["kind"] == want
    assert d["config"]["workload"].startswith("DiT-B sampling")


def test_bench_reference_arm_other_ranks_exit_quietly():
    """Under torchrun (N > 1) only rank 0 runs the CPU arm; every other rank exits 0 without output or work."""
    import subprocess
    import sys
    env = dict(os.environ, RANK="1", LOCAL_RANK="1", WORLD_SIZE="2", MASTER_ADDR="127.0.0.1", MASTER_PORT="29533")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2",
                        "--steps", "1", "--warmup", "0"], capture_output=True, text=True, timeout=300, cwd=ROOT, env=env)
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_missing_library_fails_loudly():
    """Without libosudit.so the product path raises on first use (no CPU / PyTorch fallback): checked in a fresh
    interpreter with OSUDIT_LIB pointing at a file that does not exist."""
    import subprocess
    import sys
    code = ("import sys; sys.path.insert(0, %r)\n"
            "from osudit import _lib\n"
            "try:\n    _lib.load()\nexcept _lib.OsuditError as e:\n    print('RAISED', 'no CPU' in str(e))\n"
            "else:\n    print('LOADED')\n") % os.path.join(ROOT, "osu-diffusion_b200")
    env = dict(os.environ, OSUDIT_LIB="/nonexistent/libosudit.so")
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300, env=env)
    assert r.stdout.strip() == "RAISED True", (r.stdout, r.stderr[-500:])
