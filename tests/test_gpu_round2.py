"""Round-2 regression and integration tests on the GPU:

* the cache-coherence / lifetime / validation defects the round-1 review found (fused optimizer vs the packed-weight
  caches, captured graphs vs evicted workspaces, unchecked shapes and labels, recycled mask addresses);
* data-parallel training: gradients after the all-reduce equal the single-process gradients of the concatenated batch
  (the reference's only collective, train.py:152,257);
* the reference's UNMODIFIED scripts (train.py, sample.py from baseline/_ref/, placed there by
  `__graft_entry__.build()`) executed against the drop-in modules with stubbed third-party imports.
"""
import copy
import math
import os
import socket
import sys
import types

import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import dit as odit  # noqa: E402
from osudit import synth  # noqa: E402

DEV = "cuda"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def _model(name="DiT-S", seed=1, dropout=0.2, std=0.05):
    import models
    shape = odit.shape_of(name)
    sd = odit.init_state_dict(shape, seed=seed, zero_init_std=std)
    m = models.DiT_models[name](num_classes=52670, context_size=144, class_dropout_prob=dropout)
    m.load_state_dict(sd, strict=True)
    return m.to(DEV)


# ------------------------------------------------------------------------------ review item: fused optimizer
def test_fused_optimizer_keeps_every_weight_cache_coherent(monkeypatch):
    """FusedAdamWEMA writes parameters and EMA tensors through raw pointers.  Eager training (no CUDA graph: the
    bf16 / transposed weight copies are cached by parameter version), sampling with the trained model and with the
    EMA copy must all see the updated weights: everything is compared with a twin driven by torch.optim.AdamW +
    the reference's update_ema loop (train.py:36-45)."""
    from diffusion import create_diffusion
    from osudit import graphs, train as otrain
    from osudit.optim import FusedAdamWEMA
    monkeypatch.setattr(otrain, "_GRAPHS_ENABLED", False)
    monkeypatch.setattr(graphs, "_ENABLED", False)
    B, T = 4, 128
    m_f = _model().eval()
    m_t = copy.deepcopy(m_f)
    ema_f, ema_t = copy.deepcopy(m_f).requires_grad_(False), copy.deepcopy(m_t).requires_grad_(False)
    opt_f = FusedAdamWEMA(m_f.parameters(), lr=3e-3, weight_decay=0)
    opt_f.attach_ema(ema_f, m_f, decay=0.5)
    opt_t = torch.optim.AdamW(m_t.parameters(), lr=3e-3, weight_decay=0)
    d = create_diffusion("", noise_schedule="squaredcos_cap_v2", use_l1=True)
    (x, o, c), y = synth.training_batch(B, T, seed=3)
    x, o, c, y = [v.to(DEV) for v in (x, o, c, y)]
    # one eval forward of every model first, so that each holds a packed-weight cache that the steps must invalidate
    ds = create_diffusion("100", noise_schedule="squaredcos_cap_v2")
    z, so, sc, sy = [v.to(DEV) for v in synth.sampling_batch(1, 128, seed=5)]
    tt = torch.full((2,), 40, device=DEV)

    def eps(model):
        with torch.no_grad():
            return model.forward_with_cfg(z, ds._tables(z.device)["tmap"][tt], o=so, c=sc, y=sy, cfg_scale=1.5).clone()

    before = eps(m_f)
    eps(ema_f)
    losses = []
    for step in range(4):
        gen = torch.Generator().manual_seed(step)
        t = torch.randint(0, 1000, (B,), generator=gen).to(DEV)
        nz = torch.randn(B, 2, T, generator=gen).to(DEV)
        pair = []
        for m, opt in ((m_f, opt_f), (m_t, opt_t)):
            loss = d.training_losses(m, x, t, dict(o=o, c=c, y=y), noise=nz)["loss"].mean()
            loss.backward()
            opt.step()
            opt.zero_grad(set_to_none=True)
            pair.append(float(loss.detach()))
        with torch.no_grad():
            for pe, pm in zip(ema_t.parameters(), m_t.parameters()):
                pe.mul_(0.5).add_(pm.detach(), alpha=0.5)
        losses.append(pair)
        # identical losses step after step <=> the fused side's forward used the weights it had just written
        assert abs(pair[0] - pair[1]) < 2e-3 * abs(pair[1]), (step, pair)
    assert all(p._version > 0 for p in m_f.parameters() if p.requires_grad)
    print(f"fused vs torch AdamW, 4 eager steps, losses {losses}")
    # (the parameters themselves are not compared element-wise: Adam turns the rounding noise of near-zero gradients
    # into +-lr steps, so two correct runs differ by O(lr) on those elements; tests/test_gpu_optim.py pins the update
    # rule on identical gradients.)  Sampling after the steps: the packed bf16 copies held by the inference engine of
    # the trained model and of the EMA copy must have been rebuilt — bit-identical to a freshly built engine.
    after, after_ema = eps(m_f), eps(ema_f)
    assert rel(after, before) > 1e-2              # the weights really moved
    m_f._engine = ema_f._engine = None            # drop every cache: what a correct re-pack must reproduce
    assert torch.equal(after, eps(m_f))
    assert torch.equal(after_ema, eps(ema_f))
    assert rel(after_ema, after) > 1e-4           # EMA (decay 0.5) is a different set of weights


def test_fused_optimizer_checkpoint_loads_into_torch_adamw():
    """train.py:287-293 saves opt.state_dict(); the fused optimizer's must load into the reference's AdamW with one
    independent `step` tensor per parameter (a shared one would be advanced once per parameter by the foreach path)."""
    from osudit.optim import FusedAdamWEMA
    net = torch.nn.Sequential(torch.nn.Linear(16, 32), torch.nn.Linear(32, 8)).to(DEV)
    twin = copy.deepcopy(net)
    fused = FusedAdamWEMA(net.parameters(), lr=1e-2, weight_decay=0)
    x = torch.randn(4, 16, device=DEV)
    for _ in range(3):
        net(x).square().mean().backward()
        fused.step()
        fused.zero_grad(set_to_none=True)
    import io
    buf = io.BytesIO()
    torch.save(fused.state_dict(), buf)  # what train.py:287-293 does; a checkpoint shares no storage with the live state
    buf.seek(0)
    sd = torch.load(buf, map_location=DEV, weights_only=True)
    steps = [st["step"] for st in sd["state"].values()]
    assert len({s.data_ptr() for s in steps}) == len(steps) and all(float(s) == 3.0 for s in steps)
    twin.load_state_dict(net.state_dict())
    ref = torch.optim.AdamW(twin.parameters(), lr=1e-2, weight_decay=0)
    ref.load_state_dict(sd)
    for opt, mod in ((fused, net), (ref, twin)):
        mod(x).square().mean().backward()
        opt.step()
    got_steps = [float(st["step"]) for st in ref.state.values()]
    assert got_steps == [4.0] * len(got_steps), got_steps
    assert max(rel(a, b) for a, b in zip(net.parameters(), twin.parameters())) < 1e-5


# ------------------------------------------------------------------------------ review item: graph lifetimes
def test_captured_step_survives_workspace_and_mask_cache_eviction():
    """A cached StepGraph replays kernels that point into the engine's workspace, packed weights and mask copy; the
    engine evicts workspaces after four shapes.  The graph must keep what it captured alive: replay after evictions
    (and an empty_cache) equals the eager step bit for bit."""
    from diffusion import create_diffusion
    from osudit import engine as eng, graphs
    m = _model(dropout=0.1).eval()
    d = create_diffusion("100", noise_schedule="squaredcos_cap_v2")
    T = 160
    z, o, c, y = [v.to(DEV) for v in synth.sampling_batch(1, T, seed=2)]
    mask = torch.rand(T, T, device=DEV) < 0.3  # generic mask: the uint8 copy lives in the mask cache only
    mask.fill_diagonal_(False)
    kw = dict(o=o, c=c, y=y, cfg_scale=1.5, attn_mask=mask)
    t = torch.full((2,), 30, device=DEV)
    graphs.release_graphs()
    torch.manual_seed(0)
    first = d.p_sample(m.forward_with_cfg, z, t, model_kwargs=kw)["pred_xstart"].clone()
    assert len(graphs._cache) == 1
    for TT in (64, 96, 130, 200, 72):  # five more shapes: the engine's four-entry workspace cache turns over
        zz, oo, cc, yy = [v.to(DEV) for v in synth.sampling_batch(1, TT, seed=3)]
        with torch.no_grad():
            m.forward(zz, torch.full((2,), 5, device=DEV), o=oo, c=cc, y=yy)
    for i in range(eng._MASK_CACHE_ENTRIES + 2):  # and the mask-classification cache
        eng.classify_mask(torch.rand(T, T, device=DEV) < 0.5, T)
    torch.cuda.empty_cache()
    junk = [torch.full((1 << 22,), float("nan"), device=DEV) for _ in range(8)]  # recycle whatever was freed
    torch.manual_seed(0)
    again = d.p_sample(m.forward_with_cfg, z, t, model_kwargs=kw)["pred_xstart"]
    assert len(graphs._cache) == 1  # it WAS a replay of the first capture
    assert torch.equal(first, again)
    del junk
    graphs.release_graphs()


def test_mask_classification_is_not_fooled_by_a_recycled_address():
    """Sweep over band widths with short-lived mask tensors of the same shape (the allocator hands the same address
    back): each forward must apply ITS band."""
    from osudit import ops
    B, T, H, hd = 1, 256, 2, 64
    qkv = torch.randn(B * T, 3 * H * hd, device=DEV).to(torch.bfloat16)
    from osudit.engine import classify_mask
    seen = set()
    for W in (8, 16, 32, 64, 100):
        mask = synth.band_mask(T, W).to(DEV)
        seen.add(mask.data_ptr())
        spec = classify_mask(mask, T)
        assert (spec.w_left, spec.w_right) == (W - 1, W), (W, spec.w_left, spec.w_right)
        out = ops.attn_band(qkv, torch.empty(B * T, H * hd, device=DEV, dtype=torch.bfloat16), B, T, H, hd,
                            spec.w_left, spec.w_right, spec.generic)
        q, k, v = [a.float().reshape(T, H, hd).transpose(0, 1) for a in qkv.float().split(H * hd, dim=1)]
        s = (q @ k.transpose(1, 2)) / math.sqrt(hd)
        s = s.masked_fill(mask[None], float("-inf"))
        ref = (torch.softmax(s, -1) @ v).transpose(0, 1).reshape(T, H * hd)
        assert rel(out.float(), ref) < 1e-2
        del mask, spec


# ------------------------------------------------------------------------------ review item: input validation
def test_shape_and_label_validation_matches_the_reference_error_behaviour():
    import models
    from diffusion import create_diffusion
    with pytest.raises(ValueError):  # 384 + 142 = 526 is not a multiple of 8 (the reference's constructor default)
        models.DiT(depth=1, hidden_size=128, num_heads=2, num_classes=10)
    m = _model(dropout=0.1).eval()
    T = 64
    z, o, c, y = [v.to(DEV) for v in synth.sampling_batch(1, T, seed=1)]
    t = torch.full((2,), 3, device=DEV)
    with torch.no_grad():
        m(z, t, o=o, c=c, y=y)  # fine
        with pytest.raises(ValueError):
            m(z, t, o=o, c=torch.cat([c, c[:, :8]], 1), y=y)     # more context channels than context_size
        with pytest.raises(ValueError):
            m(z[:, :, :32], t, o=o, c=c, y=y)                    # x / o length mismatch
        with pytest.raises(ValueError):
            m(z, t[:1], o=o, c=c, y=y)                           # t batch mismatch
        with pytest.raises(ValueError):
            m(z, t, o=o[:1], c=c, y=y)                           # o batch mismatch
        with pytest.raises(IndexError):
            m(z, t, o=o, c=c, y=torch.tensor([3, 52671], device=DEV))  # label beyond the table (52670 = null class)
        with pytest.raises(IndexError):
            m(z, t, o=o, c=c, y=torch.tensor([-1, 5], device=DEV))
    d = create_diffusion("", noise_schedule="squaredcos_cap_v2", use_l1=True)
    with pytest.raises(ValueError):
        d.training_losses(m, z, torch.zeros(2, dtype=torch.long, device=DEV), dict(o=o, c=c[:, :100], y=y))


# ------------------------------------------------------------------------------ step-invariant work hoisted
@torch.no_grad()
def test_sampling_loop_hoists_step_invariant_work_bit_exactly(monkeypatch):
    """Within one sampling loop only x changes: the loop computes the conditioning of all steps up front in one batched
    pass and rewrites only the x columns of the first-layer operand after the first step.  Both must
    be invisible: the loop equals step-by-step p_sample calls (no precomputed conditioning) bit for bit, also when two
    loops with different o / c / y alternate on the same workspace."""
    import diffusion.gaussian_diffusion as gd
    from diffusion import create_diffusion
    from osudit import graphs
    monkeypatch.setattr(graphs, "_ENABLED", False)  # the GPU-bound path (what BASELINE config 2 runs), at a test size
    m = _model(dropout=0.1).eval()
    d = create_diffusion("8", noise_schedule="squaredcos_cap_v2")
    T = 200
    mask = synth.band_mask(T, 128).to(DEV)
    sets = []
    for seed in (1, 2):
        z, o, c, y = [v.to(DEV) for v in synth.sampling_batch(2, T, seed=seed)]
        sets.append((z, dict(o=o, c=c, y=y, cfg_scale=1.5, attn_mask=mask)))
    g = torch.Generator().manual_seed(0)
    noises = [torch.randn(4, 2, T, generator=g).to(DEV) for _ in range(8)]

    def with_noise(fn):
        it = iter(noises)
        real = gd.th.randn_like
        gd.th.randn_like = lambda v: next(it)
        try:
            return fn()
        finally:
            gd.th.randn_like = real

    def stepwise(z, kw):  # plain p_sample calls: conditioning computed inside every forward, fresh engine each time
        m._engine = None
        x = z
        for i in reversed(range(8)):
            m.engine().reuse_oc_columns = False
            x = d.p_sample(m.forward_with_cfg, x, torch.full((4,), i, device=DEV), model_kwargs=kw)["sample"]
        m._engine = None
        return x

    want = [with_noise(lambda: stepwise(z, kw)) for z, kw in sets]
    calls = {"x": 0, "xoc": 0}
    from osudit import ops
    real_x, real_xoc = ops.embed_x, ops.embed_xoc
    monkeypatch.setattr(ops, "embed_x", lambda *a, **k: (calls.__setitem__("x", calls["x"] + 1), real_x(*a, **k))[1])
    monkeypatch.setattr(ops, "embed_xoc", lambda *a, **k: (calls.__setitem__("xoc", calls["xoc"] + 1), real_xoc(*a, **k))[1])
    got = [with_noise(lambda: d.p_sample_loop(m.forward_with_cfg, z.shape, z, model_kwargs=kw, device=DEV))
           for z, kw in sets]
    assert calls == {"x": 14, "xoc": 2}  # one full operand per loop, then x columns only
    for a, b in zip(got, want):
        assert torch.equal(a, b)
    # alternating the two conditioning sets step by step on the same workspace: every switch rewrites everything
    calls.update(x=0, xoc=0)
    for i in (7, 6):
        for z, kw in sets:
            d.p_sample(m.forward_with_cfg, z, torch.full((4,), i, device=DEV), model_kwargs=kw)
    assert calls == {"x": 0, "xoc": 4}


# ------------------------------------------------------------------------------ data-parallel gradients
def _ddp_worker(rank, world, port, backend, q, use_wrap):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch.distributed as dist
    from torch.nn.parallel import DistributedDataParallel as DDP
    dev_index = rank if backend == "nccl" else 0
    torch.cuda.set_device(dev_index)
    dist.init_process_group(backend, rank=rank, world_size=world)
    try:
        from diffusion import create_diffusion
        from osudit import ddp
        m = _model().eval()  # eval: no label dropout, every rank sees its own rows of one global batch
        B, T = 4, 128
        (x, o, c), y = synth.training_batch(world * B, T, seed=3)
        gen = torch.Generator().manual_seed(5)
        t = torch.randint(0, 1000, (world * B,), generator=gen)
        nz = torch.randn(world * B, 2, T, generator=gen)
        sl = slice(rank * B, (rank + 1) * B)
        d = create_diffusion("", noise_schedule="squaredcos_cap_v2", use_l1=True)
        if use_wrap:
            net = ddp.wrap(m, device_ids=[dev_index])
        else:
            net = DDP(m, device_ids=[dev_index])  # train.py:152
        loss = d.training_losses(net, x[sl].to(DEV), t[sl].to(DEV), dict(o=o[sl].to(DEV), c=c[sl].to(DEV), y=y[sl].to(DEV)),
                                 noise=nz[sl].to(DEV))["loss"].mean()
        loss.backward()  # DDP averages over ranks: mean of per-rank means = mean over the global batch
        torch.cuda.synchronize()
        if rank == 0:  # numpy arrays travel by value (tensors would be shared through fds of a process about to exit)
            q.put({k: p.grad.detach().cpu().numpy() for k, p in m.named_parameters() if p.grad is not None})
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("use_wrap", [False, True])
def test_ddp_gradients_equal_single_process_gradients_of_the_global_batch(use_wrap):
    """SURVEY §4 (vi) / train.py:152: after the all-reduce every rank holds the gradient of the mean loss over the
    GLOBAL batch.  Two ranks (NCCL on two GPUs when the box has them, otherwise gloo with both ranks on one GPU) vs one
    process on the concatenated batch; `use_wrap` adds osudit.ddp.wrap's settings (128 MB bucket views, SM cap)."""
    import torch.multiprocessing as mp
    from diffusion import create_diffusion
    world = 2
    backend = "nccl" if torch.cuda.device_count() >= world else "gloo"
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_ddp_worker, args=(r, world, port, backend, q, use_wrap)) for r in range(world)]
    for p in procs:
        p.start()
    got = {k: torch.from_numpy(v) for k, v in q.get(timeout=600).items()}
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    m = _model().eval()
    B, T = 4, 128
    (x, o, c), y = synth.training_batch(world * B, T, seed=3)
    gen = torch.Generator().manual_seed(5)
    t = torch.randint(0, 1000, (world * B,), generator=gen)
    nz = torch.randn(world * B, 2, T, generator=gen)
    d = create_diffusion("", noise_schedule="squaredcos_cap_v2", use_l1=True)
    d.training_losses(m, x.to(DEV), t.to(DEV), dict(o=o.to(DEV), c=c.to(DEV), y=y.to(DEV)),
                      noise=nz.to(DEV))["loss"].mean().backward()
    worst = max((rel(got[k], p.grad), k) for k, p in m.named_parameters() if p.grad is not None)
    print(f"DDP ({backend}, wrap={use_wrap}) vs single process on the global batch: worst gradient rel-L2 {worst[0]:.2e} ({worst[1]})")
    assert set(got) == {k for k, p in m.named_parameters() if p.grad is not None}
    # fp32 buckets: only the batch split changes the bf16 roundings inside the kernels
    assert worst[0] < 1e-2, worst


# ------------------------------------------------------------------------------ the unmodified scripts
def _script(name):
    for root in (os.environ.get("OSU_DIFFUSION_REF"), os.path.join(ROOT, "baseline", "_ref")):
        if root and os.path.exists(os.path.join(root, name)):
            return os.path.join(root, name)
    pytest.skip(f"{name} of the reference is not installed under baseline/_ref (run __graft_entry__.build() where "
                "/root/reference exists)")


class _Stubs:
    """Third-party / data-pipeline modules the scripts import but that are outside the hot path (SURVEY §8 out of
    scope: `slider`, matplotlib, the .osu data loader and exporter), replaced for the duration of one test.  The
    scripts also flip process-wide torch switches at import (TF32 on: train.py:7-8, sample.py:25-26; sample.py:42
    turns autograd off): restored on exit so that later tests' fp32 torch references stay fp32."""

    def __init__(self, mods):
        self.mods, self.saved = mods, {}

    def __enter__(self):
        self.flags = (torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32, torch.is_grad_enabled())
        for k, v in self.mods.items():
            self.saved[k] = sys.modules.get(k)
            sys.modules[k] = v
        return self

    def __exit__(self, *exc):
        torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = self.flags[:2]
        torch.set_grad_enabled(self.flags[2])
        for k, v in self.saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v


def _fake_data_loading(batches, seq):
    dl = types.ModuleType("data_loading")
    dl.feature_size = 19
    dl.get_data_loader = lambda **kw: iter(batches)
    dl.window_and_relative_time = dl.load_and_process_beatmap = lambda *a, **k: None
    dl.BeatmapDatasetIterableFactory = lambda *a, **k: None
    dl.beatmap_to_sequence = lambda beatmap: seq
    dl.get_beatmap_idx = lambda path: {1234: 7}

    def split_and_process_sequence(s):  # data_loading.py:154-169 without the random flip, on the device-side builder
        from osudit import data
        (x, o, c), n = data.split_and_process_sequence_no_augment(s.to(DEV))
        return (x.cpu(), o.cpu(), c.cpu()), n
    dl.split_and_process_sequence = split_and_process_sequence
    return dl


def test_unmodified_train_script_runs_and_resumes(tmp_path, monkeypatch, caplog):
    """`train.py` of the reference, byte for byte, through runpy: DDP wrap (one NCCL rank), fp16-autocast context,
    GradScaler, AdamW, EMA, logging all-reduce, checkpoint save at step 3 — then a second launch that restores the
    checkpoint with `relearn_embeds` (drops optimizer state 7 = the embedding table) and keeps training."""
    import argparse
    import logging
    import runpy
    import torch.distributed as dist
    path = _script("train.py")
    caplog.set_level(logging.INFO)  # the script logs through logging.getLogger(__name__); pytest owns the root handlers
    B, T = 8, 128
    batches = [synth.training_batch(B, T, seed=40 + i) for i in range(3)]
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    for k, v in dict(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK="0", WORLD_SIZE="1", LOCAL_RANK="0").items():
        monkeypatch.setenv(k, v)
    args = argparse.Namespace(
        data_path="unused", num_classes=52670, data_end=8, data_start=0, results_dir=str(tmp_path / "results"),
        model="DiT-S", epochs=1, global_batch_size=B, global_seed=0, num_workers=0, log_every=1, ckpt_every=3,
        seq_len=T, stride=16, use_amp=True, ckpt=None, dist="nccl", fine_tune_ids=None,
        noise_schedule="squaredcos_cap_v2", l1_loss=True, lr=1e-4, relearn_embeds=False, embed_only_epochs=0)
    with _Stubs({"data_loading": _fake_data_loading(batches, None)}):
        g = runpy.run_path(path, run_name="reference_train")
        assert g["DiT_models"]["DiT-S"].__module__ == "models" and "osu-diffusion_b200" in sys.modules["models"].__file__
        g["main"](args)
        assert not dist.is_initialized()  # cleanup() ran: the script reached its end
        ckpts = sorted((tmp_path / "results").glob("*/checkpoints/*.pt"))
        assert [p.name for p in ckpts] == ["0000003.pt"]
        ck = torch.load(ckpts[0], map_location="cpu", weights_only=False)
        assert set(ck) == {"model", "ema", "opt", "scaler", "args"}
        assert len(ck["model"]) == 132 and len(ck["opt"]["state"]) == 131  # playfield_size has no optimizer state
        assert ck["opt"]["state"][7]["exp_avg"].shape == (52671, 384)      # index 7 = y_embedder.embedding_table
        moved = rel(ck["ema"]["blocks.0.mlp.fc1.weight"], ck["model"]["blocks.0.mlp.fc1.weight"])
        assert 0 < moved < 5e-2  # EMA (decay 0.9999) trails three AdamW steps
        losses = [float(r.getMessage().split("Train Loss: ")[1].split(",")[0]) for r in caplog.records
                  if "Train Loss" in r.getMessage()]
        assert len(losses) == 3 and all(math.isfinite(v) and 0.1 < v < 10 for v in losses), losses
        # resume, re-learning the embeddings
        with socket.socket() as s:
            s.bind(("127.0.0.1", 0))
            monkeypatch.setenv("MASTER_PORT", str(s.getsockname()[1]))
        args2 = copy.copy(args)
        args2.ckpt, args2.relearn_embeds, args2.ckpt_every = str(ckpts[0]), True, 1000
        caplog.clear()
        g = runpy.run_path(path, run_name="reference_train")
        # train.py:207 calls torch.load without `weights_only`; torch >= 2.6 then refuses the argparse.Namespace the
        # script itself stored under "args" (the reference has the same problem with this torch): allow-list it
        with torch.serialization.safe_globals([argparse.Namespace]):
            g["main"](args2)
        msgs = [r.getMessage() for r in caplog.records]
        assert any("Restored from checkpoint" in v for v in msgs)
        resumed = [float(v.split("Train Loss: ")[1].split(",")[0]) for v in msgs if "Train Loss" in v]
        assert len(resumed) == 3 and all(math.isfinite(v) for v in resumed), resumed


def test_unmodified_sample_script_runs_with_refinement(tmp_path, monkeypatch):
    """`sample.py` of the reference, byte for byte: loads a train.py-style checkpoint ({"ema": ...}) into
    DiT_models[...], builds the band mask loop of sample.py:81-84, samples 2 variants with CFG through
    p_sample_loop(progress=True), then the refine loop with a second checkpoint (sample.py:151-172)."""
    import argparse
    import runpy
    path = _script("sample.py")
    T = 300
    g0 = torch.Generator().manual_seed(3)
    seq = torch.zeros(19, T)
    seq[0], seq[1] = torch.rand(T, generator=g0) * 512, torch.rand(T, generator=g0) * 384
    seq[2] = torch.cumsum(torch.randint(50, 400, (T,), generator=g0).float(), 0)
    seq[3 + torch.randint(0, 16, (T,), generator=g0), torch.arange(T)] = 1
    written = []

    class Beatmap:
        beatmap_id, artist, title = 4242, "synthetic", "band: mask?"

        @classmethod
        def from_path(cls, p):
            return cls()

    def create_beatmap(s, beatmap, name):
        assert s.shape == (19, T) and torch.isfinite(s).all()
        obj = types.SimpleNamespace(write_path=lambda p: (written.append((p, s.clone())), open(p, "w").write(name)))
        return obj

    slider = types.ModuleType("slider")
    slider.Beatmap = Beatmap
    export = types.ModuleType("export")
    cb = types.ModuleType("export.create_beatmap")
    cb.create_beatmap, cb.plot_beatmap = create_beatmap, lambda *a, **k: []
    export.create_beatmap = cb
    mpl = types.ModuleType("matplotlib")
    plt = types.ModuleType("matplotlib.pyplot")
    mpl.pyplot, mpl.animation = plt, types.ModuleType("matplotlib.animation")
    m = _model(dropout=0.1).eval()
    torch.save({"ema": m.state_dict()}, tmp_path / "main.pt")
    with torch.no_grad():
        m.final_layer.linear.weight.mul_(0.9)
    torch.save(m.state_dict(), tmp_path / "refine.pt")
    monkeypatch.chdir(tmp_path)
    args = argparse.Namespace(beatmap="x.osu", ckpt=str(tmp_path / "main.pt"), model="DiT-S", num_classes=52670,
                              beatmap_idx="unused", cfg_scale=1.5, num_sampling_steps=10, seed=0, seq_len=128,
                              use_amp=True, style_id=1234, plot_time=None, plot_width=2000, num_variants=2,
                              make_animation=False, refine_ckpt=str(tmp_path / "refine.pt"), refine_iters=3)
    try:
        with _Stubs({"data_loading": _fake_data_loading(None, seq), "slider": slider, "export": export,
                     "export.create_beatmap": cb, "matplotlib": mpl, "matplotlib.pyplot": plt,
                     "matplotlib.animation": mpl.animation}):
            g = runpy.run_path(path, run_name="reference_sample")
            assert "osu-diffusion_b200" in sys.modules["models"].__file__
            g["main"](args)
    finally:
        torch.set_grad_enabled(True)  # sample.py:42 switches autograd off process-wide
    assert len(written) == 4  # 2 variants after sampling + 2 after refinement
    first, refined = written[0][1], written[2][1]
    assert torch.equal(first[2:], seq[2:])                       # time / type rows pass through (sample.py:111-113)
    assert float(first[:2].min()) >= -1.0 and float(first[:2].max()) <= 2.0   # clip_denoised range of x0
    assert not torch.equal(first[:2], refined[:2])               # the refine checkpoint changed the result
    assert not torch.equal(written[0][1][:2], written[1][1][:2])  # two style labels, two results
