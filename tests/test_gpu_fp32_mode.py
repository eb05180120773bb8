"""fp32 mode (`model.precision = "fp32"`, osudit/fp32.py + csrc/fp32_mode.cu): the north star's second
tolerance — predicted epsilon within 1e-5 relative L2 of the fp32 reference."""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import diffusion as odiff  # noqa: E402
from oracle import dit as odit  # noqa: E402
from osudit import fp32, synth  # noqa: E402

DEV = "cuda"
FP32_TOL = 1e-5


def rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).norm() / b.norm())


def test_split3_is_exact_to_24_bits():
    g = torch.Generator().manual_seed(0)
    v = (torch.randn(37, 72, generator=g) * torch.logspace(-6, 6, 72)[None]).to(DEV)
    out3 = fp32.split3(v, torch.empty(37, 216, dtype=torch.bfloat16, device=DEV))
    s = out3[:, :72].double() + out3[:, 72:144].double() + out3[:, 144:].double()
    assert float(((s - v.double()).abs() / v.double().abs()).max()) < 2.0 ** -23
    # activations: GELU(tanh) and SiLU(+table) against torch fp64
    x = torch.randn(50, 64, generator=g).to(DEV) * 3
    for act, fn in ((1, lambda z: torch.nn.functional.gelu(z, approximate="tanh")), (2, torch.nn.functional.silu)):
        o3 = fp32.split3(x, torch.empty(50, 192, dtype=torch.bfloat16, device=DEV), act=act)
        got = o3[:, :64].double() + o3[:, 64:128].double() + o3[:, 128:].double()
        assert float((got - fn(x.double())).abs().max()) < 1e-6
    table = torch.randn(9, 64, generator=g).to(DEV)
    y = torch.randint(0, 9, (50,), generator=g).to(DEV)
    o3 = fp32.split3(x, torch.empty(50, 192, dtype=torch.bfloat16, device=DEV), act=2, table=table, y=y)
    got = o3[:, :64].double() + o3[:, 64:128].double() + o3[:, 128:].double()
    assert float((got - torch.nn.functional.silu((x + table[y]).double())).abs().max()) < 1e-6


@pytest.mark.parametrize("M,N,K", [(300, 768, 528), (4096, 2304, 768), (2, 768, 256), (1000, 384, 1536)])
def test_gemm_f32_matches_fp64(M, N, K):
    g = torch.Generator().manual_seed(M + N + K)
    a = torch.randn(M, K, generator=g).to(DEV)
    w = (torch.randn(N, K, generator=g) / math.sqrt(K)).to(DEV)
    bias = torch.randn(N, generator=g).to(DEV)
    a3 = fp32.split3(a, torch.empty(M, 3 * K, dtype=torch.bfloat16, device=DEV))
    w6 = fp32.pack_weight6(w)
    ref = a.double() @ w.double().t() + bias.double()
    errs = {kb: rel(fp32.gemm_f32(a3, w6, bias, torch.empty(M, N, device=DEV), kb_per_split=kb), ref)
            for kb in (0, 8, 4, 2, 1)}  # 0 = one tensor-core accumulation chain over the whole K
    e = rel(fp32.gemm_f32(a3, w6, bias, torch.full((M, N), 7.0, device=DEV)), ref)  # default; zeroes `out` itself
    e_torch = rel(a @ w.t() + bias, ref)  # cuBLAS fp32 (no TF32) on the same inputs, for scale
    print(f"gemm_f32 {M}x{N}x{K}: rel-L2 {e:.2e} (torch fp32: {e_torch:.2e}); by chain length "
          + " ".join(f"{kb}:{v:.1e}" for kb, v in errs.items()))
    assert e < 1e-6


@pytest.mark.parametrize("D,has_branch", [(384, True), (768, False), (1152, True)])
def test_ln_modulate_f32(D, has_branch):
    g = torch.Generator().manual_seed(D)
    B, T = 3, 50
    x = torch.randn(B * T, D, generator=g).to(DEV) * 2 + 0.3
    br = torch.randn(B * T, D, generator=g).to(DEV)
    mod = torch.randn(B, 3 * D, generator=g).to(DEV) * 0.5
    x1 = x.clone()
    h3 = fp32.ln_modulate(x1, br if has_branch else None, mod, 0, D, 2 * D, T, torch.empty(B * T, 3 * D, dtype=torch.bfloat16, device=DEV))
    xr = x.double()
    if has_branch:
        xr = xr + mod[:, :D].double().repeat_interleave(T, 0) * br.double()
        assert rel(x1, xr) < 1e-7
    hr = torch.nn.functional.layer_norm(xr, (D,), eps=1e-6) * (1 + mod[:, 2 * D:].double().repeat_interleave(T, 0)) \
        + mod[:, D:2 * D].double().repeat_interleave(T, 0)
    got = h3[:, :D].double() + h3[:, D:2 * D].double() + h3[:, 2 * D:].double()
    assert rel(got, hr) < 3e-7
    # final layer on the same inputs
    w = torch.randn(4, D, generator=g).to(DEV) / math.sqrt(D)
    bias = torch.randn(4, generator=g).to(DEV)
    x2 = x.clone()
    out = fp32.final_layer(x2, br if has_branch else None, mod, 0, D, 2 * D, T, w, bias, torch.empty(B, 4, T, device=DEV))
    ref = (hr @ w.double().t() + bias.double()).reshape(B, T, 4).transpose(1, 2)
    assert rel(out, ref) < 1e-6


@pytest.mark.parametrize("hd,T,wl,wr,generic", [(64, 300, 127, 128, False), (72, 200, 15, 16, False),
                                                (64, 130, -1, -1, False), (64, 96, -1, -1, True)])
def test_attn_band_f32(hd, T, wl, wr, generic):
    g = torch.Generator().manual_seed(hd + T)
    B, H = 2, 3
    D = H * hd
    qkv = torch.randn(B * T, 3 * D, generator=g).to(DEV)
    idx = torch.arange(T)
    d = idx[None, :] - idx[:, None]
    if generic:
        blocked = torch.rand(T, T, generator=g) < 0.5
        blocked[idx, idx] = False
    elif wl >= 0:
        blocked = ~((d >= -wl) & (d <= wr))
    else:
        blocked = torch.zeros(T, T, dtype=torch.bool)
    out = fp32.attn_band(qkv, torch.empty(B * T, D, device=DEV), B, T, H, hd, wl, wr,
                         blocked.to(torch.uint8).to(DEV) if generic else None)
    q, k, v = (z.reshape(B, T, H, hd).transpose(1, 2).double() for z in qkv.split(D, dim=-1))
    s = (q @ k.transpose(-1, -2)) / math.sqrt(hd)
    s = s.masked_fill(blocked.to(DEV), float("-inf"))
    ref = (torch.softmax(s, -1) @ v).transpose(1, 2).reshape(B * T, D)
    assert rel(out, ref) < 5e-7


@pytest.mark.parametrize("name,T,W", [("DiT-S", 256, 128), ("DiT-B", 256, 128), ("DiT-S", 200, None)])
@torch.no_grad()
def test_fp32_mode_forward_within_1e5(name, T, W):
    import models
    shape = odit.shape_of(name)
    sd = odit.init_state_dict(shape, seed=1, zero_init_std=0.02)
    m = models.DiT_models[name](num_classes=52670, context_size=144)
    m.load_state_dict(sd, strict=True)
    m = m.to(DEV).eval()
    m.precision = "fp32"
    z, o, c, y = synth.sampling_batch(1, T, seed=0)
    x = torch.randn(2, 2, T, generator=torch.Generator().manual_seed(2))
    t = torch.tensor([505, 20])
    mask = synth.band_mask(T, W) if W else None
    ref32 = odit.forward(sd, shape.heads, x, t, o, c, y, mask)
    ref64 = odit.forward(sd, shape.heads, x.double(), t, o.double(), c.double(), y, mask, dtype=torch.float64)
    out = m(x.to(DEV), t.to(DEV), o=o.to(DEV), c=c.to(DEV), y=y.to(DEV), attn_mask=mask.to(DEV) if W else None)
    e32, e64 = rel(out[:, :2], ref32[:, :2]), rel(out[:, :2], ref64[:, :2])
    print(f"{name} T={T} fp32 mode: eps rel-L2 vs fp32 oracle {e32:.2e}, vs fp64 oracle {e64:.2e}; "
          f"fp32 oracle vs fp64 {rel(ref32[:, :2], ref64[:, :2]):.2e}")
    assert e32 < FP32_TOL and rel(out, ref32) < FP32_TOL
    # and the default mode on the same module is the bf16 schedule again
    m.precision = "bf16"
    out_bf = m(x.to(DEV), t.to(DEV), o=o.to(DEV), c=c.to(DEV), y=y.to(DEV), attn_mask=mask.to(DEV) if W else None)
    assert 1e-5 < rel(out_bf[:, :2], ref32[:, :2]) < 2e-3


@torch.no_grad()
def test_fp32_mode_cfg_sampling_step(monkeypatch):
    import models
    from diffusion import create_diffusion
    from osudit import graphs
    monkeypatch.setattr(graphs, "_ENABLED", False)
    shape = odit.shape_of("DiT-S")
    sd = odit.init_state_dict(shape, seed=4, zero_init_std=0.02)
    m = models.DiT_models["DiT-S"](num_classes=52670, context_size=144)
    m.load_state_dict(sd, strict=True)
    m = m.to(DEV).eval()
    m.precision = "fp32"
    T = 192
    z, o, c, y = synth.sampling_batch(1, T, seed=3)
    mask = synth.band_mask(T, 64)
    s = odiff.Schedule("100")
    d = create_diffusion("100", noise_schedule="squaredcos_cap_v2")
    t = torch.full((2,), 40)
    noise = torch.randn(2, 2, T, generator=torch.Generator().manual_seed(8))
    ref_out = odit.forward_with_cfg(sd, shape.heads, z, odiff.original_timesteps(s, t), o, c, y, 1.5, mask)
    ref = odiff.p_sample(s, ref_out, z, t, noise)
    import diffusion.gaussian_diffusion as gd
    monkeypatch.setattr(gd.th, "randn_like", lambda x: noise.to(x.device))
    got = d.p_sample(m.forward_with_cfg, z.to(DEV), t.to(DEV), clip_denoised=True,
                     model_kwargs=dict(o=o.to(DEV), c=c.to(DEV), y=y.to(DEV), cfg_scale=1.5, attn_mask=mask.to(DEV)))
    assert rel(got["pred_xstart"], ref["pred_xstart"]) < 5e-5
    assert rel(got["sample"], ref["sample"]) < 5e-5
