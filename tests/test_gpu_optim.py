"""osudit.optim.FusedAdamWEMA against torch.optim.AdamW + the reference's update_ema loop
(train.py:36-45,154,258-261), with and without GradScaler, including a skipped (inf) step."""
import copy

import pytest
import torch

pytestmark = pytest.mark.gpu

from osudit.optim import FusedAdamWEMA  # noqa: E402

DEV = "cuda"


class Net(torch.nn.Module):
    def __init__(self):
        super().__init__()
        g = torch.Generator().manual_seed(0)
        shapes = [(768, 528), (2304,), (1,), (3, 5), (4097,), (52, 768), (9000, 33)]
        self.ps = torch.nn.ParameterList([torch.nn.Parameter(torch.randn(*s, generator=g)) for s in shapes])
        self.frozen = torch.nn.Parameter(torch.randn(7, generator=g), requires_grad=False)


def _grads(net, step, scale=1.0):
    g = torch.Generator().manual_seed(100 + step)
    for p in net.ps:
        p.grad = (torch.randn(p.shape, generator=g) * 0.1).to(DEV) * scale


@torch.no_grad()
def _ema_ref(ema, net, decay):  # train.py:36-45
    for pe, pm in zip(ema.parameters(), net.parameters()):
        pe.mul_(decay).add_(pm.detach(), alpha=1 - decay)


def _close(a, b, tol):
    return float((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-30)) < tol


@pytest.mark.parametrize("wd", [0.0, 0.05])
def test_fused_adamw_ema_matches_torch(wd):
    ref = Net().to(DEV)
    net = copy.deepcopy(ref)
    ema_ref, ema = copy.deepcopy(ref).requires_grad_(False), copy.deepcopy(ref).requires_grad_(False)
    o_ref = torch.optim.AdamW(ref.parameters(), lr=1e-2, betas=(0.9, 0.99), eps=1e-8, weight_decay=wd)
    o = FusedAdamWEMA(net.parameters(), lr=1e-2, betas=(0.9, 0.99), eps=1e-8, weight_decay=wd)
    o.attach_ema(ema, net, decay=0.99)
    for step in range(6):
        _grads(ref, step)
        _grads(net, step)
        o_ref.step()
        _ema_ref(ema_ref, ref, 0.99)
        o.step()
        o_ref.zero_grad(set_to_none=True)
        o.zero_grad(set_to_none=True)
    for a, b in zip(net.parameters(), ref.parameters()):
        assert _close(a, b, 2e-6)
    for a, b in zip(ema.parameters(), ema_ref.parameters()):
        assert _close(a, b, 2e-6)
    assert torch.equal(net.frozen, ref.frozen)
    sd, sd_ref = o.state_dict(), o_ref.state_dict()
    assert sd["state"].keys() == sd_ref["state"].keys()
    for k in sd_ref["state"]:
        assert set(sd["state"][k]) == {"step", "exp_avg", "exp_avg_sq"}
        assert float(sd["state"][k]["step"]) == float(sd_ref["state"][k]["step"]) == 6.0
        assert _close(sd["state"][k]["exp_avg_sq"], sd_ref["state"][k]["exp_avg_sq"].to(DEV), 2e-6)
    # a checkpoint written by torch.optim.AdamW resumes in the fused optimizer
    net2 = copy.deepcopy(ref)
    o2 = FusedAdamWEMA(net2.parameters(), lr=1e-2, betas=(0.9, 0.99), eps=1e-8, weight_decay=wd)
    o2.load_state_dict(copy.deepcopy(sd_ref))
    _grads(ref, 50)
    _grads(net2, 50)
    o_ref.step()
    o2.step()
    for a, b in zip(net2.parameters(), ref.parameters()):
        assert _close(a, b, 2e-6)


def test_fused_adamw_under_gradscaler_skips_inf_steps_without_sync():
    ref = Net().to(DEV)
    net = copy.deepcopy(ref)
    ema = copy.deepcopy(ref).requires_grad_(False)
    ema0 = copy.deepcopy(ema)
    o_ref = torch.optim.AdamW(ref.parameters(), lr=1e-2, weight_decay=0)
    o = FusedAdamWEMA(net.parameters(), lr=1e-2, weight_decay=0)
    o.attach_ema(ema, net, decay=0.5)
    s_ref, s = torch.amp.GradScaler("cuda", init_scale=1024.0), torch.amp.GradScaler("cuda", init_scale=1024.0)
    for step in range(5):
        for m, opt, sc in ((ref, o_ref, s_ref), (net, o, s)):
            sc.scale(torch.ones((), device=DEV))  # what scaler.scale(loss) does first: creates the scale tensor
            _grads(m, step, scale=float(sc.get_scale()))
            if step == 2:
                m.ps[3].grad[0, 0] = float("inf")
            sc.step(opt)
            sc.update()
            opt.zero_grad(set_to_none=True)
        if step == 2:  # skipped on both sides: parameters untouched, scale halved
            assert float(s.get_scale()) == float(s_ref.get_scale()) == 512.0
    for a, b in zip(net.parameters(), ref.parameters()):
        assert _close(a, b, 2e-6)
    assert float(o.state[net.ps[0]]["step"]) == 4.0  # the skipped step did not advance the counter
    assert not torch.equal(ema.ps[0], ema0.ps[0])
