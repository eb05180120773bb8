"""World-size-2 gloo test (CPU) of the multi-GPU sampling logic: beatmaps are sharded with their CFG
pairs intact, every rank works independently (no data-path collective), results reassemble in order."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from osudit import shard


def test_shard_range_covers_everything_once():
    for n in (1, 2, 7, 64, 513):
        for world in (1, 2, 3, 8):
            spans = [shard.shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard.shard_range(4, 2, 2)


def _fake_sampler(z, o, c, y):
    """Stands in for p_sample_loop on a CPU rank: a per-row function of all inputs of that row and of
    its CFG partner, so a broken pairing or ordering changes the result."""
    n = z.shape[0] // 2
    partner = torch.cat([z[n:], z[:n]])
    return z * 2 + o[:, None, :] * 1e-3 + c.sum(1, keepdim=True) + y[:, None, None] + partner.flip(-1) * 0.5


def _worker(rank, world, port, n, T, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = torch.Generator().manual_seed(0)
        z = torch.randn(2 * n, 2, T, generator=g)
        o = torch.rand(2 * n, T, generator=g)
        c = torch.randn(2 * n, 5, T, generator=g)
        y = torch.arange(2 * n)
        zs, os_, cs, ys = shard.shard_cfg_batch([z, o, c, y], rank, world)
        a, b = shard.shard_range(n, rank, world)
        assert zs.shape[0] == 2 * (b - a)
        assert torch.equal(ys, torch.cat([y[a:b], y[n + a:n + b]]))  # pairs stay together
        local = _fake_sampler(zs, os_, cs, ys)
        full = shard.gather_cfg_samples(local, n, rank, world)
        if rank == 0:
            q.put(torch.equal(full, _fake_sampler(z, o, c, y)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n", [5, 8])
def test_two_rank_sampling_matches_single_rank(n):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n, 16, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    assert q.get(timeout=10) is True
