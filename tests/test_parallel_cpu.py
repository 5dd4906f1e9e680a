"""World-size-2 gloo tests (CPU) of the ray-sharding host logic (star_b200.parallel)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import star_b200  # noqa: F401  (registers the package)
from star_b200 import parallel as P


def test_shard_bounds_cover_and_balance():
    for n in (0, 1, 7, 8, 640000, 4097):
        for w in (1, 2, 3, 8):
            b = [P.shard_bounds(n, r, w) for r in range(w)]
            assert b[0][0] == 0 and b[-1][1] == n
            assert all(b[i][1] == b[i + 1][0] for i in range(w - 1))
            sizes = [e - s for s, e in b]
            assert max(sizes) - min(sizes) <= 1


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _run_ranks(worker, world=2, attempts=2):
    """Spawn `world` gloo ranks and collect one result per rank.  The children import torch from a possibly cold page
    cache (a minute on a fresh container), so the waits are generous, and a lost rendezvous (port taken between
    _free_port and init_process_group) is retried once on a fresh port."""
    import queue as _q
    ctx = mp.get_context("spawn")
    last = None
    for _ in range(attempts):
        port = _free_port()
        q = ctx.Queue()
        procs = [ctx.Process(target=worker, args=(r, world, port, q)) for r in range(world)]
        for p in procs:
            p.start()
        try:
            res = sorted([q.get(timeout=600) for _ in range(world)], key=lambda t: t[0])
            for p in procs:
                p.join(timeout=120)
            if all(p.exitcode == 0 for p in procs):
                return res
            last = RuntimeError(f"rank exit codes {[p.exitcode for p in procs]}")
        except _q.Empty as e:
            last = e
        finally:
            for p in procs:
                if p.is_alive():
                    p.kill()
                    p.join(timeout=30)
    raise AssertionError(f"gloo ranks did not finish: {last!r}")


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.manual_seed(0)
        # a "model": rgb = sigmoid(d . w) per ray; every rank holds the same weights and its slice of the rays
        w = torch.nn.Parameter(torch.randn(3, 3))
        pose = torch.nn.Parameter(torch.randn(2, 7))
        unused = torch.nn.Parameter(torch.zeros(4))
        g = torch.Generator().manual_seed(1)
        rays_o, rays_d, target = torch.randn(10, 3, generator=g), torch.randn(10, 3, generator=g), torch.rand(10, 3, generator=g)

        def render(ro, rd):
            return {"rgb": torch.sigmoid(rd @ w + pose[:, :3].sum(0)), "depth": rd.norm(dim=-1), "acc": ro[:, 0]}

        ro, rd, tg = P.shard_rays(rays_o, rays_d, target)
        loss = ((render(ro, rd)["rgb"] - tg) ** 2).mean()
        loss.backward()
        P.allreduce_gradients([w, pose, unused])
        tot = P.allreduce_scalars([loss])
        with torch.no_grad():
            full = P.render_sharded(render, rays_o, rays_d)
        q.put((rank, w.grad.clone(), pose.grad.clone(), unused.grad, tot.clone(),
               {k: v.clone() for k, v in full.items()}))
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_matches_single_process():
    res = _run_ranks(_worker)
    # single-process reference
    torch.manual_seed(0)
    w = torch.nn.Parameter(torch.randn(3, 3))
    pose = torch.nn.Parameter(torch.randn(2, 7))
    g = torch.Generator().manual_seed(1)
    rays_o, rays_d, target = torch.randn(10, 3, generator=g), torch.randn(10, 3, generator=g), torch.rand(10, 3, generator=g)
    rgb = torch.sigmoid(rays_d @ w + pose[:, :3].sum(0))
    loss = ((rgb - target) ** 2).mean()        # equal shards (5 + 5): mean of rank means == global mean
    loss.backward()
    for rank, gw, gp, gu, tot, full in res:
        assert torch.allclose(gw, w.grad, atol=1e-6) and torch.allclose(gp, pose.grad, atol=1e-6)
        assert gu is None      # no rank had a gradient for it: stays None, as in the single-process reference
        assert abs(float(tot[0]) - float(loss)) < 1e-6
        assert torch.allclose(full["rgb"], rgb.detach(), atol=1e-6)
        assert torch.equal(full["depth"], rays_d.norm(dim=-1)) and torch.equal(full["acc"], rays_o[:, 0])
    assert torch.equal(res[0][1], res[1][1])    # ranks agree bit for bit after the all-reduce


def test_uneven_gather():
    assert P.shard_bounds(7, 0, 2) == (0, 4) and P.shard_bounds(7, 1, 2) == (4, 7)
    t = torch.arange(12.0).view(4, 3)
    assert torch.equal(P.gather_rays(t, 4), t)   # world size 1: identity


def _worker_flat(rank, world, port, q):
    """Gradients lying back to back (a net's flat gradient): reduced in place, one collective per run."""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from star_b200 import optim as O_
        torch.manual_seed(0)
        holder = torch.nn.ParameterList([torch.nn.Parameter(torch.randn(5, 3)), torch.nn.Parameter(torch.randn(5)),
                                         torch.nn.Parameter(torch.randn(3, 5)), torch.nn.Parameter(torch.randn(3))])
        O_.flatten_parameters(holder)
        ps = list(holder)
        n = sum(p.numel() for p in ps)
        gbuf = torch.arange(n, dtype=torch.float32) * (rank + 1)          # rank 0: k, rank 1: 2k  -> mean 1.5 k
        lone = torch.nn.Parameter(torch.zeros(2))
        lone.grad = torch.full((2,), float(rank))
        off = 0
        for p in ps:
            p.grad = gbuf[off:off + p.numel()].view(p.shape)
            off += p.numel()
        runs = P.allreduce_gradients(ps + [lone])
        q.put((rank, len(runs), gbuf.clone(), lone.grad.clone(), ps[2].grad.clone()))
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_flat_gradient_runs_are_reduced_in_place():
    res = _run_ranks(_worker_flat)
    n = 15 + 5 + 15 + 3
    for rank, n_runs, gbuf, lone, g2 in res:
        assert n_runs == 2                                               # the flat run + the lone parameter
        assert torch.equal(gbuf, 1.5 * torch.arange(n, dtype=torch.float32))   # reduced where it lies
        assert torch.equal(lone, torch.full((2,), 0.5))
        assert torch.equal(g2, 1.5 * torch.arange(20, 35, dtype=torch.float32).view(3, 5))


class _Toy(torch.nn.Module):
    def __init__(self):
        super().__init__()
        self.a = torch.nn.Linear(3, 4)
        self.b = torch.nn.Linear(4, 3)
        self.unused = torch.nn.Parameter(torch.zeros(2))

    def forward(self, x):
        return torch.sigmoid(self.b(torch.relu(self.a(x))))


def _worker_sync(rank, world, port, q):
    """parallel.GradSync on a module without kernel-side gradient sinks: flat parameter / gradient buffers, the loss
    scaled by 1 / world, one all-reduce of the flat gradient in finish()."""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.manual_seed(0)
        net = _Toy()
        pose = torch.nn.Parameter(torch.randn(2, 7))
        sync = P.GradSync(net, extra_params=[pose])
        g = torch.Generator().manual_seed(1)
        x, target = torch.randn(10, 3, generator=g), torch.rand(10, 3, generator=g)
        grads = []
        for step in range(2):           # two optimiser steps: the buffer is re-zeroed between them
            xs, ts = P.shard_rays(x, target)
            net.zero_grad(set_to_none=True)     # what a trainer does; GradSync re-attaches its views in finish()
            sync.zero_grad()
            loss = ((net(xs + pose[:, :3].sum(0)) - ts) ** 2).mean()
            sync.scale_loss(loss).backward()
            flat = sync.finish()
            grads.append((flat.clone(), pose.grad.clone()))
        ok_views = all(p.grad.data_ptr() == v.data_ptr() for p, v in zip(sync.order, sync.views))
        q.put((rank, grads, ok_views, net.a.weight.grad.clone()))
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_gradsync_matches_single_process():
    res = _run_ranks(_worker_sync)
    torch.manual_seed(0)
    net = _Toy()
    pose = torch.nn.Parameter(torch.randn(2, 7))
    g = torch.Generator().manual_seed(1)
    x, target = torch.randn(10, 3, generator=g), torch.rand(10, 3, generator=g)
    loss = ((net(x + pose[:, :3].sum(0)) - target) ** 2).mean()
    loss.backward()
    ref_flat = torch.cat([p.grad.reshape(-1) if p.grad is not None else torch.zeros(p.numel()) for p in net.parameters()])
    for rank, grads, ok_views, ga in res:
        assert ok_views
        for flat, gp in grads:                                   # both steps: same data, same gradient (no carry-over)
            assert torch.allclose(flat, ref_flat, atol=1e-6)
            assert torch.allclose(gp, pose.grad, atol=1e-6)
        assert torch.allclose(ga, net.a.weight.grad, atol=1e-6)
    assert torch.equal(res[0][1][1][0], res[1][1][1][0])
