"""CPU, world_size 2 over gloo: the data-parallel host logic (unit sharding, weight sync, bucketed gradient averaging)."""
import os
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import sdface_gan_b200 as sg
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.manual_seed(100 + rank)                      # deliberately different initial weights per rank
        # the real parameter set of the hot path (13.66 M elements incl. the 50.6 MB hash table) -- parameters only, no kernels
        mo, ro = sg.default_options("ngp", renderer_res=8)
        net = sg.Generator(mo, ro, full_pipeline=False)
        sg.distributed.sync_parameters(net)
        w0 = torch.cat([p.detach().reshape(-1)[:64] for p in net.parameters()])
        # units: 7 images over 2 ranks -> 4 + 3, contiguous, disjoint, covering
        g = torch.Generator().manual_seed(0)
        z, cam = torch.randn(7, 256, generator=g), torch.randn(7, 3, 4, generator=g)
        zs, cs = sg.distributed.shard_batch((z, cam))
        # synthetic per-rank gradients: grad = rank-dependent value; the average must be the mean over ranks
        for i, p in enumerate(net.parameters()):
            p.grad = torch.full_like(p, float(rank + 1) * (i + 1))
        calls = sg.distributed.average_gradients(net)
        ok = all(torch.allclose(p.grad, torch.full_like(p, 1.5 * (i + 1))) for i, p in enumerate(net.parameters()))
        q.put((rank, w0, zs.shape[0], float(zs.sum()), ok, calls))
    finally:
        dist.destroy_process_group()


def test_shard_sync_and_gradient_average_world2():
    sys.path.insert(0, ROOT)
    import sdface_gan_b200 as sg
    assert [sg.distributed.shard_range(7, r, 2) for r in range(2)] == [(0, 4), (4, 7)]
    assert [sg.distributed.shard_range(64, r, 8) for r in range(8)] == [(8 * r, 8 * r + 8) for r in range(8)]
    assert sum(b - a for a, b in (sg.distributed.shard_range(5, r, 8) for r in range(8))) == 5
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=300) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, w0, n0, s0, ok0, c0), (r1, w1, n1, s1, ok1, c1) = res
    assert torch.equal(w0, w1)                      # identical weights after sync
    assert (n0, n1) == (4, 3)
    g = torch.Generator().manual_seed(0)
    z = torch.randn(7, 256, generator=g)
    assert abs(s0 - float(z[:4].sum())) < 1e-4 and abs(s1 - float(z[4:].sum())) < 1e-4
    assert ok0 and ok1
    assert c0 == c1 and 1 <= c0 <= 4                # 54.6 MB of fp32 gradients in a handful of bucketed collectives


class _Toy(torch.nn.Module):
    """Same parameter naming as the renderer: `...encoder.embeddings` is the table the field node exchanges itself."""

    _table_exchange = None

    def __init__(self):
        super().__init__()
        self.encoder = torch.nn.Module()
        self.encoder._table_exchange = None
        self.encoder.embeddings = torch.nn.Parameter(torch.ones(16, 2))
        self.lin = torch.nn.Linear(2, 1)

    def forward(self, idx, scale):
        return (self.lin(self.encoder.embeddings[idx]) * scale).sum()


def _worker_dp(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import sdface_gan_b200 as sg
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.manual_seed(7 + rank)
        toy = _Toy()
        with torch.no_grad():
            toy.encoder.embeddings.add_(float(rank))              # differs per rank before wrapping
        model = sg.distributed.data_parallel(toy, early_table_exchange=True)
        emb0 = toy.encoder.embeddings.detach().clone()             # rank 0's values everywhere after the wrap
        w0 = toy.lin.weight.detach().clone()
        loss = model(torch.arange(4) + 4 * rank, float(rank + 1))
        loss.backward()
        # the switch is scoped to the wrapped module (owner of the table + its encoder), nothing process-global
        scoped = toy._table_exchange is not None and toy.encoder._table_exchange is not None and _Toy()._table_exchange is None
        # plain lists: a tensor would travel as a shared-memory handle that dies with this process
        q.put((rank, emb0.tolist(), w0.tolist(), toy.encoder.embeddings.grad.tolist(), toy.lin.weight.grad.tolist(), scoped))
    finally:
        dist.destroy_process_group()


def test_data_parallel_takes_the_table_out_of_ddp_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker_dp, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=300) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (_, e0, w0, ge0, gw0, on0), (_, e1, w1, ge1, gw1, on1) = [tuple(torch.tensor(x) if isinstance(x, list) else x for x in r) for r in res]
    assert on0 and on1
    assert torch.equal(e0, e1) and torch.equal(w0, w1)            # table broadcast by the helper, the rest by DDP
    assert torch.allclose(gw0, gw1)                               # DDP averaged the ordinary parameter
    assert not torch.allclose(ge0, ge1)                           # ...and left the table gradient to the field node (rank-local here)
    assert ge0[:4].abs().sum() > 0 and ge0[4:].abs().sum() == 0 and ge1[4:8].abs().sum() > 0
