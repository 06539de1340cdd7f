"""GPU, world_size 2: data-parallel training THROUGH the fused field node (`sdf_model._field.backward`): gradient chain -> table
scatter -> async all-reduce of the table gradient -> weight-gradient kernels, plus DistributedDataParallel for everything else.

Both ranks share cuda:0 (the driver's GPU test box has one GPU) and exchange over gloo -- what is tested is the host-side
ordering, the SUM + divide averaging (ReduceOp.AVG is NCCL-only) and that the scoped switch does not leak; NCCL itself is
exercised by `bench.py --gpus N` under torchrun.  Oracle: one process rendering both images with the mean of the two losses.
"""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
RES, S = 16, 24


def _build(sg, dev):
    torch.manual_seed(11)
    mo, ro = sg.default_options("ngp", renderer_res=RES, n_samples=S, perturb=0., return_sdf=True)
    g = sg.Generator(mo, ro, full_pipeline=False).to(dev)
    g.renderer.network.encoder.embeddings.data.uniform_(-0.5, 0.5)
    g.renderer.network.precision = "tc16"
    cam, focal, near, far, _ = sg.generate_camera_params(RES, dev, batch=2)
    z = torch.randn(2, 256, device=dev)
    return g, (z, cam, focal, near, far)


def _loss(g_or_ddp, inp, sl):
    z, cam, focal, near, far = (t[sl] for t in inp)
    _, thumb, sdf, eik = g_or_ddp([z], cam, focal, near, far, return_sdf=True, return_eikonal=True)
    w = torch.linspace(-1, 1, thumb[0].numel(), device=thumb.device).view(thumb[0].shape)
    return (thumb * w).sum(dim=(1, 2, 3)).mean() + 10 * sdf.square().mean() + 0 * eik.sum()


PICK = ["renderer.network.encoder.embeddings", "renderer.network.pts_linears.1.weight", "renderer.network.views_linears.gamma.weight",
        "renderer.network.input_linear.weight", "style.0.weight", "renderer.sigmoid_beta"]


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import sdface_gan_b200 as sg
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        dev = torch.device("cuda", 0)
        g, inp = _build(sg, dev)
        with torch.no_grad():
            g.renderer.network.encoder.embeddings.add_(0.01 * rank)      # differs per rank before the wrap: the helper must broadcast it
        model = sg.distributed.data_parallel(g, device_ids=[0])
        assert g.renderer.network._table_exchange is not None and g.renderer.network.encoder._table_exchange is not None
        loss = _loss(model, inp, slice(rank, rank + 1))
        loss.backward()
        torch.cuda.synchronize()
        params = dict(g.named_parameters())
        q.put((rank, {n: params[n].grad.double().norm().item() for n in PICK},
               params[PICK[0]].grad.reshape(-1)[:: 9973].cpu().tolist(), params[PICK[1]].grad.reshape(-1)[:: 97].cpu().tolist()))
    finally:
        dist.destroy_process_group()


def test_data_parallel_step_through_the_field_node_world2():
    sys.path.insert(0, ROOT)
    import sdface_gan_b200 as sg
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 33500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=600) for _ in procs)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    # single process, both images, mean of the two per-image losses == average of the two ranks' gradients
    dev = torch.device("cuda", 0)
    g, inp = _build(sg, dev)
    assert g.renderer.network._table_exchange is None                    # nothing process-global leaked into a fresh model
    (0.5 * (_loss(g, inp, slice(0, 1)) + _loss(g, inp, slice(1, 2)))).backward()
    torch.cuda.synchronize()
    params = dict(g.named_parameters())
    ref_norm = {n: params[n].grad.double().norm().item() for n in PICK}
    ref_tab = torch.tensor(params[PICK[0]].grad.reshape(-1)[:: 9973].cpu().tolist())
    ref_w = torch.tensor(params[PICK[1]].grad.reshape(-1)[:: 97].cpu().tolist())
    for rank, norms, tab, w in res:
        for n in PICK:
            assert abs(norms[n] - ref_norm[n]) <= 1e-3 * ref_norm[n] + 1e-12, (rank, n, norms[n], ref_norm[n])
        assert torch.allclose(torch.tensor(tab), ref_tab, rtol=1e-3, atol=1e-6 * float(ref_tab.abs().max()))
        assert torch.allclose(torch.tensor(w), ref_w, rtol=2e-3, atol=1e-4 * float(ref_w.abs().max()))
    assert res[0][2] == res[1][2]                                        # both ranks hold the SAME averaged table gradient
