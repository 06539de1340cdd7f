"""GPU parity of the tcgen05 path (fp16 operands -- activations, weights, loss-scaled gradients -- with fp32 accumulation in TMEM).

Tolerances (north star): rendered maps / features / sdf within 2e-2 relative where the MLP runs in 16-bit operands; GRADIENTS within
1e-2 relative of the REFERENCE (fixture digests, test_tc_training_step_matches_reference_fixture; full-size oracle comparison in
test_gpu_fullsize.py).  The raw GEMM probe is compared against an fp32 matmul of the same fp16-rounded operands, where only the
accumulation order differs (1e-3 relative to the row scale).  Why fp16 and not bf16: tests/test_operand_precision.py."""
import os

import numpy as np
import pytest
import torch

import helpers as H

pytestmark = pytest.mark.gpu
DEV = "cuda"


def _sg():
    import sdface_gan_b200 as sg
    return sg


@pytest.mark.parametrize("M,K,N", [(128, 64, 256), (1000, 256, 256), (4096 + 77, 272, 256), (300, 32, 256), (513, 256, 32), (129, 8, 256),
                                   (20000, 256, 256)])
def test_tc_linear_probe_matches_f16_matmul(M, K, N):
    sg = _sg()
    torch.manual_seed(M + K + N)
    x = torch.randn(M, K, device=DEV)
    w = torch.randn(N, K, device=DEV) / K ** 0.5
    out = sg.ops.tc_linear_probe(x, w)
    ref = x.half().float() @ w.half().float().t()
    torch.cuda.synchronize()
    err = (out - ref).abs().max().item()
    assert err < 2e-3 * ref.abs().max().item(), (err, ref.abs().max().item())
    # exactness of the data path: a one-hot x row picks out single (fp16-rounded) weights
    e = torch.zeros(M, K, device=DEV)
    idx = torch.arange(M, device=DEV) % K
    e[torch.arange(M, device=DEV), idx] = 1.0
    out = sg.ops.tc_linear_probe(e, w)
    assert torch.equal(out, w.half().float().t()[idx])


@pytest.mark.parametrize("name", ["ngp_fwd_tab1", "ngp_fwd_init", "ngp_mesh", "siren_fwd"])
def test_tc_generator_forward_close_to_reference_fixture(name):
    z = H.load_fixture(name)
    g = H.product_generator(z, DEV, precision="tc16")
    inp = H.fixture_inputs(z, DEV)
    kw = {}
    if "out_sdf" in z.files:
        kw["return_sdf"] = True
    if "out_xyz" in z.files:
        kw["return_xyz"] = True
    with torch.no_grad():
        out = g([inp["z"]], inp["cam"], inp["focal"], inp["near"], inp["far"], t_rand=inp["t_rand"], **kw)
        style = g.style(inp["z"])
        _, feat, _, _, _, _ = g.renderer(inp["cam"], inp["focal"], inp["near"], inp["far"], styles=style, t_rand=inp["t_rand"])
    thumb = out[1]
    # the rgb map is -1 + 2*sum(w*sigmoid(.)): a difference of O(1) terms, so its error is measured against its [-1, 1] range;
    # the direct MLP outputs (features, sdf) are compared in relative L2
    assert H.max_abs(thumb, z["out_thumb_rgb"]) < 2e-2
    if "features" in z.files:
        assert H.rel_err(feat, z["features"]) < 2e-2
    if "out_sdf" in z.files:
        sdf = out[3] if "out_xyz" in z.files else out[2]
        assert H.rel_err(sdf, z["out_sdf"]) < 2e-2


def test_tc_field_matches_fp32_field_large():
    """98 304 samples per image (the real 64x64x24 layout), 3 images: per-sample field outputs of the two CUDA paths."""
    sg = _sg()
    torch.manual_seed(0)
    mo, ro = sg.default_options("ngp", renderer_res=64, n_samples=24, perturb=0.)
    g = sg.Generator(mo, ro, full_pipeline=False).to(DEV)
    net = g.renderer.network
    net.encoder.embeddings.data.uniform_(-0.5, 0.5)
    B = 3
    npts = torch.rand(B, 64, 64, 24, 3, device=DEV) * 2 - 1
    vd = torch.nn.functional.normalize(torch.randn(B, 64, 64, 3, device=DEV), dim=-1)
    style = torch.randn(B, 256, device=DEV) * 0.5
    with torch.no_grad():
        net.precision = "fp32"
        sdf0, rgb0, feat0, _ = net.forward_rays(npts, vd, style)
        net.precision = "tc16"
        sdf1, rgb1, feat1, _ = net.forward_rays(npts, vd, style)
    assert H.rel_err(sdf1, sdf0) < 2e-2 and H.rel_err(rgb1, rgb0) < 2e-2 and H.rel_err(feat1, feat0) < 2e-2


def test_tc_rejects_unaligned_images_loudly():
    sg = _sg()
    mo, ro = sg.default_options("ngp", renderer_res=6, n_samples=24, perturb=0.)       # 864 samples per image
    g = sg.Generator(mo, ro, full_pipeline=False).to(DEV)
    g.renderer.network.precision = "tc16"
    cam, focal, near, far, _ = sg.generate_camera_params(6, DEV, batch=2)
    with pytest.raises(RuntimeError, match="multiple of 128"):
        with torch.no_grad():
            g([torch.randn(2, 256, device=DEV)], cam, focal, near, far)


@pytest.mark.parametrize("fmt", [0, 1])   # both operands share one 16-bit type: mixed bf16 x fp16 is an illegal instruction on sm_100a (measured)
@pytest.mark.parametrize("N,Kx,rpi", [(256, 256, 128), (1024, 32, 256), (4096, 272, 1024), (128 * 40, 256, 128 * 8)])
def test_tc_wgrad_probe_matches_matmul(N, Kx, rpi, fmt):
    """The MN-major, sample-axis contraction (fp16 x fp16 as used by the backward, or bf16 x bf16) against an fp32 matmul of the rounded operands."""
    sg = _sg()
    torch.manual_seed(N + Kx)
    dz = torch.randn(N, 256, device=DEV) * 1e-3
    x = torch.randn(N, Kx, device=DEV)
    G, colsum = sg.ops.tc_wgrad_probe(dz, x, rpi, x_fmt=fmt)
    dzr = (dz.bfloat16() if fmt == 1 else dz.half()).float()
    xr = (x.bfloat16() if fmt == 1 else x.half()).float()
    B = N // rpi
    ref = torch.einsum("bnj,bnk->bjk", dzr.view(B, rpi, 256), xr.view(B, rpi, Kx))
    torch.cuda.synchronize()
    assert (G - ref).abs().max().item() < 2e-3 * ref.abs().max().item()
    cs = dzr.view(B, rpi, 256).sum(1)
    assert (colsum - cs).abs().max().item() < 2e-3 * cs.abs().max().item()


@pytest.mark.parametrize("name", ["ngp_train", "ngp_train_feat8"])
def test_tc_training_step_matches_reference_fixture(name):
    """The BENCHMARKED path (tc16) against the reference's own gradient digests at the north star's 1e-2 -- not against the repo's
    fp32 kernels.  8 x 8 rays x 24 samples = 1536 = 12 x 128 samples per image: tensor-core eligible.
    Per tensor the L2 norm, a seeded random projection of the gradient and 16 sampled single entries (against max(|entry|, rms))
    are held to 1e-2 of the reference -- the same bar as the fp32 kernels in test_gpu_render.py.  Measured worst: norm 5.3e-3
    (sigmoid_beta), sampled entry 4.8e-3 (pts_linears.0.weight); before the saved sines carried their rounding bit the
    sampled entries reached 1.2e-2.  Tensors whose reference gradient norm is below 1e-3 of the largest one (sigma_linear.bias under
    a loss without an sdf term: 3e-4 vs 3.5, a residue of cancelling per-sample terms) are held to 1e-2 of that floor instead of
    their own norm."""
    from test_gpu_render import check_training_fixture
    worst, worst_val = check_training_fixture(name, "tc16", 2e-2, 2e-2, val_tol=1e-2,
                                              norm_floor_frac=1e-3)
    print("worst |grad norm| deviation vs reference: %.3e (%s); worst sampled-entry deviation: %.3e (%s)" % (worst + worst_val))


def _train_step(g, z, inp, kw):
    names = ["rgb", "thumb_rgb"] + (["sdf"] if kw.get("return_sdf") else []) + (["eikonal"] if kw.get("return_eikonal") else [])
    out = dict(zip(names, g([inp["z"]], inp["cam"], inp["focal"], inp["near"], inp["far"], t_rand=inp["t_rand"], **kw)))
    loss = 0
    for k in ("thumb_rgb", "sdf"):
        if "lossw_" + k in z.files and k in out:
            loss = loss + (torch.from_numpy(z["lossw_" + k]).to(DEV) * out[k]).sum() / out[k].numel() ** 0.5
    g.zero_grad()
    loss.backward()
    return out, {n: p.grad.detach().clone() for n, p in g.named_parameters() if p.grad is not None}


@pytest.mark.parametrize("name", ["ngp_train", "ngp_fwd_tab1"])
def test_tc_training_step_gradients_close_to_fp32_path(name):
    """Stage-1 step (sdf + eikonal + thumbnail loss; and a with-features variant) through the tensor-core backward (recompute,
    dgrad, MN-major wgrad) against the fp32 CUDA path on identical inputs.  Gradient tolerance: 2e-2 relative L2 per tensor
    (loss-scaled fp16 gradients x fp16 activations; the north star's 1e-2 is met by the fp32 path, see test_gpu_render.py)."""
    z = H.load_fixture(name)
    inp = H.fixture_inputs(z, DEV)
    kw = dict(return_sdf=True, return_eikonal=True) if name == "ngp_train" else {}
    res = {}
    for prec in ("fp32", "tc16"):
        g = H.product_generator(z, DEV, precision=prec)
        if name != "ngp_train":
            z = dict(z.items()) if not isinstance(z, dict) else z
            z.setdefault("lossw_thumb_rgb", np.random.RandomState(1).standard_normal(z["out_thumb_rgb"].shape).astype(np.float32))
            class _Z(dict):
                @property
                def files(self):
                    return list(self.keys())
            z = _Z(z)
        res[prec] = _train_step(g, z, inp, kw)
    out0, g0 = res["fp32"]
    out1, g1 = res["tc16"]
    assert H.max_abs(out1["thumb_rgb"], out0["thumb_rgb"]) < 2e-2
    if "eikonal" in out0:
        assert H.rel_err(out1["eikonal"], out0["eikonal"]) < 2e-2
        assert H.rel_err(out1["sdf"], out0["sdf"]) < 2e-2
    assert set(g0) == set(g1)
    worst = max((H.rel_err(g1[n], g0[n]), n) for n in g0 if g0[n].abs().max() > 0)
    assert worst[0] < 2e-2, worst


_VARIANT_SCRIPT = r"""
import json, sys, torch
sys.path.insert(0, %r)
import sdface_gan_b200 as sg
torch.manual_seed(0)
mo, ro = sg.default_options("ngp", renderer_res=16, n_samples=24, perturb=0., return_sdf=True)
g = sg.Generator(mo, ro, full_pipeline=False).to("cuda")
g.renderer.network.encoder.embeddings.data.uniform_(-0.5, 0.5)
g.renderer.network.precision = "tc16"
cam, focal, near, far, _ = sg.generate_camera_params(16, "cuda", batch=2)
z = torch.randn(2, 256, device="cuda")
_, thumb, sdf, eik = g([z], cam, focal, near, far, return_sdf=True, return_eikonal=True)
loss = (thumb * torch.linspace(-1, 1, thumb.numel(), device="cuda").view_as(thumb)).sum() + sdf.square().mean() * 10
loss.backward()
out = {"thumb": thumb.detach().flatten()[::7].tolist(), "eik": eik.flatten()[::997].tolist()}
for n, p in g.named_parameters():
    if p.grad is not None:
        out["g:" + n] = float(p.grad.norm())
print("RESULT" + json.dumps(out))
"""


def test_tc_kernel_variants_agree():
    '''The fallback kernels must not rot: fused chains on CTA pairs (default), on single CTAs (SDFG_TC_CG=1, what odd tile counts
    get), the two-tiles-in-flight backward chain for every pass / for none (SDFG_TC_PP=2 / 0) and the per-layer kernels
    (SDFG_TC_CHAIN=0, what shapes outside the chains get) are the same computation in the same number formats -- outputs and gradient norms of a
    small training step agree to 1e-2 relative (different accumulation orders, atomics).  Env switches are read once per
    process, hence the subprocesses.'''
    import json, os, subprocess, sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    res = {}
    for name, env in (("pairs", {}), ("single", {"SDFG_TC_CG": "1"}), ("per_layer", {"SDFG_TC_CHAIN": "0"}),
                      ("pingpong_all", {"SDFG_TC_PP": "2"}), ("pingpong_off", {"SDFG_TC_PP": "0"})):
        e = dict(os.environ, **env)
        r = subprocess.run([sys.executable, "-c", _VARIANT_SCRIPT % root], capture_output=True, text=True, env=e, timeout=300)
        assert r.returncode == 0, (name, r.stderr[-2000:])
        line = [l for l in r.stdout.splitlines() if l.startswith("RESULT")][-1]
        res[name] = json.loads(line[len("RESULT"):])
    ref = res["pairs"]
    for name in ("single", "per_layer", "pingpong_all", "pingpong_off"):
        got = res[name]
        assert set(got) == set(ref)
        for k in ref:
            a, b = np.asarray(ref[k], np.float64), np.asarray(got[k], np.float64)
            # the per-layer kernels keep input_linear as its own fp16 layer (no collapse, no hi/lo split of the first layer): same
            # network, coarser rounding where gamma ~ 30 amplifies it -- 3e-2 instead of 1e-2
            tol = 3e-2 if name == "per_layer" else 1e-2
            assert np.linalg.norm(a - b) <= tol * max(np.linalg.norm(a), 1e-12), (name, k, a, b)


@pytest.mark.parametrize("B,R,S", [(1, 8, 24), (3, 16, 24), (2, 32, 12), (5, 8, 48)])
def test_tc_training_step_small_and_ragged_shapes(B, R, S):
    """Fused chains at shapes where the persistent grids are ragged (fewer tile pairs than SM pairs, odd image counts, one tile
    pair per image boundary ...): outputs and every parameter gradient against the fp32 CUDA path, 2e-2 relative."""
    sg = _sg()
    res = {}
    for prec in ("fp32", "tc16"):
        torch.manual_seed(7)
        mo, ro = sg.default_options("ngp", renderer_res=R, n_samples=S, perturb=0., return_sdf=True)
        g = sg.Generator(mo, ro, full_pipeline=False).to(DEV)
        g.renderer.network.encoder.embeddings.data.uniform_(-0.5, 0.5)
        g.renderer.network.precision = prec
        cam, focal, near, far, _ = sg.generate_camera_params(R, DEV, batch=B)
        z = torch.randn(B, 256, device=DEV)
        _, thumb, sdf, eik = g([z], cam, focal, near, far, return_sdf=True, return_eikonal=True)
        w = torch.linspace(-1, 1, thumb.numel(), device=DEV).view_as(thumb)
        loss = (thumb * w).sum() + 10 * sdf.square().mean() + ((eik.norm(dim=-1) - 1) ** 2).mean()
        loss.backward()
        res[prec] = (thumb.detach(), sdf.detach(), eik.detach(), {n: p.grad.detach().clone() for n, p in g.named_parameters() if p.grad is not None})
    t0, s0, e0, g0 = res["fp32"]
    t1, s1, e1, g1 = res["tc16"]
    assert H.max_abs(t1, t0) < 2e-2 and H.rel_err(s1, s0) < 2e-2 and H.rel_err(e1, e0) < 2e-2
    assert set(g0) == set(g1)
    worst = max((H.rel_err(g1[n], g0[n]), n) for n in g0 if g0[n].abs().max() > 0)
    assert worst[0] < 2e-2, worst


def test_fused_eikonal_contraction_matches_the_two_kernel_form(monkeypatch):
    """sdfg_field_eikonal applies the hash encoder's chain rule (dy_dx) inside the eikonal chain's last epilogue; SDFG_EIK_FUSE=0 runs the
    older form -- d sdf / d feature [N,32] written by the chain, contracted by grid_input_backward_kernel.  Same arithmetic up to the
    order of one fp32 sum: 1e-5 relative.  The eikonal term stays detached from the parameters (ngp mode, SURVEY finding 4)."""
    z = H.load_fixture("ngp_train")
    inp = H.fixture_inputs(z, DEV)
    g = H.product_generator(z, DEV, precision="tc16")
    res = {}
    for fuse in ("1", "0"):
        monkeypatch.setenv("SDFG_EIK_FUSE", fuse)
        out = g([inp["z"]], inp["cam"], inp["focal"], inp["near"], inp["far"], t_rand=inp["t_rand"], return_sdf=True, return_eikonal=True)
        res[fuse] = out[3].detach().clone()
        assert not out[3].requires_grad
    assert torch.isfinite(res["1"]).all() and float(res["0"].abs().max()) > 0
    assert H.rel_err(res["1"], res["0"]) < 1e-5
