"""GPU parity at BASELINE size against the CPU oracle (oracle/field_oracle.py, pinned by the reference fixtures).

The reference fixtures are tiny (B = 2, 6-8^2 rays); here the REAL shapes run: configs[1]/[2] = 64 x 64 rays x 24 samples per image
(98 304 samples, 768 tiles of 128 per image) with cameras from `generate_camera_params`, and the configs[4] mesh-query shape
(128 x 128 rays x 128 samples, static view directions, forced background).  Both CUDA paths -- the fp32 kernels and the
benchmarked tensor-core path -- are compared with the oracle on the same seeded inputs: every output map, the eikonal term
and every parameter gradient.

Tolerances (north star): fp32 path max-abs 1e-3 on rendered maps; tensor-core path 2e-2 relative (rgb map: 2e-2 of its [-1, 1]
range); gradients 1e-2 relative L2 per parameter tensor for both.  The oracle needs ~2 s per image forward + backward.
"""
import numpy as np
import pytest
import torch

import helpers as H
import oracle
from oracle import field_oracle as fo

pytestmark = pytest.mark.gpu
DEV = "cuda"
R, S, B = 64, 24, 2

_CACHE = {}


def _sg():
    import sdface_gan_b200 as sg
    return sg


def _setup(table_amp, features):
    """One seeded generator + inputs + the oracle's outputs and gradients (cached per configuration: the oracle is the slow part)."""
    key = (table_amp, features)
    if key in _CACHE:
        return _CACHE[key]
    sg = _sg()
    torch.manual_seed(2024)
    over = dict(perturb=1.0, return_sdf=True)
    if not features:
        over["no_features_output"] = True
    mo, ro = sg.default_options("ngp", renderer_res=R, n_samples=S, **over)
    g = sg.Generator(mo, ro, full_pipeline=False)
    g.renderer.network.encoder.embeddings.data.uniform_(-table_amp, table_amp)
    cam, focal, near, far, _ = sg.generate_camera_params(R, "cpu", batch=B)
    z = torch.randn(B, 256)
    t_rand = torch.rand(B, R, R)
    gen = torch.Generator().manual_seed(7)
    lw = dict(thumb=torch.randn(B, 3, R, R, generator=gen), sdf=torch.randn(B, R, R, S, 1, generator=gen),
              feat=torch.randn(B, 256, R, R, generator=gen) if features else None)
    # oracle (CPU): same parameters by name
    params = {n: p.detach().clone().requires_grad_(True) for n, p in g.named_parameters()}
    rp, sp = H.oracle_param_dicts(params)
    style = fo.mapping(sp, z)
    rgb, feat, sdf, _, _, eik = fo.render(rp, cam, focal, near, far, style, res=R, S=S, t_rand=t_rand, output_features=features,
                                          return_sdf=True, return_eikonal=True)
    loss = _loss(rgb, sdf, feat, lw)
    loss.backward()
    ref = dict(rgb=rgb.detach(), feat=None if feat is None else feat.detach(), sdf=sdf.detach(), eik=eik.detach(), loss=float(loss),
               grads={n: p.grad for n, p in params.items() if p.grad is not None})
    _CACHE[key] = (g, (cam, focal, near, far, z, t_rand), lw, ref)
    return _CACHE[key]


def _loss(rgb, sdf, feat, lw):
    dev = rgb.device
    loss = (lw["thumb"].to(dev) * rgb).sum() / rgb.numel() ** 0.5 + (lw["sdf"].to(dev) * sdf).sum() / sdf.numel() ** 0.5
    if feat is not None:
        loss = loss + (lw["feat"].to(dev) * feat).sum() / feat.numel() ** 0.5
    return loss


@pytest.mark.parametrize("precision", ["fp32", "tc16"])
@pytest.mark.parametrize("table_amp,features", [(1.0, True), (1e-4, False)])
def test_full_size_forward_backward_matches_oracle(table_amp, features, precision):
    """B = 2 images at 64^2 x 24 with jittered depths: table U(-1, 1) with the feature map (configs[2]-style outputs, all gradient
    paths live) and the reference's init U(-1e-4, 1e-4) in the stage-1 configuration (configs[1]: no feature output)."""
    g, (cam, focal, near, far, z, t_rand), lw, ref = _setup(table_amp, features)
    g = g.to(DEV)
    g.renderer.network.precision = precision
    g.zero_grad()
    d = lambda t: t.to(DEV)
    _, thumb, sdf, eik = g([d(z)], d(cam), d(focal), d(near), d(far), return_sdf=True, return_eikonal=True, t_rand=d(t_rand))
    feat = None
    if features:
        style = g.style(d(z))
        _, feat, _, _, _, _ = g.renderer(d(cam), d(focal), d(near), d(far), styles=style, t_rand=d(t_rand))
    fp32 = precision == "fp32"
    assert H.max_abs(thumb, ref["rgb"]) < (1e-3 if fp32 else 2e-2)
    if fp32:
        assert H.max_abs(sdf, ref["sdf"]) < 1e-3
        assert H.max_abs(eik, ref["eik"]) < 1e-3 * max(1.0, float(ref["eik"].abs().max()))
        if features:
            assert H.max_abs(feat, ref["feat"]) < 1e-3
    else:
        assert H.rel_err(sdf, ref["sdf"]) < 2e-2
        assert H.rel_err(eik, ref["eik"]) < 2e-2
        if features:
            assert H.rel_err(feat, ref["feat"]) < 2e-2
    loss = _loss(thumb, sdf, feat, lw)
    loss.backward()
    errs = {}
    # floor of the relative error: a gradient that is itself the rounding-level residue of cancelling terms is measured against a
    # fraction of the LARGEST gradient norm instead of its own.  Only d sigmoid_beta with the reference's 1e-4 table init is such a
    # case: |d beta| = 2e-7 next to gradient norms of ~1 (the field is almost constant, every sample's contribution cancels), and
    # already the fp32 CUDA path and the CPU oracle differ by 1e-8 absolute = 8 % of it.
    gmax = max(float(r.double().norm()) for r in ref["grads"].values())
    floor = 1e-6 * gmax
    if table_amp < 1e-3:
        assert float(ref["grads"]["renderer.sigmoid_beta"].abs().max()) < 1e-5 * gmax      # the premise of the special case
    for n, p in g.named_parameters():
        r = ref["grads"].get(n)
        if r is None or float(r.abs().max()) == 0.0:
            continue
        assert p.grad is not None, n
        fl = 2e-4 * gmax if (n == "renderer.sigmoid_beta" and table_amp < 1e-3) else floor
        errs[n] = float((p.grad.detach().cpu().double() - r.double()).norm() / max(float(r.double().norm()), fl))
    g.to("cpu")
    worst = max(errs.items(), key=lambda kv: kv[1])
    print("%s table %g features %d: worst gradient rel err %.3e (%s) over %d tensors" % (precision, table_amp, features, worst[1], worst[0], len(errs)))
    assert len(errs) >= 30
    # 1e-2 (north star) for both CUDA paths.  The hard case is tensor-core path x reference init: with a 1e-4 table every sample of
    # an image sees the same activations, so the backward's per-element rounding (cos rebuilt from the saved fp16 sine: coarse
    # where |sin| -> 1) no longer averages out over the 98 304 samples of an image but hits whole rows coherently.  Measured worst
    # 6.1e-3 (gamma head of FiLM layer 1); 1.6e-2 before the saved sines carried their rounding bit.
    tol = 1e-2
    assert worst[1] < tol, [(n, e, float(ref["grads"][n].double().norm())) for n, e in sorted(errs.items(), key=lambda kv: -kv[1])[:5]]


@pytest.mark.parametrize("precision", ["fp32", "tc16"])
def test_mesh_query_shape_matches_oracle(precision):
    """configs[4] / sdf_mesh.py:243-253: 128 x 128 rays x 128 samples, static view directions, forced background, perturb = 0,
    return sdf + xyz + mask (2.1 M samples of one identity)."""
    sg = _sg()
    torch.manual_seed(5)
    Rm = 128
    mo, ro = sg.default_options("ngp", renderer_res=Rm, n_samples=Rm, perturb=0., return_sdf=True, return_xyz=True, static_viewdirs=True,
                                force_background=True)
    g = sg.Generator(mo, ro, full_pipeline=False)
    g.renderer.network.encoder.embeddings.data.uniform_(-1.0, 1.0)
    cam, focal, near, far, _ = sg.generate_camera_params(Rm, "cpu", batch=1)
    z = torch.randn(1, 256)
    params = {n: p.detach() for n, p in g.named_parameters()}
    rp, sp = H.oracle_param_dicts(params)
    with torch.no_grad():
        o_rgb, _, o_sdf, o_mask, o_xyz, _ = fo.render(rp, cam, focal, near, far, fo.mapping(sp, z), res=Rm, S=Rm, static_viewdirs=True,
                                                      force_background=True, return_sdf=True, return_xyz=True, output_features=True)
        g = g.to(DEV)
        g.renderer.network.precision = precision
        d = lambda t: t.to(DEV)
        _, thumb, xyz, sdf, mask = g([d(z)], d(cam), d(focal), d(near), d(far), return_sdf=True, return_xyz=True)
    if precision == "fp32":
        assert H.max_abs(thumb, o_rgb) < 1e-3 and H.max_abs(sdf, o_sdf) < 1e-3 and H.max_abs(xyz, o_xyz) < 1e-3 and H.max_abs(mask, o_mask) < 1e-3
    else:
        assert H.max_abs(thumb, o_rgb) < 2e-2 and H.rel_err(sdf, o_sdf) < 2e-2 and H.rel_err(xyz, o_xyz) < 2e-2 and H.max_abs(mask, o_mask) < 2e-2
