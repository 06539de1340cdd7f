"""CPU: the oracle (oracle/) against the golden vectors produced by running the reference itself (tests/golden/make_golden.py).

This is what pins the restatement.  Tolerances: the fixtures were produced by the reference's torch code on CPU and the oracle
is torch code on CPU with the same operation order, so agreement is at float32 rounding level (1e-5 abs).
"""
import numpy as np
import pytest
import torch

import helpers as H
import param_fill as pf
import oracle
from oracle import field_oracle as fo

FWD_CASES = ["siren_fwd", "ngp_fwd_init", "ngp_fwd_tab1", "ngp_mesh", "ngp_nosdf_strat", "fc_fwd"]


def _run_oracle(z, return_eikonal=False, requires_grad=False):
    cfg = H.fixture_cfg(z)
    params = H.fixture_params(z, requires_grad=requires_grad)
    rp, sp = H.oracle_param_dicts(params)
    inp = H.fixture_inputs(z)
    style = fo.mapping(sp, inp["z"])
    kw = H.render_kwargs(cfg)
    out = fo.render(rp, inp["cam"], inp["focal"], inp["near"], inp["far"], style, t_rand=inp["t_rand"] if cfg["perturb"] > 0 else None,
                    return_eikonal=return_eikonal, **kw)
    return params, style, out


@pytest.mark.parametrize("name", FWD_CASES)
def test_forward_matches_reference(name):
    z = H.load_fixture(name)
    with torch.no_grad():
        _, style, (rgb, feat, sdf, mask, xyz, _) = _run_oracle(z)
    assert H.max_abs(style, z["style"]) < 1e-5
    assert H.max_abs(rgb, z["out_thumb_rgb"]) < 2e-5
    if "features" in z.files:
        assert H.max_abs(feat, z["features"]) < 2e-5
    if "out_sdf" in z.files:
        assert H.max_abs(sdf, z["out_sdf"]) < 2e-5
    if "out_xyz" in z.files:
        assert H.max_abs(xyz, z["out_xyz"]) < 2e-5
        assert H.max_abs(mask, z["out_mask"]) < 2e-5


@pytest.mark.parametrize("name", ["ngp_train", "ngp_train_feat", "ngp_train_feat8", "siren_train"])
def test_training_step_matches_reference(name):
    z = H.load_fixture(name)
    want_eik = "out_eikonal" in z.files
    params, style, (rgb, feat, sdf, mask, xyz, eik) = _run_oracle(z, return_eikonal=want_eik, requires_grad=True)
    assert H.max_abs(rgb, z["out_thumb_rgb"]) < 2e-5
    if want_eik:
        assert H.max_abs(sdf, z["out_sdf"]) < 2e-5
        scale = max(1.0, float(np.abs(z["out_eikonal"]).max()))
        assert H.max_abs(eik, z["out_eikonal"]) < 1e-4 * scale
    loss = 0
    for k, v in (("thumb_rgb", rgb), ("sdf", sdf)):
        if "lossw_" + k in z.files and v is not None and v.requires_grad:
            loss = loss + (torch.from_numpy(z["lossw_" + k]) * v).sum() / v.numel() ** 0.5
    assert abs(float(loss) - float(z["loss"])) < 1e-5 * max(1.0, abs(float(z["loss"])))
    loss.backward()
    checked = 0
    for pname, p in params.items():
        key = "g_norm_" + pname
        if key not in z.files:
            continue
        assert p.grad is not None, pname
        d = pf.grad_digest(pname, p.grad.numpy())
        ref_norm = float(z[key])
        tol = 2e-4 * max(ref_norm, 1e-12) + 1e-9
        assert abs(d["norm"] - ref_norm) < tol, (pname, d["norm"], ref_norm)
        assert abs(d["proj"] - float(z["g_proj_" + pname])) < 5e-4 * max(ref_norm, 1e-12) + 1e-9, pname
        checked += 1
    assert checked >= 20


def test_init_pass_matches_reference():
    z = H.load_fixture("ngp_init_pass")
    params = H.fixture_params(z)
    rp, sp = H.oracle_param_dicts(params)
    inp = H.fixture_inputs(z)
    with torch.no_grad():
        sdf, target = fo.mlp_init_pass(rp, inp["cam"], inp["focal"], inp["near"], inp["far"], fo.mapping(sp, inp["z"]), inp["t_rand"],
                                       res=int(z["cfg_res"]), S=int(z["cfg_S"]))
    assert H.max_abs(sdf, z["init_sdf"]) < 2e-5
    assert H.max_abs(target, z["init_target"]) < 2e-6


def test_sh_oracle_matches_reference_polynomials():
    z = H.load_fixture("sh_deg8")
    for deg in (1, 2, 4, 8):
        out, dd = oracle.sh_encode_forward(z["dirs"], deg, calc_dy_dx=True)
        assert np.abs(out - z["outputs"][:, :deg * deg]).max() < 2e-5
        assert np.abs(dd - z["dy_dx"][:, :, :deg * deg]).max() < 2e-4


def test_camera_matches_reference():
    import sdface_gan_b200 as sg
    z = H.load_fixture("camera")
    cam, focal, near, far, vp = sg.generate_camera_params(64, "cpu", locations=torch.from_numpy(z["loc"]), fov_ang=6, dist_radius=0.12)
    assert H.max_abs(cam, z["cam"]) < 1e-6
    assert H.max_abs(focal, z["focal"]) < 1e-3
    assert H.max_abs(near, z["near"]) == 0 and H.max_abs(far, z["far"]) == 0
    assert H.max_abs(vp, z["vp"]) == 0


def test_grid_oracle_level_table_and_properties():
    """Level table of grid.py:97-131 and linearity / out-of-bounds behaviour of the C restatement."""
    offsets, pls = oracle.grid_offsets(**fo.NGP_GRID)
    assert offsets[-1] == 6328848 and abs(pls - 1.447269) < 1e-6
    sizes = np.diff(offsets)
    assert list(sizes[:5]) == [4920, 15632, 42880, 125000, 373248] and all(s == 524288 for s in sizes[5:])
    rs = np.random.RandomState(0)
    x = rs.uniform(0.2, 0.8, (257, 3)).astype(np.float32)
    x[0] = [1.5, 0.5, 0.5]                      # out of bounds -> zeros
    x[1] = [0.0, 1.0, 0.5]                      # boundary stays inside
    t1 = rs.uniform(-1, 1, (offsets[-1], 2)).astype(np.float32)
    t2 = rs.uniform(-1, 1, (offsets[-1], 2)).astype(np.float32)
    S = float(np.float32(np.log2(pls)))
    f1 = oracle.grid_encode_forward(x, t1, offsets, S, 16, want_corners=True)
    f2 = oracle.grid_encode_forward(x, t2, offsets, S, 16)["outputs"]
    f12 = oracle.grid_encode_forward(x, t1 + t2, offsets, S, 16)["outputs"]
    assert np.abs(f12 - (f1["outputs"] + f2)).max() < 1e-5       # linear in the table
    assert np.all(f1["outputs"][:, 0] == 0) and np.any(f1["outputs"][:, 1] != 0)
    w = f1["corner_w"][1:]
    assert np.abs(w.sum(-1) - 1).max() < 1e-5                     # trilinear weights sum to one
    idx = f1["corner_idx"][1:]
    assert idx.max() < 524288
    # backward is the transpose of forward:  <g, F(t)> == <F^T(g), t>
    g = rs.standard_normal(f1["outputs"].shape).astype(np.float32)
    gt, _ = oracle.grid_encode_backward(g, x, t1, offsets, S, 16)
    lhs = float((g.astype(np.float64) * f1["outputs"]).sum())
    rhs = float((gt.astype(np.float64) * t1).sum())
    assert abs(lhs - rhs) < 1e-3 * max(1.0, abs(lhs))
