"""configs[4] / SURVEY 8(f-3): the sdf-only mesh-extraction query and the frustum -> box resampling (align_volume).

CPU: the oracle's align_volume against the golden produced by the reference's own function (tests/golden/make_golden.py).
GPU: sdfg_align_volume against that golden and the oracle; the sdf-only renderer pass against the full pass."""
import numpy as np
import pytest
import torch

import helpers as H
from oracle import field_oracle as fo


def test_oracle_align_volume_matches_reference_golden():
    z = H.load_fixture("align_volume")
    v = torch.from_numpy(z["volume"])
    assert H.max_abs(fo.align_volume(v), z["out"]) < 1e-6
    assert H.max_abs(fo.align_volume(v, near=0.7, far=1.3), z["out_near07_far13"]) < 1e-6
    assert 0.2 < float((z["out"] == 1).mean()) < 0.5           # the out-of-frustum fill is exercised


@pytest.mark.gpu
def test_align_volume_kernel_matches_reference_golden_and_oracle():
    import sdface_gan_b200 as sg
    z = H.load_fixture("align_volume")
    v = torch.from_numpy(z["volume"]).cuda()
    assert H.max_abs(sg.align_volume(v), z["out"]) < 1e-5
    assert H.max_abs(sg.align_volume(v, near=0.7, far=1.3), z["out_near07_far13"]) < 1e-5
    # the renderer's output shape ([1, R, R, S, 1]) at a mesh-extraction size, batched (every element resampled alike), C = 2
    torch.manual_seed(0)
    vol = torch.randn(1, 64, 64, 64, 1)
    out = sg.align_volume(vol.cuda())
    assert H.max_abs(out, fo.align_volume(vol)) < 1e-5
    vol2 = torch.randn(3, 20, 24, 28, 2)
    out2 = sg.align_volume(vol2.cuda())
    ref2 = torch.cat([torch.cat([fo.align_volume(vol2[b:b + 1, ..., c:c + 1]) for c in range(2)], -1) for b in range(3)], 0)
    assert H.max_abs(out2, ref2) < 1e-5
    with pytest.raises(RuntimeError):
        sg.align_volume(vol)                                   # CPU tensor: no fallback


@pytest.mark.gpu
@pytest.mark.parametrize("precision", ["fp32", "tc16"])
def test_sdf_only_query_equals_full_pass(precision):
    """renderer.sdf_only skips the view layer, rgb and features; sdf / xyz / mask must equal the full pass (same trunk, same weights)."""
    import sdface_gan_b200 as sg
    torch.manual_seed(3)
    Rm = 32
    mo, ro = sg.default_options("ngp", renderer_res=Rm, n_samples=Rm, perturb=0., return_sdf=True, return_xyz=True, static_viewdirs=True,
                                force_background=True)
    g = sg.Generator(mo, ro, full_pipeline=False, ema=True).cuda().eval()      # ema: Generator.forward runs the renderer without grad
    g.renderer.network.encoder.embeddings.data.uniform_(-1, 1)
    g.renderer.network.precision = precision
    cam, focal, near, far, _ = sg.generate_camera_params(Rm, "cuda", batch=2)
    zl = torch.randn(2, 256, device="cuda")
    with torch.no_grad():
        _, thumb, xyz, sdf, mask = g([zl], cam, focal, near, far, return_sdf=True, return_xyz=True)
        g.renderer.sdf_only = True
        sg._lib.launch_count_reset()
        _, thumb2, xyz2, sdf2, mask2 = g([zl], cam, focal, near, far, return_sdf=True, return_xyz=True)
        n_launch = sg._lib.launch_count()
    assert thumb is not None and thumb2 is None
    tol = 1e-6 if precision == "fp32" else 1e-6        # the trunk kernels are the same in both passes
    assert H.max_abs(sdf2, sdf) <= tol and H.max_abs(xyz2, xyz) <= tol and H.max_abs(mask2, mask) <= tol
    assert n_launch > 0
    gg = sg.Generator(mo, ro, full_pipeline=False).cuda()              # training-mode generator: grad is on inside forward
    gg.renderer.sdf_only = True
    with pytest.raises(RuntimeError, match="inference"):
        gg([zl], cam, focal, near, far, return_sdf=True, return_xyz=True)
