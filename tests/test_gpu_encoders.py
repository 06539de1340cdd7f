"""GPU parity of the encoders and the ray sampler against the oracle (and, where present, the reference's own CUDA extension)."""
import importlib.util
import os

import numpy as np
import pytest
import torch

import helpers as H
import oracle
from oracle import field_oracle as fo

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _sg():
    import sdface_gan_b200 as sg
    return sg


def _ngp_grid():
    offsets, pls = oracle.grid_offsets(**fo.NGP_GRID)
    return offsets, pls, float(np.float32(np.log2(pls)))


def _inputs(n, seed=0, edge=True):
    rs = np.random.RandomState(seed)
    x = rs.uniform(-1.15, 1.15, (n, 3)).astype(np.float32)      # renderer range: (x+2)/4 in [0.21, 0.79]
    if edge:
        x[0] = [2.0, -2.0, 0.0]      # exactly on the boundary -> inside
        x[1] = [2.5, 0.0, 0.0]       # outside -> zeros
        x[2] = [-2.0000002, 0, 0]    # just outside
        x[3] = [0, 0, 0]
    return x


def test_level_scales_and_corner_indices_bit_exact():
    sg = _sg()
    offsets, pls, S = _ngp_grid()
    dev = "cuda"
    scales = sg.ops.grid_level_scales(16, S, 16, dev).cpu().numpy()
    x = _inputs(20000)
    u = ((torch.from_numpy(x) + 2.0) / 4.0).numpy()              # grid.py:149 in torch fp32
    ref = oracle.grid_encode_forward(u, np.zeros((offsets[-1], 2), np.float32), offsets, S, 16, level_scales=scales, want_corners=True)
    idx, w = sg.ops.grid_corner_indices(torch.from_numpy(x).to(dev), torch.from_numpy(offsets).to(dev), 2, S, 16, bound=2.0)
    idx = idx.cpu().numpy().view(np.uint32)
    inside = ~((u < 0) | (u > 1)).any(1)
    assert inside.sum() > 19000 and (~inside).sum() >= 2
    assert np.array_equal(idx[inside], ref["corner_idx"][inside])            # bit-exact rows
    assert np.all(idx[~inside] == 0xFFFFFFFF)
    assert np.array_equal(w.cpu().numpy()[inside], ref["corner_w"][inside])  # bit-exact trilinear weights
    # libm's exp2f and CUDA's agree on this table to the last bit or differ by <= 1 ulp; record it
    lib_scales = oracle.grid_level_scales(16, S, 16)
    assert np.max(np.abs(scales - lib_scales) / lib_scales) < 2e-7


@pytest.mark.parametrize("layout", [0, 1])
def test_grid_forward_and_dydx_match_oracle(layout):
    sg = _sg()
    offsets, pls, S = _ngp_grid()
    dev = "cuda"
    rs = np.random.RandomState(1)
    table = rs.uniform(-1, 1, (offsets[-1], 2)).astype(np.float32)
    x = _inputs(8191)
    u = ((torch.from_numpy(x) + 2.0) / 4.0).numpy()
    scales = sg.ops.grid_level_scales(16, S, 16, dev).cpu().numpy()
    ref = oracle.grid_encode_forward(u, table, offsets, S, 16, calc_dy_dx=True, level_scales=scales)
    out, dd = sg.ops.grid_encode_forward(torch.from_numpy(x).to(dev), torch.from_numpy(table).to(dev), torch.from_numpy(offsets).to(dev),
                                         S, 16, bound=2.0, calc_dy_dx=True, layout=layout)
    out = out.cpu().numpy()
    want = ref["outputs"] if layout == 1 else ref["outputs"].transpose(1, 0, 2).reshape(len(x), -1)
    assert np.array_equal(out, want)          # same fmaf chain, same order -> bit-exact features
    dd = dd.cpu().numpy().reshape(16, 3, 2, len(x)).transpose(3, 0, 1, 2)     # component-major on the device
    scale = np.abs(ref["dy_dx"]).max()
    assert np.abs(dd - ref["dy_dx"]).max() <= 2e-6 * scale


def test_grid_backward_matches_oracle_and_transpose_property():
    sg = _sg()
    offsets, pls, S = _ngp_grid()
    dev = "cuda"
    rs = np.random.RandomState(2)
    table = rs.uniform(-1, 1, (offsets[-1], 2)).astype(np.float32)
    x = _inputs(6000)
    u = ((torch.from_numpy(x) + 2.0) / 4.0).numpy()
    scales = sg.ops.grid_level_scales(16, S, 16, dev).cpu().numpy()
    g = rs.standard_normal((len(x), 32)).astype(np.float32)
    ref_f = oracle.grid_encode_forward(u, table, offsets, S, 16, calc_dy_dx=True, level_scales=scales)
    g_lnc = np.ascontiguousarray(g.reshape(len(x), 16, 2).transpose(1, 0, 2))
    ref_ge, ref_gi = oracle.grid_encode_backward(g_lnc, u, table, offsets, S, 16, dy_dx=ref_f["dy_dx"], level_scales=scales)
    xt, tt, ot = torch.from_numpy(x).to(dev), torch.from_numpy(table).to(dev), torch.from_numpy(offsets).to(dev)
    out, dd = sg.ops.grid_encode_forward(xt, tt, ot, S, 16, bound=2.0, calc_dy_dx=True)
    ge = torch.zeros_like(tt)
    _, gi = sg.ops.grid_encode_backward(torch.from_numpy(g).to(dev), xt, tt, ot, S, 16, bound=2.0, dy_dx=dd, grad_embeddings=ge,
                                        want_grad_inputs=True)
    ge = ge.cpu().numpy()
    # atomics reorder the sums: compare with a tolerance relative to the accumulated magnitude
    assert np.abs(ge - ref_ge).max() <= 1e-5 * max(1.0, np.abs(ref_ge).max())
    assert np.array_equal(ge != 0, ref_ge != 0)                   # exactly the same rows are touched
    # oracle grad_inputs is w.r.t. u; the fused kernel returns it w.r.t. x = 4u - 2
    assert np.abs(gi.cpu().numpy() * 4.0 - ref_gi).max() <= 1e-4 * max(1.0, np.abs(ref_gi).max())
    lhs = float((g.astype(np.float64) * out.cpu().numpy()).sum())
    rhs = float((ge.astype(np.float64) * table).sum())
    assert abs(lhs - rhs) < 1e-3 * max(1.0, abs(lhs))


def test_grid_module_autograd_and_2d_c4_variants():
    sg = _sg()
    dev = "cuda"
    torch.manual_seed(0)
    for D, C, gridtype, align in ((2, 4, "hash", False), (3, 1, "tiled", True), (3, 8, "hash", False), (2, 2, "tiled", False)):
        enc = sg.GridEncoder(input_dim=D, num_levels=6, level_dim=C, base_resolution=4, log2_hashmap_size=10, desired_resolution=64,
                             gridtype=gridtype, align_corners=align).to(dev)
        enc.embeddings.data.uniform_(-1, 1)
        x = (torch.rand(999, D, device=dev) * 2 - 1).requires_grad_(True)
        y = enc(x, bound=1)
        w = torch.randn_like(y)
        (y * w).sum().backward()
        offsets = enc.offsets.cpu().numpy()
        S = float(np.float32(np.log2(enc.per_level_scale)))
        scales = sg.ops.grid_level_scales(6, S, 4, dev).cpu().numpy()
        u = ((x.detach().cpu() + 1) / 2).numpy()
        gid = {"hash": 0, "tiled": 1}[gridtype]
        ref = oracle.grid_encode_forward(u, enc.embeddings.detach().cpu().numpy(), offsets, S, 4, calc_dy_dx=True, gridtype=gid,
                                         align_corners=align, level_scales=scales)
        want = ref["outputs"].transpose(1, 0, 2).reshape(999, -1)
        assert np.abs(y.detach().cpu().numpy() - want).max() < 1e-6
        g_lnc = np.ascontiguousarray(w.cpu().numpy().reshape(999, 6, C).transpose(1, 0, 2))
        rge, rgi = oracle.grid_encode_backward(g_lnc, u, enc.embeddings.detach().cpu().numpy(), offsets, S, 4, dy_dx=ref["dy_dx"],
                                               gridtype=gid, align_corners=align, level_scales=scales)
        assert np.abs(enc.embeddings.grad.cpu().numpy() - rge).max() < 1e-4 * max(1.0, np.abs(rge).max())
        assert np.abs(x.grad.cpu().numpy() * 2.0 - rgi).max() < 1e-4 * max(1.0, np.abs(rgi).max())


def test_grid_empty_and_unsupported():
    sg = _sg()
    enc = sg.GridEncoder(num_levels=2, log2_hashmap_size=8).cuda()
    assert enc(torch.zeros(0, 3, device="cuda")).shape == (0, 4)
    with pytest.raises(RuntimeError, match="input_dim"):
        sg.GridEncoder(input_dim=4, num_levels=2, log2_hashmap_size=8).cuda()(torch.zeros(2, 4, device="cuda"))


def test_grid_total_variation_matches_oracle():
    sg = _sg()
    dev = "cuda"
    enc = sg.GridEncoder(num_levels=5, level_dim=2, base_resolution=4, log2_hashmap_size=9, desired_resolution=32).to(dev)
    enc.embeddings.data.uniform_(-1, 1)
    x = torch.rand(500, 3, device=dev) * 2 - 1
    enc.embeddings.grad = torch.zeros_like(enc.embeddings)
    enc.grad_total_variation(weight=0.5, inputs=x, bound=1)
    S = float(np.float32(np.log2(enc.per_level_scale)))
    scales = sg.ops.grid_level_scales(5, S, 4, dev).cpu().numpy()
    ref = np.zeros(tuple(enc.embeddings.shape), np.float32)
    oracle.grad_total_variation(((x.cpu() + 1) / 2).numpy(), enc.embeddings.detach().cpu().numpy(), ref, enc.offsets.cpu().numpy(), 0.5, S, 4,
                                level_scales=scales)
    assert np.abs(enc.embeddings.grad.cpu().numpy() - ref).max() < 1e-4 * max(1.0, np.abs(ref).max())


def _load_ref(name):
    path = os.path.join(ROOT, "oracle", "_ref", name + ".so")
    if not os.path.exists(path):
        pytest.skip("oracle/_ref/%s.so not built (needs /root/reference at build time)" % name)
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def test_grid_matches_unmodified_reference_extension():
    """oracle/_ref/_gridencoder_ref.so = the reference's gridencoder.cu compiled unchanged: the GPU ground truth."""
    ref = _load_ref("_gridencoder_ref")
    sg = _sg()
    offsets, pls, S64 = _ngp_grid()
    dev = "cuda"
    rs = np.random.RandomState(3)
    table = torch.from_numpy(rs.uniform(-1, 1, (offsets[-1], 2)).astype(np.float32)).to(dev)
    x = torch.from_numpy(_inputs(100003, seed=4)).to(dev)
    ot = torch.from_numpy(offsets).to(dev)
    u = (x + 2.0) / 4.0
    B = x.shape[0]
    out_ref = torch.empty(16, B, 2, device=dev)
    dd_ref = torch.empty(B, 16 * 3 * 2, device=dev)
    ref.grid_encode_forward(u, table, ot, out_ref, B, 3, 2, 16, float(np.log2(pls)), 16, dd_ref, 0, False, 0)
    out, dd = sg.ops.grid_encode_forward(x, table, ot, S64, 16, bound=2.0, calc_dy_dx=True, layout=1)
    torch.cuda.synchronize()
    assert torch.equal(out, out_ref)                                # bit-exact features => bit-exact indices and weights
    assert (dd.view(96, B).t() - dd_ref).abs().max().item() <= 2e-6 * dd_ref.abs().max().item()
    g = torch.randn(16, B, 2, device=dev)
    ge_ref = torch.zeros_like(table)
    gi_ref = torch.zeros(B, 3, device=dev)
    ref.grid_encode_backward(g, u, table, ot, ge_ref, B, 3, 2, 16, float(np.log2(pls)), 16, dd_ref, gi_ref, 0, False, 0)
    ge = torch.zeros_like(table)
    _, gi = sg.ops.grid_encode_backward(g, x, table, ot, S64, 16, bound=2.0, dy_dx=dd, grad_embeddings=ge, want_grad_inputs=True, layout=1)
    torch.cuda.synchronize()
    assert (ge - ge_ref).abs().max().item() <= 1e-5 * max(1.0, ge_ref.abs().max().item())
    assert (gi * 4.0 - gi_ref).abs().max().item() <= 1e-4 * max(1.0, gi_ref.abs().max().item())


def test_sh_matches_unmodified_reference_extension():
    ref = _load_ref("_shencoder_ref")
    sg = _sg()
    dev = "cuda"
    d = torch.nn.functional.normalize(torch.randn(50000, 3, device=dev), dim=-1)
    for deg in (1, 4, 8):
        out_ref = torch.empty(50000, deg * deg, device=dev)
        dd_ref = torch.empty(50000, 3 * deg * deg, device=dev)
        ref.sh_encode_forward(d, out_ref, 50000, 3, deg, dd_ref)
        out, dd = sg.ops.sh_encode_forward(d, deg, True)
        torch.cuda.synchronize()
        assert (out - out_ref).abs().max().item() < 1e-5      # fp32 Horner vs the reference's expanded polynomials
        assert (dd - dd_ref).abs().max().item() < 1e-4


def _ref_style_grid_function(backend, inputs, embeddings, offsets, per_level_scale, H, grad_out):
    """The call sequence of the reference's `_grid_encode` autograd.Function (gridencoder/grid.py:24-89) against a `_backend`
    object: forward with dy_dx, permute to [B, L*C], then backward of `grad_out` incl. grad_inputs.  Written from the reference's
    documented contract (caller allocates, [L,B,C] layout, pre-zeroed sinks) so that any `_backend` can be dropped in."""
    B, D = inputs.shape
    L, C = offsets.shape[0] - 1, embeddings.shape[1]
    S = float(np.log2(per_level_scale))
    outputs = torch.empty(L, B, C, device=inputs.device)
    dy_dx = torch.empty(B, L * D * C, device=inputs.device)
    backend.grid_encode_forward(inputs, embeddings, offsets, outputs, B, D, C, L, S, H, dy_dx, 0, False, 0)
    out = outputs.permute(1, 0, 2).reshape(B, L * C)
    grad = grad_out.view(B, L, C).permute(1, 0, 2).contiguous()
    grad_embeddings = torch.zeros_like(embeddings)
    grad_inputs = torch.zeros_like(inputs)
    backend.grid_encode_backward(grad, inputs, embeddings, offsets, grad_embeddings, B, D, C, L, S, H, dy_dx, grad_inputs, 0, False, 0)
    return out, grad_embeddings, grad_inputs


def test_compat_backend_is_a_drop_in_for_the_reference_extension_modules():
    """INTEGRATION.md section 2, executed: `compat_backend.grid_backend` / `sh_backend` (ctypes over libsdfg.so) driven through the
    reference wrappers' call sequence side by side with the UNMODIFIED reference extensions (oracle/_ref, the same pybind API)."""
    ref = _load_ref("_gridencoder_ref")
    sg = _sg()
    offsets, pls, _ = _ngp_grid()
    dev = "cuda"
    rs = np.random.RandomState(11)
    table = torch.from_numpy(rs.uniform(-1, 1, (offsets[-1], 2)).astype(np.float32)).to(dev)
    u = ((torch.from_numpy(_inputs(60001, seed=5)).to(dev) + 2.0) / 4.0).contiguous()
    ot = torch.from_numpy(offsets).to(dev)
    g = torch.randn(u.shape[0], 32, device=dev)
    o_r, ge_r, gi_r = _ref_style_grid_function(ref, u, table, ot, pls, 16, g)
    o_s, ge_s, gi_s = _ref_style_grid_function(sg.compat_backend.grid_backend, u, table, ot, pls, 16, g)
    torch.cuda.synchronize()
    assert torch.equal(o_s, o_r)
    assert (ge_s - ge_r).abs().max().item() <= 1e-5 * max(1.0, ge_r.abs().max().item())
    assert (gi_s - gi_r).abs().max().item() <= 1e-4 * max(1.0, gi_r.abs().max().item())
    # total-variation gradient (grid.py:165-185)
    tv_r, tv_s = torch.zeros_like(table), torch.zeros_like(table)
    B = 50000
    ref.grad_total_variation(u[:B].contiguous(), table, tv_r, ot, 0.5, B, 3, 2, 16, float(np.log2(pls)), 16, 0, False)
    sg.compat_backend.grid_backend.grad_total_variation(u[:B].contiguous(), table, tv_s, ot, 0.5, B, 3, 2, 16, float(np.log2(pls)), 16, 0, False)
    torch.cuda.synchronize()
    assert (tv_s - tv_r).abs().max().item() <= 1e-4 * max(1.0, tv_r.abs().max().item())
    # spherical harmonics: forward + dy_dx + backward (shencoder/sphere_harmonics.py:14-58)
    refs = _load_ref("_shencoder_ref")
    d = torch.nn.functional.normalize(torch.randn(30000, 3, device=dev), dim=-1)
    gs = torch.randn(30000, 16, device=dev)
    res = []
    for be in (refs, sg.compat_backend.sh_backend):
        out = torch.empty(30000, 16, device=dev)
        dd = torch.empty(30000, 48, device=dev)
        be.sh_encode_forward(d, out, 30000, 3, 4, dd)
        gi = torch.zeros(30000, 3, device=dev)
        be.sh_encode_backward(gs, d, 30000, 3, 4, dd, gi)
        res.append((out, gi))
    torch.cuda.synchronize()
    assert (res[1][0] - res[0][0]).abs().max().item() < 1e-5
    assert (res[1][1] - res[0][1]).abs().max().item() < 1e-4 * max(1.0, res[0][1].abs().max().item())
    # error behaviour: CPU tensors raise RuntimeError, like TORCH_CHECK in the reference (gridencoder.cu:15-18)
    with pytest.raises(RuntimeError):
        sg.compat_backend.grid_backend.grid_encode_forward(u.cpu(), table, ot, torch.empty(16, u.shape[0], 2, device=dev), u.shape[0], 3, 2, 16,
                                                           float(np.log2(pls)), 16, None, 0, False, 0)


def test_sh_matches_oracle_and_golden():
    sg = _sg()
    z = H.load_fixture("sh_deg8")
    d = torch.from_numpy(z["dirs"]).cuda()
    for deg in (1, 2, 3, 4, 5, 8):
        out, dd = sg.ops.sh_encode_forward(d, deg, True)
        assert np.abs(out.cpu().numpy() - z["outputs"][:, :deg * deg]).max() < 1e-5
        assert np.abs(dd.cpu().numpy().reshape(-1, 3, deg * deg) - z["dy_dx"][:, :, :deg * deg]).max() < 1e-4
    enc = sg.SHEncoder(degree=4).cuda()
    x = d.clone().requires_grad_(True)
    y = enc(x)
    w = torch.randn_like(y)
    (y * w).sum().backward()
    o, dd = oracle.sh_encode_forward(z["dirs"], 4, True)
    gi = oracle.sh_encode_backward(w.cpu().numpy(), 4, dd)
    assert np.abs(x.grad.cpu().numpy() - gi).max() < 1e-4
    assert enc(torch.zeros(0, 3, device="cuda")).shape == (0, 16)


@pytest.mark.parametrize("mode", ["plain", "offset_jitter", "stratified", "static_nonorm"])
def test_ray_sampler_matches_oracle(mode):
    sg = _sg()
    dev = "cuda"
    B, R, S = 3, 16, 24
    g = torch.Generator().manual_seed(5)
    loc = torch.stack([0.3 * torch.randn(B, generator=g), 0.15 * torch.randn(B, generator=g)], 1)
    cam, focal, near, far, _ = sg.generate_camera_params(R, "cpu", locations=loc)
    offset = mode != "stratified"
    t_rand = None
    jitter = 0
    if mode == "offset_jitter":
        t_rand, jitter = torch.rand(B, R, R, generator=g), 1
    elif mode == "stratified":
        t_rand, jitter = torch.rand(B, R, R, S, generator=g), 2
    static, znorm = (mode == "static_nonorm"), (mode != "static_nonorm")
    ro, rd, vd = fo.get_rays(focal, cam, R, static)
    z = fo.sample_depths(near, far, R, S, offset, t_rand)
    pts = ro.unsqueeze(3) + rd.unsqueeze(3) * z.unsqueeze(-1)
    npts = pts * 2 / (far - near).view(-1, 1, 1, 1, 1) if znorm else pts
    t_vals = torch.linspace(0., 1. - 1 / S, S) if offset else torch.linspace(0., 1., S)
    r = sg.ops.sample_rays(cam.to(dev), focal.to(dev), near.to(dev), far.to(dev), t_vals.to(dev), None if t_rand is None else t_rand.to(dev),
                           jitter, static, znorm, R, S)
    assert H.max_abs(r["z_vals"], z) < 2e-7
    assert H.max_abs(r["rays_d"], rd) < 2e-7
    assert H.max_abs(r["viewdirs"], vd) < 2e-7
    assert H.max_abs(r["pts"], pts) < 3e-7
    assert H.max_abs(r["npts"], npts) < 3e-6
