"""Tensor-level wrappers over the C-ABI (include/sdfg.h): argument checking, output allocation, current-stream launch.

Every function here takes CUDA float32 tensors, calls exactly one C entry point on torch's current stream and raises
RuntimeError on a non-zero status.  Nothing in this module (or anywhere in the package) computes on the CPU.
"""
import ctypes
import os

import numpy as np
import torch

from . import _lib


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _ptr(t):
    if t is None:
        return None
    return ctypes.c_void_p(t.data_ptr())


def _chk(t, name, dtype=torch.float32):
    if t is None:
        return None
    if not t.is_cuda:
        raise RuntimeError("%s must be a CUDA tensor (this path has no CPU fallback)" % name)
    if t.dtype != dtype:
        raise RuntimeError("%s must be %s (got %s)" % (name, dtype, t.dtype))
    if not t.is_contiguous():
        raise RuntimeError("%s must be contiguous" % name)
    return t


# ------------------------------------------------------------------------------------------------------------------
# rays

def sample_rays(c2w, focal, near, far, t_vals, t_rand, jitter_mode, static_viewdirs, z_normalize, R, S, want_pts=True):
    """sdfg_sample_rays.  Returns dict(z_vals [B,R,R,S], pts, npts [B,R,R,S,3], viewdirs, rays_d [B,R,R,3])."""
    lib = _lib.load()
    B = c2w.shape[0]
    c2w = _chk(c2w.reshape(B, 12).contiguous().float(), "c2w")
    focal = _chk(focal.reshape(B).contiguous().float(), "focal")
    near = _chk(near.reshape(B).contiguous().float(), "near")
    far = _chk(far.reshape(B).contiguous().float(), "far")
    t_vals = _chk(t_vals.reshape(S).contiguous().float(), "t_vals")
    if t_rand is not None:
        t_rand = _chk(t_rand.contiguous().float(), "t_rand")
        want = B * R * R * (S if jitter_mode == 2 else 1)
        if t_rand.numel() != want:
            raise RuntimeError("t_rand has %d elements, expected %d" % (t_rand.numel(), want))
    dev = c2w.device
    z = torch.empty(B, R, R, S, device=dev)
    pts = torch.empty(B, R, R, S, 3, device=dev) if want_pts else None
    npts = torch.empty(B, R, R, S, 3, device=dev)
    vd = torch.empty(B, R, R, 3, device=dev)
    rd = torch.empty(B, R, R, 3, device=dev)
    with torch.cuda.device(dev):
        _lib.check(lib.sdfg_sample_rays(_ptr(c2w), _ptr(focal), _ptr(near), _ptr(far), _ptr(t_vals), _ptr(t_rand),
                                        int(jitter_mode), int(bool(static_viewdirs)), int(bool(z_normalize)), B, R, S,
                                        _ptr(z), _ptr(pts), _ptr(npts), _ptr(vd), _ptr(rd), _stream()), "sdfg_sample_rays")
    return dict(z_vals=z, pts=pts, npts=npts, viewdirs=vd, rays_d=rd)


# ------------------------------------------------------------------------------------------------------------------
# hash grid

def grid_encode_forward(inputs, embeddings, offsets, S, H, bound=0.0, calc_dy_dx=False, gridtype=0, align_corners=False,
                        interp=0, layout=_lib.LAYOUT_NLC, outputs=None, dy_dx=None):
    lib = _lib.load()
    _chk(inputs, "inputs"); _chk(embeddings, "embeddings"); _chk(offsets, "offsets", torch.int32)
    N, D = inputs.shape
    C = embeddings.shape[1]
    L = offsets.shape[0] - 1
    if outputs is None:
        outputs = torch.empty((N, L * C) if layout == _lib.LAYOUT_NLC else (L, N, C), device=inputs.device)
    if calc_dy_dx and dy_dx is None:
        dy_dx = torch.empty(L * D * C, N, device=inputs.device)      # component-major (include/sdfg.h)
    with torch.cuda.device(inputs.device):
        _lib.check(lib.sdfg_grid_encode_forward(_ptr(inputs), _ptr(embeddings), _ptr(offsets), _ptr(_chk(outputs, "outputs")), N, D, C, L,
                                                float(S), int(H), float(bound), _ptr(_chk(dy_dx, "dy_dx")), int(gridtype), int(bool(align_corners)),
                                                int(interp), int(layout), _stream()), "sdfg_grid_encode_forward")
    return outputs, dy_dx


def grid_encode_backward(grad, inputs, embeddings, offsets, S, H, bound=0.0, dy_dx=None, grad_embeddings=None, want_grad_inputs=False,
                         gridtype=0, align_corners=False, interp=0, layout=_lib.LAYOUT_NLC):
    """Scatter `grad` into grad_embeddings (accumulating; pass None to skip) and/or reduce grad_inputs."""
    lib = _lib.load()
    _chk(grad, "grad"); _chk(inputs, "inputs"); _chk(offsets, "offsets", torch.int32)
    N, D = inputs.shape
    C = embeddings.shape[1]
    L = offsets.shape[0] - 1
    gi = torch.empty(N, D, device=inputs.device) if want_grad_inputs else None
    with torch.cuda.device(inputs.device):
        _lib.check(lib.sdfg_grid_encode_backward(_ptr(grad), _ptr(inputs), _ptr(embeddings), _ptr(offsets), _ptr(_chk(grad_embeddings, "grad_embeddings")),
                                                 N, D, C, L, float(S), int(H), float(bound), _ptr(_chk(dy_dx, "dy_dx")), _ptr(gi), int(gridtype),
                                                 int(bool(align_corners)), int(interp), int(layout), _stream()), "sdfg_grid_encode_backward")
    return grad_embeddings, gi


def grad_total_variation(inputs, embeddings, grad, offsets, weight, S, H, gridtype=0, align_corners=False):
    lib = _lib.load()
    _chk(inputs, "inputs"); _chk(embeddings, "embeddings"); _chk(grad, "grad"); _chk(offsets, "offsets", torch.int32)
    N, D = inputs.shape
    C = embeddings.shape[1]
    L = offsets.shape[0] - 1
    with torch.cuda.device(inputs.device):
        _lib.check(lib.sdfg_grad_total_variation(_ptr(inputs), _ptr(embeddings), _ptr(grad), _ptr(offsets), float(weight), N, D, C, L,
                                                 float(S), int(H), int(gridtype), int(bool(align_corners)), _stream()), "sdfg_grad_total_variation")


def grid_level_scales(L, S, H, device):
    lib = _lib.load()
    out = torch.empty(L, device=device)
    with torch.cuda.device(out.device):
        _lib.check(lib.sdfg_grid_level_scales(_ptr(out), L, float(S), int(H), _stream()), "sdfg_grid_level_scales")
    return out


def grid_corner_indices(inputs, offsets, C, S, H, bound=0.0, gridtype=0, align_corners=False):
    lib = _lib.load()
    _chk(inputs, "inputs"); _chk(offsets, "offsets", torch.int32)
    N, D = inputs.shape
    L = offsets.shape[0] - 1
    idx = torch.empty(N, L, 1 << D, device=inputs.device, dtype=torch.int32)
    w = torch.empty(N, L, 1 << D, device=inputs.device)
    with torch.cuda.device(inputs.device):
        _lib.check(lib.sdfg_grid_corner_indices(_ptr(inputs), _ptr(offsets), _ptr(idx), _ptr(w), N, D, C, L, float(S), int(H), float(bound),
                                                int(gridtype), int(bool(align_corners)), _stream()), "sdfg_grid_corner_indices")
    return idx, w


def l2_gather_probe(buf, threads, rounds):
    """Issue threads*8*rounds random 8-byte gathers from buf [rows, 2] (roofline probe; time it with CUDA events)."""
    lib = _lib.load()
    _chk(buf, "buf")
    sink = torch.zeros(1, device=buf.device)
    with torch.cuda.device(buf.device):
        _lib.check(lib.sdfg_l2_gather_probe(_ptr(buf), buf.shape[0], int(threads), int(rounds), _ptr(sink), _stream()), "sdfg_l2_gather_probe")
    return threads * 8 * rounds


# ------------------------------------------------------------------------------------------------------------------
# spherical harmonics

def sh_encode_forward(inputs, degree, calc_dy_dx=False):
    lib = _lib.load()
    _chk(inputs, "inputs")
    N = inputs.shape[0]
    out = torch.empty(N, degree * degree, device=inputs.device)
    dy_dx = torch.empty(N, 3 * degree * degree, device=inputs.device) if calc_dy_dx else None
    with torch.cuda.device(inputs.device):
        _lib.check(lib.sdfg_sh_encode_forward(_ptr(inputs), _ptr(out), N, int(degree), _ptr(dy_dx), _stream()), "sdfg_sh_encode_forward")
    return out, dy_dx


def sh_encode_backward(grad, dy_dx, degree):
    lib = _lib.load()
    _chk(grad, "grad"); _chk(dy_dx, "dy_dx")
    N = grad.shape[0]
    gi = torch.zeros(N, 3, device=grad.device)
    with torch.cuda.device(grad.device):
        _lib.check(lib.sdfg_sh_encode_backward(_ptr(grad), _ptr(dy_dx), _ptr(gi), N, int(degree), _stream()), "sdfg_sh_encode_backward")
    return gi


# ------------------------------------------------------------------------------------------------------------------
# field

class FieldSpec:
    """Host description of one network (weights are read at call time, so optimizer updates are seen)."""

    def __init__(self, width, in_dim, view_dim, n_film, has_input_linear):
        self.width, self.in_dim, self.view_dim, self.n_film, self.has_input_linear = width, in_dim, view_dim, n_film, has_input_linear


def _field_params(spec, samples_per_image, samples_per_ray, gamma, beta, weights):
    """weights: dict(input_w, input_b, film_w [n+1], film_b [n+1], sigma_w, sigma_b, rgb_w, rgb_b) of CUDA tensors."""
    p = _lib.FieldParams()
    p.width, p.in_dim, p.view_dim, p.n_film = spec.width, spec.in_dim, spec.view_dim, spec.n_film
    p.has_input_linear = int(spec.has_input_linear)
    p.samples_per_image, p.samples_per_ray = int(samples_per_image), int(samples_per_ray)
    if spec.has_input_linear:
        p.input_w = _chk(weights["input_w"], "input_w").data_ptr()
        p.input_b = _chk(weights["input_b"], "input_b").data_ptr()
    for i in range(spec.n_film + 1):
        p.film_w[i] = _chk(weights["film_w"][i], "film_w").data_ptr()
        p.film_b[i] = _chk(weights["film_b"][i], "film_b").data_ptr()
    p.gamma = _chk(gamma, "gamma").data_ptr()
    p.beta = _chk(beta, "beta").data_ptr()
    p.sigma_w = _chk(weights["sigma_w"], "sigma_w").data_ptr()
    p.sigma_b = _chk(weights["sigma_b"], "sigma_b").data_ptr()
    if weights.get("rgb_w") is not None:
        p.rgb_w = _chk(weights["rgb_w"], "rgb_w").data_ptr()
        p.rgb_b = _chk(weights["rgb_b"], "rgb_b").data_ptr()
    return p


def field_forward(spec, x_in, view_feat, gamma, beta, weights, samples_per_image, samples_per_ray, want_rgb=True, want_feat=True,
                  save_for_backward=False, precision=_lib.PRECISION_FP32, feat_f16=False):
    """Returns (sdf [N], rgb [N,3]|None, feat [N,W]|None, workspace).  feat_f16 (tensor-core path, inference only): the
    features come back as a float16 tensor (sdfg_field_forward_h), to be consumed by composite_forward."""
    lib = _lib.load()
    _chk(x_in, "x_in"); _chk(view_feat, "view_feat")
    N = x_in.shape[0]
    dev = x_in.device
    p = _field_params(spec, samples_per_image, samples_per_ray, gamma, beta, weights)
    nbytes = int(lib.sdfg_field_workspace_bytes(ctypes.byref(p), N, int(save_for_backward), int(precision)))
    ws = torch.empty(max(nbytes, 16) // 4, device=dev, dtype=torch.float32)
    sdf = torch.empty(N, device=dev)
    rgb = torch.empty(N, 3, device=dev) if want_rgb else None
    if feat_f16 and want_feat:
        if save_for_backward or precision != _lib.PRECISION_TC16:
            raise RuntimeError("fp16 features are an inference output of the tensor-core path")
        feat = torch.empty(N, spec.width, device=dev, dtype=torch.float16)
        with torch.cuda.device(dev):
            _lib.check(lib.sdfg_field_forward_h(ctypes.byref(p), _ptr(x_in), _ptr(view_feat), N, _ptr(sdf), _ptr(rgb), _ptr(feat), _ptr(ws),
                                                _stream()), "sdfg_field_forward_h")
        return sdf, rgb, feat, ws
    feat = torch.empty(N, spec.width, device=dev) if want_feat else None
    with torch.cuda.device(dev):
        _lib.check(lib.sdfg_field_forward(ctypes.byref(p), _ptr(x_in), _ptr(view_feat), N, _ptr(sdf), _ptr(rgb), _ptr(feat), _ptr(ws),
                                          int(save_for_backward), int(precision), _stream()), "sdfg_field_forward")
    return sdf, rgb, feat, ws


def field_backward(spec, x_in, view_feat, gamma, beta, weights, samples_per_image, samples_per_ray, workspace, out_feat,
                   d_sdf, d_rgb, d_feat, grads=None, want_dx=False, precision=_lib.PRECISION_FP32, phases=_lib.BWD_BOTH, state=None):
    """grads: None (no parameter gradients) or dict like `weights` + gamma/beta of pre-zeroed (or live .grad) buffers that are
    accumulated into.  Returns d_x_in [N,in_dim] or None.
    phases (sdfg_field_backward_phase): BWD_CHAIN returns (d_x_in, state); a later call with phases=BWD_WGRAD, state=state and
    otherwise identical arguments produces the parameter gradients (the caller enqueues whatever should run in between)."""
    lib = _lib.load()
    N = x_in.shape[0]
    dev = x_in.device
    p = _field_params(spec, samples_per_image, samples_per_ray, gamma, beta, weights)
    g = None
    if grads is not None:
        g = _lib.FieldGrads()
        if spec.has_input_linear:
            g.input_w = _chk(grads["input_w"], "d_input_w").data_ptr()
            g.input_b = _chk(grads["input_b"], "d_input_b").data_ptr()
        for i in range(spec.n_film + 1):
            g.film_w[i] = _chk(grads["film_w"][i], "d_film_w").data_ptr()
            g.film_b[i] = _chk(grads["film_b"][i], "d_film_b").data_ptr()
        g.gamma = _chk(grads["gamma"], "d_gamma").data_ptr()
        g.beta = _chk(grads["beta"], "d_beta").data_ptr()
        g.sigma_w = _chk(grads["sigma_w"], "d_sigma_w").data_ptr()
        g.sigma_b = _chk(grads["sigma_b"], "d_sigma_b").data_ptr()
        if grads.get("rgb_w") is not None:
            g.rgb_w = _chk(grads["rgb_w"], "d_rgb_w").data_ptr()
            g.rgb_b = _chk(grads["rgb_b"], "d_rgb_b").data_ptr()
    if state is not None:
        scratch, dx = state
    else:
        nbytes = int(lib.sdfg_field_backward_scratch_bytes(ctypes.byref(p), N, int(precision)))
        scratch = torch.empty(max(nbytes, 16) // 4, device=dev, dtype=torch.float32)
        dx = torch.empty(N, spec.in_dim, device=dev) if want_dx else None
    gref = ctypes.byref(g) if g is not None else None
    with torch.cuda.device(dev):
        if phases != _lib.BWD_BOTH:
            _lib.check(lib.sdfg_field_backward_phase(ctypes.byref(p), gref, _ptr(x_in), _ptr(view_feat), N,
                                                     _ptr(_chk(d_sdf, "d_sdf")), _ptr(_chk(d_rgb, "d_rgb")), _ptr(_chk(d_feat, "d_feat")),
                                                     _ptr(out_feat), _ptr(workspace), _ptr(scratch), _ptr(dx), int(precision), int(phases),
                                                     _stream()), "sdfg_field_backward_phase")
            return dx, (scratch, dx)
        _lib.check(lib.sdfg_field_backward(ctypes.byref(p), gref, _ptr(x_in), _ptr(view_feat), N,
                                           _ptr(_chk(d_sdf, "d_sdf")), _ptr(_chk(d_rgb, "d_rgb")), _ptr(_chk(d_feat, "d_feat")),
                                           _ptr(out_feat), _ptr(workspace), _ptr(scratch), _ptr(dx), int(precision), _stream()),
                   "sdfg_field_backward")
    return dx


def field_eikonal(spec, x_in, view_feat, gamma, beta, weights, samples_per_image, samples_per_ray, workspace, d_sdf, dy_dx, scale,
                  precision=_lib.PRECISION_TC16):
    """The eikonal pass with the encoder's chain rule fused in (sdfg_field_eikonal): d sdf / d point [N,3] from the saved forward state
    and dy_dx [L*3*2, N], without materialising d sdf / d feature [N,in_dim].  Returns None when the shape / precision has no fused
    kernel (the caller then runs field_backward(want_dx=True) + grid_encode_backward(want_grad_inputs=True))."""
    lib = _lib.load()
    if int(precision) != _lib.PRECISION_TC16 or dy_dx is None or dy_dx.shape[0] * 2 != spec.in_dim * 6 or os.environ.get("SDFG_EIK_FUSE", "1") == "0":
        return None
    N = x_in.shape[0]
    dev = x_in.device
    p = _field_params(spec, samples_per_image, samples_per_ray, gamma, beta, weights)
    nbytes = int(lib.sdfg_field_backward_scratch_bytes(ctypes.byref(p), N, int(precision)))
    scratch = torch.empty(max(nbytes, 16) // 4, device=dev, dtype=torch.float32)
    d_pts = torch.zeros(N, 3, device=dev)
    with torch.cuda.device(dev):
        code = lib.sdfg_field_eikonal(ctypes.byref(p), _ptr(x_in), _ptr(view_feat), N, _ptr(_chk(d_sdf, "d_sdf")), _ptr(workspace), _ptr(scratch),
                                      _ptr(_chk(dy_dx, "dy_dx")), 3, 2, float(scale), _ptr(d_pts), int(precision), _stream())
    if code == _lib.ERR_UNSUPPORTED:
        return None
    _lib.check(code, "sdfg_field_eikonal")
    return d_pts


def tc_linear_probe(x, w):
    """out = fp16(x) @ fp16(w).T with fp32 accumulation, through the tcgen05 layer pipeline (parity-test probe)."""
    lib = _lib.load()
    _chk(x, "x"); _chk(w, "w")
    M, K = x.shape
    N = w.shape[0]
    out = torch.empty(M, N, device=x.device)
    ws = torch.empty(int(lib.sdfg_tc_linear_probe_workspace_bytes(M, K, N)), device=x.device, dtype=torch.uint8)
    with torch.cuda.device(x.device):
        _lib.check(lib.sdfg_tc_linear_probe(_ptr(x), _ptr(w), _ptr(out), M, K, N, _ptr(ws), _stream()), "sdfg_tc_linear_probe")
    return out


def tc_wgrad_probe(dz, x, rows_per_image, x_fmt=0):
    """G[b, j, k] = sum_{n in image b} x16(dz)[n, j] * x16(x)[n, k] plus the ones column; returns (G [B,256,Kx], colsum [B,256])."""
    lib = _lib.load()
    _chk(dz, "dz"); _chk(x, "x")
    N, Kx = x.shape
    B = N // rows_per_image
    G = torch.zeros(B, 256, 336, device=x.device)
    ws = torch.empty(2 * N * (256 + Kx + 8) + 512, device=x.device, dtype=torch.uint8)
    ldg, ones = ctypes.c_uint32(0), ctypes.c_uint32(0)
    # the kernel is told the pitch it must use through G's allocation: probe once for ldg, then view
    with torch.cuda.device(x.device):
        _lib.check(lib.sdfg_tc_wgrad_probe(_ptr(dz), _ptr(x), _ptr(G), N, Kx, int(rows_per_image), int(x_fmt), ctypes.byref(ldg), ctypes.byref(ones),
                                           _ptr(ws), _stream()), "sdfg_tc_wgrad_probe")
    flat = G.reshape(-1)[:B * 256 * ldg.value].view(B, 256, ldg.value)
    return flat[:, :, :Kx], flat[:, :, ones.value]


# ------------------------------------------------------------------------------------------------------------------
# compositing

def composite_forward(sdf, rgb, feat, z_vals, rays_d, pts, noise, sigmoid_beta, S, with_sdf, force_background, want_xyz):
    """sdf [NR*S]; rgb [NR*S,3] (None: sdf-only query, no rgb map); feat [NR*S,F]|None; z_vals [NR*S]; rays_d [NR,3]; pts [NR*S,3]|None.
    Returns (rgb_map [NR,3]|None, feat_map [NR,F]|None, xyz [NR,3]|None, mask [NR]|None)."""
    lib = _lib.load()
    NR = rays_d.shape[0]
    dev = sdf.device
    F = feat.shape[-1] if feat is not None else 0
    rgb_map = torch.empty(NR, 3, device=dev) if rgb is not None else None
    feat_map = torch.empty(NR, F, device=dev) if feat is not None else None
    xyz = torch.empty(NR, 3, device=dev) if want_xyz else None
    mask = torch.empty(NR, device=dev) if want_xyz else None
    if feat is not None and feat.dtype == torch.float16:
        with torch.cuda.device(dev):
            _lib.check(lib.sdfg_composite_forward_h(_ptr(_chk(sdf, "sdf")), _ptr(_chk(rgb, "rgb")), _ptr(_chk(feat, "feat", torch.float16)),
                                                    _ptr(_chk(z_vals, "z_vals")), _ptr(_chk(rays_d, "rays_d")), _ptr(_chk(pts, "pts")),
                                                    _ptr(_chk(noise, "noise")), _ptr(_chk(sigmoid_beta, "sigmoid_beta")), NR, int(S), int(F),
                                                    int(bool(with_sdf)), int(bool(force_background)), _ptr(rgb_map), _ptr(feat_map), _ptr(xyz),
                                                    _ptr(mask), None, _stream()), "sdfg_composite_forward_h")
        return rgb_map, feat_map, xyz, mask
    with torch.cuda.device(dev):
        _lib.check(lib.sdfg_composite_forward(_ptr(_chk(sdf, "sdf")), _ptr(_chk(rgb, "rgb")), _ptr(_chk(feat, "feat")), _ptr(_chk(z_vals, "z_vals")),
                                              _ptr(_chk(rays_d, "rays_d")), _ptr(_chk(pts, "pts")), _ptr(_chk(noise, "noise")),
                                              _ptr(_chk(sigmoid_beta, "sigmoid_beta")), NR, int(S), int(F), int(bool(with_sdf)),
                                              int(bool(force_background)), _ptr(rgb_map), _ptr(feat_map), _ptr(xyz), _ptr(mask), None,
                                              _stream()), "sdfg_composite_forward")
    return rgb_map, feat_map, xyz, mask


def composite_backward(sdf, rgb, feat, z_vals, rays_d, pts, noise, sigmoid_beta, S, with_sdf, force_background,
                       d_rgb_map, d_feat_map, d_xyz, d_mask, want_d_feat):
    """Returns (d_sdf, d_rgb, d_feat|None, d_sigmoid_beta [1]|None)."""
    lib = _lib.load()
    NR = rays_d.shape[0]
    dev = sdf.device
    F = feat.shape[-1] if feat is not None else 0
    d_sdf = torch.empty_like(sdf)
    d_rgb = torch.empty_like(rgb)
    d_feat = torch.empty_like(feat) if (feat is not None and want_d_feat) else None
    d_beta = torch.zeros(1, device=dev) if with_sdf else None
    use_feat = feat if (d_feat is not None or d_feat_map is not None) else None
    with torch.cuda.device(dev):
        _lib.check(lib.sdfg_composite_backward(_ptr(sdf), _ptr(rgb), _ptr(use_feat), _ptr(z_vals), _ptr(rays_d), _ptr(pts), _ptr(noise),
                                               _ptr(sigmoid_beta), NR, int(S), int(F), int(bool(with_sdf)), int(bool(force_background)),
                                               _ptr(_chk(d_rgb_map, "d_rgb_map")), _ptr(_chk(d_feat_map, "d_feat_map")),
                                               _ptr(_chk(d_xyz, "d_xyz")), _ptr(_chk(d_mask, "d_mask")), _ptr(d_sdf), _ptr(d_rgb),
                                               _ptr(d_feat), None, _ptr(d_beta), _stream()), "sdfg_composite_backward")
    return d_sdf, d_rgb, d_feat, d_beta


def align_volume(volume, near=0.88, far=1.12):
    """sdfg_align_volume: volume [B,H,W,D,C] (frustum-sampled sdf) -> the box-aligned volume marching cubes wants (ref sdf_utils.py:164-184)."""
    lib = _lib.load()
    _chk(volume, "volume")
    if volume.dim() != 5:
        raise RuntimeError("volume must be [B, H, W, D, C]")
    B, H, W, D, C = volume.shape
    out = torch.empty_like(volume)
    with torch.cuda.device(volume.device):
        _lib.check(lib.sdfg_align_volume(_ptr(volume), _ptr(out), B, H, W, D, C, float(near), float(far), _stream()), "sdfg_align_volume")
    return out


def log2_scale(per_level_scale):
    """S = log2(per_level_scale), rounded to float32 as the reference's pybind call does (gridencoder/grid.py:38)."""
    return float(np.float32(np.log2(per_level_scale)))


# ------------------------------------------------------------------------------------------------------------------
# StyleGAN2 decoder (forward): channels-last fp16 activations

def nhwc16(x):
    """fp32 [..., C] (channels last, contiguous) -> fp16, same shape (sdfg_nhwc16)."""
    lib = _lib.load()
    _chk(x, "x")
    out = torch.empty(x.shape, device=x.device, dtype=torch.float16)
    with torch.cuda.device(x.device):
        _lib.check(lib.sdfg_nhwc16(_ptr(x), _ptr(out), x.numel(), _stream()), "sdfg_nhwc16")
    return out


def modconv_fold(weight, style, scale, demodulate=True):
    """weight [1, Cout, Cin, k, k] (reference ModulatedConv2d parameter), style [B, Cin] -> per-sample fp16 weights [B, k*k, Cout, Cin]."""
    lib = _lib.load()
    _, Cout, Cin, k, _ = weight.shape
    w = _chk(weight.reshape(Cout, Cin, k * k).contiguous(), "weight")
    style = _chk(style.contiguous().float(), "style")
    B = style.shape[0]
    out = torch.empty(B, k * k, Cout, Cin, device=w.device, dtype=torch.float16)
    demod = torch.empty(B, Cout, device=w.device) if demodulate else None
    with torch.cuda.device(w.device):
        _lib.check(lib.sdfg_modconv_fold(_ptr(w), _ptr(style), float(scale), B, Cin, Cout, k * k, int(bool(demodulate)), _ptr(demod), _ptr(out),
                                         _stream()), "sdfg_modconv_fold")
    return out


def conv_forward(x, wf, bias=None, noise=None, noise_w=None):
    """x [B,H,W,Cin] fp16; wf [B,taps,Cout,Cin] fp16 -> fp16 [B,H,W,Cout] (3x3 / 1x1 convolution + noise + bias + leaky ReLU * sqrt 2)."""
    lib = _lib.load()
    _chk(x, "x", torch.float16); _chk(wf, "wf", torch.float16)
    B, H, W, Cin = x.shape
    _, taps, Cout, _ = wf.shape
    out = torch.empty(B, H, W, Cout, device=x.device, dtype=torch.float16)
    with torch.cuda.device(x.device):
        _lib.check(lib.sdfg_conv_forward(_ptr(x), _ptr(wf), B, H, W, Cin, Cout, taps, _ptr(_chk(bias, "bias")),
                                         _ptr(_chk(noise, "noise")), _ptr(_chk(noise_w, "noise_w")), _ptr(out), _stream()), "sdfg_conv_forward")
    return out


def upconv_forward(x, wf, bias=None, noise=None, noise_w=None):
    """x [B,H,W,Cin] fp16; wf [B,9,Cout,Cin] fp16 -> fp16 [B,2H,2W,Cout]: transposed convolution (stride 2) + blur + noise + bias + act."""
    lib = _lib.load()
    _chk(x, "x", torch.float16); _chk(wf, "wf", torch.float16)
    B, H, W, Cin = x.shape
    _, taps, Cout, _ = wf.shape
    if taps != 9:
        raise ValueError("upconv_forward: 3 x 3 kernels only")
    t = torch.empty(B, 2 * H + 1, 2 * W + 1, Cout, device=x.device, dtype=torch.float16)
    out = torch.empty(B, 2 * H, 2 * W, Cout, device=x.device, dtype=torch.float16)
    with torch.cuda.device(x.device):
        _lib.check(lib.sdfg_upconv_forward(_ptr(x), _ptr(wf), B, H, W, Cin, Cout, _ptr(t), _ptr(_chk(bias, "bias")), _ptr(_chk(noise, "noise")),
                                           _ptr(_chk(noise_w, "noise_w")), _ptr(out), _stream()), "sdfg_upconv_forward")
    return out


def to_rgb(x, weight, style, scale, bias, skip=None, want_nhwc=True, want_nchw=False):
    """x [B,H,W,C] fp16; weight [1,3,C,1,1]; style [B,C]; bias [1,3,1,1]; skip [B,H/2,W/2,3] fp32|None -> (nhwc [B,H,W,3]|None, nchw [B,3,H,W]|None) fp32."""
    lib = _lib.load()
    _chk(x, "x", torch.float16)
    B, H, W, C = x.shape
    w = _chk(weight.reshape(3, C).contiguous(), "weight")
    style = _chk(style.contiguous().float(), "style")
    b = _chk(bias.reshape(3).contiguous(), "bias")
    scratch = torch.empty(B, C, 4, device=x.device)
    o1 = torch.empty(B, H, W, 3, device=x.device) if want_nhwc else None
    o2 = torch.empty(B, 3, H, W, device=x.device) if want_nchw else None
    with torch.cuda.device(x.device):
        _lib.check(lib.sdfg_to_rgb(_ptr(x), _ptr(w), _ptr(style), float(scale), _ptr(b), _ptr(_chk(skip, "skip")), B, H, W, C, _ptr(scratch),
                                   _ptr(o1), _ptr(o2), _stream()), "sdfg_to_rgb")
    return o1, o2
