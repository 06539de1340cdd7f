"""oracle/field_oracle.py -- TEST INFRASTRUCTURE ONLY (never imported by the product).

Torch-CPU, functional restatement of the reference's per-ray-sample field evaluation and volume rendering
(SURVEY.md section 8a rows a2-a14).  All citations are /root/reference/im2scene/sdf/models/sdf_model.py unless noted.
Parameters are passed as a flat dict keyed like the reference ``state_dict`` of ``Generator.renderer``
(e.g. ``network.pts_linears.0.gamma.weight``), so one weight set drives the reference, this oracle and the CUDA path.

Pinned by tests/golden/*.npz, produced by running the reference's own classes (tests/golden/make_golden.py).
Gradients come from torch autograd over this restatement; the hash grid differentiates through the C restatement.
"""
import math

import numpy as np
import torch
import torch.nn.functional as F

from . import grid_encode_backward, grid_encode_forward, grid_offsets, sh_encode_forward

# hash-grid hyper-parameters of NGPSIRENGenerator: sdf_model.py:1512-1531,1545 (desired_resolution = 2048*bound)
NGP_GRID = dict(input_dim=3, num_levels=16, level_dim=2, base_resolution=16, log2_hashmap_size=19,
                desired_resolution=4096, align_corners=False)
NGP_BOUND = 2.0            # sdf_model.py:1540
NGP_SH_DEGREE = 4          # get_encoder default, sdf_model.py:1514


class _GridEncodeOracle(torch.autograd.Function):
    """gridencoder/grid.py:24-89 with the C restatement as backend (outputs [B, L*C])."""

    @staticmethod
    def forward(ctx, inputs, embeddings, offsets, S, H, level_scales):
        calc = inputs.requires_grad
        r = grid_encode_forward(inputs.detach().numpy(), embeddings.detach().numpy(), offsets.numpy(), S, H,
                                calc_dy_dx=calc, level_scales=level_scales)
        out = torch.from_numpy(r["outputs"])            # [L,B,C]
        L, B, C = out.shape
        ctx.save_for_backward(inputs, embeddings, offsets)
        ctx.dy_dx = r["dy_dx"]
        ctx.meta = (S, H, level_scales)
        return out.permute(1, 0, 2).reshape(B, L * C)   # grid.py:57

    @staticmethod
    def backward(ctx, grad):
        inputs, embeddings, offsets = ctx.saved_tensors
        S, H, level_scales = ctx.meta
        B = inputs.shape[0]
        L = offsets.shape[0] - 1
        C = embeddings.shape[1]
        g = grad.detach().reshape(B, L, C).permute(1, 0, 2).contiguous().numpy()     # grid.py:75 (opaque kernel: no double backward)
        ge, gi = grid_encode_backward(g, inputs.detach().numpy(), embeddings.detach().numpy(), offsets.numpy(), S, H,
                                      dy_dx=ctx.dy_dx, level_scales=level_scales)
        return (torch.from_numpy(gi) if gi is not None else None), torch.from_numpy(ge), None, None, None, None


def hash_encode(x, embeddings, offsets, per_level_scale, base_resolution, bound, level_scales=None):
    """GridEncoder.forward, gridencoder/grid.py:145-161: map [-bound,bound] -> [0,1], encode, restore prefix shape."""
    u = (x + bound) / (2 * bound)
    prefix = list(u.shape[:-1])
    flat = u.reshape(-1, u.shape[-1])
    S = float(np.log2(per_level_scale))                 # grid.py:38 (python float -> C float at the binding)
    out = _GridEncodeOracle.apply(flat, embeddings, offsets, S, base_resolution, level_scales)
    return out.view(prefix + [out.shape[-1]])


def sh_encode(d, degree=NGP_SH_DEGREE):
    """SHEncoder.forward, shencoder/sphere_harmonics.py:75-87 (view dirs never require grad on this path)."""
    prefix = list(d.shape[:-1])
    out, _ = sh_encode_forward(d.detach().reshape(-1, 3).numpy(), degree)
    return torch.from_numpy(out).view(prefix + [degree * degree])


def linear_layer(p, name, x, std_init=1.0, bias_init=0.0):
    """LinearLayer.forward :38-41."""
    return std_init * F.linear(x, p[name + ".weight"], p[name + ".bias"]) + bias_init


def film_siren(p, name, x, style):
    """FiLMSiren.forward :61-69; gamma = 15*Lin(w)+30, beta = 0.25*Lin(w) (:58-59)."""
    batch = style.shape[0]
    out = F.linear(x, p[name + ".weight"], p[name + ".bias"])
    gamma = linear_layer(p, name + ".gamma", style, 15.0, 30.0).view(batch, 1, 1, 1, -1)
    beta = linear_layer(p, name + ".beta", style, 0.25, 0.0).view(batch, 1, 1, 1, -1)
    return torch.sin(gamma * out + beta)


def field_ngp(p, pts, viewdirs, style, output_features=True, level_scales=None, grid_cfg=None):
    """NGPSIRENGenerator.forward :1566-1592.  pts/viewdirs [B,H,W,S,3] -> raw [B,H,W,S,3+1(+W)]."""
    cfg = dict(NGP_GRID if grid_cfg is None else grid_cfg)
    offsets = p["network.encoder.offsets"]
    _, pls = grid_offsets(**cfg)
    feat = hash_encode(pts, p["network.encoder.embeddings"], offsets, pls, cfg["base_resolution"], NGP_BOUND, level_scales)
    sh = sh_encode(viewdirs)
    h = linear_layer(p, "network.input_linear", feat)
    n_pts = len([k for k in p if k.startswith("network.pts_linears.") and k.endswith(".gamma.weight")])
    for i in range(n_pts):
        h = film_siren(p, f"network.pts_linears.{i}", h, style)
    sdf = linear_layer(p, "network.sigma_linear", h)
    hv = film_siren(p, "network.views_linears", torch.cat([h, sh], -1), style)
    rgb = linear_layer(p, "network.rgb_linear", hv)
    out = torch.cat([rgb, sdf], -1)
    if output_features:
        out = torch.cat([out, hv], -1)
    return out


def field_siren(p, pts, viewdirs, style, output_features=True):
    """SirenGenerator.forward :121-139 (--ngp 0 --fc 0)."""
    h = pts
    n_pts = len([k for k in p if k.startswith("network.pts_linears.") and k.endswith(".gamma.weight")])
    for i in range(n_pts):
        h = film_siren(p, f"network.pts_linears.{i}", h, style)
    sdf = linear_layer(p, "network.sigma_linear", h)
    hv = film_siren(p, "network.views_linears", torch.cat([h, viewdirs], -1), style)
    rgb = linear_layer(p, "network.rgb_linear", hv)
    out = torch.cat([rgb, sdf], -1)
    if output_features:
        out = torch.cat([out, hv], -1)
    return out


def _posenc(x, L):
    """FCGenerator.transform_points :1625-1638."""
    x = x / 2
    return torch.cat([torch.cat([torch.sin((2 ** i) * math.pi * x), torch.cos((2 ** i) * math.pi * x)], -1)
                      for i in range(L)], -1)


def field_fc(p, pts, viewdirs, style, output_features=True):
    """FCGenerator.forward :1640-1670 (--fc 1): posenc + ReLU MLP, style added once after the first layer."""
    h = F.linear(_posenc(pts, 10), p["network.x_in.weight"], p["network.x_in.bias"])
    s = F.linear(style, p["network.style_in.weight"], p["network.style_in.bias"])
    h = F.relu(h + s[:, None, None, None, :])
    n = len([k for k in p if k.startswith("network.pts_linears.") and k.endswith(".weight")])
    for i in range(n):
        h = F.relu(F.linear(h, p[f"network.pts_linears.{i}.weight"], p[f"network.pts_linears.{i}.bias"]))
    sdf = F.linear(h, p["network.sigma_linear.weight"], p["network.sigma_linear.bias"])
    hv = F.linear(torch.cat([h, _posenc(viewdirs, 4)], -1), p["network.views_linears.weight"], p["network.views_linears.bias"])
    rgb = F.linear(hv, p["network.rgb_linear.weight"], p["network.rgb_linear.bias"])
    out = torch.cat([rgb, sdf], -1)
    if output_features:
        out = torch.cat([out, hv], -1)
    return out


def get_rays(focal, c2w, res, static_viewdirs=False):
    """VolumeFeatureRenderer.get_rays :207-222 with the i/j buffers of :167-171 (i = column+0.5, j = row+0.5)."""
    lin = torch.linspace(0.5, res - 0.5, res)
    i = lin.view(1, 1, res).expand(1, res, res)            # varies along x (columns)
    j = lin.view(1, res, 1).expand(1, res, res)            # varies along y (rows)
    B = focal.shape[0]
    dirs = torch.stack([(i - res * .5) / focal, -(j - res * .5) / focal, -torch.ones(B, res, res)], -1)
    rays_d = torch.sum(dirs[..., None, :] * c2w[:, None, None, :3, :3], -1)
    rays_o = c2w[:, None, None, :3, -1].expand(rays_d.shape)
    viewdirs = dirs if static_viewdirs else rays_d
    viewdirs = viewdirs / torch.norm(viewdirs, dim=-1, keepdim=True)           # :367
    return rays_o, rays_d, viewdirs


def sample_depths(near, far, res, S, offset_sampling=True, t_rand=None):
    """render_rays :321-340.  near/far [B,1,1]; t_rand [B,res,res] (offset mode) or [B,res,res,S] (stratified) or None."""
    B = near.shape[0]
    near = near.view(B, 1, 1, 1).expand(B, res, res, 1)
    far = far.view(B, 1, 1, 1).expand(B, res, res, 1)
    if offset_sampling:
        t = torch.linspace(0., 1. - 1 / S, steps=S).view(1, 1, 1, -1)           # :175
    else:
        t = torch.linspace(0., 1., steps=S).view(1, 1, 1, -1)                   # :177
    z = near * (1. - t) + far * t                                               # :324
    if t_rand is not None:
        if offset_sampling:
            upper = torch.cat([z[..., 1:], far], -1)
            lower = z
            u = t_rand.unsqueeze(-1)
        else:
            mids = .5 * (z[..., 1:] + z[..., :-1])
            upper = torch.cat([mids, z[..., -1:]], -1)
            lower = torch.cat([z[..., :1], mids], -1)
            u = t_rand
        z = lower + (upper - lower) * u                                          # :340
    return z


def volume_integration(raw, z_vals, rays_d, pts, sigmoid_beta, *, with_sdf=True, output_features=True,
                       force_background=False, return_sdf=False, return_xyz=False, raw_noise=None, feature_dim=256):
    """volume_integration :236-301 (+ sdf_activation :231-234).  Shapes as in the reference (samples on dim 3)."""
    dists = z_vals[..., 1:] - z_vals[..., :-1]
    d_norm = torch.norm(rays_d.unsqueeze(3), dim=-1)
    dists = torch.cat([dists, torch.full_like(d_norm, 1e10)], -1) * d_norm      # :240-241
    if output_features:
        rgb, sdf, features = torch.split(raw, [3, 1, feature_dim], dim=-1)
    else:
        rgb, sdf = torch.split(raw, [3, 1], dim=-1)
        features = None
    if with_sdf:
        sigma = torch.sigmoid(-sdf / sigmoid_beta) / sigmoid_beta               # :232,255
        alpha = 1 - torch.exp(-sigma * dists.unsqueeze(-1))                     # :262
    else:
        noise = 0. if raw_noise is None else raw_noise
        alpha = 1 - torch.exp(-F.softplus(sdf + noise) * dists.unsqueeze(-1))   # :267
    trans = torch.cumprod(torch.cat([torch.ones_like(alpha[:, :, :, :1]), 1. - alpha + 1e-10], 3), 3)[:, :, :, :-1]
    weights = alpha * trans                                                     # :269-272
    if force_background:                                                        # :279-280
        weights = torch.cat([weights[:, :, :, :-1], 1 - weights[:, :, :, :-1].sum(3, keepdim=True)], 3)
    rgb_map = -1 + 2 * torch.sum(weights * torch.sigmoid(rgb), 3)               # :282
    feat_map = torch.sum(weights * features, 3) if output_features else None    # :285
    xyz = torch.sum(weights * pts, 3) if return_xyz else None                   # :295
    mask = weights[:, :, :, -1] if return_xyz else None                         # :296
    return rgb_map, feat_map, (sdf if return_sdf else None), mask, xyz


def render(p, cam_poses, focal, near, far, style, *, res=64, S=24, net_type="ngp", fc=False, t_rand=None,
           offset_sampling=True, z_normalize=True, static_viewdirs=False, with_sdf=True, output_features=True,
           force_background=False, return_sdf=False, return_xyz=False, return_eikonal=False, raw_noise=None,
           level_scales=None, return_raw=False, grid_cfg=None):
    """VolumeFeatureRenderer.forward :411-423 = render :363-378 + render_rays :310-361 + NCHW permutes.

    Returns (rgb [B,3,R,R], features [B,W,R,R]|None, sdf [B,R,R,S,1]|None, mask, xyz, eikonal [B,R,R,S,3]|None)
    and, if return_raw, additionally the dict of intermediates (pts, z_vals, raw, viewdirs).
    """
    rays_o, rays_d, viewdirs = get_rays(focal, cam_poses, res, static_viewdirs)
    z_vals = sample_depths(near, far, res, S, offset_sampling, t_rand)
    pts = rays_o.unsqueeze(3) + rays_d.unsqueeze(3) * z_vals.unsqueeze(-1)      # :343
    if return_eikonal:
        pts = pts.detach().requires_grad_(True)                                 # :345-346
    npts = pts * 2 / (far - near).view(-1, 1, 1, 1, 1) if z_normalize else pts  # :348-351
    vd = viewdirs.unsqueeze(3).expand(npts.shape)                               # :304
    if net_type == "ngp":
        raw = field_ngp(p, npts, vd, style, output_features, level_scales, grid_cfg)
        fdim = p["network.views_linears.weight"].shape[0]
    elif fc:
        raw = field_fc(p, npts, vd, style, output_features)
        fdim = p["network.views_linears.weight"].shape[0]
    else:
        raw = field_siren(p, npts, vd, style, output_features)
        fdim = p["network.views_linears.weight"].shape[0]
    sb = p.get("sigmoid_beta", None)
    rgb, feat, sdf, mask, xyz = volume_integration(
        raw, z_vals, rays_d, pts, sb, with_sdf=with_sdf, output_features=output_features,
        force_background=force_background, return_sdf=return_sdf, return_xyz=return_xyz, raw_noise=raw_noise,
        feature_dim=fdim)
    eik = None
    if return_eikonal and with_sdf:
        sdf_raw = raw[..., 3:4]
        eik = torch.autograd.grad(sdf_raw, pts, torch.ones_like(sdf_raw), create_graph=True)[0]   # :224-229
    rgb = rgb.permute(0, 3, 1, 2).contiguous()
    if feat is not None:
        feat = feat.permute(0, 3, 1, 2).contiguous()
    if xyz is not None:
        xyz = xyz.permute(0, 3, 1, 2).contiguous()
        mask = mask.permute(0, 3, 1, 2).contiguous()
    out = (rgb, feat, sdf, mask, xyz, eik)
    if return_raw:
        return out, dict(pts=pts, z_vals=z_vals, raw=raw, viewdirs=viewdirs, rays_d=rays_d, npts=npts)
    return out


def mapping(p_style, z):
    """Generator.style: 3x MappingLinear + fused_leaky_relu(scale=1) :437-461, sdf_op.py:106-117 (CPU branch)."""
    h = z
    for i in range(3):
        h = F.linear(h, p_style[f"{i}.weight"])
        h = F.leaky_relu(h + p_style[f"{i}.bias"].view(1, -1), negative_slope=0.2) * 1.0
    return h


def mlp_init_pass(p, cam_poses, focal, near, far, style, t_rand, *, res=64, S=24, net_type="ngp", fc=False,
                  z_normalize=True, static_viewdirs=False, offset_sampling=True, level_scales=None):
    """mlp_init_pass :380-409: stratified jitter (t_rand [B,R,R,S] injected), returns (sdf [B,R,R,S], target)."""
    rays_o, rays_d, viewdirs = get_rays(focal, cam_poses, res, static_viewdirs)
    B = near.shape[0]
    n = near.view(B, 1, 1, 1).expand(B, res, res, 1)
    f = far.view(B, 1, 1, 1).expand(B, res, res, 1)
    t = torch.linspace(0., 1. - 1 / S, steps=S) if offset_sampling else torch.linspace(0., 1., steps=S)
    z = n * (1. - t.view(1, 1, 1, -1)) + f * t.view(1, 1, 1, -1)
    mids = .5 * (z[..., 1:] + z[..., :-1])
    upper = torch.cat([mids, z[..., -1:]], -1)
    lower = torch.cat([z[..., :1], mids], -1)
    z = lower + (upper - lower) * t_rand
    pts = rays_o.unsqueeze(3) + rays_d.unsqueeze(3) * z.unsqueeze(-1)
    npts = pts * 2 / (far - near).view(-1, 1, 1, 1, 1) if z_normalize else pts
    vd = viewdirs.unsqueeze(3).expand(npts.shape)
    if net_type == "ngp":
        raw = field_ngp(p, npts, vd, style, True, level_scales)
    elif fc:
        raw = field_fc(p, npts, vd, style, True)
    else:
        raw = field_siren(p, npts, vd, style, True)
    sdf = raw[..., 3]
    target = pts.detach().norm(dim=-1) - ((far - near) / 4).view(-1, 1, 1, 1)
    return sdf, target


def align_volume(volume, near=0.88, far=1.12):
    """align_volume, sdf_utils.py:164-184: frustum -> box resampling of an sdf volume [b,h,w,d,c] for marching cubes.
    x/y sampling positions are stretched by linspace(far/near, 1, d) along depth, the volume is resampled trilinearly
    (grid_sample, align_corners=True, border padding) and cells whose stretched position leaves [-1, 1] are set to 1."""
    b, h, w, d, c = volume.shape
    ys, xs, zs = torch.meshgrid(torch.linspace(-1, 1, h), torch.linspace(-1, 1, w), torch.linspace(-1, 1, d), indexing="ij")
    stretch = torch.linspace(far / near, 1, d).view(1, 1, 1, d)
    pos = torch.stack([xs.unsqueeze(0) * stretch, ys.unsqueeze(0) * stretch, zs.unsqueeze(0).expand(1, h, w, d)], -1)   # [1,h,w,d,(x,y,z)]
    outside = ((pos < -1) | (pos > 1)).any(-1, keepdim=True)                                      # :174
    sample_at = pos.permute(0, 3, 1, 2, 4).contiguous().expand(b, d, h, w, 3)                     # :175  (grid_sample wants [N,D,H,W,3])
    vol = volume.permute(0, 4, 3, 1, 2).contiguous()                                              # :176  [b,c,d,h,w]
    res = F.grid_sample(vol, sample_at, padding_mode="border", align_corners=True)                # :177
    res = res.permute(0, 3, 4, 2, 1).contiguous()                                                 # :178  back to [b,h,w,d,c]
    res[outside.expand_as(res)] = 1                                                               # :181
    return res
