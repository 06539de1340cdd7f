"""Host-side helpers of the SDF generator path: camera sampling and option trees.

`generate_camera_params` follows /root/reference/im2scene/sdf/models/sdf_utils.py:97-159 (same arguments, same outputs);
it is ~25 tiny torch launches on [B,1] tensors and stays in torch.
"""
import math

import torch
import torch.nn.functional as F


class Munch(dict):
    """Attribute-access dict, the subset of `munch.Munch` the option trees need (SDFOptions, ref sdf_utils.py:447-594)."""

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e

    def __setattr__(self, k, v):
        self[k] = v


def default_options(net_type="ngp", size=256, renderer_res=64, n_samples=24, style_dim=256, **rendering_overrides):
    """(model, rendering) option trees with the defaults of SDFOptions (ref sdf_utils.py:447-594) as mutated by
    get_vol_render_opt (ref im2scene/training_utils.py:144-193)."""
    model = Munch(size=size, style_dim=style_dim, channel_multiplier=2, n_mlp=8, lr_mapping=0.01, renderer_spatial_output_dim=renderer_res,
                  project_noise=False, freeze_renderer=False, psp=0, no_viewpoint_loss=False)
    rendering = Munch(depth=8, width=256, no_sdf=False, no_z_normalize=False, static_viewdirs=False, N_samples=n_samples,
                      no_offset_sampling=False, perturb=1., raw_noise_std=0., force_background=False, return_xyz=False, return_sdf=False,
                      type=net_type, fc=0)
    for k, v in rendering_overrides.items():
        rendering[k] = v
    return model, rendering


def generate_camera_params(resolution, device, batch=1, locations=None, sweep=False, uniform=False, azim_range=0.3, elev_range=0.15,
                           fov_ang=6, dist_radius=0.12):
    """-> (extrinsics [B,3,4] camera-to-world, focal [B,1,1], near [B,1,1], far [B,1,1], viewpoint [B,2])."""
    if locations is not None:
        azim = locations[:, 0].view(-1, 1)
        elev = locations[:, 1].view(-1, 1)
        n = azim.shape[0]
    elif sweep:
        azim = (-azim_range + (2 * azim_range / 7) * torch.arange(8, device=device)).view(-1, 1).repeat(batch, 1)
        elev = (-elev_range + 2 * elev_range * torch.rand(batch, 1, device=device).repeat(1, 8).view(-1, 1))
        n = batch * 8
    else:
        if uniform:
            azim = -azim_range + 2 * azim_range * torch.rand(batch, 1, device=device)
            elev = -elev_range + 2 * elev_range * torch.rand(batch, 1, device=device)
        else:
            azim = azim_range * torch.randn(batch, 1, device=device)
            elev = elev_range * torch.randn(batch, 1, device=device)
        n = batch
    dist = torch.ones(n, 1, device=device)                       # cameras sit on the unit sphere
    near, far = (dist - dist_radius).unsqueeze(-1), (dist + dist_radius).unsqueeze(-1)
    fov_angle = fov_ang * torch.ones(n, 1, device=device) * math.pi / 180
    focal = 0.5 * resolution / torch.tan(fov_angle).unsqueeze(-1)
    viewpoint = torch.cat([azim, elev], 1)

    camera_dir = torch.stack([torch.cos(elev) * torch.sin(azim), torch.sin(elev), torch.cos(elev) * torch.cos(azim)], dim=1).view(-1, 3)
    camera_loc = dist * camera_dir
    up = torch.tensor([[0., 1., 0.]], device=device) * torch.ones_like(dist)
    z_axis = F.normalize(camera_dir, eps=1e-5)                   # -z points into the screen
    x_axis = F.normalize(torch.cross(up, z_axis, dim=1), eps=1e-5)
    y_axis = F.normalize(torch.cross(z_axis, x_axis, dim=1), eps=1e-5)
    degenerate = torch.isclose(x_axis, torch.tensor(0.0, device=device), atol=5e-3).all(dim=1, keepdim=True)
    if degenerate.any():
        x_axis = torch.where(degenerate, F.normalize(torch.cross(y_axis, z_axis, dim=1), eps=1e-5), x_axis)
    R = torch.stack((x_axis, y_axis, z_axis), dim=1)
    extrinsics = torch.cat((R.transpose(1, 2), camera_loc[:, :, None]), -1)
    return extrinsics, focal, near, far, viewpoint


def align_volume(volume, near=0.88, far=1.12):
    """Frustum -> box resampling of the renderer's sdf volume [B,H,W,D,C] on the GPU (ref align_volume sdf_utils.py:164-184, same
    signature; the reference runs torch.meshgrid + grid_sample on whatever device the volume is on -- sdf_mesh.py moves it to the
    CPU first, :152).  CUDA tensors only."""
    from . import ops
    return ops.align_volume(volume.contiguous().float(), near, far)
