"""Import the UNMODIFIED reference renderer classes on a CPU-only box (build container only).

Recipe from SURVEY.md section 8(c):
  1. register empty namespace modules for im2scene / im2scene.sdf / im2scene.sdf.models (bypasses the package
     __init__ files, which pull in CUDA-only / missing third-party code);
  2. stub the off-path third-party imports (pytorch3d, trimesh, lmdb, skimage, configargparse, munch);
  3. make torch.utils.cpp_extension.load a no-op so sdf_op.py's CPU branches serve (sdf_op.py:106-117);
  4. provide `_gridencoder` / `_shencoder` backend modules whose functions run oracle/liboracle.so (the reference has
     no CPU implementation of those two extensions), so that the reference's OWN gridencoder/grid.py and
     shencoder/sphere_harmonics.py autograd wrappers and nn.Modules execute unchanged on CPU tensors.

Nothing here travels to the GPU box at run time: it is used only by make_golden.py to write tests/golden/*.npz.
"""
import os
import sys
import types

import numpy as np
import torch

REF = os.environ.get("SDFGAN_REFERENCE_ROOT", "/root/reference")
REPO = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


class Munch(dict):
    """10-line stand-in for munch.Munch (attribute access over a dict)."""

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e

    def __setattr__(self, k, v):
        self[k] = v


def _stub(name, **attrs):
    m = types.ModuleType(name)
    for k, v in attrs.items():
        setattr(m, k, v)
    sys.modules[name] = m
    return m


class _Anything:
    def __init__(self, *a, **k):
        pass

    def __call__(self, *a, **k):
        return _Anything()

    def __getattr__(self, k):
        return _Anything()


def install():
    if "im2scene.sdf.models.sdf_model" in sys.modules:
        return sys.modules["im2scene.sdf.models.sdf_model"]
    if REPO not in sys.path:
        sys.path.insert(0, REPO)
    import oracle

    # 1. namespace packages
    for name, rel in (("im2scene", "im2scene"), ("im2scene.sdf", "im2scene/sdf"), ("im2scene.sdf.models", "im2scene/sdf/models")):
        m = types.ModuleType(name)
        m.__path__ = [os.path.join(REF, rel)]
        sys.modules[name] = m
    # 2. third-party stubs
    _stub("trimesh")
    _stub("lmdb")
    _stub("skimage")
    _stub("skimage.measure", marching_cubes=_Anything())
    _stub("configargparse", ArgumentParser=__import__("argparse").ArgumentParser)
    _stub("munch", Munch=Munch, __all__=["Munch"])
    _stub("pytorch3d")
    _stub("pytorch3d.io")
    _stub("pytorch3d.structures", Meshes=_Anything)
    _stub("pytorch3d.transforms", matrix_to_euler_angles=_Anything())
    _stub("pytorch3d.renderer", **{k: _Anything for k in (
        "look_at_view_transform", "FoVPerspectiveCameras", "PointLights", "RasterizationSettings", "MeshRenderer",
        "MeshRasterizer", "SoftPhongShader", "TexturesVertex")})
    # 3. no JIT compilation of the decoder ops
    import torch.utils.cpp_extension as ce
    ce.load = lambda *a, **k: _Anything()

    # 4. CPU backends for the two CUDA-only extensions, same call signature as the pybind modules
    def grid_encode_forward(inputs, embeddings, offsets, outputs, B, D, C, L, S, H, dy_dx, gridtype, align_corners, interp):
        r = oracle.grid_encode_forward(inputs.detach().numpy(), embeddings.detach().numpy(), offsets.numpy(), float(np.float32(S)), H,
                                       calc_dy_dx=dy_dx is not None, gridtype=gridtype, align_corners=align_corners, interp=interp)
        outputs.copy_(torch.from_numpy(r["outputs"]))
        if dy_dx is not None:
            dy_dx.copy_(torch.from_numpy(r["dy_dx"]).reshape(dy_dx.shape))

    def grid_encode_backward(grad, inputs, embeddings, offsets, grad_embeddings, B, D, C, L, S, H, dy_dx, grad_inputs, gridtype, align_corners, interp):
        ge = grad_embeddings.numpy()
        _, gi = oracle.grid_encode_backward(grad.detach().numpy(), inputs.detach().numpy(), embeddings.detach().numpy(), offsets.numpy(),
                                            float(np.float32(S)), H, dy_dx=None if dy_dx is None else dy_dx.detach().numpy().reshape(B, L, D, C),
                                            gridtype=gridtype, align_corners=align_corners, interp=interp, grad_embeddings=ge)
        if grad_inputs is not None:
            grad_inputs.copy_(torch.from_numpy(gi))

    def grad_total_variation(inputs, embeddings, grad, offsets, weight, B, D, C, L, S, H, gridtype, align_corners):
        oracle.grad_total_variation(inputs.detach().numpy(), embeddings.detach().numpy(), grad.numpy(), offsets.numpy(), weight,
                                    float(np.float32(S)), H, gridtype=gridtype, align_corners=align_corners)

    _stub("_gridencoder", grid_encode_forward=grid_encode_forward, grid_encode_backward=grid_encode_backward,
          grad_total_variation=grad_total_variation)

    def sh_encode_forward(inputs, outputs, B, D, C, dy_dx):
        o, dd = oracle.sh_encode_forward(inputs.detach().numpy(), C, dy_dx is not None)
        outputs.copy_(torch.from_numpy(o))
        if dy_dx is not None:
            dy_dx.copy_(torch.from_numpy(dd).reshape(dy_dx.shape))

    def sh_encode_backward(grad, inputs, B, D, C, dy_dx, grad_inputs):
        grad_inputs.add_(torch.from_numpy(oracle.sh_encode_backward(grad.detach().numpy(), C, dy_dx.detach().numpy().reshape(B, 3, C * C))))

    _stub("_shencoder", sh_encode_forward=sh_encode_forward, sh_encode_backward=sh_encode_backward)

    import importlib
    return importlib.import_module("im2scene.sdf.models.sdf_model")


def default_opts(net_type="ngp", res=64, S=24, style_dim=256, size=256, **rendering_overrides):
    """model / rendering option trees with the defaults of SDFOptions (sdf_utils.py:447-594) as mutated by
    get_vol_render_opt (training_utils.py:144-193)."""
    model = Munch(size=size, style_dim=style_dim, channel_multiplier=2, n_mlp=8, lr_mapping=0.01,
                  renderer_spatial_output_dim=res, project_noise=False, freeze_renderer=False, psp=0,
                  no_viewpoint_loss=False)
    rendering = Munch(depth=8, width=256, no_sdf=False, no_z_normalize=False, static_viewdirs=False, N_samples=S,
                      no_offset_sampling=False, perturb=0., raw_noise_std=0., force_background=False, return_xyz=False,
                      return_sdf=False, type=net_type, fc=0)
    for k, v in rendering_overrides.items():
        rendering[k] = v
    return model, rendering
