"""The reference's two pybind extension modules, served by libsdfg.so (operator-level drop-in).

The reference's Python wrappers call `_backend.grid_encode_forward/backward`, `_backend.grad_total_variation`
(/root/reference/im2scene/sdf/models/gridencoder/grid.py:53,84,183; bindings gridencoder/src/bindings.cpp:6-8, signatures
gridencoder.h:12-15) and `_backend.sh_encode_forward/backward` (shencoder/sphere_harmonics.py:33,54; shencoder.h:9-10) with torch
tensors and plain scalars.  `grid_backend` / `sh_backend` below take exactly those argument lists -- same order, same buffer
layouts ([L,B,C] features, caller-allocated outputs, pre-zeroed gradient sinks) -- and forward them to the C ABI on torch's
current stream.  A maintainer replaces the JIT `load(...)` in `gridencoder/backend.py` / `shencoder/backend.py` by

    from sdface_gan_b200.compat_backend import grid_backend as _backend      # resp. sh_backend

Errors surface as RuntimeError, like TORCH_CHECK does in the reference (gridencoder.cu:15-18).  float32 only (the reference's
half/double dispatch, gridencoder.cu:467,498, is never exercised by SDFace-GAN).  `dy_dx` is an opaque scratch tensor of
B*L*D*C floats that is only ever handed back to `grid_encode_backward`; its internal layout here is component-major.
"""
import ctypes

import torch

from . import _lib


def _p(t):
    return ctypes.c_void_p(0 if t is None else t.data_ptr())


def _st():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _dev(*ts):
    for t in ts:
        if t is not None:
            if not t.is_cuda:
                raise RuntimeError("compat_backend: tensors must be CUDA tensors (no CPU fallback)")
            if t.dtype not in (torch.float32, torch.int32):
                raise RuntimeError("compat_backend: float32 tensors (int32 offsets) only, got %s" % t.dtype)
            if not t.is_contiguous():
                raise RuntimeError("compat_backend: tensors must be contiguous")
    return torch.cuda.device(next(t for t in ts if t is not None).device)


class grid_backend:
    """stands in for the `_gridencoder` pybind module"""

    @staticmethod
    def grid_encode_forward(inputs, embeddings, offsets, outputs, B, D, C, L, S, H, dy_dx, gridtype, align_corners, interp=0):
        lib = _lib.load()
        with _dev(inputs, embeddings, offsets, outputs, dy_dx):
            _lib.check(lib.sdfg_grid_encode_forward(_p(inputs), _p(embeddings), _p(offsets), _p(outputs), B, D, C, L, float(S), int(H), 0.0,
                                                    _p(dy_dx), int(gridtype), int(bool(align_corners)), int(interp), _lib.LAYOUT_LNC, _st()),
                       "grid_encode_forward")

    @staticmethod
    def grid_encode_backward(grad, inputs, embeddings, offsets, grad_embeddings, B, D, C, L, S, H, dy_dx, grad_inputs, gridtype, align_corners,
                             interp=0):
        lib = _lib.load()
        with _dev(grad, inputs, embeddings, offsets, grad_embeddings, dy_dx, grad_inputs):
            _lib.check(lib.sdfg_grid_encode_backward(_p(grad), _p(inputs), _p(embeddings), _p(offsets), _p(grad_embeddings), B, D, C, L, float(S),
                                                     int(H), 0.0, _p(dy_dx), _p(grad_inputs), int(gridtype), int(bool(align_corners)), int(interp),
                                                     _lib.LAYOUT_LNC, _st()), "grid_encode_backward")

    @staticmethod
    def grad_total_variation(inputs, embeddings, grad, offsets, weight, B, D, C, L, S, H, gridtype, align_corners):
        lib = _lib.load()
        with _dev(inputs, embeddings, grad, offsets):
            _lib.check(lib.sdfg_grad_total_variation(_p(inputs), _p(embeddings), _p(grad), _p(offsets), float(weight), B, D, C, L, float(S), int(H),
                                                     int(gridtype), int(bool(align_corners)), _st()), "grad_total_variation")


class sh_backend:
    """stands in for the `_shencoder` pybind module"""

    @staticmethod
    def sh_encode_forward(inputs, outputs, B, D, C, dy_dx):
        lib = _lib.load()
        if D != 3:
            raise RuntimeError("sh_encode_forward: input_dim must be 3")
        with _dev(inputs, outputs, dy_dx):
            _lib.check(lib.sdfg_sh_encode_forward(_p(inputs), _p(outputs), B, int(C), _p(dy_dx), _st()), "sh_encode_forward")

    @staticmethod
    def sh_encode_backward(grad, inputs, B, D, C, dy_dx, grad_inputs):
        lib = _lib.load()
        with _dev(grad, inputs, dy_dx, grad_inputs):
            _lib.check(lib.sdfg_sh_encode_backward(_p(grad), _p(dy_dx), _p(grad_inputs), B, int(C), _st()), "sh_encode_backward")
