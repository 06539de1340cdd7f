"""oracle/ -- TEST INFRASTRUCTURE ONLY.

CPU restatement of the reference hot path, used as the checker by tests/,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs.
The product package (``sdface-gan_b200/``) never imports anything from here.

  hashgrid_oracle.c / sh_oracle.c   plain-C restatement of the CUDA-only encoders (gridencoder.cu, shencoder.cu)
  field_oracle.py                   torch-CPU restatement of rays / sampling / FiLM-SIREN field / volume integration
  build_ref.py                      compiles the UNMODIFIED reference CUDA extensions into oracle/_ref/ (GPU ground truth)

Parity status: the reference has no tests or golden vectors.  The pure-torch part of this oracle is
pinned by tests/golden/*.npz, which were produced by importing and running the reference's own Python classes in the
build container (tests/golden/make_golden.py); the C encoders are pinned on the GPU box against oracle/_ref.
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


def build(force=False):
    """Compile liboracle.so with gcc (seconds)."""
    so = os.path.join(_HERE, "liboracle.so")
    srcs = [os.path.join(_HERE, f) for f in ("hashgrid_oracle.c", "sh_oracle.c")]
    if force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.check_call(["make", "-C", _HERE, "-B", "liboracle.so"], stdout=subprocess.DEVNULL)
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = ctypes.CDLL(build())
    return _LIB


def _p(a, ty=ctypes.c_float):
    if a is None:
        return None
    assert a.flags["C_CONTIGUOUS"]
    return a.ctypes.data_as(ctypes.POINTER(ty))


def cuda_level_scales(L, S, H):
    """The per-level scale table as the CUDA device computes it (the reference evaluates `exp2f(level*S)*H - 1` on the GPU,
    gridencoder.cu:138, and CUDA's exp2f is 1 ulp off libm's on some levels), from the values recorded on a B200 in
    oracle/cuda_level_scales.json.  Returns None for configurations that were never recorded."""
    import json
    tabs = json.load(open(os.path.join(_HERE, "cuda_level_scales.json")))["tables"]
    s_hex = int(np.float32(S).view(np.uint32))
    for t in tabs:
        if t["S_hex"] == s_hex and t["H"] == H and t["L"] >= L:
            return np.array(t["scale_hex"][:L], np.uint32).view(np.float32).copy()
    return None


def grid_level_scales(L, S, H):
    out = np.empty(L, np.float32)
    lib().oracle_grid_level_scales(ctypes.c_uint32(L), ctypes.c_float(S), ctypes.c_uint32(H), _p(out))
    return out


def grid_encode_forward(inputs, embeddings, offsets, S, H, calc_dy_dx=False, gridtype=0, align_corners=False,
                        interp=0, level_scales=None, want_corners=False):
    """inputs [B,D] f32 in [0,1]; embeddings [sO,C] f32; offsets [L+1] i32.
    Returns dict(outputs [L,B,C], dy_dx [B,L,D,C]|None, corner_idx [B,L,2^D]|None, corner_w|None)."""
    inputs = np.ascontiguousarray(inputs, np.float32)
    embeddings = np.ascontiguousarray(embeddings, np.float32)
    offsets = np.ascontiguousarray(offsets, np.int32)
    B, D = inputs.shape
    C = embeddings.shape[1]
    L = offsets.shape[0] - 1
    out = np.empty((L, B, C), np.float32)
    dy_dx = np.empty((B, L, D, C), np.float32) if calc_dy_dx else None
    cidx = np.empty((B, L, 1 << D), np.uint32) if want_corners else None
    cw = np.empty((B, L, 1 << D), np.float32) if want_corners else None
    if level_scales is None:
        level_scales = cuda_level_scales(L, S, H)
    ls = None if level_scales is None else np.ascontiguousarray(level_scales, np.float32)
    lib().oracle_grid_encode_forward(
        _p(inputs), _p(embeddings), _p(offsets, ctypes.c_int), _p(out),
        ctypes.c_uint32(B), ctypes.c_uint32(D), ctypes.c_uint32(C), ctypes.c_uint32(L), ctypes.c_float(S),
        ctypes.c_uint32(H), _p(dy_dx), ctypes.c_uint32(gridtype), ctypes.c_int(int(align_corners)),
        ctypes.c_uint32(interp), _p(ls), _p(cidx, ctypes.c_uint32), _p(cw))
    return dict(outputs=out, dy_dx=dy_dx, corner_idx=cidx, corner_w=cw)


def grid_encode_backward(grad, inputs, embeddings, offsets, S, H, dy_dx=None, gridtype=0, align_corners=False,
                         interp=0, level_scales=None, grad_embeddings=None):
    """grad [L,B,C].  Returns (grad_embeddings [sO,C] (accumulated into if given), grad_inputs [B,D]|None)."""
    grad = np.ascontiguousarray(grad, np.float32)
    inputs = np.ascontiguousarray(inputs, np.float32)
    embeddings = np.ascontiguousarray(embeddings, np.float32)
    offsets = np.ascontiguousarray(offsets, np.int32)
    B, D = inputs.shape
    C = embeddings.shape[1]
    L = offsets.shape[0] - 1
    if grad_embeddings is None:
        grad_embeddings = np.zeros_like(embeddings)
    gi = np.zeros((B, D), np.float32) if dy_dx is not None else None
    dd = None if dy_dx is None else np.ascontiguousarray(dy_dx, np.float32)
    if level_scales is None:
        level_scales = cuda_level_scales(L, S, H)
    ls = None if level_scales is None else np.ascontiguousarray(level_scales, np.float32)
    lib().oracle_grid_encode_backward(
        _p(grad), _p(inputs), _p(embeddings), _p(offsets, ctypes.c_int), _p(grad_embeddings),
        ctypes.c_uint32(B), ctypes.c_uint32(D), ctypes.c_uint32(C), ctypes.c_uint32(L), ctypes.c_float(S),
        ctypes.c_uint32(H), _p(dd), _p(gi), ctypes.c_uint32(gridtype), ctypes.c_int(int(align_corners)),
        ctypes.c_uint32(interp), _p(ls))
    return grad_embeddings, gi


def grad_total_variation(inputs, embeddings, grad, offsets, weight, S, H, gridtype=0, align_corners=False,
                         level_scales=None):
    inputs = np.ascontiguousarray(inputs, np.float32)
    embeddings = np.ascontiguousarray(embeddings, np.float32)
    offsets = np.ascontiguousarray(offsets, np.int32)
    assert grad.dtype == np.float32 and grad.flags["C_CONTIGUOUS"]
    B, D = inputs.shape
    C = embeddings.shape[1]
    L = offsets.shape[0] - 1
    if level_scales is None:
        level_scales = cuda_level_scales(L, S, H)
    ls = None if level_scales is None else np.ascontiguousarray(level_scales, np.float32)
    lib().oracle_grad_total_variation(
        _p(inputs), _p(embeddings), _p(grad), _p(offsets, ctypes.c_int), ctypes.c_float(weight),
        ctypes.c_uint32(B), ctypes.c_uint32(D), ctypes.c_uint32(C), ctypes.c_uint32(L), ctypes.c_float(S),
        ctypes.c_uint32(H), ctypes.c_uint32(gridtype), ctypes.c_int(int(align_corners)), _p(ls))
    return grad


def sh_encode_forward(inputs, degree, calc_dy_dx=False):
    inputs = np.ascontiguousarray(inputs, np.float32)
    B = inputs.shape[0]
    out = np.empty((B, degree * degree), np.float32)
    dy_dx = np.empty((B, 3, degree * degree), np.float32) if calc_dy_dx else None
    lib().oracle_sh_encode_forward(_p(inputs), _p(out), ctypes.c_uint32(B), ctypes.c_uint32(degree), _p(dy_dx))
    return out, dy_dx


def sh_encode_backward(grad, degree, dy_dx):
    grad = np.ascontiguousarray(grad, np.float32)
    dy_dx = np.ascontiguousarray(dy_dx, np.float32)
    B = grad.shape[0]
    gi = np.zeros((B, 3), np.float32)
    lib().oracle_sh_encode_backward(_p(grad), ctypes.c_uint32(B), ctypes.c_uint32(degree), _p(dy_dx), _p(gi))
    return gi


def grid_offsets(input_dim=3, num_levels=16, level_dim=2, per_level_scale=2, base_resolution=16,
                 log2_hashmap_size=19, desired_resolution=None, align_corners=False):
    """Level table exactly as gridencoder/grid.py:97-131 builds it.  Returns (offsets int32 [L+1], per_level_scale)."""
    if desired_resolution is not None:
        per_level_scale = np.exp2(np.log2(desired_resolution / base_resolution) / (num_levels - 1))
    offsets, offset = [], 0
    max_params = 2 ** log2_hashmap_size
    for i in range(num_levels):
        resolution = int(np.ceil(base_resolution * per_level_scale ** i))
        n = min(max_params, (resolution if align_corners else resolution + 1) ** input_dim)
        n = int(np.ceil(n / 8) * 8)
        offsets.append(offset)
        offset += n
    offsets.append(offset)
    return np.array(offsets, dtype=np.int32), per_level_scale
