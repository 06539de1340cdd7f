"""ctypes binding of lib/libsdfg.so (the C-ABI declared in include/sdfg.h).

The product path has NO fallback: if the shared library is missing it is built in-tree with nvcc (``_build.py``); if that is
impossible, or a call returns non-zero, a RuntimeError is raised -- as TORCH_CHECK does in the reference extensions
(/root/reference/im2scene/sdf/models/gridencoder/src/gridencoder.cu:15-18).
"""
import ctypes
import os
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libsdfg.so")

SDFG_MAX_FILM = 9
LAYOUT_NLC, LAYOUT_LNC = 0, 1
PRECISION_FP32, PRECISION_TC16 = 0, 1
BWD_CHAIN, BWD_WGRAD, BWD_BOTH = 1, 2, 3
ERR_INVALID, ERR_UNSUPPORTED, ERR_CUDA = -1, -2, -3

c_f = ctypes.POINTER(ctypes.c_float)
vp = ctypes.c_void_p
u32, u64, i32, f32 = ctypes.c_uint32, ctypes.c_uint64, ctypes.c_int, ctypes.c_float


class FieldParams(ctypes.Structure):
    """sdfg_field_params (include/sdfg.h)."""
    _fields_ = [("width", u32), ("in_dim", u32), ("view_dim", u32), ("n_film", u32), ("has_input_linear", u32),
                ("samples_per_image", u32), ("samples_per_ray", u32), ("reserved", u32),
                ("input_w", vp), ("input_b", vp),
                ("film_w", vp * SDFG_MAX_FILM), ("film_b", vp * SDFG_MAX_FILM),
                ("gamma", vp), ("beta", vp), ("sigma_w", vp), ("sigma_b", vp), ("rgb_w", vp), ("rgb_b", vp)]


class FieldGrads(ctypes.Structure):
    """sdfg_field_grads (include/sdfg.h)."""
    _fields_ = [("input_w", vp), ("input_b", vp),
                ("film_w", vp * SDFG_MAX_FILM), ("film_b", vp * SDFG_MAX_FILM),
                ("gamma", vp), ("beta", vp), ("sigma_w", vp), ("sigma_b", vp), ("rgb_w", vp), ("rgb_b", vp)]


# name -> (restype, argtypes); every symbol include/sdfg.h declares
PROTOTYPES = {
    "sdfg_last_error": (ctypes.c_char_p, []),
    "sdfg_version": (i32, []),
    "sdfg_launch_count": (ctypes.c_int64, []),
    "sdfg_launch_count_reset": (None, []),
    "sdfg_prof_enable": (None, [i32, ctypes.c_char_p]),
    "sdfg_prof_collect": (i32, [ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_int64)]),
    "sdfg_sample_rays": (i32, [vp, vp, vp, vp, vp, vp, i32, i32, i32, u32, u32, u32, vp, vp, vp, vp, vp, vp]),
    "sdfg_grid_encode_forward": (i32, [vp, vp, vp, vp, u32, u32, u32, u32, f32, u32, f32, vp, u32, i32, u32, i32, vp]),
    "sdfg_grid_encode_backward": (i32, [vp, vp, vp, vp, vp, u32, u32, u32, u32, f32, u32, f32, vp, vp, u32, i32, u32, i32, vp]),
    "sdfg_grad_total_variation": (i32, [vp, vp, vp, vp, f32, u32, u32, u32, u32, f32, u32, u32, i32, vp]),
    "sdfg_grid_level_scales": (i32, [vp, u32, f32, u32, vp]),
    "sdfg_grid_corner_indices": (i32, [vp, vp, vp, vp, u32, u32, u32, u32, f32, u32, f32, u32, i32, vp]),
    "sdfg_l2_gather_probe": (i32, [vp, u32, u32, u32, vp, vp]),
    "sdfg_sh_encode_forward": (i32, [vp, vp, u32, u32, vp, vp]),
    "sdfg_sh_encode_backward": (i32, [vp, vp, vp, u32, u32, vp]),
    "sdfg_field_workspace_bytes": (u64, [ctypes.POINTER(FieldParams), u64, i32, i32]),
    "sdfg_field_backward_scratch_bytes": (u64, [ctypes.POINTER(FieldParams), u64, i32]),
    "sdfg_field_forward": (i32, [ctypes.POINTER(FieldParams), vp, vp, u64, vp, vp, vp, vp, i32, i32, vp]),
    "sdfg_field_forward_h": (i32, [ctypes.POINTER(FieldParams), vp, vp, u64, vp, vp, vp, vp, vp]),
    "sdfg_field_backward": (i32, [ctypes.POINTER(FieldParams), ctypes.POINTER(FieldGrads), vp, vp, u64, vp, vp, vp, vp, vp, vp,
                                  vp, i32, vp]),
    "sdfg_field_backward_phase": (i32, [ctypes.POINTER(FieldParams), ctypes.POINTER(FieldGrads), vp, vp, u64, vp, vp, vp, vp, vp, vp,
                                        vp, i32, i32, vp]),
    "sdfg_field_eikonal": (i32, [ctypes.POINTER(FieldParams), vp, vp, u64, vp, vp, vp, vp, u32, u32, ctypes.c_float, vp, i32, vp]),
    "sdfg_tc_linear_probe_workspace_bytes": (u64, [u32, u32, u32]),
    "sdfg_tc_linear_probe": (i32, [vp, vp, vp, u32, u32, u32, vp, vp]),
    "sdfg_tc_wgrad_probe": (i32, [vp, vp, vp, u32, u32, u32, u32, ctypes.POINTER(u32), ctypes.POINTER(u32), vp, vp]),
    "sdfg_composite_forward": (i32, [vp, vp, vp, vp, vp, vp, vp, vp, u64, u32, u32, i32, i32, vp, vp, vp, vp, vp, vp]),
    "sdfg_composite_forward_h": (i32, [vp, vp, vp, vp, vp, vp, vp, vp, u64, u32, u32, i32, i32, vp, vp, vp, vp, vp, vp]),
    "sdfg_nhwc16": (i32, [vp, vp, u64, vp]),
    "sdfg_modconv_fold": (i32, [vp, vp, f32, u32, u32, u32, u32, i32, vp, vp, vp]),
    "sdfg_conv_forward": (i32, [vp, vp, u32, u32, u32, u32, u32, u32, vp, vp, vp, vp, vp]),
    "sdfg_upconv_forward": (i32, [vp, vp, u32, u32, u32, u32, u32, vp, vp, vp, vp, vp, vp]),
    "sdfg_to_rgb": (i32, [vp, vp, vp, f32, vp, vp, u32, u32, u32, u32, vp, vp, vp, vp]),
    "sdfg_align_volume": (i32, [vp, vp, u32, u32, u32, u32, u32, f32, f32, vp]),
    "sdfg_composite_backward": (i32, [vp, vp, vp, vp, vp, vp, vp, vp, u64, u32, u32, i32, i32, vp, vp, vp, vp, vp, vp, vp, vp,
                                      vp, vp]),
}

_lib = None
_lock = threading.Lock()


def _declare(lib):
    for name, (res, args) in PROTOTYPES.items():
        fn = getattr(lib, name)      # AttributeError if the library does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    return lib


def _build_module():
    import importlib.util
    spec = importlib.util.spec_from_file_location("_sdfg_build", os.path.join(_HERE, "_build.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def load(build_if_missing=True):
    """Return the loaded library.  A missing OR STALE binary (lib/libsdfg.stamp does not match the hash of csrc/, include/sdfg.h
    and the build flags) is rebuilt in-tree first -- under a file lock, so the ranks of one job build it once -- and refused
    with a RuntimeError when that is impossible (no nvcc) or not allowed (build_if_missing=False)."""
    global _lib
    with _lock:
        if _lib is not None:
            return _lib
        mod = _build_module()
        if mod.is_stale():
            if not build_if_missing:
                raise RuntimeError("libsdfg.so is missing or stale (%s); run `python sdface-gan_b200/_build.py`" % LIB_PATH)
            mod.build()
        try:
            _lib = _declare(ctypes.CDLL(LIB_PATH))
        except OSError as e:
            raise RuntimeError("cannot load %s: %s (no CPU / eager fallback exists for this path)" % (LIB_PATH, e)) from e
        return _lib


def check(code, what):
    if code != 0:
        msg = load().sdfg_last_error()
        raise RuntimeError("%s failed (%d): %s" % (what, code, msg.decode() if msg else "?"))


def launch_count():
    return int(load().sdfg_launch_count())


def launch_count_reset():
    load().sdfg_launch_count_reset()


def prof_enable(on, tag=""):
    load().sdfg_prof_enable(int(bool(on)), tag.encode())


def prof_collect():
    """-> (total milliseconds, launches) of the profiled kernels since the last collect."""
    ms, n = ctypes.c_double(0), ctypes.c_int64(0)
    load().sdfg_prof_collect(ctypes.byref(ms), ctypes.byref(n))
    return ms.value, n.value
