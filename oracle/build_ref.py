#!/usr/bin/env python
"""Build the UNMODIFIED reference CUDA extensions into oracle/_ref/ (test infrastructure only).

Compiles, from the sources where they lie under /root/reference (never copied into this repo):
  im2scene/sdf/models/gridencoder/src/{gridencoder.cu,bindings.cpp}  -> oracle/_ref/_gridencoder_ref.so
  im2scene/sdf/models/shencoder/src/{shencoder.cu,bindings.cpp}      -> oracle/_ref/_shencoder_ref.so
for sm_100a with nvcc + the torch headers.  The only deviation from the reference's own recipe
(gridencoder/backend.py:6-9) is -std=c++17 instead of -std=c++14, which torch >= 2.1 headers demand.

The resulting modules are the GPU-side ground truth for hash indices / trilinear weights / SH values.
They are imported ONLY by tests/ (and never by the product package).  oracle/_ref/ is git-ignored but
travels to the GPU box with the gpurun snapshot.  /root/reference does not exist on the GPU box, so this
script is a no-op there.
"""
import os
import subprocess
import sys
import sysconfig

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")
REF = os.environ.get("SDFGAN_REFERENCE_ROOT", "/root/reference")
MODELS = os.path.join(REF, "im2scene", "sdf", "models")

EXTS = {
    "_gridencoder_ref": [os.path.join(MODELS, "gridencoder", "src", f) for f in ("gridencoder.cu", "bindings.cpp")],
    "_shencoder_ref": [os.path.join(MODELS, "shencoder", "src", f) for f in ("shencoder.cu", "bindings.cpp")],
}


def build(name, sources, force=False):
    from torch.utils import cpp_extension as ce
    import torch

    out_so = os.path.join(OUT, name + ".so")
    if os.path.exists(out_so) and not force:
        return out_so
    if not all(os.path.exists(s) for s in sources):
        return None  # GPU box / no reference checkout: use the prebuilt file if any
    os.makedirs(OUT, exist_ok=True)
    inc = []
    for p in ce.include_paths("cuda") if "device_type" in ce.include_paths.__code__.co_varnames else ce.include_paths(True):
        inc += ["-I", p]
    inc += ["-I", sysconfig.get_paths()["include"]]
    abi = int(torch._C._GLIBCXX_USE_CXX11_ABI)
    common = [f"-DTORCH_EXTENSION_NAME={name}", "-DTORCH_API_INCLUDE_EXTENSION_H",
              f"-D_GLIBCXX_USE_CXX11_ABI={abi}"]
    objs = []
    for src in sources:
        obj = os.path.join(OUT, name + "_" + os.path.basename(src) + ".o")
        if src.endswith(".cu"):
            cmd = ["nvcc", "-c", src, "-o", obj, "-O3", "-std=c++17", "--threads", "8",
                   "-gencode", "arch=compute_100a,code=sm_100a",
                   "-U__CUDA_NO_HALF_OPERATORS__", "-U__CUDA_NO_HALF_CONVERSIONS__",
                   "-U__CUDA_NO_HALF2_OPERATORS__", "--expt-relaxed-constexpr",
                   "-Xcompiler", "-fPIC"] + common + inc
        else:
            cmd = ["g++", "-c", src, "-o", obj, "-O3", "-std=c++17", "-fPIC"] + common + inc
        print(" ".join(cmd), flush=True)
        subprocess.check_call(cmd)
        objs.append(obj)
    libdir = os.path.join(os.path.dirname(torch.__file__), "lib")
    cmd = ["g++", "-shared", "-o", out_so] + objs + [
        "-L", libdir, "-lc10", "-lc10_cuda", "-ltorch_cpu", "-ltorch_cuda", "-ltorch", "-ltorch_python",
        "-L", "/usr/local/cuda/lib64", "-lcudart", f"-Wl,-rpath,{libdir}"]
    print(" ".join(cmd), flush=True)
    subprocess.check_call(cmd)
    for o in objs:
        os.remove(o)
    return out_so


def main():
    force = "--force" in sys.argv
    for name, srcs in EXTS.items():
        print(name, "->", build(name, srcs, force))


if __name__ == "__main__":
    main()
