"""In-tree build of libsdfg.so (the C-ABI CUDA library) with nvcc for sm_100a.

    python sdface-gan_b200/_build.py [--force]

Objects are cached under csrc/build/ keyed by a HASH of (source, every header, compiler flags, nvcc version) -- a debug build
(`SDFG_BUILD_DEFS=-DSDFG_CHAIN_DEBUG`) or an edited header can therefore never be mistaken for the default objects.  The shared
library lands in sdface-gan_b200/lib/libsdfg.so together with lib/libsdfg.stamp (the hash of everything it was built from), so
that `_lib.load()` can refuse a stale binary; it travels with the repo snapshot to the GPU box (a JIT cache under ~/.cache would
not).  The whole build runs under an exclusive file lock and links to a temporary file that is renamed into place: N ranks
importing the package at once build it exactly once.  No torch headers are involved: the library's boundary is plain C
(include/sdfg.h).
"""
import fcntl
import hashlib
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(CSRC, "build")
LIBDIR = os.path.join(HERE, "lib")
LIB = os.path.join(LIBDIR, "libsdfg.so")
STAMP = os.path.join(LIBDIR, "libsdfg.stamp")

NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
BASE_FLAGS = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr", "-Xptxas", "-v"]


def _flags():
    # e.g. SDFG_BUILD_DEFS=-DSDFG_CHAIN_DEBUG for the chain kernels' event log: part of the cache key, no --force needed
    return BASE_FLAGS + os.environ.get("SDFG_BUILD_DEFS", "").split()


def _sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def _headers():
    hs = sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h")))
    hs.append(os.path.join(os.path.dirname(HERE), "include", "sdfg.h"))
    return hs


def _file_digest(path):
    with open(path, "rb") as f:
        return hashlib.sha256(f.read()).hexdigest()


_NVCC_VERSION = None


def _nvcc_version():
    global _NVCC_VERSION
    if _NVCC_VERSION is None:
        try:
            _NVCC_VERSION = subprocess.run([NVCC, "--version"], capture_output=True, text=True).stdout.strip().splitlines()[-1]
        except Exception:
            _NVCC_VERSION = "unknown"
    return _NVCC_VERSION


def _common_key(with_compiler=True):
    h = hashlib.sha256()
    for p in _headers():
        h.update(os.path.basename(p).encode())
        h.update(_file_digest(p).encode())
    h.update(" ".join(ARCH + _flags()).encode())
    if with_compiler:
        h.update(_nvcc_version().encode())
    return h.hexdigest()


def source_stamp():
    """Hash of everything libsdfg.so is built from (sources, headers, flags) -- compiler version excluded, so that a box without
    nvcc can still verify a travelling binary against the sources lying next to it."""
    h = hashlib.sha256(_common_key(with_compiler=False).encode())
    for s in _sources():
        h.update(os.path.basename(s).encode())
        h.update(_file_digest(s).encode())
    return h.hexdigest()


def is_stale():
    """True when lib/libsdfg.so is missing or was built from different sources / flags than the ones in the tree."""
    if not (os.path.exists(LIB) and os.path.exists(STAMP)):
        return True
    with open(STAMP) as f:
        return f.read().strip() != source_stamp()


def _compile(src, common, force):
    key = hashlib.sha256((common + _file_digest(src)).encode()).hexdigest()[:16]
    base = os.path.basename(src)[:-3]
    obj = os.path.join(OBJ, "%s.%s.o" % (base, key))
    if not force and os.path.exists(obj):
        return obj, ""
    for old in os.listdir(OBJ):                      # drop objects of this source built from other contents / flags
        if old.startswith(base + ".") and old.endswith(".o"):
            os.unlink(os.path.join(OBJ, old))
    tmp = obj + ".tmp%d" % os.getpid()
    cmd = [NVCC, "-c", src, "-o", tmp] + ARCH + _flags()
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed for %s:\n%s\n%s" % (src, r.stdout, r.stderr))
    os.replace(tmp, obj)
    return obj, r.stderr


def build(force=False, verbose=False):
    """Compile every csrc/*.cu and link lib/libsdfg.so.  Returns the library path."""
    if not os.path.exists(NVCC):
        if os.path.exists(LIB) and not is_stale():
            return LIB
        raise RuntimeError("nvcc not found at %s and no up-to-date prebuilt %s" % (NVCC, LIB))
    os.makedirs(OBJ, exist_ok=True)
    os.makedirs(LIBDIR, exist_ok=True)
    with open(os.path.join(LIBDIR, ".build.lock"), "w") as lock:
        fcntl.flock(lock, fcntl.LOCK_EX)             # one builder at a time (ranks of one job, pytest-xdist workers ...)
        try:
            if not force and not is_stale():
                return LIB
            srcs = _sources()
            common = _common_key()
            with ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
                res = list(ex.map(lambda s: _compile(s, common, force), srcs))
            objs = [o for o, _ in res]
            if verbose:
                for _, log in res:
                    if log:
                        sys.stderr.write(log)
            tmp = LIB + ".tmp%d" % os.getpid()
            cmd = [NVCC, "-shared", "-o", tmp] + objs + ARCH + ["-lcudart"]
            r = subprocess.run(cmd, capture_output=True, text=True)
            if r.returncode != 0:
                raise RuntimeError("link failed:\n%s\n%s" % (r.stdout, r.stderr))
            os.replace(tmp, LIB)
            with open(STAMP + ".tmp", "w") as f:
                f.write(source_stamp() + "\n")
            os.replace(STAMP + ".tmp", STAMP)
        finally:
            fcntl.flock(lock, fcntl.LOCK_UN)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
