"""CPU: the C-ABI library loads and exports every symbol include/sdfg.h declares (no compute calls)."""
import ctypes
import os
import re

import conftest  # noqa: F401

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "sdfg.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(sdfg_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    import sdface_gan_b200 as sg
    lib = sg._lib.load()
    names = _declared_symbols()
    assert len(names) >= 18
    for n in names:
        assert hasattr(lib, n), n
    assert set(names) == set(sg._lib.PROTOTYPES), set(names) ^ set(sg._lib.PROTOTYPES)
    assert lib.sdfg_version() >= 100


def test_error_reporting_without_gpu():
    """Argument validation happens before any CUDA call, so it can be exercised on a CPU-only box."""
    import sdface_gan_b200 as sg
    lib = sg._lib.load()
    rc = lib.sdfg_sh_encode_forward(None, None, 4, 9, None, None)
    assert rc == -2 and b"degree" in lib.sdfg_last_error()
    rc = lib.sdfg_grid_encode_forward(None, None, None, None, 1, 7, 2, 16, 0.5, 16, 0.0, None, 0, 0, 0, 0, None)
    assert rc == -2 and b"input_dim" in lib.sdfg_last_error()
    rc = lib.sdfg_composite_forward(None, None, None, None, None, None, None, None, 1, 24, 0, 1, 0, None, None, None, None, None, None)
    assert rc == -1


def test_cpu_tensors_are_rejected_loudly():
    import pytest
    import torch
    import sdface_gan_b200 as sg
    enc = sg.GridEncoder(num_levels=2, log2_hashmap_size=8)
    with pytest.raises(RuntimeError, match="CUDA"):
        enc(torch.zeros(4, 3))
    with pytest.raises(RuntimeError, match="CUDA"):
        sg.SHEncoder()(torch.zeros(4, 3))
