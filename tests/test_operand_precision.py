"""CPU evidence for the operand format of the tensor-core path: fp16, not the bf16 the north star mentions.

The tcgen05 kernels multiply 16-bit operands and accumulate in fp32.  This test emulates exactly that on the oracle (operands of
every per-sample `F.linear` rounded to the 16-bit type, fp32 accumulation) on a reference-generated fixture:
  * bf16 operands put the rendered features outside the north star's 2e-2 relative band (gamma ~ 30 amplifies the 2^-9
    pre-activation rounding) -- recorded as a strict expected failure;
  * fp16 operands (|sin| <= 1, SIREN weights << 1, so range is no concern) stay well inside it.
"""
import numpy as np
import pytest
import torch

import helpers as H
from oracle import field_oracle as fo


def _features_with_operands(dtype, name):
    z = H.load_fixture(name)
    params = H.fixture_params(z)
    rp, sp = H.oracle_param_dicts(params)
    inp = H.fixture_inputs(z)
    style = fo.mapping(sp, inp["z"])
    orig = fo.F.linear

    def rounded(x, w, b=None):
        if dtype is None or x.dim() <= 2:          # [B, style_dim] gamma/beta heads stay fp32 (torch GEMMs in the product, too)
            return orig(x, w, b)
        return orig(x.to(dtype).float(), w.to(dtype).float(), b)

    fo.F.linear = rounded
    try:
        with torch.no_grad():
            rgb, feat, sdf, _, _, _ = fo.render(rp, inp["cam"], inp["focal"], inp["near"], inp["far"], style, t_rand=inp["t_rand"],
                                                **dict(H.render_kwargs(H.fixture_cfg(z)), return_sdf=True))
    finally:
        fo.F.linear = orig
    return z, rgb, feat, sdf


@pytest.mark.parametrize("name", ["ngp_fwd_tab1", "ngp_fwd_init"])
def test_fp16_operands_meet_the_band(name):
    """measured: 0.24 % (table U(-1,1)) / 0.55 % (reference init U(-1e-4,1e-4)) relative feature error"""
    z, rgb, feat, sdf = _features_with_operands(torch.float16, name)
    assert H.rel_err(feat, z["features"]) < 1e-2
    assert H.max_abs(rgb, z["out_thumb_rgb"]) < 1e-2


@pytest.mark.xfail(strict=True, reason="bf16 operands: gamma ~ 30 amplifies the 8-bit mantissa's pre-activation rounding beyond 2e-2 relative")
def test_bf16_operands_meet_the_band():
    """measured: 3.5 % relative feature error with the reference's table init (1.9 % with a U(-1,1) table)"""
    z, rgb, feat, sdf = _features_with_operands(torch.bfloat16, "ngp_fwd_init")
    assert H.rel_err(feat, z["features"]) < 2e-2
