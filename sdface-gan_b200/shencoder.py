"""Spherical-harmonics direction encoder with the reference's module / function API.

Mirrors /root/reference/im2scene/sdf/models/shencoder/sphere_harmonics.py: `_sh_encoder` (:14-58), `sh_encode` (:62),
`SHEncoder` (:65-87), on top of csrc/sh.cu.
"""
import torch
import torch.nn as nn
from torch.autograd import Function

from . import ops


class _sh_encoder(Function):
    @staticmethod
    def forward(ctx, inputs, degree, calc_grad_inputs=False):
        inputs = inputs.contiguous().float()
        outputs, dy_dx = ops.sh_encode_forward(inputs, degree, calc_grad_inputs)
        ctx.save_for_backward(dy_dx)
        ctx.degree = degree
        return outputs

    @staticmethod
    def backward(ctx, grad):
        (dy_dx,) = ctx.saved_tensors
        if dy_dx is None:
            return None, None, None
        return ops.sh_encode_backward(grad.contiguous(), dy_dx, ctx.degree), None, None


sh_encode = _sh_encoder.apply


class SHEncoder(nn.Module):
    def __init__(self, input_dim=3, degree=4):
        super().__init__()
        self.input_dim = input_dim
        self.degree = degree
        self.output_dim = degree ** 2
        assert self.input_dim == 3, "SH encoder only support input dim == 3"
        assert self.degree > 0 and self.degree <= 8, "SH encoder only supports degree in [1, 8]"

    def __repr__(self):
        return f"SHEncoder: input_dim={self.input_dim} degree={self.degree}"

    def forward(self, inputs, size=1):
        inputs = inputs / size
        prefix_shape = list(inputs.shape[:-1])
        flat = inputs.reshape(-1, self.input_dim)
        outputs = sh_encode(flat, self.degree, flat.requires_grad)
        return outputs.reshape(prefix_shape + [self.output_dim])
