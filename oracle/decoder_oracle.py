"""oracle/decoder_oracle.py -- TEST INFRASTRUCTURE ONLY (never imported by the product).

Torch-CPU, functional restatement of the reference's StyleGAN2 decoder forward.  Citations are
/root/reference/im2scene/sdf/models/sdf_model.py unless noted.  Parameters come as a flat dict keyed like the reference
``Decoder.state_dict()`` (``conv1.conv.weight``, ``to_rgbs.0.conv.modulation.bias`` ...).

Pinned by tests/golden/decoder.npz, produced by running the reference's own ``Decoder`` (tests/golden/make_golden.py).
"""
import math

import torch
import torch.nn.functional as F


def _fused_lrelu(x, bias, scale=2 ** 0.5):
    """fused_leaky_relu, CPU branch sdf_op.py:106-117: leaky_relu(x + bias, 0.2) * scale, bias along dim 1"""
    shape = [1, -1] + [1] * (x.ndim - 2)
    return F.leaky_relu(x + bias.view(shape), negative_slope=0.2) * scale


def equal_linear(p, name, x, lr_mul=1.0, activation=False):
    """EqualLinear.forward :596-606"""
    w = p[name + ".weight"]
    scale = (1 / math.sqrt(w.shape[1])) * lr_mul
    if activation:
        return _fused_lrelu(F.linear(x, w * scale), p[name + ".bias"] * lr_mul)
    return F.linear(x, w * scale, bias=p[name + ".bias"] * lr_mul)


def mapping(p, z, lr_mul=0.01):
    """Decoder.style :893-911: PixelNorm + 5 x EqualLinear(fused_lrelu); lr_mul = model_opt.lr_mapping"""
    h = z * torch.rsqrt(torch.mean(z ** 2, dim=1, keepdim=True) + 1e-8)
    for i in range(1, 6):
        h = equal_linear(p, f"style.{i}", h, lr_mul, activation=True)
    return h


def _blur_kernel(factor_sq):
    k = torch.tensor([1., 3., 3., 1.])
    k = k[None, :] * k[:, None]
    return k / k.sum() * factor_sq          # make_kernel :469-477 (* upsample_factor ** 2, :527-528 / :485)


def upfirdn2d(x, kernel, up=1, pad=(0, 0)):
    """upfirdn2d_native sdf_op.py:273-314 with down = 1: zero-stuff, pad, correlate with the flipped kernel"""
    b, c, h, w = x.shape
    if up > 1:
        z = x.new_zeros(b, c, h, up, w, up)
        z[:, :, :, 0, :, 0] = x
        x = z.view(b, c, h * up, w * up)
    x = F.pad(x, [pad[0], pad[1], pad[0], pad[1]])
    k = torch.flip(kernel, [0, 1]).view(1, 1, *kernel.shape).expand(c, 1, -1, -1)
    return F.conv2d(x, k, groups=c)


def modulated_conv(p, name, x, style, upsample=False, demodulate=True):
    """ModulatedConv2d.forward :655-704 (per-sample weights, grouped convolution; up-sampling = conv_transpose2d + Blur)"""
    w = p[name + ".weight"]                                   # [1, out, in, k, k]
    _, cout, cin, k, _ = w.shape
    b, _, h, wd = x.shape
    s = equal_linear(p, name + ".modulation", style).view(b, 1, cin, 1, 1)
    weight = (1 / math.sqrt(cin * k * k)) * w * s
    if demodulate:
        weight = weight * torch.rsqrt(weight.pow(2).sum([2, 3, 4]) + 1e-8).view(b, cout, 1, 1, 1)
    if upsample:
        wt = weight.transpose(1, 2).reshape(b * cin, cout, k, k)
        out = F.conv_transpose2d(x.reshape(1, b * cin, h, wd), wt, padding=0, stride=2, groups=b)
        out = out.view(b, cout, out.shape[-2], out.shape[-1])
        pq = (4 - 2) - (k - 1)
        return upfirdn2d(out, _blur_kernel(4.0), pad=((pq + 1) // 2 + 1, pq // 2 + 1))       # Blur pad (:628-631)
    out = F.conv2d(x.reshape(1, b * cin, h, wd), weight.view(b * cout, cin, k, k), padding=k // 2, groups=b)
    return out.view(b, cout, h, wd)


def styled_conv(p, name, x, style, noise, upsample=False):
    """StyledConv.forward :812-818: conv -> + noise.weight * noise -> FusedLeakyReLU(bias)"""
    out = modulated_conv(p, name + ".conv", x, style, upsample=upsample)
    out = out + p[name + ".noise.weight"] * noise
    return _fused_lrelu(out, p[name + ".activate.bias"])


def to_rgb(p, name, x, style, skip=None):
    """ToRGB.forward :833-843"""
    out = modulated_conv(p, name + ".conv", x, style, demodulate=False) + p[name + ".bias"]
    if skip is not None:
        out = out + upfirdn2d(skip, _blur_kernel(4.0), up=2, pad=(2, 1))                      # Upsample :480-499
    return out


def decoder_forward(p, features, z, noise, lr_mapping=0.01):
    """Decoder.forward :1027-1056 with one style (latent repeated n_latent times, :1005-1008) and explicit per-layer noise"""
    w = mapping(p, z, lr_mapping)
    n_up = len({k.split(".")[1] for k in p if k.startswith("to_rgbs.")})
    out = styled_conv(p, "conv1", features, w, noise[0])
    skip = to_rgb(p, "to_rgb1", out, w)
    for k in range(n_up):
        out = styled_conv(p, f"convs.{2 * k}", out, w, noise[2 * k + 1], upsample=True)
        out = styled_conv(p, f"convs.{2 * k + 1}", out, w, noise[2 * k + 2])
        skip = to_rgb(p, f"to_rgbs.{k}", out, w, skip)
    return skip
