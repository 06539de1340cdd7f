"""StyleGAN2 decoder of the SDF generator with the reference's module API, forward on the sm_100a kernels.

Mirrors /root/reference/im2scene/sdf/models/sdf_model.py: PixelNorm :429-434, Upsample :480-499, Blur :522-538, EqualLinear :578-611,
ModulatedConv2d :614-704, NoiseInjection :707-790, StyledConv :793-818, ToRGB :821-843, Decoder :883-1056 -- same class names,
constructor arguments, parameter / buffer names and shapes (a reference `full_pipeline.pt` loads unchanged), same forward
signature and return values.  What differs is HOW an image is decoded:
  * activations are channels-last fp16 and never leave that layout between layers;
  * a ModulatedConv2d is a per-sample weight fold (style modulation + demodulation folded into fp16 weights, like the reference's
    grouped-convolution trick) followed by an implicit-GEMM 3 x 3 convolution on tcgen05 (csrc/tc_conv.cuh) whose epilogue applies
    NoiseInjection, the FusedLeakyReLU bias and the activation;
  * an up-sampling StyledConv runs conv_transpose2d as its four output-parity classes in the same convolution kernel (per-class tap
    tables, stride-2 store into the fp16 (2H+1)^2 intermediate) and one kernel that blurs (upfirdn2d, separable), adds noise + bias and
    activates;
  * ToRGB is a 16-column GEMM in the same kernel whose epilogue adds the bias and the up-sampled skip (upfirdn2d): one launch per
    resolution.
The mapping network (5 EqualLinear on [B, 512]) stays torch.  Forward / inference only: with autograd enabled the decoder raises
(the generator's stage-2 training step, BASELINE configs[3], is not built).  `project_noise` (pytorch3d) is not supported.
"""
import math
import random

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import ops


def _lrelu(x, bias, scale=2 ** 0.5):
    # fused_leaky_relu on [B, C] tensors (mapping network): bias + leaky_relu(0.2) * scale (sdf_op.py:83-117)
    return F.leaky_relu(x + bias.view(1, -1), negative_slope=0.2) * scale


class PixelNorm(nn.Module):
    def forward(self, input):
        return input * torch.rsqrt(torch.mean(input ** 2, dim=1, keepdim=True) + 1e-8)


def make_kernel(k):
    k = torch.tensor(k, dtype=torch.float32)
    if k.ndim == 1:
        k = k[None, :] * k[:, None]
    return k / k.sum()


class Upsample(nn.Module):
    """parameters of the skip up-sampling (the filter itself runs inside the ToRGB kernel)"""

    def __init__(self, kernel, factor=2):
        super().__init__()
        if list(kernel) != [1, 3, 3, 1] or factor != 2:
            raise NotImplementedError("the fused ToRGB kernel implements the [1,3,3,1] x2 up-sampling filter only")
        self.factor = factor
        self.register_buffer("kernel", make_kernel(kernel) * (factor ** 2))
        p = self.kernel.shape[0] - factor
        self.pad = ((p + 1) // 2 + factor - 1, p // 2)


class Blur(nn.Module):
    """parameters of the blur behind an up-sampling convolution (runs inside sdfg_upconv_forward)"""

    def __init__(self, kernel, pad, upsample_factor=1):
        super().__init__()
        if list(kernel) != [1, 3, 3, 1]:
            raise NotImplementedError("the fused up-sampling kernel implements the [1,3,3,1] blur only")
        k = make_kernel(kernel)
        if upsample_factor > 1:
            k = k * (upsample_factor ** 2)
        self.register_buffer("kernel", k)
        self.pad = pad


class EqualLinear(nn.Module):
    def __init__(self, in_dim, out_dim, bias=True, bias_init=0, lr_mul=1, activation=None):
        super().__init__()
        self.weight = nn.Parameter(torch.randn(out_dim, in_dim).div_(lr_mul))
        self.bias = nn.Parameter(torch.zeros(out_dim).fill_(bias_init)) if bias else None
        self.activation = activation
        self.scale = (1 / math.sqrt(in_dim)) * lr_mul
        self.lr_mul = lr_mul

    def forward(self, input):
        if self.activation:
            return _lrelu(F.linear(input, self.weight * self.scale), self.bias * self.lr_mul)
        return F.linear(input, self.weight * self.scale, bias=self.bias * self.lr_mul)

    def __repr__(self):
        return f"{self.__class__.__name__}({self.weight.shape[1]}, {self.weight.shape[0]})"


class ModulatedConv2d(nn.Module):
    def __init__(self, in_channel, out_channel, kernel_size, style_dim, demodulate=True, upsample=False, downsample=False,
                 blur_kernel=[1, 3, 3, 1]):
        super().__init__()
        if downsample:
            raise NotImplementedError("down-sampling modulated convolutions are not part of the decoder")
        self.eps = 1e-8
        self.kernel_size = kernel_size
        self.in_channel = in_channel
        self.out_channel = out_channel
        self.upsample = upsample
        self.downsample = downsample
        if upsample:
            factor = 2
            p = (len(blur_kernel) - factor) - (kernel_size - 1)
            self.blur = Blur(blur_kernel, pad=((p + 1) // 2 + factor - 1, p // 2 + 1), upsample_factor=factor)
        self.scale = 1 / math.sqrt(in_channel * kernel_size ** 2)
        self.padding = kernel_size // 2
        self.weight = nn.Parameter(torch.randn(1, out_channel, in_channel, kernel_size, kernel_size))
        self.modulation = EqualLinear(style_dim, in_channel, bias_init=1)
        self.demodulate = demodulate

    def __repr__(self):
        return (f"{self.__class__.__name__}({self.in_channel}, {self.out_channel}, {self.kernel_size}, "
                f"upsample={self.upsample}, downsample={self.downsample})")

    def folded_weights(self, style):
        """per-sample fp16 weights [B, k*k, Cout, Cin]: scale * weight * modulation(style), demodulated (ref :655-662)"""
        return ops.modconv_fold(self.weight, self.modulation(style), self.scale, self.demodulate)


class NoiseInjection(nn.Module):
    def __init__(self, project=False):
        super().__init__()
        if project:
            raise NotImplementedError("project_noise (pytorch3d mesh projection, ref :713-782) is not supported")
        self.project = project
        self.weight = nn.Parameter(torch.zeros(1))

    def noise_for(self, batch, height, width, noise, device):
        """[B, H, W] fp32, contiguous: drawn when None (ref :785-786), broadcast over the batch when a [1,1,H,W] buffer is given"""
        if noise is None:
            return torch.randn(batch, height, width, device=device)
        return noise.to(device=device, dtype=torch.float32).expand(batch, 1, height, width).reshape(batch, height, width).contiguous()


class FusedLeakyReLU(nn.Module):
    def __init__(self, channel, bias=True, negative_slope=0.2, scale=2 ** 0.5):
        super().__init__()
        self.bias = nn.Parameter(torch.zeros(channel)) if bias else None
        self.negative_slope = negative_slope
        self.scale = scale


class StyledConv(nn.Module):
    def __init__(self, in_channel, out_channel, kernel_size, style_dim, upsample=False, blur_kernel=[1, 3, 3, 1], project_noise=False):
        super().__init__()
        self.conv = ModulatedConv2d(in_channel, out_channel, kernel_size, style_dim, upsample=upsample, blur_kernel=blur_kernel)
        self.noise = NoiseInjection(project=project_noise)
        self.bias = nn.Parameter(torch.zeros(1, out_channel, 1, 1))       # unused by the reference's forward as well (kept: state_dict)
        self.activate = FusedLeakyReLU(out_channel)

    def forward(self, input, style, noise=None, transform=None, mesh_path=None):
        """input: channels-last fp16 [B, H, W, Cin] -> [B, H', W', Cout] (conv -> noise -> bias + leaky ReLU * sqrt 2, ref :812-816)"""
        B, H, W, _ = input.shape
        wf = self.conv.folded_weights(style)
        up = 2 if self.conv.upsample else 1
        nz = self.noise.noise_for(B, H * up, W * up, noise, input.device)
        if self.conv.upsample:
            return ops.upconv_forward(input, wf, bias=self.activate.bias, noise=nz, noise_w=self.noise.weight)
        return ops.conv_forward(input, wf, bias=self.activate.bias, noise=nz, noise_w=self.noise.weight)


class ToRGB(nn.Module):
    def __init__(self, in_channel, style_dim, upsample=True, blur_kernel=[1, 3, 3, 1]):
        super().__init__()
        self.upsample = upsample
        if upsample:
            self.upsample = Upsample(blur_kernel)
        self.conv = ModulatedConv2d(in_channel, 3, 1, style_dim, demodulate=False)
        self.bias = nn.Parameter(torch.zeros(1, 3, 1, 1))

    def forward(self, input, style, skip=None, final=False):
        """input fp16 [B,H,W,C]; skip fp32 [B,H/2,W/2,3] (channels last) -> rgb fp32: channels last (next level's skip), or NCHW when final"""
        if skip is not None and not self.upsample:
            raise NotImplementedError("a skip input needs the up-sampling ToRGB")
        nhwc, nchw = ops.to_rgb(input, self.conv.weight, self.conv.modulation(style), self.conv.scale, self.bias, skip,
                                want_nhwc=not final, want_nchw=final)
        return nchw if final else nhwc


class Decoder(nn.Module):
    def __init__(self, model_opt, blur_kernel=[1, 3, 3, 1]):
        super().__init__()
        self.size = model_opt.size
        self.style_dim = model_opt.style_dim * 2
        layers = [PixelNorm(), EqualLinear(self.style_dim // 2 if not model_opt.psp else self.style_dim, self.style_dim,
                                           lr_mul=model_opt.lr_mapping, activation="fused_lrelu")]
        for _ in range(4):
            layers.append(EqualLinear(self.style_dim, self.style_dim, lr_mul=model_opt.lr_mapping, activation="fused_lrelu"))
        self.style = nn.Sequential(*layers)
        self.channels = {4: 512, 8: 512, 16: 512, 32: 512, 64: 256 * model_opt.channel_multiplier, 128: 128 * model_opt.channel_multiplier,
                         256: 64 * model_opt.channel_multiplier, 512: 32 * model_opt.channel_multiplier, 1024: 16 * model_opt.channel_multiplier}
        decoder_in_size = model_opt.renderer_spatial_output_dim
        self.log_size = int(math.log(self.size, 2))
        self.log_in_size = int(math.log(decoder_in_size, 2))
        input_feature_channels = model_opt.feature_encoder_in_channels if not model_opt.psp else self.style_dim
        self.conv1 = StyledConv(input_feature_channels, self.channels[decoder_in_size], 3, self.style_dim, blur_kernel=blur_kernel,
                                project_noise=model_opt.project_noise)
        self.to_rgb1 = ToRGB(self.channels[decoder_in_size], self.style_dim, upsample=False)
        self.num_layers = (self.log_size - self.log_in_size) * 2 + 1
        self.convs = nn.ModuleList()
        self.upsamples = nn.ModuleList()
        self.to_rgbs = nn.ModuleList()
        self.noises = nn.Module()
        in_channel = self.channels[decoder_in_size]
        for layer_idx in range(self.num_layers):
            res = (layer_idx + 2 * self.log_in_size + 1) // 2
            self.noises.register_buffer(f"noise_{layer_idx}", torch.randn(1, 1, 2 ** res, 2 ** res))
        for i in range(self.log_in_size + 1, self.log_size + 1):
            out_channel = self.channels[2 ** i]
            self.convs.append(StyledConv(in_channel, out_channel, 3, self.style_dim, upsample=True, blur_kernel=blur_kernel,
                                         project_noise=model_opt.project_noise))
            self.convs.append(StyledConv(out_channel, out_channel, 3, self.style_dim, blur_kernel=blur_kernel,
                                         project_noise=model_opt.project_noise))
            self.to_rgbs.append(ToRGB(out_channel, self.style_dim))
            in_channel = out_channel
        self.n_latent = (self.log_size - self.log_in_size) * 2 + 2

    def mean_latent(self, renderer_latent):
        return self.style(renderer_latent).mean(0, keepdim=True)

    def get_latent(self, input):
        return self.style(input)

    def styles_and_noise_forward(self, styles, noise, inject_index=None, truncation=1, truncation_latent=None, input_is_latent=False,
                                 randomize_noise=True):
        if not input_is_latent:
            styles = [self.style(s) for s in styles]
        if noise is None:
            noise = [None] * self.num_layers if randomize_noise else [getattr(self.noises, f"noise_{i}") for i in range(self.num_layers)]
        if truncation < 1:
            styles = [truncation_latent[1] + truncation * (s - truncation_latent[1]) for s in styles]
        if len(styles) < 2:
            inject_index = self.n_latent
            latent = styles[0].unsqueeze(1).repeat(1, inject_index, 1) if styles[0].ndim < 3 else styles[0]
        else:
            if inject_index is None:
                inject_index = random.randint(1, self.n_latent - 1)
            latent = torch.cat([styles[0].unsqueeze(1).repeat(1, inject_index, 1),
                                styles[1].unsqueeze(1).repeat(1, self.n_latent - inject_index, 1)], 1)
        return latent, noise

    def forward(self, features, styles, rgbd_in=None, transform=None, return_latents=False, inject_index=None, truncation=1,
                truncation_latent=None, input_is_latent=False, noise=None, randomize_noise=True, mesh_path=None):
        """features [B, C, H, W] fp32 (any memory format; the renderer hands over a channels-last view, so no transposing copy happens)
        -> (image [B, 3, size, size] fp32, latent | None)"""
        if torch.is_grad_enabled() and (features.requires_grad or any(p.requires_grad for p in self.parameters())):
            raise NotImplementedError("the B200 decoder kernels are forward-only: run the decoder under torch.no_grad() "
                                      "(stage-2 training, BASELINE configs[3], is not built)")
        if not features.is_cuda:
            raise RuntimeError("features must be a CUDA tensor (this path has no CPU fallback)")
        if rgbd_in is not None or transform is not None:
            raise NotImplementedError("rgbd_in / project_noise are not supported")
        latent, noise = self.styles_and_noise_forward(styles, noise, inject_index, truncation, truncation_latent, input_is_latent, randomize_noise)
        x = ops.nhwc16(features.permute(0, 2, 3, 1).contiguous().float())
        n_up = len(self.to_rgbs)
        out = self.conv1(x, latent[:, 0], noise=noise[0])
        skip = self.to_rgb1(out, latent[:, 1], final=n_up == 0)
        i = 1
        for k, (conv1, conv2, to_rgb) in enumerate(zip(self.convs[::2], self.convs[1::2], self.to_rgbs)):
            out = conv1(out, latent[:, i], noise=noise[2 * k + 1])
            out = conv2(out, latent[:, i + 1], noise=noise[2 * k + 2])
            skip = to_rgb(out, latent[:, i + 2], skip=skip, final=k + 1 == n_up)
            i += 2
        return skip, (latent if return_latents else None)
