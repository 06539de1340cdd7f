"""SDF volume renderer + style-modulated field with the reference's module API, on the sm_100a kernels.

Mirrors /root/reference/im2scene/sdf/models/sdf_model.py (class names, constructor arguments, attribute names, option keys,
`state_dict` keys/shapes, forward signatures and output tuple protocol):
  LinearLayer :23-41, FiLMSiren :44-69, SirenGenerator :101-139, VolumeFeatureRenderer :143-423, MappingLinear :437-466,
  Generator :1059-1216, get_encoder :1512-1531, NGPSIRENGenerator :1534-1596.
What is different is HOW a batch is rendered: one ray/sample kernel, one hash-grid kernel, one SH kernel per RAY (not per
sample), the field as a fused sequence of layer kernels (csrc/field_*.cu) and one compositing kernel -- about 12 launches
instead of ~45 + the [N,272]/[N,260] concatenations.  There is no CPU path: tensors must live on a CUDA device.
"""
import math
import warnings

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F
from torch.autograd import Function
from torch.autograd.function import once_differentiable

from . import _lib, ops
from .gridencoder import GridEncoder
from .shencoder import SHEncoder

_PRECISIONS = {"fp32": _lib.PRECISION_FP32, "tc16": _lib.PRECISION_TC16, "bf16": _lib.PRECISION_TC16}


class LinearLayer(nn.Module):
    """std_init * (x W^T + b) + bias_init  (ref :23-41).  Only evaluated by torch for [B, style_dim] inputs (gamma/beta heads);
    the per-sample layers are consumed as raw weights by the field kernels."""

    def __init__(self, in_dim, out_dim, bias=True, bias_init=0, std_init=1, freq_init=False, is_first=False):
        super().__init__()
        if is_first:
            self.weight = nn.Parameter(torch.empty(out_dim, in_dim).uniform_(-1 / in_dim, 1 / in_dim))
        elif freq_init:
            lim = np.sqrt(6 / in_dim) / 25
            self.weight = nn.Parameter(torch.empty(out_dim, in_dim).uniform_(-lim, lim))
        else:
            self.weight = nn.Parameter(0.25 * nn.init.kaiming_normal_(torch.randn(out_dim, in_dim), a=0.2, mode="fan_in",
                                                                      nonlinearity="leaky_relu"))
        self.bias = nn.Parameter(nn.init.uniform_(torch.empty(out_dim), a=-np.sqrt(1 / in_dim), b=np.sqrt(1 / in_dim)))
        self.bias_init = bias_init
        self.std_init = std_init

    def forward(self, input):
        return self.std_init * F.linear(input, self.weight, bias=self.bias) + self.bias_init


class FiLMSiren(nn.Module):
    """sin(gamma * (x W^T + b) + beta) with gamma = 15 Lin(w) + 30, beta = 0.25 Lin(w)  (ref :44-69)."""

    def __init__(self, in_channel, out_channel, style_dim, is_first=False):
        super().__init__()
        self.in_channel = in_channel
        self.out_channel = out_channel
        lim = 1 / 3 if is_first else np.sqrt(6 / in_channel) / 25
        self.weight = nn.Parameter(torch.empty(out_channel, in_channel).uniform_(-lim, lim))
        self.bias = nn.Parameter(nn.init.uniform_(torch.empty(out_channel), a=-np.sqrt(1 / in_channel), b=np.sqrt(1 / in_channel)))
        self.activation = torch.sin
        self.gamma = LinearLayer(style_dim, out_channel, bias_init=30, std_init=15)
        self.beta = LinearLayer(style_dim, out_channel, bias_init=0, std_init=0.25)

    def modulation(self, style):
        return self.gamma(style), self.beta(style)

    def forward(self, input, style):
        # stand-alone evaluation of one layer (the generators below never call this; they hand all layers to the field kernels)
        batch, features = style.shape
        out = F.linear(input, self.weight, bias=self.bias)
        gamma, beta = self.modulation(style)
        shape = [batch] + [1] * (out.dim() - 2) + [-1]
        return torch.sin(gamma.view(shape) * out + beta.view(shape))


# --------------------------------------------------------------------------------------------------------------------------
# the fused field as one autograd node

class _field(Function):
    """(x_in [N,in_dim], view_feat [N/S,V], gamma/beta [B,n+1,W], emb, weights...) -> (sdf [N], rgb [N,3], feat [N,W], dsdf_dx [N,in_dim]).

    `emb` is the hash table when the caller computed x_in = grid_encode(meta["grid"]["pts"], emb) itself (without autograd):
    the node then also owns the table gradient, so its backward can order the work as
        gradient chain -> table scatter -> [async all-reduce of the table gradient] -> weight-gradient kernels
    on ONE stream: in data-parallel training (distributed.data_parallel) the 50.6 MB table gradient -- 93 % of the bytes
    exchanged per step, and the LAST gradient autograd would hand to a DistributedDataParallel bucket -- travels while the
    HBM-bound weight-gradient contractions run, and the L2-atomic-bound scatter never shares the chip with them.
    emb = None: x_in is an ordinary differentiable input.

    First order only: the backward is `once_differentiable` (a double backward raises instead of silently returning zeros), and
    the view feature gets no gradient (asking for one raises)."""

    @staticmethod
    def forward(ctx, spec, meta, x_in, view_feat, gamma, beta, emb, *wts):
        ctx.set_materialize_grads(False)
        if ctx.needs_input_grad[3]:
            raise RuntimeError("the fused field has no gradient for the view feature (view directions / SH features must not "
                               "require grad; the reference never differentiates them either, SURVEY.md 2.3 kernel_sh_backward)")
        x_in = x_in.contiguous()
        view_feat = view_feat.contiguous()
        gamma = gamma.contiguous()
        beta = beta.contiguous()
        wts = tuple(w.contiguous() for w in wts)
        weights = _unpack_weights(spec, wts)
        need_bwd = (meta.get("grad", True) and any(ctx.needs_input_grad[2:])) or meta["want_dsdf"]
        # inference on the tensor-core path: features stay fp16 between the field and the compositing kernel (half the HBM bytes)
        feat_f16 = bool(meta.get("feat_f16")) and not need_bwd and meta["precision"] == _lib.PRECISION_TC16
        sdf, rgb, feat, ws = ops.field_forward(spec, x_in, view_feat, gamma, beta, weights, meta["spi"], meta["spr"],
                                               want_rgb=meta["want_rgb"], want_feat=meta["want_feat"],
                                               save_for_backward=need_bwd, precision=meta["precision"], feat_f16=feat_f16)
        dsdf = None
        if meta["want_dsdf"]:
            # d sdf / d x_in for the eikonal term (ref get_eikonal_term :224-229): a trunk-only backward with d_sdf = 1
            ones = torch.ones_like(sdf)
            eik = meta.get("eik")
            if eik is not None:
                # hash-grid field: the encoder's chain rule (d feature / d point, dy_dx) is applied inside the chain's last epilogue --
                # d sdf / d point [N,3] comes back directly, d sdf / d feature [N,32] never exists.  None = no fused kernel for this shape.
                dsdf = ops.field_eikonal(spec, x_in, view_feat, gamma, beta, weights, meta["spi"], meta["spr"], ws, ones, eik["dy_dx"],
                                         eik["scale"], precision=meta["precision"])
            if dsdf is None:
                dsdf = ops.field_backward(spec, x_in, view_feat, gamma, beta, weights, meta["spi"], meta["spr"], ws, feat, ones, None, None,
                                          grads=None, want_dx=True, precision=meta["precision"])
        if need_bwd:
            ctx.save_for_backward(x_in, view_feat, gamma, beta, feat, ws, emb, *wts)
        ctx.spec, ctx.meta = spec, meta
        empty = x_in.new_empty(0)
        outs = (sdf, rgb if rgb is not None else empty, feat if feat is not None else empty, dsdf if dsdf is not None else empty)
        # one call: mark_non_differentiable REPLACES the set on every call.  dsdf is constant w.r.t. the parameters (SURVEY finding 4)
        ctx.mark_non_differentiable(*([outs[3]] + ([outs[1]] if rgb is None else []) + ([outs[2]] if feat is None else [])))
        return outs

    @staticmethod
    @once_differentiable
    def backward(ctx, d_sdf, d_rgb, d_feat, _d_dsdf):
        spec, meta = ctx.spec, ctx.meta
        x_in, view_feat, gamma, beta, feat, ws, emb = ctx.saved_tensors[:7]
        wts = ctx.saved_tensors[7:]
        weights = _unpack_weights(spec, wts)
        if d_sdf is None and d_rgb is None and d_feat is None:
            return (None,) * (7 + len(wts))
        if not meta["want_rgb"]:
            d_rgb = None
        if not meta["want_feat"]:
            d_feat = None
        cont = lambda t: None if t is None else t.contiguous()
        need_param = any(ctx.needs_input_grad[4:6]) or any(ctx.needs_input_grad[7:])
        grid = meta.get("grid")
        need_table = grid is not None and emb is not None and ctx.needs_input_grad[6]
        grads = None
        if need_param:
            gw = tuple(torch.zeros_like(w) for w in wts)
            grads = _unpack_weights(spec, gw)
            grads["gamma"] = torch.zeros_like(gamma)
            grads["beta"] = torch.zeros_like(beta)
        d_sdf, d_rgb, d_feat = cont(d_sdf), cont(d_rgb), cont(d_feat)
        want_dx = bool(ctx.needs_input_grad[2] or need_table)
        args = (spec, x_in, view_feat, gamma, beta, weights, meta["spi"], meta["spr"], ws, feat if meta["want_feat"] else None,
                d_sdf, d_rgb, d_feat)
        # phase 1: loss scale + gradient chain through the layers -> d x_in (and the gradient tiles phase 2 contracts)
        dx, state = ops.field_backward(*args, grads=grads, want_dx=want_dx, precision=meta["precision"], phases=_lib.BWD_CHAIN)
        d_emb, work, world = None, None, 1
        if need_table:
            # The scatter runs ALONE between the two phases.  Overlapping it with the weight-gradient contractions was measured both ways
            # (DESIGN 6): with equal stream priorities its 12 k pending blocks starve the other stream's kernels (no overlap), with the
            # contractions on a high-priority stream both run together and both slow down 2-3x (the contractions stream 13 GB through
            # the L2 that the scatter's 50 MB gradient table wants to stay resident in): 14.8-14.9 ms per step against 14.3-14.4.
            d_emb = torch.zeros_like(emb)
            ops.grid_encode_backward(dx, grid["pts"], emb, grid["offsets"], grid["S"], grid["H"], bound=grid["bound"], grad_embeddings=d_emb,
                                     gridtype=grid["gridtype"], align_corners=grid["align_corners"], interp=grid["interp"])
            ex = meta.get("exchange")
            if ex is not None and torch.distributed.is_available() and torch.distributed.is_initialized():
                world = torch.distributed.get_world_size(ex["group"])
                if world > 1:      # SUM + divide: ReduceOp.AVG exists for NCCL only
                    work = torch.distributed.all_reduce(d_emb, op=torch.distributed.ReduceOp.SUM, group=ex["group"], async_op=True)
        # phase 2: parameter gradients (HBM-bound contractions) while the table gradient travels
        if need_param:
            ops.field_backward(*args, grads=grads, want_dx=want_dx, precision=meta["precision"], phases=_lib.BWD_WGRAD, state=state)
        if work is not None:
            work.wait()                                        # stream-level for NCCL: the current stream waits for the collective
            d_emb.div_(world)
        del state
        out = [None, None, dx if ctx.needs_input_grad[2] else None, None]
        if need_param:
            out += [grads["gamma"], grads["beta"], d_emb] + list(gw)
        else:
            out += [None, None, d_emb] + [None] * len(wts)
        return tuple(out)


def _pack_weights(spec, net):
    w = []
    if spec.has_input_linear:
        w += [net.input_linear.weight, net.input_linear.bias]
    layers = list(net.pts_linears) + [net.views_linears]
    w += [l.weight for l in layers] + [l.bias for l in layers]
    w += [net.sigma_linear.weight, net.sigma_linear.bias, net.rgb_linear.weight, net.rgb_linear.bias]
    return w


def _unpack_weights(spec, w):
    d, i = {}, 0
    if spec.has_input_linear:
        d["input_w"], d["input_b"] = w[0], w[1]
        i = 2
    n = spec.n_film + 1
    d["film_w"] = list(w[i:i + n])
    d["film_b"] = list(w[i + n:i + 2 * n])
    i += 2 * n
    d["sigma_w"], d["sigma_b"], d["rgb_w"], d["rgb_b"] = w[i], w[i + 1], w[i + 2], w[i + 3]
    return d


_WARNED = set()


def _warn_once(key, msg):
    if key not in _WARNED:
        _WARNED.add(key)
        warnings.warn(msg, RuntimeWarning, stacklevel=3)


class _FieldNetwork(nn.Module):
    """Shared driver of SirenGenerator / NGPSIRENGenerator: evaluates the whole network through `_field`."""

    # "auto": the tcgen05 path (fp16 activations and loss-scaled fp16 gradients, fp32 accumulate) whenever its shape constraints
    # hold (width 256, samples per image a multiple of 128, and -- when d sdf / d x_in is wanted -- an input_linear layer with
    # in_dim % 32 == 0), else the fp32 SIMT path, with a one-time warning (the fp32 path is ~13x slower).  Both are CUDA; neither
    # falls back to the CPU.
    precision = "auto"
    _table_exchange = None      # set by distributed.data_parallel(early_table_exchange=True): {"group": process group or None}

    def _pick_precision(self, samples_per_image, want_dx=False):
        if self.precision != "auto":
            return _PRECISIONS[self.precision]
        spec = self._spec
        why = None
        if spec.width != 256:
            why = "width %d != 256" % spec.width
        elif samples_per_image % 128 != 0:
            why = "samples per image (%d) not a multiple of 128" % samples_per_image
        elif spec.has_input_linear and spec.in_dim % 32 != 0:
            why = "in_dim %d not a multiple of 32" % spec.in_dim
        elif want_dx and not spec.has_input_linear:
            why = "d sdf / d points through a network without input_linear"
        if why is None:
            return _lib.PRECISION_TC16
        _warn_once(("fp32", why), "sdface-gan_b200: precision='auto' uses the fp32 SIMT field kernels instead of the tcgen05 path (%s); "
                   "this is the slow, <= 1e-3-parity path" % why)
        return _lib.PRECISION_FP32

    def _modulation(self, styles):
        """gamma = 15 Lin(w) + 30, beta = 0.25 Lin(w) of every FiLM layer (ref :58-59) as ONE [B, style] x [style, 2 (n+1) W] product
        instead of 2 (n+1) skinny GEMMs with their scale / shift launches (forward and backward: ~100 launches of a training step)."""
        layers = list(self.pts_linears) + [self.views_linears]
        n, W = len(layers), layers[0].out_channel
        heads = [l.gamma for l in layers] + [l.beta for l in layers]
        weight = torch.cat([h.weight for h in heads], 0)
        bias = torch.cat([h.bias for h in heads], 0)
        key = (styles.device, styles.dtype)
        consts = self.__dict__.setdefault("_mod_consts", {})
        if key not in consts:                                             # per-column scale / shift, built once per device (no per-call H2D copy)
            consts[key] = (styles.new_tensor([h.std_init for h in heads]).repeat_interleave(W),
                           styles.new_tensor([h.bias_init for h in heads]).repeat_interleave(W))
        mul, add = consts[key]
        out = torch.addcmul(add, F.linear(styles, weight, bias), mul).view(styles.shape[0], 2, n, W)
        return out[:, 0], out[:, 1]                                      # gamma, beta: [B, n+1, W] (made contiguous by the field node)

    def _run_field(self, x_in, view_feat, styles, samples_per_image, samples_per_ray, want_rgb=True, want_feat=True, want_dsdf=False,
                   feat_f16=False, emb=None, grid=None, eik=None):
        spec = self._spec
        gamma, beta = self._modulation(styles)
        want_dx = bool(want_dsdf) or (torch.is_grad_enabled() and x_in.requires_grad)
        meta = dict(spi=int(samples_per_image), spr=int(samples_per_ray), want_rgb=bool(want_rgb), want_feat=bool(want_feat),
                    want_dsdf=bool(want_dsdf), precision=self._pick_precision(int(samples_per_image), want_dx),
                    grad=torch.is_grad_enabled(), feat_f16=bool(feat_f16), grid=grid,   # Function.forward itself always runs with grad mode off
                    exchange=self._table_exchange, eik=eik if want_dsdf else None)
        sdf, rgb, feat, dsdf = _field.apply(spec, meta, x_in, view_feat, gamma, beta, emb, *_pack_weights(spec, self))
        return sdf, (rgb if rgb.numel() else None), (feat if feat.numel() else None), (dsdf if dsdf.numel() else None)

    def forward(self, x, styles):
        """Reference-compatible entry: x [B, ..., 6] = (point, view direction) per SAMPLE -> [B, ..., 3 + 1 (+ W)]."""
        prefix = list(x.shape[:-1])
        B = x.shape[0]
        flat = x.reshape(-1, x.shape[-1])
        pts, dirs = flat[:, :self.input_ch], flat[:, self.input_ch:self.input_ch + self.input_ch_views]
        x_in, view_feat = self._encode(pts, dirs.contiguous())
        sdf, rgb, feat, _ = self._run_field(x_in, view_feat, styles, flat.shape[0] // B, 1, want_feat=self.output_features)
        out = [rgb, sdf.unsqueeze(-1)]
        if self.output_features:
            out.append(feat)
        return torch.cat(out, -1).view(prefix + [-1])

    def _eikonal_fusion(self, grid_ctx):
        """what the field node needs to apply the encoder's chain rule itself (None: the caller does it in _dsdf_to_points)"""
        return None

    def forward_rays(self, npts, viewdirs, styles, want_rgb=True, want_feat=True, want_dsdf=False, feat_f16=False):
        """Renderer entry: npts [B,R,R,S,3], viewdirs [B,R,R,3] (one per ray) -> (sdf [N], rgb [N,3], feat [N,W], dsdf_dnpts [N,3])."""
        B, R1, R2, S, _ = npts.shape
        flat = npts.reshape(-1, 3)
        x_in, view_feat, grid_ctx, emb, grid = self._encode_rays(flat, viewdirs.reshape(-1, 3), want_dsdf)
        sdf, rgb, feat, dsdf = self._run_field(x_in, view_feat, styles, R1 * R2 * S, S, want_rgb, want_feat, want_dsdf, feat_f16,
                                               emb=emb, grid=grid, eik=self._eikonal_fusion(grid_ctx))
        if want_dsdf and not (grid_ctx is not None and dsdf.shape[1] == 3):      # [N,3] from a hash-grid field: the chain rule was fused in
            dsdf = self._dsdf_to_points(dsdf, flat, grid_ctx)
        return sdf, rgb, feat, dsdf


class SirenGenerator(_FieldNetwork):
    """8 x FiLM-SIREN(256) on raw points, view branch on raw directions (ref :101-139)."""

    def __init__(self, D=8, W=256, style_dim=256, input_ch=3, input_ch_views=3, output_ch=4, output_features=True):
        super().__init__()
        self.D, self.W = D, W
        self.input_ch, self.input_ch_views = input_ch, input_ch_views
        self.style_dim = style_dim
        self.output_features = output_features
        self.pts_linears = nn.ModuleList([FiLMSiren(3, W, style_dim=style_dim, is_first=True)] +
                                         [FiLMSiren(W, W, style_dim=style_dim) for _ in range(D - 1)])
        self.views_linears = FiLMSiren(input_ch_views + W, W, style_dim=style_dim)
        self.rgb_linear = LinearLayer(W, 3, freq_init=True)
        self.sigma_linear = LinearLayer(W, 1, freq_init=True)
        self._spec = ops.FieldSpec(W, input_ch, input_ch_views, D, False)

    def _encode(self, pts, dirs):
        return pts.contiguous(), dirs

    def _encode_rays(self, flat_pts, ray_dirs, want_dsdf):
        return flat_pts.contiguous(), ray_dirs.contiguous(), None, None, None

    def _dsdf_to_points(self, dsdf, flat_pts, grid_ctx):
        return dsdf

    def _trunk_sdf_torch(self, pts, styles, spi):
        """sdf(pts) through plain torch ops (ref :121-131): the second-order-capable evaluation behind the eikonal LOSS GRADIENT."""
        B = styles.shape[0]
        h = pts.view(B, spi, -1)
        for l in self.pts_linears:
            g, b = l.modulation(styles)
            h = torch.sin(g[:, None, :] * F.linear(h, l.weight, l.bias) + b[:, None, :])
        return self.sigma_linear(h).reshape(-1)

    def forward_rays(self, npts, viewdirs, styles, want_rgb=True, want_feat=True, want_dsdf=False, feat_f16=False):
        """With `--ngp 0` the reference's eikonal term carries a real second-order gradient into the SIREN weights
        (autograd.grad(..., create_graph=True), ref :224-229) -- unlike `--ngp 1`, where the opaque hash-grid backward cuts it
        (SURVEY.md finding 4).  The fused kernels are first order, so while gradients are enabled the term is evaluated here
        through torch autograd on the trunk (cuBLAS, differentiable twice); outputs and first-order gradients still come from
        the fused kernels.  Under no_grad the fused first-order kernel path provides it (fp32 kernels: no input_linear)."""
        if not (want_dsdf and torch.is_grad_enabled()):
            return super().forward_rays(npts, viewdirs, styles, want_rgb, want_feat, want_dsdf, feat_f16)
        _warn_once("siren-eik", "sdface-gan_b200: SirenGenerator evaluates the DIFFERENTIABLE eikonal term (--ngp 0 training) through torch "
                   "autograd on the trunk; the fused kernels provide outputs and first-order gradients only")
        sdf, rgb, feat, _ = super().forward_rays(npts, viewdirs, styles, want_rgb, want_feat, False, feat_f16)
        B, R1, R2, S, _ = npts.shape
        pts = npts.detach().reshape(-1, 3).requires_grad_(True)
        sdf_t = self._trunk_sdf_torch(pts, styles, R1 * R2 * S)
        dsdf = torch.autograd.grad(sdf_t, pts, torch.ones_like(sdf_t), create_graph=True)[0]
        return sdf, rgb, feat, dsdf


def get_encoder(encoding, input_dim=3, multires=6, degree=4, num_levels=16, level_dim=2, base_resolution=16, log2_hashmap_size=19,
                desired_resolution=2048, align_corners=False, **kwargs):
    """ref :1512-1531"""
    if encoding == "sphere_harmonics":
        encoder = SHEncoder(input_dim=input_dim, degree=degree)
    elif encoding == "hashgrid":
        encoder = GridEncoder(input_dim=input_dim, num_levels=num_levels, level_dim=level_dim, base_resolution=base_resolution,
                              log2_hashmap_size=log2_hashmap_size, desired_resolution=desired_resolution, gridtype="hash",
                              align_corners=align_corners)
    else:
        raise NotImplementedError("Unknown encoding mode, choose from [None, frequency, sphere_harmonics, hashgrid, tiledgrid]")
    return encoder, encoder.output_dim


class NGPSIRENGenerator(_FieldNetwork):
    """hash-grid(x) -> Linear(32->256) -> 3 x FiLM-SIREN -> {sdf head, FiLM-SIREN(256+SH16 -> 256) -> rgb head}  (ref :1534-1596)."""

    def __init__(self, D=2, W=256, style_dim=256, output_features=True):
        super().__init__()
        self.D, self.W = D, W
        self.bound = 2
        self.style_dim = style_dim
        self.input_ch, self.input_ch_views = 3, 3
        self.output_features = output_features
        self.encoder, self.in_dim = get_encoder("hashgrid", desired_resolution=2048 * self.bound)
        self.encoder_dir, self.in_dim_dir = get_encoder("sphere_harmonics")
        self.input_linear = LinearLayer(self.in_dim, W, freq_init=True)
        self.pts_linears = nn.ModuleList([FiLMSiren(W, W, style_dim=style_dim, is_first=True)] +
                                         [FiLMSiren(W, W, style_dim=style_dim) for _ in range(D)])
        self.views_linears = FiLMSiren(self.in_dim_dir + W, W, style_dim=style_dim)
        self.rgb_linear = LinearLayer(W, 3, freq_init=True)
        self.sigma_linear = LinearLayer(W, 1, freq_init=True)
        self._spec = ops.FieldSpec(W, self.in_dim, self.in_dim_dir, D + 1, True)

    def _encode(self, pts, dirs):
        return self.encoder(pts, bound=self.bound), self.encoder_dir(dirs)

    def _encode_rays(self, flat_pts, ray_dirs, want_dsdf):
        # The features are computed here WITHOUT an autograd node: the field node (`_field`) receives the table as an extra input and
        # owns its gradient, so that its backward can run the table scatter and the weight-gradient contractions concurrently.
        # dy_dx is kept for the eikonal chain rule only.
        enc = self.encoder
        pts = flat_pts.detach().contiguous().float()
        grid = dict(pts=pts, offsets=enc.offsets, S=ops.log2_scale(enc.per_level_scale), H=enc.base_resolution, bound=float(self.bound),
                    gridtype=enc.gridtype_id, align_corners=enc.align_corners, interp=enc.interp_id)
        with torch.no_grad():
            feats, dy_dx = ops.grid_encode_forward(pts, enc.embeddings.detach(), enc.offsets, grid["S"], grid["H"], bound=grid["bound"],
                                                   calc_dy_dx=bool(want_dsdf), gridtype=grid["gridtype"], align_corners=grid["align_corners"],
                                                   interp=grid["interp"])
            sh = self.encoder_dir(ray_dirs)
        return feats, sh, dy_dx, enc.embeddings, grid

    def _eikonal_fusion(self, grid_ctx):
        enc = self.encoder
        if grid_ctx is None or enc.input_dim != 3 or enc.level_dim != 2:
            return None
        return dict(dy_dx=grid_ctx, scale=1.0 / (2.0 * float(self.bound)))      # the 1 / (2 bound) of the affine map folded into the encoder (grid.py:149)

    def _dsdf_to_points(self, dsdf, flat_pts, dy_dx):
        enc = self.encoder
        _, gi = ops.grid_encode_backward(dsdf.contiguous(), flat_pts.contiguous(), enc.embeddings.detach(), enc.offsets,
                                         ops.log2_scale(enc.per_level_scale), enc.base_resolution, bound=float(self.bound), dy_dx=dy_dx,
                                         grad_embeddings=None, want_grad_inputs=True, gridtype=enc.gridtype_id,
                                         align_corners=enc.align_corners, interp=enc.interp_id)
        return gi

    def query_sdf(self, input_pts, styles):
        # ref :1594-1596 -- returns the hash EMBEDDING (the "smoothness" loss is a TV on features)
        return self.encoder(input_pts, bound=self.bound)


class _grid_encode_keep(Function):
    """grid_encode that also returns dy_dx (non-differentiable) for the eikonal pass; inputs themselves get no gradient."""

    @staticmethod
    def forward(ctx, inputs, embeddings, offsets, per_level_scale, base_resolution, gridtype, align_corners, interpolation, bound):
        inputs = inputs.contiguous().float()
        S = ops.log2_scale(per_level_scale)
        outputs, dy_dx = ops.grid_encode_forward(inputs, embeddings.contiguous(), offsets, S, base_resolution, bound=bound, calc_dy_dx=True,
                                                 gridtype=gridtype, align_corners=align_corners, interp=interpolation)
        ctx.save_for_backward(inputs, embeddings, offsets)
        ctx.meta = (S, base_resolution, gridtype, align_corners, interpolation, bound)
        ctx.mark_non_differentiable(dy_dx)
        return outputs, dy_dx

    @staticmethod
    def backward(ctx, grad, _unused):
        inputs, embeddings, offsets = ctx.saved_tensors
        S, H, gridtype, align_corners, interpolation, bound = ctx.meta
        ge = torch.zeros_like(embeddings)
        ops.grid_encode_backward(grad.contiguous(), inputs, embeddings, offsets, S, H, bound=bound, grad_embeddings=ge, gridtype=gridtype,
                                 align_corners=align_corners, interp=interpolation)
        return None, ge, None, None, None, None, None, None, None


class FCGenerator(nn.Module):
    """`--fc 1` ablation (ref :1599-1670): positional encoding -> ReLU MLP with the style added after the first layer.  Kept
    for drop-in completeness: it runs on the GPU through cuBLAS (torch) between this package's ray-sampling and compositing
    kernels; it is NOT one of the fused tensor-core kernels (the north-star configuration is `--ngp 1 --fc 0`)."""

    def __init__(self, D=8, W=256, style_dim=256, input_ch=3, input_ch_views=3, output_ch=4, output_features=True):
        super().__init__()
        self.D, self.W = D, W
        self.input_ch, self.input_ch_views = input_ch, input_ch_views
        self.style_dim = style_dim
        self.output_features = output_features
        self.n_freq_posenc, self.n_freq_posenc_views = 10, 4
        self.x_in = nn.Linear(3 * self.n_freq_posenc * 2, W)
        self.style_in = nn.Linear(style_dim, W)
        self.pts_linears = nn.ModuleList([nn.Linear(W, W) for _ in range(D - 1)])
        self.views_linears = nn.Linear(3 * self.n_freq_posenc_views * 2 + W, W)
        self.rgb_linear = nn.Linear(W, 3)
        self.sigma_linear = nn.Linear(W, 1)

    def transform_points(self, p, views=False):
        p = p / 2
        L = self.n_freq_posenc_views if views else self.n_freq_posenc
        return torch.cat([torch.cat([torch.sin((2 ** i) * math.pi * p), torch.cos((2 ** i) * math.pi * p)], dim=-1) for i in range(L)], dim=-1)

    def forward(self, x, styles):
        if not x.is_cuda:
            raise RuntimeError("x must be a CUDA tensor (this path has no CPU fallback)")
        input_pts, input_views = torch.split(x, [self.input_ch, self.input_ch_views], dim=-1)
        input_pts = self.transform_points(input_pts)
        input_views = self.transform_points(input_views, True)
        style_in = self.style_in(styles).view([styles.shape[0]] + [1] * (x.dim() - 2) + [-1])
        h = F.relu(self.x_in(input_pts) + style_in)
        for lin in self.pts_linears:
            h = F.relu(lin(h))
        sdf = self.sigma_linear(h)
        feat = self.views_linears(torch.cat([h, input_views], -1))
        out = [self.rgb_linear(feat), sdf]
        if self.output_features:
            out.append(feat)
        return torch.cat(out, -1)

    def forward_rays(self, npts, viewdirs, styles, want_rgb=True, want_feat=True, want_dsdf=False, feat_f16=False):
        B, R1, R2, S, _ = npts.shape
        if want_dsdf:
            npts = npts.detach().requires_grad_(True)
        x = torch.cat([npts, viewdirs[:, :, :, None, :].expand(B, R1, R2, S, 3)], -1)
        out = self.forward(x, styles)
        sdf = out[..., 3].reshape(-1)
        rgb = out[..., :3].reshape(-1, 3).contiguous()
        feat = out[..., 4:].reshape(-1, self.W).contiguous() if (want_feat and self.output_features) else None
        dsdf = None
        if want_dsdf:       # ref get_eikonal_term :224-229 (second-order capable: plain torch ops)
            dsdf = torch.autograd.grad(sdf, npts, torch.ones_like(sdf), create_graph=torch.is_grad_enabled())[0].reshape(-1, 3)
        return sdf, rgb, feat, dsdf


# --------------------------------------------------------------------------------------------------------------------------
# compositing as one autograd node

class _composite(Function):
    @staticmethod
    def forward(ctx, sdf, rgb, feat, sigmoid_beta, z_vals, rays_d, pts, noise, S, with_sdf, force_background, want_xyz):
        ctx.set_materialize_grads(False)
        sdf = sdf.contiguous()
        rgb = rgb.contiguous() if rgb is not None else None
        feat = feat.contiguous() if feat is not None else None
        rgb_map, feat_map, xyz, mask = ops.composite_forward(sdf, rgb, feat, z_vals, rays_d, pts if want_xyz else None, noise,
                                                             sigmoid_beta, S, with_sdf, force_background, want_xyz)
        ctx.save_for_backward(sdf, rgb, feat, sigmoid_beta, z_vals, rays_d, pts if want_xyz else None, noise)
        ctx.cfg = (S, with_sdf, force_background, want_xyz)
        empty = sdf.new_empty(0)
        outs = (rgb_map if rgb_map is not None else empty, feat_map if feat_map is not None else empty, xyz if xyz is not None else empty,
                mask if mask is not None else empty)
        if rgb_map is None:
            ctx.mark_non_differentiable(outs[0])
        if feat_map is None:
            ctx.mark_non_differentiable(outs[1])
        if not want_xyz:
            ctx.mark_non_differentiable(outs[2], outs[3])
        return outs

    @staticmethod
    @once_differentiable
    def backward(ctx, d_rgb_map, d_feat_map, d_xyz, d_mask):
        sdf, rgb, feat, sigmoid_beta, z_vals, rays_d, pts, noise = ctx.saved_tensors
        S, with_sdf, force_background, want_xyz = ctx.cfg
        cont = lambda t: None if t is None else t.contiguous()
        if feat is None:
            d_feat_map = None
        if not want_xyz:
            d_xyz, d_mask = None, None
        d_sdf, d_rgb, d_feat, d_beta = ops.composite_backward(sdf, rgb, feat, z_vals, rays_d, pts, noise, sigmoid_beta, S, with_sdf,
                                                              force_background, cont(d_rgb_map), cont(d_feat_map), cont(d_xyz), cont(d_mask),
                                                              want_d_feat=feat is not None and ctx.needs_input_grad[2])
        return d_sdf, d_rgb, d_feat, (d_beta if with_sdf and ctx.needs_input_grad[3] else None), None, None, None, None, None, None, None, None


# --------------------------------------------------------------------------------------------------------------------------

class VolumeFeatureRenderer(nn.Module):
    """Full volume renderer (ref :143-423): rays -> samples -> field -> SDF/density -> alpha compositing."""

    def __init__(self, opt, style_dim=256, out_im_res=64, mode="train"):
        super().__init__()
        self.test = mode != "train"
        self.perturb = opt.perturb
        self.offset_sampling = not opt.no_offset_sampling
        self.N_samples = opt.N_samples
        self.raw_noise_std = opt.raw_noise_std
        self.return_xyz = opt.return_xyz
        self.return_sdf = opt.return_sdf
        self.static_viewdirs = opt.static_viewdirs
        self.z_normalize = not opt.no_z_normalize
        self.out_im_res = out_im_res
        self.force_background = opt.force_background
        self.with_sdf = not opt.no_sdf
        self.output_features = "no_features_output" not in opt.keys()     # by key presence, as the reference (:156-159)
        # Not in the reference: sdf-only query for mesh extraction (sdf_mesh.py:138-182 keeps sdf / xyz / mask of a 128^3 frustum
        # and throws rgb and the 256 features away).  When set, the view layer, both rgb outputs and the feature map are skipped:
        # forward() returns None for them.  Inference only.
        self.sdf_only = False
        if self.with_sdf:
            self.sigmoid_beta = nn.Parameter(0.1 * torch.ones(1))

        lin = torch.linspace(0.5, out_im_res - 0.5, out_im_res)
        self.register_buffer("i", lin.view(1, 1, -1).expand(1, out_im_res, out_im_res).clone(), persistent=False)
        self.register_buffer("j", lin.view(1, -1, 1).expand(1, out_im_res, out_im_res).clone(), persistent=False)
        if self.offset_sampling:
            t_vals = torch.linspace(0., 1. - 1 / self.N_samples, steps=self.N_samples).view(1, 1, 1, -1)
        else:
            t_vals = torch.linspace(0., 1., steps=self.N_samples).view(1, 1, 1, -1)
        self.register_buffer("t_vals", t_vals, persistent=False)
        self.register_buffer("inf", torch.Tensor([1e10]), persistent=False)
        self.register_buffer("zero_idx", torch.LongTensor([0]), persistent=False)
        if self.test:
            self.perturb = False
            self.raw_noise_std = 0.
        self.channel_dim = -1
        self.samples_dim = 3
        self.input_ch = 3
        self.input_ch_views = 3
        self.feature_out_size = opt.width if not opt.type == "ngp" else style_dim

        if opt.type == "ngp":
            self.network = NGPSIRENGenerator(D=2, W=style_dim, style_dim=style_dim, output_features=self.output_features)
        elif opt.fc:
            self.network = FCGenerator(D=opt.depth, W=opt.width, style_dim=style_dim, input_ch=self.input_ch, output_ch=4,
                                       input_ch_views=self.input_ch_views, output_features=self.output_features)
        else:
            self.network = SirenGenerator(D=opt.depth, W=opt.width, style_dim=style_dim, input_ch=self.input_ch, output_ch=4,
                                          input_ch_views=self.input_ch_views, output_features=self.output_features)

    # -- sampling ---------------------------------------------------------------------------------------------------
    def _bounds(self, v, B, device):
        if not torch.is_tensor(v):
            v = torch.full((B,), float(v), device=device)
        return v.reshape(B).to(device=device, dtype=torch.float32)

    def _sample(self, cam_poses, focal, near, far, stratified=False, t_rand=None):
        B = cam_poses.shape[0]
        dev = cam_poses.device
        R, S = self.out_im_res, self.N_samples
        near, far = self._bounds(near, B, dev), self._bounds(far, B, dev)
        jitter = 0
        if stratified or (self.perturb > 0. and not self.offset_sampling):
            jitter = 2
            if t_rand is None:
                t_rand = torch.rand(B, R, R, S, device=dev)
        elif self.perturb > 0.:
            jitter = 1
            if t_rand is None:
                t_rand = torch.rand(B, R, R, device=dev)      # ref draws on the CPU and copies (:331); same distribution
        else:
            t_rand = None
        return ops.sample_rays(cam_poses[:, :3, :4], focal, near, far, self.t_vals, t_rand, jitter, self.static_viewdirs,
                               self.z_normalize, R, S), near, far

    def get_rays(self, focal, c2w):
        """ref :207-222 -- returns (rays_o, rays_d, viewdirs [un-normalised]) as [B,R,R,3]."""
        B = c2w.shape[0]
        r, _, _ = self._sample(c2w, focal, torch.ones(B, device=c2w.device), 2 * torch.ones(B, device=c2w.device))
        rays_o = c2w[:, None, None, :3, -1].expand(r["rays_d"].shape)
        if self.static_viewdirs:
            dirs = torch.stack([(self.i - self.out_im_res * .5) / focal, -(self.j - self.out_im_res * .5) / focal,
                                -torch.ones_like(self.i).expand(B, self.out_im_res, self.out_im_res)], -1)
            return rays_o, r["rays_d"], dirs
        return rays_o, r["rays_d"], r["rays_d"]

    # -- rendering --------------------------------------------------------------------------------------------------
    def render(self, focal, c2w, near, far, styles, c2w_staticcam=None, return_eikonal=False, t_rand=None):
        """ref :363-378 + :310-361; returns channel-last maps (rgb [B,R,R,3], features [B,R,R,W], sdf [B,R,R,S,1], mask, xyz, eikonal)."""
        B = c2w.shape[0]
        R, S = self.out_im_res, self.N_samples
        smp, near, far = self._sample(c2w, focal, near, far, t_rand=t_rand)
        want_eik = bool(return_eikonal and self.with_sdf)
        want_rgb = not self.sdf_only
        if self.sdf_only and torch.is_grad_enabled():
            raise RuntimeError("renderer.sdf_only is an inference path: build the Generator with ema=True (or call the renderer under "
                               "torch.no_grad()) -- note that Generator.forward re-enables grad for a training-mode generator")
        sdf, rgb, feat, dsdf = self.network.forward_rays(smp["npts"], smp["viewdirs"], styles, want_rgb=want_rgb,
                                                         want_feat=self.output_features and want_rgb, want_dsdf=want_eik, feat_f16=True)
        noise = None
        if (not self.with_sdf) and self.raw_noise_std > 0.:
            noise = torch.randn_like(sdf) * self.raw_noise_std
        sb = self.sigmoid_beta if self.with_sdf else None
        rgb_map, feat_map, xyz, mask = _composite.apply(sdf, rgb, feat, sb, smp["z_vals"].reshape(-1), smp["rays_d"].reshape(-1, 3),
                                                        smp["pts"].reshape(-1, 3), noise, S, self.with_sdf, self.force_background,
                                                        bool(self.return_xyz))
        rgb_map = rgb_map.view(B, R, R, 3) if want_rgb else None
        feat_map = feat_map.view(B, R, R, -1) if (self.output_features and want_rgb) else None
        sdf_out = sdf.view(B, R, R, S, 1) if self.return_sdf else None
        if self.return_xyz:
            xyz, mask = xyz.view(B, R, R, 3), mask.view(B, R, R, 1)
        else:
            xyz, mask = None, None
        eik = None
        if want_eik:
            scale = (2.0 / (far - near)).view(B, 1, 1, 1, 1) if self.z_normalize else 1.0
            eik = dsdf.view(B, R, R, S, 3) * scale
        return rgb_map, feat_map, sdf_out, mask, xyz, eik

    def mlp_init_pass(self, cam_poses, focal, near, far, styles=None, t_rand=None):
        """ref :380-409 -- stratified samples, returns (sdf [B,R,R,S], |pts| - (far-near)/4)."""
        B = cam_poses.shape[0]
        R, S = self.out_im_res, self.N_samples
        smp, near, far = self._sample(cam_poses, focal, near, far, stratified=True, t_rand=t_rand)
        sdf, _, _, _ = self.network.forward_rays(smp["npts"], smp["viewdirs"], styles, want_rgb=False, want_feat=False)
        target = smp["pts"].norm(dim=-1) - ((far - near) / 4).view(B, 1, 1, 1)
        return sdf.view(B, R, R, S), target

    def forward(self, cam_poses, focal, near, far, styles=None, return_eikonal=False, t_rand=None):
        rgb, features, sdf, mask, xyz, eikonal_term = self.render(focal, c2w=cam_poses, near=near, far=far, styles=styles,
                                                                  return_eikonal=return_eikonal, t_rand=t_rand)
        if rgb is not None:
            rgb = rgb.permute(0, 3, 1, 2).contiguous()
        if features is not None:
            # logically [B, W, R, R] like the reference (:415-421); the memory stays channels-last (what the compositing kernel wrote and
            # what the decoder kernels read): no transposing copy in either direction
            features = features.permute(0, 3, 1, 2)
        if xyz is not None:
            xyz = xyz.permute(0, 3, 1, 2).contiguous()
            mask = mask.permute(0, 3, 1, 2).contiguous()
        return rgb, features, sdf, mask, xyz, eikonal_term


class MappingLinear(nn.Module):
    """ref :437-466 (fused_leaky_relu with scale = 1 is leaky_relu(x + b, 0.2))."""

    def __init__(self, in_dim, out_dim, bias=True, activation=None, is_last=False):
        super().__init__()
        weight_std = 0.25 if is_last else 1
        self.weight = nn.Parameter(weight_std * nn.init.kaiming_normal_(torch.empty(out_dim, in_dim), a=0.2, mode="fan_in",
                                                                        nonlinearity="leaky_relu"))
        self.bias = nn.Parameter(nn.init.uniform_(torch.empty(out_dim), a=-np.sqrt(1 / in_dim), b=np.sqrt(1 / in_dim))) if bias else None
        self.activation = activation

    def forward(self, input):
        if self.activation is not None:
            return F.leaky_relu(F.linear(input, self.weight) + self.bias, negative_slope=0.2)
        return F.linear(input, self.weight, bias=self.bias)

    def __repr__(self):
        return f"{self.__class__.__name__}({self.weight.shape[1]}, {self.weight.shape[0]})"


_DECODER_FACTORY = None


def register_decoder(factory):
    """Use another decoder class than this package's (decoder.Decoder, forward-only kernels) -- e.g. the host project's own
    torch StyleGAN2 Decoder for stage-2 training: factory(model_opt) -> nn.Module with the reference Decoder's
    forward / mean_latent signature (ref :883-1056).  None restores the default."""
    global _DECODER_FACTORY
    _DECODER_FACTORY = factory


class Generator(nn.Module):
    """Mapping network + volume renderer (+ decoder) with the reference signature (ref :1059-1216)."""

    def __init__(self, model_opt, renderer_opt, blur_kernel=[1, 3, 3, 1], ema=False, full_pipeline=True):
        super().__init__()
        self.size = model_opt.size
        self.style_dim = model_opt.style_dim * 2 if model_opt.psp else model_opt.style_dim
        self.num_layers = 1
        self.train_renderer = not model_opt.freeze_renderer
        self.full_pipeline = full_pipeline
        model_opt.feature_encoder_in_channels = renderer_opt.width
        self.is_train = not (ema or "is_test" in model_opt.keys())
        self.style = nn.Sequential(*[MappingLinear(self.style_dim, self.style_dim, activation="fused_lrelu") for _ in range(3)])
        self.renderer = VolumeFeatureRenderer(renderer_opt, style_dim=self.style_dim, out_im_res=model_opt.renderer_spatial_output_dim)
        if self.full_pipeline:
            if _DECODER_FACTORY is not None:
                self.decoder = _DECODER_FACTORY(model_opt)
            else:
                from .decoder import Decoder
                self.decoder = Decoder(model_opt)

    def mean_latent(self, n_latent, device, z=None):
        if z is None:
            renderer_latent = self.style(torch.randn(n_latent, self.style_dim, device=device))
            renderer_latent_mean = renderer_latent.mean(0, keepdim=True)
        else:
            renderer_latent = None
            renderer_latent_mean = self.style(z)
        decoder_latent_mean = self.decoder.mean_latent(renderer_latent) if self.full_pipeline else None
        return [renderer_latent_mean, decoder_latent_mean]

    def get_latent(self, input):
        return self.style(input)

    def styles_and_noise_forward(self, styles, inject_index=None, truncation=1, truncation_latent=None, input_is_latent=False):
        if not input_is_latent:
            styles = [self.style(s) for s in styles]
        if truncation < 1:
            styles = [truncation_latent[0] + truncation * (s - truncation_latent[0]) for s in styles]
        return styles

    def init_forward(self, styles, cam_poses, focals, near=0.88, far=1.12, t_rand=None):
        latent = self.styles_and_noise_forward(styles)
        return self.renderer.mlp_init_pass(cam_poses, focals, near, far, styles=latent[0], t_rand=t_rand)

    def forward(self, styles, cam_poses, focals, near=0.88, far=1.12, return_latents=False, inject_index=None, truncation=1,
                truncation_latent=None, input_is_latent=False, noise=None, randomize_noise=True, return_sdf=False, return_xyz=False,
                return_eikonal=False, project_noise=False, mesh_path=None, t_rand=None):
        with torch.set_grad_enabled(self.is_train and self.train_renderer):
            latent = self.styles_and_noise_forward(styles, inject_index, truncation, truncation_latent, input_is_latent)
            style0 = latent[0][:, 0] if input_is_latent else latent[0]
            thumb_rgb, features, sdf, mask, xyz, eikonal_term = self.renderer(cam_poses, focals, near, far, styles=style0,
                                                                              return_eikonal=return_eikonal, t_rand=t_rand)
        if self.full_pipeline:
            rgb, decoder_latent = self.decoder(features, latent, transform=cam_poses if project_noise else None,
                                               return_latents=return_latents, inject_index=inject_index, truncation=truncation,
                                               truncation_latent=truncation_latent, noise=noise, input_is_latent=input_is_latent,
                                               randomize_noise=randomize_noise, mesh_path=mesh_path)
        else:
            rgb = None
        if return_latents:
            return rgb, decoder_latent
        out = (rgb, thumb_rgb)
        if return_xyz:
            out += (xyz,)
        if return_sdf:
            out += (sdf,)
        if return_eikonal:
            out += (eikonal_term,)
        if return_xyz:
            out += (mask,)
        return out
