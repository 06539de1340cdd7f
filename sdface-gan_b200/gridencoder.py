"""Multi-resolution hash-grid encoder with the reference's module / function API.

Mirrors /root/reference/im2scene/sdf/models/gridencoder/grid.py: `_grid_encode` (:24-93), `grid_encode` (:93),
`GridEncoder` (:96-185) -- same constructor arguments, `state_dict` keys (`embeddings`, `offsets`), level table and
call signatures -- on top of the sm_100a kernels in csrc/hashgrid.cu.  Differences that do not change results:
  * features are produced sample-major [B, L*C] directly (no [L,B,C] buffer + permute copy, grid.py:47,57);
  * `GridEncoder.forward` folds the affine map (x+bound)/(2*bound) (grid.py:149) into the kernel;
  * the backward scatters straight into a zero buffer that torch then adds to `.grad` (same contract as grid.py:77);
  * launches go to torch's current stream (the reference uses the legacy default stream).
"""
import numpy as np
import torch
import torch.nn as nn
from torch.autograd import Function
from torch.autograd.function import once_differentiable

from . import ops

_gridtype_to_id = {"hash": 0, "tiled": 1}
_interp_to_id = {"linear": 0, "smoothstep": 1}


class _grid_encode(Function):
    @staticmethod
    def forward(ctx, inputs, embeddings, offsets, per_level_scale, base_resolution, calc_grad_inputs=False, gridtype=0,
                align_corners=False, interpolation=0, bound=0.0, exchange=None):
        # inputs [B, D] in [0,1] (or in [-bound, bound] when bound > 0); embeddings [sO, C]; offsets [L+1] int32 -> [B, L*C]
        inputs = inputs.contiguous().float()
        S = ops.log2_scale(per_level_scale)
        outputs, dy_dx = ops.grid_encode_forward(inputs, embeddings.contiguous(), offsets, S, base_resolution, bound=bound,
                                                 calc_dy_dx=calc_grad_inputs, gridtype=gridtype, align_corners=align_corners,
                                                 interp=interpolation)
        ctx.save_for_backward(inputs, embeddings, offsets, dy_dx)
        ctx.meta = (S, base_resolution, gridtype, align_corners, interpolation, bound)
        ctx.exchange = exchange
        return outputs

    @staticmethod
    @once_differentiable
    def backward(ctx, grad):
        inputs, embeddings, offsets, dy_dx = ctx.saved_tensors
        S, H, gridtype, align_corners, interpolation, bound = ctx.meta
        grad = grad.contiguous()
        grad_embeddings = torch.zeros_like(embeddings) if ctx.needs_input_grad[1] else None
        want_gi = dy_dx is not None and ctx.needs_input_grad[0]
        _, grad_inputs = ops.grid_encode_backward(grad, inputs, embeddings, offsets, S, H, bound=bound, dy_dx=dy_dx if want_gi else None,
                                                  grad_embeddings=grad_embeddings, want_grad_inputs=want_gi, gridtype=gridtype,
                                                  align_corners=align_corners, interp=interpolation)
        ex = ctx.exchange
        if ex is not None and grad_embeddings is not None and torch.distributed.is_available() and torch.distributed.is_initialized():
            # data-parallel training with the table outside DistributedDataParallel's buckets (distributed.data_parallel)
            world = torch.distributed.get_world_size(ex["group"])
            if world > 1:
                torch.distributed.all_reduce(grad_embeddings, op=torch.distributed.ReduceOp.SUM, group=ex["group"])
                grad_embeddings.div_(world)
        return grad_inputs, grad_embeddings, None, None, None, None, None, None, None, None, None


grid_encode = _grid_encode.apply


class GridEncoder(nn.Module):
    def __init__(self, input_dim=3, num_levels=16, level_dim=2, per_level_scale=2, base_resolution=16, log2_hashmap_size=19,
                 desired_resolution=None, gridtype="hash", align_corners=False, interpolation="linear"):
        super().__init__()
        if desired_resolution is not None:   # finest resolution overrides per_level_scale (grid.py:101-102)
            per_level_scale = np.exp2(np.log2(desired_resolution / base_resolution) / (num_levels - 1))
        self.input_dim = input_dim
        self.num_levels = num_levels
        self.level_dim = level_dim
        self.per_level_scale = per_level_scale
        self.log2_hashmap_size = log2_hashmap_size
        self.base_resolution = base_resolution
        self.output_dim = num_levels * level_dim
        self.gridtype = gridtype
        self.gridtype_id = _gridtype_to_id[gridtype]
        self.interpolation = interpolation
        self.interp_id = _interp_to_id[interpolation]
        self.align_corners = align_corners
        self._table_exchange = None       # distributed.data_parallel(early_table_exchange=True) sets {"group": ...}

        # level table (grid.py:117-131): entries per level capped at 2^log2_hashmap_size and rounded up to a multiple of 8
        self.max_params = 2 ** log2_hashmap_size
        offsets, offset = [], 0
        for i in range(num_levels):
            resolution = int(np.ceil(base_resolution * per_level_scale ** i))
            n = min(self.max_params, (resolution if align_corners else resolution + 1) ** input_dim)
            n = int(np.ceil(n / 8) * 8)
            offsets.append(offset)
            offset += n
        offsets.append(offset)
        self.register_buffer("offsets", torch.from_numpy(np.array(offsets, dtype=np.int32)))
        self.n_params = offsets[-1] * level_dim
        self.embeddings = nn.Parameter(torch.empty(offset, level_dim))
        self.reset_parameters()

    def reset_parameters(self):
        self.embeddings.data.uniform_(-1e-4, 1e-4)      # grid.py:138-140

    def __repr__(self):
        return (f"GridEncoder: input_dim={self.input_dim} num_levels={self.num_levels} level_dim={self.level_dim} "
                f"resolution={self.base_resolution} -> {int(round(self.base_resolution * self.per_level_scale ** (self.num_levels - 1)))} "
                f"per_level_scale={self.per_level_scale:.4f} params={tuple(self.embeddings.shape)} gridtype={self.gridtype} "
                f"align_corners={self.align_corners} interpolation={self.interpolation}")

    def forward(self, inputs, bound=1):
        # inputs [..., input_dim] in [-bound, bound] -> [..., num_levels * level_dim]
        prefix_shape = list(inputs.shape[:-1])
        flat = inputs.reshape(-1, self.input_dim)
        outputs = grid_encode(flat, self.embeddings, self.offsets, self.per_level_scale, self.base_resolution, flat.requires_grad,
                              self.gridtype_id, self.align_corners, self.interp_id, float(bound), self._table_exchange)
        return outputs.view(prefix_shape + [self.output_dim])

    @torch.no_grad()
    def grad_total_variation(self, weight=1e-7, inputs=None, bound=1, B=1000000):
        # adds the TV gradient at the cells of `inputs` into embeddings.grad (grid.py:165-185)
        if inputs is None:
            inputs = torch.rand(B, self.input_dim, device=self.embeddings.device)
        else:
            inputs = ((inputs + bound) / (2 * bound)).reshape(-1, self.input_dim).contiguous()
        if self.embeddings.grad is None:
            raise ValueError("grad is None, should be called after loss.backward() and before optimizer.step()!")
        ops.grad_total_variation(inputs.float(), self.embeddings.detach(), self.embeddings.grad, self.offsets, weight,
                                 ops.log2_scale(self.per_level_scale), self.base_resolution, self.gridtype_id, self.align_corners)
