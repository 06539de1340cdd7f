"""Data-parallel plumbing for the SDF generator path: one process per GPU, units (images) sharded by rank, no data-path
collective; training adds ONE exchange step, the gradient average.

The reference never initialises a process group (SURVEY.md finding 2: `opt.distributed` is derived from WORLD_SIZE,
im2scene/training_utils.py:186-187, but nothing is wrapped); this module is the missing piece.  `DistributedDataParallel` works
unchanged on the package's modules (their custom autograd nodes deposit gradients on ordinary parameters); `average_gradients`
is the explicit, bucketed alternative used when the caller accumulates several micro-batches before exchanging.
"""
import os

import torch
import torch.distributed as dist


def shard_range(total, rank, world):
    """Contiguous [start, stop) of `total` units owned by `rank` (remainder spread over the first ranks)."""
    base, rem = divmod(total, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def shard_batch(tensors, rank=None, world=None):
    """Slice every tensor of a (latents, cam_poses, focal, near, far) tuple along dim 0 for this rank."""
    rank = dist.get_rank() if rank is None else rank
    world = dist.get_world_size() if world is None else world
    a, b = shard_range(tensors[0].shape[0], rank, world)
    return tuple(t[a:b] for t in tensors)


@torch.no_grad()
def sync_parameters(module, src=0):
    """Broadcast parameters and buffers from `src` so that every rank starts from identical weights."""
    for t in list(module.parameters()) + list(module.buffers()):
        dist.broadcast(t.data, src)


@torch.no_grad()
def average_gradients(module, bucket_bytes=64 << 20):
    """All-reduce (mean) every existing gradient, flattened into buckets of ~bucket_bytes so that the 50.6 MB hash-table gradient
    and the ~4 MB of MLP gradients travel in two or three NCCL calls instead of 38."""
    world = dist.get_world_size()
    if world == 1:
        return 0
    params = [p for p in module.parameters() if p.grad is not None]
    calls, i = 0, 0
    while i < len(params):
        bucket, size = [], 0
        while i < len(params) and (not bucket or size + params[i].grad.numel() * params[i].grad.element_size() <= bucket_bytes):
            if bucket and params[i].grad.dtype != bucket[0].grad.dtype:
                break
            bucket.append(params[i])
            size += params[i].grad.numel() * params[i].grad.element_size()
            i += 1
        flat = torch.cat([p.grad.reshape(-1) for p in bucket])
        dist.all_reduce(flat, op=dist.ReduceOp.SUM)
        flat.div_(world)
        off = 0
        for p in bucket:
            n = p.grad.numel()
            p.grad.copy_(flat[off:off + n].view_as(p.grad))
            off += n
        calls += 1
    return calls


def data_parallel(module, device_ids=None, early_table_exchange=True, process_group=None, **ddp_kwargs):
    """Wrap a Generator (or any module holding the SDF renderer) in DistributedDataParallel for one-process-per-GPU training.

    With `early_table_exchange` (default) the hash-table parameter(s) (`*.encoder.embeddings`, 93 % of the bytes exchanged in
    stage 1) are taken out of DDP's buckets: autograd produces that gradient last, so its bucket could not overlap any compute.
    The field's backward node all-reduces (averages) it itself right after the scatter kernel is enqueued, while the
    weight-gradient kernels run behind it on the same stream (sdf_model._field.backward).  The switch is an attribute of the
    wrapped module's field networks (`_table_exchange`), not process-global state: other models in the process, and backward
    passes that only some ranks run, are unaffected.  The parameter is broadcast from rank 0 here, as DDP would have done.
    Both autograd nodes that produce a table gradient exchange it in that mode: the renderer's fused field node and
    `GridEncoder`'s own node (`network.query_sdf`, the reference's smoothness term, smoothLoss.py:5-25)."""
    ddp = torch.nn.parallel.DistributedDataParallel
    world = dist.get_world_size(process_group)
    names = [n for n, _ in module.named_parameters() if n.endswith("encoder.embeddings")] if early_table_exchange else []
    if names and world > 1:
        ddp._set_params_and_buffers_to_ignore_for_model(module, names)
        lookup = dict(module.named_parameters())
        mods = dict(module.named_modules())
        with torch.no_grad():
            for n in names:
                dist.broadcast(lookup[n].data, 0, group=process_group)
                owner = n[:-len("encoder.embeddings")].rstrip(".")
                mods[owner]._table_exchange = {"group": process_group}                       # the field network (sdf_model._field)
                mods[(owner + "." if owner else "") + "encoder"]._table_exchange = {"group": process_group}   # GridEncoder's own node (query_sdf)
    # (with the table handled above, 4 MB of MLP / mapping gradients remain) small buckets: all but the last leave while backward
    # still runs; buffers are constants (level offsets, pixel grids): no per-forward broadcast
    ddp_kwargs.setdefault("bucket_cap_mb", int(os.environ.get("SDFG_DDP_BUCKET_MB", "2")) if names else 64)
    ddp_kwargs.setdefault("gradient_as_bucket_view", True)
    ddp_kwargs.setdefault("broadcast_buffers", False)
    if process_group is not None:
        ddp_kwargs.setdefault("process_group", process_group)
    return ddp(module, device_ids=device_ids, **ddp_kwargs)
