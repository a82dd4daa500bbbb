"""``gather_features`` -- drop-in for reference ``utils/dist_utils.py:280-306``.

The reference issues two autograd-aware all-gathers of [B_loc,768] per call (image, text).  The payload is tiny
(<= 196 KB/rank at B=256 over 8 GPUs) so the exchange is latency-bound: here image and text embeddings are packed
into one [B_loc, 2, D] buffer and exchanged with a SINGLE all_gather_into_tensor (NCCL over NVLink on the GPU box,
gloo in the CPU tests); the backward of the gather is one reduce-scatter of the packed gradient.
Same signature, same return values (rank-ordered concatenation), same gradient semantics.
"""
from __future__ import annotations

import torch

try:
    import torch.distributed as dist
    has_distributed = True
except ImportError:  # pragma: no cover
    dist = None
    has_distributed = False


def _world():
    if has_distributed and dist.is_available() and dist.is_initialized():
        return dist.get_world_size(), dist.get_rank()
    return 1, 0


class _PackedAllGather(torch.autograd.Function):
    @staticmethod
    def forward(ctx, packed):
        world, rank = _world()
        ctx.world, ctx.rank = world, rank
        out = torch.empty((world,) + tuple(packed.shape), dtype=packed.dtype, device=packed.device)
        dist.all_gather_into_tensor(out.view(-1), packed.contiguous().view(-1))
        return out

    @staticmethod
    def backward(ctx, grad):
        grad = grad.contiguous()
        world, rank = ctx.world, ctx.rank
        if dist.get_backend() == "nccl":
            mine = torch.empty_like(grad[0])
            dist.reduce_scatter_tensor(mine.view(-1), grad.view(-1), op=dist.ReduceOp.SUM)
            return mine
        dist.all_reduce(grad, op=dist.ReduceOp.SUM)      # gloo has no reduce_scatter
        return grad[rank].clone()


def gather_features(image_features, text_features, local_loss=False, gather_with_grad=True, rank=0, world_size=1):
    """Rank-ordered concatenation of every rank's image / text features (dist_utils.py:280-306).

    ``rank`` / ``world_size`` are accepted for signature compatibility; like every reference call site (which passes
    the process group's own values, CLIP_stage1.py:143) the exchange uses the initialised process group's rank and
    size.  Each returned tensor keeps its input's dtype (mixed dtypes -- fp32 image features from the cls head with
    bf16 text features under autocast -- travel as the wider one and are cast back)."""
    assert has_distributed, 'torch.distributed did not import correctly, please use a PyTorch version with support.'
    world, my_rank = _world()
    if world == 1:
        # the reference bootstraps a 1-process group and gathers a single block (train_CLIP_stage1.py:29-38)
        return image_features, text_features
    if image_features.dim() != 2 or image_features.shape != text_features.shape:
        raise ValueError(f"gather_features packs image and text features into one buffer: both must be [B_loc, D] of "
                         f"the same shape, got {tuple(image_features.shape)} and {tuple(text_features.shape)}")
    b = image_features.shape[0]
    dt_i, dt_t = image_features.dtype, text_features.dtype
    wire = torch.promote_types(dt_i, dt_t)
    packed = torch.stack([image_features.to(wire), text_features.to(wire)], dim=1)          # [B_loc, 2, D]
    if gather_with_grad:
        allp = _PackedAllGather.apply(packed)                             # [W, B_loc, 2, D]
    else:
        with torch.no_grad():
            allp = _PackedAllGather.apply(packed.detach())
        if not local_loss:
            # keep the local block differentiable, as dist_utils.py:300-303 does
            allp = allp.clone()
            allp[my_rank] = packed
    allp = allp.reshape(world * b, 2, -1)
    return allp[:, 0].to(dt_i), allp[:, 1].to(dt_t)
