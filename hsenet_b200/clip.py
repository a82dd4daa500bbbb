"""CLIP-side pieces of the hot path (reference ``model/CLIP_stage1.py``):

    clip_image_head   <- encode_image tail + ``[:, 0]``   (CLIP_stage1.py:100-101, 117)
    contrastive_logits <- image_text_contrastive_learning  (CLIP_stage1.py:141-155), using gather_features
"""
from __future__ import annotations

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import _lib
from . import runtime as rt
from .dist_utils import gather_features


class ClipImageHead(nn.Module):
    """``mm_vision_proj`` + L2 normalise + cls row.  Only the cls row is projected (the reference projects all
    2049 rows and discards 2048 of them; row 0 is numerically the same)."""

    def __init__(self, hidden_size: int = 768):
        super().__init__()
        if hidden_size != 768:
            raise ValueError("hidden_size must be 768")
        self.mm_vision_proj = nn.Linear(hidden_size, hidden_size)
        self._cache = rt.WeightCache()

    def forward(self, tokens: torch.Tensor) -> torch.Tensor:
        return clip_image_head(tokens, self.mm_vision_proj, self._cache)


def clip_image_head(tokens: torch.Tensor, proj: nn.Linear, cache: rt.WeightCache | None = None) -> torch.Tensor:
    """tokens [B,2049,768] (act dtype of the current precision) -> unit-norm cls embeddings fp32 [B,768]."""
    rt.require_cuda(tokens, "tokens")
    if torch.is_grad_enabled() and (tokens.requires_grad or any(p.requires_grad for p in proj.parameters())):
        # training (CLIP_stage1.py:100-101,117): 1.2 MFLOP on the cls row -- plain autograd ops, like the logits
        cls = tokens[:, 0].float()
        return F.normalize(F.linear(cls, proj.weight.float(), proj.bias.float()), dim=-1)
    prec = rt.get_precision()
    act = rt.act_dtype(prec)
    if tokens.dim() != 3 or tokens.shape[1] != 2049 or tokens.shape[2] != 768:
        raise ValueError(f"expected tokens [B,2049,768], got {tuple(tokens.shape)}")
    t = tokens.detach()
    if t.dtype != act or not t.is_contiguous():
        t = t.to(act).contiguous()
    build = lambda p: {"w": rt.cast_weight(proj.weight, p), "b": rt.f32(proj.bias)}
    B = t.shape[0]
    out = torch.empty(B, 768, dtype=torch.float32, device=t.device)
    ws = rt.workspace(t.device, B * 768 * 4, "clip_head")
    with torch.cuda.device(t.device):
        pl = cache.get(proj.parameters(), prec, build) if cache is not None else build(prec)
        rc = _lib.load().hsenet_clip_image_head(t.data_ptr(), pl["w"].data_ptr(), pl["b"].data_ptr(), B,
                                                rt.precision_code(prec), out.data_ptr(), ws.data_ptr(), ws.numel(),
                                                rt.stream_ptr(t.device))
    _lib.check(rc, "clip_image_head")
    return out


def contrastive_logits(image_features, text_features, logit_scale, labels=None, gather_loss=True, local_loss=False,
                       rank=0, world_size=1):
    """CLIP_stage1.py:141-155.  The [B,B] logits / cross-entropy are tiny and stay in PyTorch (plumbing); the
    exchange step is the packed all-gather of hsenet_b200.dist_utils.gather_features."""
    if gather_loss:
        all_i, all_t = gather_features(image_features, text_features, local_loss=local_loss, rank=rank,
                                       world_size=world_size)
        if local_loss:
            lpi = logit_scale * image_features @ all_t.T
            lpt = logit_scale * text_features @ all_i.T
        else:
            lpi = logit_scale * all_i @ all_t.T
            lpt = lpi.T
    else:
        lpi = logit_scale * image_features @ text_features.T
        lpt = logit_scale * text_features @ image_features.T
    loss = None
    if labels is not None:
        loss = (F.cross_entropy(lpi, labels) + F.cross_entropy(lpt, labels)) / 2
    return loss, lpi, lpt
