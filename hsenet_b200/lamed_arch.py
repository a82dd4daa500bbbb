"""The VLM-side contract of the hot path: ``encode_images`` (reference ``model/lamed_arch.py:122-141``).

``encode_images(model, images, text_feature, images_2d)`` can be bound onto an unmodified LamedMetaForCausalLM
(see INTEGRATION.md); ``HSENetVisualEncoder`` bundles tower + packers for stand-alone use (bench.py, tests).
Both write the two packers' outputs straight into one [B,256,out_dim] buffer instead of torch.cat.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from . import runtime as rt
from .builder import build_mm_projector, build_vision_tower


def encode_images_with(tower, mm_projector, mm_projector2, images, images_2d, out_dtype=None):
    feats = tower(images, images_2d)
    if isinstance(feats, tuple):
        if feats[-1].shape[1] != 2048:
            raise ValueError("dual-tower encode_images expects select_feature == 'patch' (2048 tokens per tower), "
                             "as the reference dispatch at lamed_arch.py:125 does")
        f1, f2 = feats
        second = mm_projector2 if mm_projector2 is not None else mm_projector   # fallback of lamed_arch.py:128-131
        n1, n2 = mm_projector.proj_out_num, second.proj_out_num
        dt = out_dtype or mm_projector.output_dtype or rt.act_dtype()
        out = torch.empty(f1.shape[0], n1 + n2, mm_projector.out_dim, dtype=dt, device=f1.device)
        mm_projector.forward_into(f1, out, 0)
        second.forward_into(f2, out, n1)
        return out
    return mm_projector(feats)


def encode_images(self, images, text_feature=None, images_2d=None):
    """Method-compatible replacement for LamedMetaForCausalLM.encode_images (lamed_arch.py:122)."""
    m = self.get_model()
    out_dtype = text_feature.dtype if isinstance(text_feature, torch.Tensor) and text_feature.is_floating_point() \
        and text_feature.dtype in (torch.float32, torch.bfloat16) else None
    return encode_images_with(m.get_vision_tower(), m.mm_projector, getattr(m, "mm_projector2", None), images,
                              images_2d, out_dtype)


class HSENetVisualEncoder(nn.Module):
    """Dual tower + the two spatial packers, assembled the way LamedMetaModel.initialize_vision_modules does
    (lamed_arch.py:41-84).  ``forward(images, images_2d) -> [B,256,out_dim]``."""

    def __init__(self, config, use_parallel_projector: bool = True):
        super().__init__()
        self.config = config
        self.vision_tower = build_vision_tower(config)
        config.mm_hidden_size = self.vision_tower.hidden_size
        self.mm_projector = build_mm_projector(config)
        if use_parallel_projector:
            self.mm_projector2 = build_mm_projector(config)

    def get_model(self):
        return self

    def get_vision_tower(self):
        return self.vision_tower

    def encode_images(self, images, text_feature=None, images_2d=None):
        return encode_images(self, images, text_feature, images_2d)

    def forward(self, images, images_2d):
        return self.encode_images(images, None, images_2d)


class VisionConfig:
    """Attribute bag with the fields the factories read (train/train_VLM.py:76-99 defaults)."""

    def __init__(self, out_dim: int = 3072, select_feature: str = "patch", remain: str = "dual_vits"):
        self.image_channel = 1
        self.image_size = (32, 256, 256)
        self.patch_size = (4, 16, 16)
        self.vision_tower = "vit_stage2_dual_encoders"
        self.vision_select_layer = -1
        self.vision_select_feature = select_feature
        self.remain_2d3d_ViT_type = remain
        self.mm_projector_type = "VisualPacker_3d_phi_v3"
        self.mm_hidden_size = 768
        self.hidden_size = out_dim
        self.proj_layer_type = "mlp"
        self.proj_layer_num = 2
        self.proj_pooling_type = "spatial"
        self.proj_pooling_size = 2
