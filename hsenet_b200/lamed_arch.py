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
    return pack_features(tower(images, images_2d), mm_projector, mm_projector2, out_dtype)


def pack_features(feats, mm_projector, mm_projector2, out_dtype=None):
    """Tower features -> LLM-space visual tokens (lamed_arch.py:125-139)."""
    if isinstance(feats, tuple):
        if feats[-1].shape[1] != 2048:
            raise ValueError("dual-tower encode_images expects select_feature == 'patch' (2048 tokens per tower), "
                             "as the reference dispatch at lamed_arch.py:125 does")
        f1, f2 = feats
        second = mm_projector2 if mm_projector2 is not None else mm_projector   # fallback of lamed_arch.py:128-131
        n1, n2 = mm_projector.proj_out_num, second.proj_out_num
        dt = out_dtype or mm_projector.output_dtype or rt.act_dtype()
        drops = any(getattr(m, "_dropout_active", lambda: False)() for m in (mm_projector, second))
        if drops or (torch.is_grad_enabled() and (f1.requires_grad or f2.requires_grad or any(
                p.requires_grad for m in (mm_projector, second) for p in m.parameters()))):
            # training (or train()-mode dropout): the packers run through their autograd Functions; the concatenation is
            # the reference's
            return torch.cat([mm_projector(f1), second(f2)], dim=1).to(dt)
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


def prepare_inputs_for_multimodal(self, input_ids, position_ids, attention_mask, past_key_values, labels, images,
                                  images_2d):
    """Method-compatible replacement for LamedMetaForCausalLM.prepare_inputs_for_multimodal (lamed_arch.py:143-155).

    The reference builds ``cat(embeds[:, :1], image_features, embeds[:, n+1:])`` (two full-tensor copies of
    [B, L, 3072]); here the two packers' GEMM epilogues write their 128 tokens each directly into positions
    1..256 of a copy of ``inputs_embeds``.  Same early-out as the reference when there
    is no tower / no image / a single decode token (lamed_arch.py:148)."""
    vision_tower = self.get_vision_tower()
    if vision_tower is None or images is None or input_ids.shape[1] == 1:
        return input_ids, position_ids, attention_mask, past_key_values, None, labels
    m = self.get_model()
    inputs_embeds = m.embed_tokens(input_ids)
    inputs_embeds = splice_visual_tokens(m.get_vision_tower(), m.mm_projector, getattr(m, "mm_projector2", None),
                                         inputs_embeds, images, images_2d)
    return None, position_ids, attention_mask, past_key_values, inputs_embeds, labels


def splice_visual_tokens(tower, mm_projector, mm_projector2, inputs_embeds, images, images_2d):
    """inputs_embeds [B, L, D] -> new tensor whose positions 1 .. 256 hold the visual tokens."""
    feats = tower(images, images_2d)
    if not (isinstance(feats, tuple) and feats[-1].shape[1] == 2048):
        # single-tower configurations: keep the reference's formulation
        f = mm_projector(feats)
        return torch.cat((inputs_embeds[:, :1, :], f.to(inputs_embeds.dtype), inputs_embeds[:, f.shape[1] + 1:, :]), 1)
    second = mm_projector2 if mm_projector2 is not None else mm_projector
    n1, n2 = mm_projector.proj_out_num, second.proj_out_num
    B, L, D = inputs_embeds.shape
    if L < 1 + n1 + n2:
        raise ValueError(f"sequence of {L} tokens cannot hold {n1 + n2} visual tokens after the first token")
    if D != mm_projector.out_dim:
        raise ValueError(f"embedding width {D} != packer out_dim {mm_projector.out_dim}")
    # The in-place epilogue writes below are invisible to autograd.  Whenever a gradient has to flow -- into the
    # embedding table (LoRA / trainable embed_tokens, incl. the added vision tokens) or into trainable packers -- build
    # the result differentiably instead: a NON-detached clone with the visual tokens assigned into their slot, which
    # has exactly the gradient of the reference's torch.cat (lamed_arch.py:153-154).
    needs_grad = torch.is_grad_enabled() and (
        inputs_embeds.requires_grad or any(p.requires_grad for m in (tower, mm_projector, second)
                                           for p in m.parameters()))
    # packers in train() mode apply dropout, which only their module forward (not forward_into) implements
    drops = any(getattr(m, "_dropout_active", lambda: False)() for m in (mm_projector, second))
    if needs_grad or drops:
        vis = pack_features(feats, mm_projector, mm_projector2)
        out = inputs_embeds.clone(memory_format=torch.contiguous_format)
        out[:, 1:1 + n1 + n2] = vis.to(out.dtype)
        return out
    direct = inputs_embeds.dtype in (torch.float32, torch.bfloat16) and not (
        rt.get_precision() == "fp32_verify" and inputs_embeds.dtype != torch.float32)
    if direct:
        out = inputs_embeds.detach().clone(memory_format=torch.contiguous_format)
        mm_projector.forward_into(feats[0], out, 1)
        second.forward_into(feats[1], out, 1 + n1)
        return out
    # fp16 (the reference's eval autocast): pack in the activation dtype, then one cast-copy into the slot
    vis = pack_features(feats, mm_projector, mm_projector2)
    out = inputs_embeds.detach().clone(memory_format=torch.contiguous_format)
    out[:, 1:1 + n1 + n2] = vis.to(out.dtype)
    return out


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
