"""String factories, drop-in for the reference's
``model/multimodal_encoder/builder.py:4-11`` (build_vision_tower) and
``model/multimodal_projector/builder.py:81-105`` (build_mm_projector).

Only the live branches exist: 'vit_stage2_dual_encoders' and 'VisualPacker_3d_phi_v3'.  The reference's 'vit3d'
tower (ViT3DTower) cannot be called by encode_images (it takes one argument, lamed_arch.py:123 passes two) and its
'baseline' projector branch raises NameError; both are rejected here with ValueError like any unknown name.
"""
from __future__ import annotations

from .spatial_pooling_projector import VisualPacker_3d_phi_v3
from .vit import ViT3DTower_dual_encoders


def build_vision_tower(config, **kwargs):
    vision_tower = getattr(config, 'vision_tower', None)
    if vision_tower is not None and 'vit_stage2_dual_encoders' in vision_tower.lower():
        return ViT3DTower_dual_encoders(config, **kwargs)
    raise ValueError(f'Unknown vision tower: {vision_tower}')


def build_mm_projector(config, delay_load=False, **kwargs):
    projector_type = getattr(config, 'mm_projector_type')
    if projector_type == 'VisualPacker_3d_phi_v3':
        return VisualPacker_3d_phi_v3(image_size=config.image_size,
                                      patch_size=config.patch_size,
                                      in_dim=config.mm_hidden_size,
                                      out_dim=config.hidden_size,
                                      layer_type=config.proj_layer_type,
                                      layer_num=config.proj_layer_num,
                                      pooling_type=config.proj_pooling_type,
                                      pooling_size=config.proj_pooling_size)
    raise ValueError(f'Unknown projector type: {projector_type}')
