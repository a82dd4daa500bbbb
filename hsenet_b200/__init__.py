"""hsenet_b200 -- B200-native (sm_100a) implementation of the HSENet visual-encoding hot path.

Public surface mirrors the reference's module interface:
    ViT_stage1, ViT_stage2, ViT3DTower_dual_encoders, regular_attention       (multimodal_encoder/vit.py)
    VisualPacker_3d_phi_v3, resolution_attention_v3                           (multimodal_projector/...)
    build_vision_tower, build_mm_projector                                    (the two builder.py factories)
    encode_images, HSENetVisualEncoder                                        (lamed_arch.py:122-141)
    prepare_inputs_for_multimodal, splice_visual_tokens                       (lamed_arch.py:143-155)
    gather_features                                                           (utils/dist_utils.py:280-306)
    ClipImageHead, clip_image_head, contrastive_logits                        (CLIP_stage1.py:97-155)
    extract_slices                                                            (vit.py:529-531)
    SliceTrunkViTB16   (online 2D-slice branch: vit.py:802-827, CT-RATE_2D_to_npy_file.py:75-98)
All arithmetic runs in libhsenet_sm100a.so (C ABI: include/hsenet_b200.h); there is no CPU fallback.
"""
from .runtime import get_precision, precision, release_workspaces, set_precision
from .vit import ViT3DTower_dual_encoders, ViT_stage1, ViT_stage2, regular_attention
from .spatial_pooling_projector import VisualPacker_3d_phi_v3, resolution_attention_v3
from .builder import build_mm_projector, build_vision_tower
from .lamed_arch import (HSENetVisualEncoder, VisionConfig, encode_images, encode_images_with,
                         prepare_inputs_for_multimodal, splice_visual_tokens)
from .dist_utils import gather_features
from .clip import ClipImageHead, clip_image_head, contrastive_logits
from .slices import extract_slices
from .slice_encoder import SliceTrunkViTB16
from .preprocess import preprocess_ct_volume

__all__ = [
    "ViT_stage1", "ViT_stage2", "ViT3DTower_dual_encoders", "regular_attention",
    "VisualPacker_3d_phi_v3", "resolution_attention_v3", "build_vision_tower", "build_mm_projector",
    "encode_images", "encode_images_with", "prepare_inputs_for_multimodal", "splice_visual_tokens",
    "HSENetVisualEncoder", "VisionConfig", "gather_features",
    "ClipImageHead", "clip_image_head", "contrastive_logits", "extract_slices", "SliceTrunkViTB16",
    "preprocess_ct_volume",
    "set_precision", "get_precision", "precision", "release_workspaces",
]
