"""TEST INFRASTRUCTURE ONLY -- never imported by the product path (hsenet_b200/).

Loads the reference's own hot-path files *unmodified* from /root/reference (this container only; the
GPU box has no /root/reference, so nothing under ``-m gpu`` tests, ``smoke()`` or ``bench.py`` calls this).

  vit.py                        -> /root/reference/Preprint/LaMed/src/model/multimodal_encoder/vit.py
  spatial_pooling_projector.py  -> .../multimodal_projector/spatial_pooling_projector.py
  dist_utils.py                 -> /root/reference/Preprint/LaMed/src/utils/dist_utils.py

The MONAI 1.3.0 blocks they import are supplied by oracle/monai_shim.py (monai is not installed and
cannot be fetched offline).  ``lamed_arch.py`` is NOT imported (it drags in the segmentation module's
deeper MONAI imports); its 20-line ``encode_images`` (lamed_arch.py:122-141) is restated in
oracle/hsenet_oracle.py.
"""
from __future__ import annotations

import contextlib
import importlib.util
import io
import os
import sys

REFERENCE_ROOT = os.environ.get("HSENET_REFERENCE_ROOT", "/root/reference")
_SRC = os.path.join(REFERENCE_ROOT, "Preprint", "LaMed", "src")

_cache: dict = {}


def available() -> bool:
    return os.path.isfile(os.path.join(_SRC, "model", "multimodal_encoder", "vit.py"))


def _load(name: str, path: str):
    if name in _cache:
        return _cache[name]
    from . import monai_shim
    monai_shim.install()
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    _cache[name] = mod
    return mod


def vit():
    return _load("hsenet_ref_vit", os.path.join(_SRC, "model", "multimodal_encoder", "vit.py"))


def packer():
    return _load("hsenet_ref_packer",
                 os.path.join(_SRC, "model", "multimodal_projector", "spatial_pooling_projector.py"))


def dist_utils():
    return _load("hsenet_ref_dist_utils", os.path.join(_SRC, "utils", "dist_utils.py"))


class TowerConfig:
    """Attribute bag with the fields ``ViT3DTower_dual_encoders.__init__`` reads (vit.py:892-925)."""

    def __init__(self, select_feature="patch", remain="dual_vits"):
        self.vision_select_layer = -1
        self.vision_select_feature = select_feature
        self.remain_2d3d_ViT_type = remain
        self.image_channel = 1
        self.image_size = (32, 256, 256)
        self.patch_size = (4, 16, 16)
        self.vision_tower = "vit_stage2_dual_encoders"
        self.mm_projector_type = "VisualPacker_3d_phi_v3"
        self.mm_hidden_size = 768
        self.hidden_size = 3072
        self.proj_layer_type = "mlp"
        self.proj_layer_num = 2
        self.proj_pooling_type = "spatial"
        self.proj_pooling_size = 2


def quiet():
    """The reference prints from constructors (vit.py:900-905); keep test logs clean."""
    return contextlib.redirect_stdout(io.StringIO())
