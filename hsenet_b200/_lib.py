"""ctypes binding of libhsenet_sm100a.so (C ABI in include/hsenet_b200.h).

There is deliberately no CPU or eager-PyTorch fallback: if the shared library is missing the import of any
compute entry point raises, and every non-zero status code from the library is turned into an exception.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("HSENET_LIB_PATH") or os.path.join(_HERE, "libhsenet_sm100a.so")   # override: A/B of two builds

OK = 0
ERR_SHAPE, ERR_ALIGN, ERR_CUDA, ERR_ARG, ERR_DRIVER = -1, -2, -3, -4, -5
PREC_BF16, PREC_FP32_VERIFY = 0, 1
DTYPE_F32, DTYPE_BF16, DTYPE_F16 = 0, 1, 2
PROFILE_CLASSES = 10      # HSENET_PROFILE_CLASSES
VIT_PATCH_DONE = 1        # HSENET_VIT_PATCH_DONE

vp = C.c_void_p


class BlockWeights(C.Structure):
    _fields_ = [(n, vp) for n in ("w_qkv", "w_out", "b_out", "w_fc1", "b_fc1", "w_fc2", "b_fc2",
                                  "ln1_g", "ln1_b", "ln2_g", "ln2_b",
                                  "w_qkv_ln", "cs_qkv", "b_qkv_ln", "w_fc1_ln", "cs_fc1", "b_fc1_ln", "b_qkv")]


class VitWeights(C.Structure):
    _fields_ = [("stage", C.c_int32), ("num_layers", C.c_int32)] + [
        (n, vp) for n in ("cls_token", "pos_embed", "w_patch", "b_patch", "blocks_host", "norm_g", "norm_b",
                          "w_sq", "b_sq", "w_skv", "b_skv", "w_so", "b_so", "sn_g", "sn_b", "w_score", "b_score",
                          "w_patch_f32")]


class TrunkWeights(C.Structure):
    _fields_ = [("num_layers", C.c_int32), ("ln_eps", C.c_float)] + [
        (n, vp) for n in ("w_patch_sum", "b_patch", "pos_patch", "cls_pos0", "blocks_host", "norm_g", "norm_b")]


class PackerWeights(C.Structure):
    _fields_ = [("out_dim", C.c_int32)] + [
        (n, vp) for n in ("w_q", "b_q", "w_kv", "b_kv", "w_o", "b_o", "ln_g", "ln_b", "w_p0", "b_p0", "w_p2", "b_p2")]


class BlockWeightsT(C.Structure):
    _fields_ = [(n, vp) for n in ("w_qkv_t", "w_out_t", "w_fc1_t", "w_fc2_t")]


class VitWeightsT(C.Structure):
    _fields_ = [(n, vp) for n in ("blocks_host", "w_sq_t", "w_so_t")]


class BlockGrads(C.Structure):
    _fields_ = [(n, vp) for n in ("w_qkv", "w_out", "b_out", "w_fc1", "b_fc1", "w_fc2", "b_fc2",
                                  "ln1_g", "ln1_b", "ln2_g", "ln2_b")]


class VitGrads(C.Structure):
    _fields_ = [(n, vp) for n in ("blocks_host", "cls_token", "pos_embed", "w_patch", "b_patch", "norm_g", "norm_b",
                                  "w_sq", "b_sq", "w_skv", "b_skv", "w_so", "b_so", "sn_g", "sn_b", "w_score", "b_score",
                          "w_patch_f32")]


class PackerWeightsT(C.Structure):
    _fields_ = [(n, vp) for n in ("w_q_t", "w_kv_t", "w_o_t", "w_p0_t", "w_p2_t")]


class Dropout(C.Structure):
    _fields_ = [("p_attn", C.c_float), ("p_out", C.c_float), ("seed_attn", C.c_ulonglong), ("seed_out", C.c_ulonglong)]


class PackerGrads(C.Structure):
    _fields_ = [(n, vp) for n in ("w_q", "b_q", "w_kv", "b_kv", "w_o", "b_o", "ln_g", "ln_b", "w_p0", "b_p0", "w_p2", "b_p2")]


# name -> (restype, argtypes); mirrors include/hsenet_b200.h one to one (tests/test_abi.py checks the header).
SIGNATURES = {
    "hsenet_version": (C.c_char_p, []),
    "hsenet_error_string": (C.c_char_p, [C.c_int]),
    "hsenet_launch_count": (C.c_uint64, []),
    "hsenet_profile_start": (None, []),
    "hsenet_profile_stop": (C.c_int, [C.POINTER(C.c_double), C.POINTER(C.c_double), C.POINTER(C.c_double),
                                      C.POINTER(C.c_uint64)]),
    "hsenet_vit_workspace_bytes": (C.c_size_t, [C.c_int, C.c_int, C.c_int]),
    "hsenet_vit_forward": (C.c_int, [C.POINTER(VitWeights), vp, vp, C.c_int, C.c_int, vp, vp, vp, vp, vp,
                                     C.c_size_t, C.c_int, vp]),
    "hsenet_patch_embed_dual": (C.c_int, [C.POINTER(VitWeights), C.POINTER(VitWeights), vp, vp, C.c_int, vp, C.c_size_t,
                                          vp, C.c_size_t, vp]),
    "hsenet_packer_workspace_bytes": (C.c_size_t, [C.c_int, C.c_int, C.c_int]),
    "hsenet_packer_forward": (C.c_int, [C.POINTER(PackerWeights), vp, C.c_int, C.c_int, vp, C.c_int, C.c_int,
                                        C.c_int, vp, C.c_size_t, vp]),
    "hsenet_clip_image_head": (C.c_int, [vp, vp, vp, C.c_int, C.c_int, vp, vp, C.c_size_t, vp]),
    "hsenet_linear": (C.c_int, [vp, C.c_int, vp, C.c_int, C.c_int, C.c_int, C.c_int, vp, vp, C.c_int, C.c_int,
                                vp, C.c_int, vp, C.c_int, C.c_int, vp]),
    "hsenet_self_attention": (C.c_int, [vp, vp, C.c_int, C.c_int, C.c_int, vp]),
    "hsenet_layernorm": (C.c_int, [vp, vp, vp, C.c_long, vp, C.c_int, vp]),
    "hsenet_patch_im2col": (C.c_int, [vp, C.c_int, vp, C.c_int, vp]),
    "hsenet_packer_pool": (C.c_int, [vp, vp, C.c_int, C.c_int, vp]),
    "hsenet_packer_window_attention": (C.c_int, [vp, vp, vp, C.c_int, C.c_int, vp]),
    "hsenet_slice_cross_attention": (C.c_int, [vp, vp, vp, vp, C.c_int, C.c_int, vp]),
    "hsenet_slice_extract": (C.c_int, [vp, vp, C.c_int, C.c_int, C.c_int, C.c_int, vp]),
    "hsenet_gather_rows": (C.c_int, [vp, C.c_int, C.c_long, C.c_long, C.c_int, C.c_int, vp, C.c_int, vp]),
    "hsenet_patch_gather_map": (C.c_int, [vp, vp]),
    "hsenet_packer_window_map": (C.c_int, [vp, vp]),
    "hsenet_cast_bf16": (C.c_int, [vp, vp, C.c_long, vp]),
    "hsenet_hu_resample": (C.c_int, [vp, C.c_int, C.c_int, C.c_int, C.c_float, C.c_float, C.c_float, C.c_float, vp,
                                     C.c_int, C.c_int, C.c_int, vp, vp]),
    "hsenet_minmax": (C.c_int, [vp, C.c_long, vp, vp, vp]),
    "hsenet_foreground_bbox": (C.c_int, [vp, C.c_int, C.c_int, C.c_int, vp, vp, vp]),
    "hsenet_crop_normalize_resize": (C.c_int, [vp, C.c_int, C.c_int, C.c_int, vp, vp, vp, C.c_int, C.c_int, C.c_int, vp]),
    "hsenet_fold_layernorm": (C.c_int, [vp, vp, vp, vp, C.c_int, C.c_int, vp, vp, vp, vp]),
    "hsenet_slice_trunk_workspace_bytes": (C.c_size_t, [C.c_int, C.c_int]),
    "hsenet_slice_trunk_forward": (C.c_int, [C.POINTER(TrunkWeights), vp, C.c_int, C.c_int, vp, vp, C.c_size_t, vp]),
    # training path
    "hsenet_transpose_weight": (C.c_int, [vp, C.c_int, C.c_int, vp, C.c_int, vp]),
    "hsenet_vit_tape_bytes": (C.c_size_t, [C.c_int, C.c_int, C.c_int, C.c_int]),
    "hsenet_vit_train_workspace_bytes": (C.c_size_t, [C.c_int, C.c_int, C.c_int]),
    "hsenet_vit_forward_train": (C.c_int, [C.POINTER(VitWeights), vp, vp, C.c_int, C.c_int, vp, vp, vp, vp, C.c_size_t,
                                           vp, C.c_size_t, C.POINTER(Dropout), vp]),
    "hsenet_vit_backward": (C.c_int, [C.POINTER(VitWeights), C.POINTER(VitWeightsT), vp, C.c_int, C.c_int, vp, vp, vp,
                                      C.c_size_t, C.POINTER(VitGrads), vp, C.c_size_t, C.POINTER(Dropout), vp]),
    "hsenet_dropout_mask": (C.c_int, [C.c_float, C.c_ulonglong, C.c_longlong, vp, vp]),
    "hsenet_packer_tape_bytes": (C.c_size_t, [C.c_int, C.c_int, C.c_int]),
    "hsenet_packer_train_workspace_bytes": (C.c_size_t, [C.c_int, C.c_int, C.c_int]),
    "hsenet_packer_forward_train": (C.c_int, [C.POINTER(PackerWeights), vp, C.c_int, C.c_int, vp, vp, C.c_size_t, vp,
                                              C.c_size_t, C.POINTER(Dropout), vp]),
    "hsenet_packer_backward": (C.c_int, [C.POINTER(PackerWeights), C.POINTER(PackerWeightsT), vp, C.c_int, C.c_int, vp,
                                         vp, C.c_size_t, C.POINTER(PackerGrads), vp, vp, C.c_size_t, C.POINTER(Dropout),
                                         vp]),
    "hsenet_self_attention_ws": (C.c_int, [vp, vp, vp, vp, C.c_int, C.c_int, C.c_int, vp]),
    "hsenet_self_attention_train": (C.c_int, [vp, vp, vp, C.c_int, C.c_int, C.c_int, vp]),
    "hsenet_self_attention_backward": (C.c_int, [vp, vp, vp, vp, vp, vp, C.c_int, C.c_int, C.c_int, vp]),
}

_lib = None


class HSENetLibraryError(ImportError):
    pass


def load() -> C.CDLL:
    """Load the shared library (once).  Raises loudly when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.isfile(LIB_PATH):
        raise HSENetLibraryError(
            f"{LIB_PATH} not found: build it with `python -m hsenet_b200.build` (nvcc, sm_100a). "
            "hsenet_b200 has no CPU / eager fallback by design.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)   # AttributeError here == ABI drift, surfaced at load time
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def check(rc: int, what: str = "") -> None:
    if rc == OK:
        return
    msg = load().hsenet_error_string(rc).decode()
    text = f"hsenet_b200 {what}: {msg} (code {rc})"
    if rc in (ERR_SHAPE, ERR_ARG):
        raise ValueError(text)
    raise RuntimeError(text)
