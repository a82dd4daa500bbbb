"""2D-slice extraction: the online resize of the reference's slice branch
(``F.interpolate(x, size=(32,224,224), mode='trilinear')`` + 3-channel expand + permute, vit.py:529-531 / 805-807).
Feeds only non-live classes in the reference; shipped as a stand-alone operator with its own parity test."""
from __future__ import annotations

import torch

from . import _lib
from . import runtime as rt


def extract_slices(images: torch.Tensor, out_hw=(224, 224), dtype: torch.dtype = torch.float32) -> torch.Tensor:
    """images [B,1,32,256,256] -> [B*32, 3, H, W]."""
    rt.require_cuda(images, "images")
    if images.dim() != 5 or tuple(images.shape[1:]) != (1, 32, 256, 256):
        raise ValueError(f"expected [B,1,32,256,256], got {tuple(images.shape)}")
    x = images.detach().float().contiguous()
    B = x.shape[0]
    out = torch.empty(B * 32, 3, out_hw[0], out_hw[1], dtype=dtype, device=x.device)
    with torch.cuda.device(x.device):
        rc = _lib.load().hsenet_slice_extract(x.data_ptr(), out.data_ptr(), B, out_hw[0], out_hw[1],
                                              rt.dtype_code(dtype), rt.stream_ptr(x.device))
    _lib.check(rc, "slice_extract")
    return out
