"""Volume ingest on the GPU: the arithmetic of the reference's offline preprocessing script
``Data/data_processing/CT-RATE/CT-RATE_nii_to_3D_volume_npy_file.py`` (``nii_img_to_tensor`` + ``transform``, lines 41-117)
from the raw NIfTI voxel array to the ``[1,32,256,256]`` fp32 network input, as four asynchronous kernels without a host
round trip (the reference goes GPU -> CPU -> GPU three times and resamples in float64 on the CPU).

Parity: every operator is checked against ``torch`` on the same inputs (tests/test_gpu_ingest.py); the CHAIN is checked
against oracle.preprocess_volume in fp32.  Against the reference script itself the chain is "parity unpinned": MONAI's
CropForeground / Resize sources are not in the reference tree, the script resamples in float64, and its foreground test
(x > min after interpolation) is sensitive to one-ulp differences."""
from __future__ import annotations

import torch

from . import _lib
from . import runtime as rt

TARGET_SPACING = (1.5, 0.75, 0.75)     # (z, x, y) mm, script lines 69-71
HU_WINDOW = (-1000.0, 200.0)           # script line 85
OUT_SIZE = (32, 256, 256)              # mtf.Resize(spatial_size=[32, 256, 256]), script line 117


def resampled_shape(raw_shape, xy_spacing: float, z_spacing: float):
    """new_shape of resize_array (script lines 30-35) for the (2,0,1)-transposed volume [z, a, b]."""
    n0, n1, n2 = raw_shape
    cur = (z_spacing, xy_spacing, xy_spacing)
    orig = (n2, n0, n1)
    return tuple(int(orig[i] * (cur[i] / TARGET_SPACING[i])) for i in range(3))


def preprocess_ct_volume(raw: torch.Tensor, slope: float, intercept: float, xy_spacing: float, z_spacing: float,
                         out_size=OUT_SIZE, return_intermediates: bool = False, two_pass: bool = True):
    """raw: CUDA fp32 ``[n0, n1, n2]`` in NIfTI array order (``nib.load(p).get_fdata()``).  Returns ``[1, *out_size]`` fp32."""
    rt.require_cuda(raw, "raw volume")
    if raw.dim() != 3:
        raise ValueError(f"expected a 3-D voxel array, got {tuple(raw.shape)}")
    x = raw.detach().float().contiguous()
    dev = x.device
    n0, n1, n2 = x.shape
    o = resampled_shape(x.shape, xy_spacing, z_spacing)
    if min(o) <= 0:
        raise ValueError(f"resampled shape {o} is empty")
    lib = _lib.load()
    st = rt.stream_ptr(dev)
    res = torch.empty(o, dtype=torch.float32, device=dev)
    mm = torch.empty(2, dtype=torch.float32, device=dev)
    scratch = torch.empty(2, dtype=torch.int32, device=dev)
    bbox = torch.empty(6, dtype=torch.int32, device=dev)
    out = torch.empty((1,) + tuple(out_size), dtype=torch.float32, device=dev)
    # two_pass: windowed tiled transpose into a scratch volume, then a coalesced resample (faster); else one gather pass
    tscratch = torch.empty_like(x) if two_pass else None
    with torch.cuda.device(dev):
        _lib.check(lib.hsenet_hu_resample(x.data_ptr(), n0, n1, n2, float(slope), float(intercept), HU_WINDOW[0],
                                          HU_WINDOW[1], res.data_ptr(), o[0], o[1], o[2], rt.ptr(tscratch), st),
                   "hu_resample")
        _lib.check(lib.hsenet_minmax(res.data_ptr(), res.numel(), mm.data_ptr(), scratch.data_ptr(), st), "minmax")
        _lib.check(lib.hsenet_foreground_bbox(res.data_ptr(), o[0], o[1], o[2], mm.data_ptr(), bbox.data_ptr(), st),
                   "foreground_bbox")
        _lib.check(lib.hsenet_crop_normalize_resize(res.data_ptr(), o[0], o[1], o[2], mm.data_ptr(), bbox.data_ptr(),
                                                    out.data_ptr(), out_size[0], out_size[1], out_size[2], st),
                   "crop_normalize_resize")
    if return_intermediates:
        return out, res, mm, bbox
    return out
