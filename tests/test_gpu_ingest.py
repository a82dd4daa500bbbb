"""Volume ingest kernels against torch / the oracle restatement of the reference's preprocessing
script.  Every operator is checked on identical inputs: resample within fp32 rounding, min / max and the foreground box
exactly, the fused crop + normalise + resize within fp32 rounding; then the chain against oracle.preprocess_volume."""
import pytest
import torch
import torch.nn.functional as F

from util import O, metrics

pytestmark = pytest.mark.gpu


def _raw(shape, seed, border=0):
    g = torch.Generator().manual_seed(seed)
    raw = torch.randint(-1200, 600, shape, generator=g).float()      # int16-like stored values
    if border:
        raw[:border] = -2000; raw[-border:] = -2000                   # air: clamps to the window minimum
        raw[:, :border] = -2000; raw[:, -border:] = -2000
    return raw


@pytest.mark.parametrize("shape,xy,z", [((96, 80, 40), 0.9, 2.0), ((64, 64, 33), 0.6, 1.0), ((50, 70, 20), 1.1, 3.0)])
def test_hu_resample_matches_interpolate(cuda, shape, xy, z):
    import hsenet_b200 as H
    from hsenet_b200 import preprocess as P
    raw = _raw(shape, 1)
    slope, intercept = 1.0, -24.0
    out, res, mm, bbox = P.preprocess_ct_volume(raw.to(cuda), slope, intercept, xy, z, return_intermediates=True)
    out1, res1, _, _ = P.preprocess_ct_volume(raw.to(cuda), slope, intercept, xy, z, return_intermediates=True,
                                              two_pass=False)
    # same arithmetic in both variants up to FMA contraction (values span +-1000)
    assert (res - res1).abs().max() <= 1e-3 and (out - out1).abs().max() <= 1e-5
    x = (slope * raw + intercept).clamp(-1000, 200).permute(2, 0, 1)[None, None]
    ref = F.interpolate(x, size=P.resampled_shape(shape, xy, z), mode="trilinear", align_corners=False)[0, 0]
    assert res.shape == ref.shape == O.preprocess_resampled_shape(shape, xy, z)
    assert (res.cpu() - ref).abs().max() <= 1e-3                       # values up to 1000: ~1 ulp of fp32 weights
    # min / max: exact on the SAME tensor
    assert torch.equal(mm.cpu(), torch.stack([res.min(), res.max()]).cpu())


def test_foreground_bbox_exact_and_crop_resize(cuda):
    from hsenet_b200 import _lib, runtime as rt
    lib = _lib.load()
    g = torch.Generator().manual_seed(3)
    d = (21, 37, 29)
    x = torch.zeros(d)
    x[4:17, 6:30, 3:25] = torch.rand(13, 24, 22, generator=g) + 0.5    # foreground strictly above the background
    x[2, 5, 27] = 0.25                                                  # a lone voxel stretches the box
    xd = x.to(cuda)
    st = rt.stream_ptr(xd.device)
    mm = torch.empty(2, device=cuda); sc = torch.empty(2, dtype=torch.int32, device=cuda)
    bb = torch.empty(6, dtype=torch.int32, device=cuda)
    _lib.check(lib.hsenet_minmax(xd.data_ptr(), xd.numel(), mm.data_ptr(), sc.data_ptr(), st), "minmax")
    _lib.check(lib.hsenet_foreground_bbox(xd.data_ptr(), *d, mm.data_ptr(), bb.data_ptr(), st), "bbox")
    xn = (x - x.min()) / torch.clamp(x.max() - x.min(), min=1e-8)
    lo, hi = O.foreground_bbox(xn)
    assert tuple(bb.cpu().tolist()) == lo + hi == (2, 5, 3, 17, 30, 28)
    out = torch.empty(1, 32, 64, 48, device=cuda)
    _lib.check(lib.hsenet_crop_normalize_resize(xd.data_ptr(), *d, mm.data_ptr(), bb.data_ptr(), out.data_ptr(),
                                                32, 64, 48, st), "crop_resize")
    crop = xn[lo[0]:hi[0], lo[1]:hi[1], lo[2]:hi[2]]
    ref = F.interpolate(crop[None, None], size=(32, 64, 48), mode="trilinear", align_corners=False)[0]
    assert (out.cpu() - ref).abs().max() <= 2e-6
    # constant volume: nothing is foreground -> full extent, output all zeros
    c = torch.full(d, 7.0, device=cuda)
    _lib.check(lib.hsenet_minmax(c.data_ptr(), c.numel(), mm.data_ptr(), sc.data_ptr(), st), "minmax")
    _lib.check(lib.hsenet_foreground_bbox(c.data_ptr(), *d, mm.data_ptr(), bb.data_ptr(), st), "bbox")
    assert tuple(bb.cpu().tolist()) == (0, 0, 0) + d


def test_ingest_chain_matches_oracle_and_feeds_the_encoder(cuda):
    import hsenet_b200 as H
    raw = _raw((96, 80, 40), 7)
    out = H.preprocess_ct_volume(raw.to(cuda), 1.0, -24.0, 0.9, 2.0)
    ref = O.preprocess_volume(raw, 1.0, -24.0, 0.9, 2.0)
    assert out.shape == ref.shape == (1, 32, 256, 256) and out.dtype == torch.float32
    m = metrics(out, ref)
    assert m["max_rel"] < 1e-4 and float(out.min()) >= 0.0 and float(out.max()) <= 1.0, m
    # an air border must be cropped away identically (window minimum is exactly representable, the border is wide
    # enough that whole planes stay exactly at the minimum after interpolation)
    raw = _raw((96, 80, 40), 9, border=8)
    out, res, mm, bbox = H.preprocess.preprocess_ct_volume(raw.to(cuda), 1.0, 0.0, 0.75, 1.5, return_intermediates=True)
    ref, xr, (mn, mx), box = O.preprocess_volume(raw, 1.0, 0.0, 0.75, 1.5, return_intermediates=True)
    assert tuple(bbox.cpu().tolist()) == box and box[1] >= 7 and box[4] <= 96 - 7
    assert metrics(out, ref)["max_rel"] < 1e-4
