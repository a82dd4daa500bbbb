#!/usr/bin/env python
"""Volume-ingest kernels at CT-RATE size: per-kernel time (CUDA events) and achieved HBM
bandwidth against the algorithmic bytes.  Usage: python tools/ingest_bench.py [--shape 512,512,303] [--xy 0.7] [--z 1.0]"""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from hsenet_b200 import _lib, preprocess as P, runtime as rt  # noqa: E402


def timeit(fn, iters=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--shape", default="512,512,303")
    ap.add_argument("--xy", type=float, default=0.7)
    ap.add_argument("--z", type=float, default=1.0)
    a = ap.parse_args()
    shape = tuple(int(v) for v in a.shape.split(","))
    dev = torch.device("cuda:0")
    lib = _lib.load()
    st = rt.stream_ptr(dev)
    raw = torch.randint(-1200, 600, shape, device=dev).float()
    o = P.resampled_shape(shape, a.xy, a.z)
    res = torch.empty(o, device=dev)
    scratch = torch.empty_like(raw)
    mm = torch.empty(2, device=dev)
    sc = torch.empty(2, dtype=torch.int32, device=dev)
    bb = torch.empty(6, dtype=torch.int32, device=dev)
    out = torch.empty(1, 32, 256, 256, device=dev)
    n_raw, n_res, n_out = raw.numel(), res.numel(), out.numel()
    cases = [
        ("hu_resample, one gather pass", 4.0 * (n_raw + n_res),
         lambda: lib.hsenet_hu_resample(raw.data_ptr(), *shape, 1.0, -24.0, -1000.0, 200.0, res.data_ptr(), *o, None,
                                        st)),
        ("hu_resample, transpose + resample (shipped)", 4.0 * (n_raw + n_res),
         lambda: lib.hsenet_hu_resample(raw.data_ptr(), *shape, 1.0, -24.0, -1000.0, 200.0, res.data_ptr(), *o,
                                        scratch.data_ptr(), st)),
        ("minmax", 4.0 * n_res, lambda: lib.hsenet_minmax(res.data_ptr(), n_res, mm.data_ptr(), sc.data_ptr(), st)),
        ("foreground_bbox", 4.0 * n_res,
         lambda: lib.hsenet_foreground_bbox(res.data_ptr(), *o, mm.data_ptr(), bb.data_ptr(), st)),
        ("crop_normalize_resize -> 32x256x256", 4.0 * (n_res + n_out),
         lambda: lib.hsenet_crop_normalize_resize(res.data_ptr(), *o, mm.data_ptr(), bb.data_ptr(), out.data_ptr(),
                                                  32, 256, 256, st)),
    ]
    print(f"raw {shape} ({4 * n_raw / 1e6:.0f} MB) -> resampled {o} ({4 * n_res / 1e6:.0f} MB) -> [1,32,256,256]")
    tot = 0.0
    for k, (name, nbytes, fn) in enumerate(cases):
        us = timeit(fn)
        if k != 0:
            tot += us
        print(f"  {name:46s} {us:8.1f} us   {nbytes / us / 1e3:7.1f} GB/s (algorithmic bytes)")
    us = timeit(lambda: P.preprocess_ct_volume(raw, 1.0, -24.0, a.xy, a.z))
    print(f"  chain through the Python entry point           {us:8.1f} us   (sum of kernels {tot:.1f} us) = "
          f"{1e6 / us:.0f} volumes/s per GPU")


if __name__ == "__main__":
    main()
