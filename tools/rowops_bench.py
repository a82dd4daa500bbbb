#!/usr/bin/env python
"""Memory-bound kernels of the path at production size (CUDA events): achieved GB/s of ALGORITHMIC bytes against the measured
HBM copy rate (MEASURED_PEAKS.json hbm_gbs).  Also the driver for the ncu captures under profiles/r2 (`--once`: each kernel
is launched exactly once after a warm-up launch, so `ncu --launch-skip`/-k can pick it).

    python tools/rowops_bench.py [--batch 32] [--iters 20] [--once]
"""
import argparse
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from hsenet_b200 import _lib  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--once", action="store_true")
    ap.add_argument("--only", default="")
    a = ap.parse_args()
    lib = _lib.load()
    dev = torch.device("cuda:0")
    st = torch.cuda.current_stream().cuda_stream
    B = a.batch
    bf = torch.bfloat16
    peak = 6555.8
    pp = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(pp):
        peak = json.load(open(pp)).get("hbm_gbs", peak)
    Mp, Mw, M = B * 2048, B * 128, B * 2049
    vol = torch.rand(B, 1, 32, 256, 256, device=dev)
    hr = torch.randn(Mp, 768, device=dev).to(bf)
    lr = torch.empty(Mw, 768, dtype=bf, device=dev)
    q = torch.randn(Mw, 768, device=dev)
    kv = torch.randn(Mp, 1536, device=dev).to(bf)
    o = torch.empty(Mw, 768, dtype=bf, device=dev)
    patches = torch.empty(Mp, 1024, dtype=bf, device=dev)
    slices = torch.empty(B * 32, 3, 224, 224, dtype=bf, device=dev)
    Q2 = torch.randn(Mp, 768, device=dev)
    SKV = torch.randn(B * 32, 1536, device=dev)
    O2 = torch.empty(Mp, 768, dtype=bf, device=dev)
    x32 = torch.randn(M, 768, device=dev)
    xn = torch.empty(M, 768, dtype=bf, device=dev)
    g = torch.ones(768, device=dev)
    cases = [
        ("layernorm_kernel", M * 768 * 6.0,
         lambda: lib.hsenet_layernorm(x32.data_ptr(), g.data_ptr(), g.data_ptr(), M, xn.data_ptr(), 1, st)),
        ("packer_pool_kernel", B * (2048 + 128.0) * 768 * 2,
         lambda: lib.hsenet_packer_pool(hr.data_ptr(), lr.data_ptr(), B, 1, st)),
        ("packer_window_attn_kernel", B * (128.0 * 768 * 6 + 2048 * 1536 * 2.0),
         lambda: lib.hsenet_packer_window_attention(q.data_ptr(), kv.data_ptr(), o.data_ptr(), B, 1, st)),
        ("im2col_kernel", B * 2048.0 * 1024 * 6,
         lambda: lib.hsenet_patch_im2col(vol.data_ptr(), B, patches.data_ptr(), 1, st)),
        ("slice_extract_kernel", B * 32.0 * (256 * 256 * 4 + 3 * 224 * 224 * 2),
         lambda: lib.hsenet_slice_extract(vol.data_ptr(), slices.data_ptr(), B, 224, 224, 1, st)),
        ("slice_xattn_kernel", B * (2048.0 * 768 * 6 + 32 * 1536 * 4.0),
         lambda: lib.hsenet_slice_cross_attention(Q2.data_ptr(), SKV.data_ptr(), O2.data_ptr(), None, B, 1, st)),
    ]
    print(f"batch {B}; HBM copy peak {peak:.1f} GB/s")
    for name, nbytes, fn in cases:
        if a.only and a.only not in name:
            continue
        assert fn() == 0, name
        torch.cuda.synchronize()
        if a.once:
            assert fn() == 0
            torch.cuda.synchronize()
            continue
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(a.iters):
            fn()
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) / a.iters * 1e3
        print(f"  {name:28s} {us:8.1f} us  {nbytes / 1e6:8.1f} MB algorithmic  {nbytes / us / 1e3:7.1f} GB/s  "
              f"= {nbytes / us / 1e3 / peak:5.2f} of HBM copy rate", flush=True)


if __name__ == "__main__":
    main()
