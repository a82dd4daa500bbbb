#!/usr/bin/env python
"""Training-step timing (CUDA events): forward + backward of ViT_stage1 / ViT_stage2 (12 layers) and of the packer through
the autograd Functions.  Usage: python tools/train_bench.py [--batch 8] [--iters 5]"""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import hsenet_b200 as H  # noqa: E402

GEOM = dict(in_channels=1, img_size=(32, 256, 256), patch_size=(4, 16, 16), pos_embed="perceptron", spatial_dims=3,
            classification=True)


def timeit(fn, iters):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--iters", type=int, default=5)
    a = ap.parse_args()
    dev = torch.device("cuda:0")
    H.set_precision("bf16")
    B = a.batch
    x = torch.rand(B, 1, 32, 256, 256, device=dev)
    s = torch.randn(B, 32, 768, device=dev)
    for name, cls, gf in (("ViT_stage1", H.ViT_stage1, 506.05), ("ViT_stage2", H.ViT_stage2, 511.17)):
        torch.manual_seed(0)
        m = cls(**GEOM).to(dev).eval()          # eval(): dropout of the slice-guided attention inactive; grads still flow
        args = (x,) if name == "ViT_stage1" else (x, s)

        def fwd_only():
            with torch.no_grad():
                m(*args)

        def step():
            m.zero_grad(set_to_none=True)
            y, _ = m(*args)
            y.float().square().mean().backward()

        f = timeit(fwd_only, a.iters)
        t = timeit(step, a.iters)
        print(f"{name}: inference forward {f:8.2f} ms ({B / f * 1e3:7.1f} vol/s) | train step (fwd+bwd) {t:8.2f} ms "
              f"({B / t * 1e3:7.1f} vol/s, {3 * gf * B / t:7.1f} TFLOP/s at 3x forward FLOPs, "
              f"{t / f:4.1f}x the inference forward)  peak mem {torch.cuda.max_memory_allocated() / 2**30:5.1f} GiB", flush=True)
        del m
        H.release_workspaces()
        torch.cuda.empty_cache()
        torch.cuda.reset_peak_memory_stats()
    p = H.VisualPacker_3d_phi_v3((32, 256, 256), (4, 16, 16), 768, 3072, "mlp", 2).to(dev).eval()
    feats = torch.randn(B, 2048, 768, device=dev).to(torch.bfloat16)

    def pstep():
        p.zero_grad(set_to_none=True)
        p(feats).float().square().mean().backward()

    def pfwd():
        with torch.no_grad():
            p(feats)
    f, t = timeit(pfwd, a.iters), timeit(pstep, a.iters)
    print(f"packer:     inference forward {f:8.3f} ms | train step {t:8.3f} ms ({t / f:4.1f}x)")


if __name__ == "__main__":
    main()
