#!/usr/bin/env python
"""Attention kernel timing sweep (CUDA events): batch sizes x values of an experiment switch (--env), several repetitions each so that
box noise is visible.  Usage: python tools/attn_sweep.py [--batches 4,8,16,32] [--env HSENET_ATT_POLY --modes 0,2,4] [--reps 3]"""
import argparse
import os
import statistics
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from hsenet_b200 import _lib  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batches", default="4,8,16,32")
    ap.add_argument("--modes", default="0,1")
    ap.add_argument("--env", default="HSENET_ATT_MODE")
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--iters", type=int, default=40)
    ap.add_argument("--seq", type=int, default=2049)
    ap.add_argument("--ws", action="store_true", help="give the kernel key-norm scratch (max-free softmax where the bound allows)")
    ap.add_argument("--scale", type=float, default=1.0, help="std of the random qkv activations")
    args = ap.parse_args()
    lib = _lib.load()
    dev = torch.device("cuda:0")
    st = torch.cuda.current_stream(dev).cuda_stream
    S = args.seq
    for B in [int(b) for b in args.batches.split(",")]:
        qkv = (torch.randn(B * S, 2304, device=dev) * args.scale).to(torch.bfloat16)
        scratch = torch.empty(B * 12, device=dev)

        def call():
            if args.ws:
                return lib.hsenet_self_attention_ws(qkv.data_ptr(), out.data_ptr(), None, scratch.data_ptr(), B, S, 0, st)
            return lib.hsenet_self_attention(qkv.data_ptr(), out.data_ptr(), B, S, 0, st)
        out = torch.empty(B * S, 768, dtype=torch.bfloat16, device=dev)
        fl = 4.0 * B * 12 * S * S * 64
        for mode in args.modes.split(","):
            os.environ[args.env] = mode
            res = []
            for _ in range(args.reps):
                for _ in range(5):
                    call()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(args.iters):
                    call()
                e1.record()
                torch.cuda.synchronize()
                res.append(e0.elapsed_time(e1) * 1e3 / args.iters)
            med = statistics.median(res)
            print(f"B={B:3d} {args.env}={mode}  us: " + " ".join(f"{r:7.1f}" for r in res) +
                  f"   median {med:7.1f}  {fl / med / 1e6:6.1f} TFLOP/s", flush=True)


if __name__ == "__main__":
    main()
