#!/usr/bin/env python
"""Per-phase cycle breakdown of one softmax warp of the attention kernel (trace build only):

    HSENET_NVCC_EXTRA=-DHSENET_ATT_TRACE python -m hsenet_b200.build --force
    python tools/attn_trace.py --batch 1 --seq 1152      # one CTA per SM
    python tools/attn_trace.py --batch 8 --seq 2049      # production shape, two CTAs per SM
    python -m hsenet_b200.build --force                  # back to the normal library

Prints average cycles per 64-key step spent in each phase of warp 2 of CTA 0."""
import argparse
import ctypes as C
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from hsenet_b200 import _lib  # noqa: E402

PHASES = ["wait s_full (+ fence::after in tri/rowwarp)", "tcgen05.ld + wait::ld", "fence::after (split kernel only)",
          "row max + lazy-rescale test", "exp2 / sum / pack", "tcgen05.st P (issue)", "O-correction vote (+rare path)",
          "wait::st + fence + arrive p_full"]
ISSUER = ["operand waits (v_full, k_full)", "wait p_full(t) + fence", "4 x P V mma + 2 commits", "4 x Q K^T mma + 2 commits"]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--seq", type=int, default=2049)
    args = ap.parse_args()
    lib = _lib.load()
    raw = C.CDLL(_lib.LIB_PATH) if hasattr(_lib, "LIB_PATH") else lib
    fn = getattr(raw, "hsenet_debug_att_trace", None)
    if fn is None:
        raise SystemExit("library was not built with -DHSENET_ATT_TRACE")
    fn.argtypes = [C.POINTER(C.c_ulonglong)]
    dev = torch.device("cuda:0")
    st = torch.cuda.current_stream(dev).cuda_stream
    B, S = args.batch, args.seq
    qkv = torch.randn(B * S, 2304, device=dev).to(torch.bfloat16)
    out = torch.empty(B * S, 768, dtype=torch.bfloat16, device=dev)
    for _ in range(3):
        lib.hsenet_self_attention(qkv.data_ptr(), out.data_ptr(), B, S, 0, st)
    torch.cuda.synchronize()
    buf = (C.c_ulonglong * 48)()
    assert fn(buf) == 0
    for title, base, names in (("softmax warp 2", 0, PHASES), ("MMA issuer, even steps", 16, ISSUER), ("MMA issuer, odd steps", 32, ISSUER)):
        n = buf[base + 15]
        if n == 0:
            continue
        print(f"B={B} S={S} {title}: {n} steps, loop {buf[base + 14]} cycles = {buf[base + 14] / n:.0f} per step")
        for i, name in enumerate(names):
            print(f"  {name:40s} {buf[base + i] / n:8.1f} cycles/step")


if __name__ == "__main__":
    main()
