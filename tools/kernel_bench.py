#!/usr/bin/env python
"""Per-kernel microbenchmark (CUDA events, B200): every GEMM shape of one transformer layer with its real epilogue, the
attention kernel, LayerNorm.  Usage: python tools/kernel_bench.py [--batch 8] [--iters 20]"""
import argparse
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from hsenet_b200 import _lib  # noqa: E402


def timeit(fn, iters):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters * 1e3   # us


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=8)
    ap.add_argument("--iters", type=int, default=20)
    args = ap.parse_args()
    lib = _lib.load()
    dev = torch.device("cuda:0")
    st = torch.cuda.current_stream().cuda_stream
    B = args.batch
    M = B * 2049
    bf = torch.bfloat16
    x32 = torch.randn(M, 768, device=dev)
    xn = torch.randn(M, 768, device=dev).to(bf)
    qkv = torch.randn(M, 2304, device=dev).to(bf)
    att = torch.randn(M, 768, device=dev).to(bf)
    hid = torch.randn(M, 3072, device=dev).to(bf)
    wqkv = torch.randn(2304, 768, device=dev).to(bf) * 0.03
    wout = torch.randn(768, 768, device=dev).to(bf) * 0.03
    w1 = torch.randn(3072, 768, device=dev).to(bf) * 0.03
    w2 = torch.randn(768, 3072, device=dev).to(bf) * 0.02
    b768 = torch.randn(768, device=dev)
    b3072 = torch.randn(3072, device=dev)
    g = torch.ones(768, device=dev)

    def lin(A, K, W, N, bias, resid, gelu, of, oa):
        rc = lib.hsenet_linear(A.data_ptr(), K, W.data_ptr(), K, M, N, K, None if bias is None else bias.data_ptr(),
                               None if resid is None else resid.data_ptr(), N, gelu,
                               None if of is None else of.data_ptr(), N, None if oa is None else oa.data_ptr(), N, 0, st)
        assert rc == 0, rc

    cases = {
        "qkv   768->2304 (bf16 out)": (lambda: lin(xn, 768, wqkv, 2304, None, None, 0, None, qkv), 2.0 * M * 2304 * 768),
        "out   768->768  (+bias+resid fp32)": (lambda: lin(att, 768, wout, 768, b768, x32, 0, x32, None), 2.0 * M * 768 * 768),
        "fc1   768->3072 (+bias+gelu bf16)": (lambda: lin(xn, 768, w1, 3072, b3072, None, 1, None, hid), 2.0 * M * 3072 * 768),
        "fc2  3072->768  (+bias+resid fp32)": (lambda: lin(hid, 3072, w2, 768, b768, x32, 0, x32, None), 2.0 * M * 768 * 3072),
    }
    # cuBLASLt beside every shape (torch.nn.functional.linear, bf16 in / bf16 out, bias fused by cuBLASLt where there is
    # one; the residual add / GELU / fp32 output of our epilogues are NOT included -- it is the bare-GEMM bar to beat)
    import torch.nn.functional as F
    lt = {
        "qkv   768->2304 (bf16 out)": lambda: F.linear(xn, wqkv),
        "out   768->768  (+bias+resid fp32)": lambda: F.linear(att, wout, b768.to(bf)),
        "fc1   768->3072 (+bias+gelu bf16)": lambda: F.linear(xn, w1, b3072.to(bf)),
        "fc2  3072->768  (+bias+resid fp32)": lambda: F.linear(hid, w2, b768.to(bf)),
    }
    b768b, b3072b = b768.to(bf), b3072.to(bf)
    lt = {
        "qkv   768->2304 (bf16 out)": lambda: F.linear(xn, wqkv),
        "out   768->768  (+bias+resid fp32)": lambda: F.linear(att, wout, b768b),
        "fc1   768->3072 (+bias+gelu bf16)": lambda: F.linear(xn, w1, b3072b),
        "fc2  3072->768  (+bias+resid fp32)": lambda: F.linear(hid, w2, b768b),
    }
    tot_t, tot_f = 0.0, 0.0
    for name, (fn, fl) in cases.items():
        us = timeit(lt[name], args.iters)
        tot_t += us
        tot_f += fl
        print(f"[cuBLASLt bare GEMM] {name:38s} {us:8.1f} us  {fl / us / 1e6:7.1f} TFLOP/s")
    print(f"[cuBLASLt bare GEMM] layer GEMMs total {tot_t:8.1f} us  {tot_f / tot_t / 1e6:7.1f} TFLOP/s")
    for variant in ("2cta", "1cta"):
        os.environ["HSENET_GEMM_1CTA"] = "1" if variant == "1cta" else "0"
        tot_t, tot_f = 0.0, 0.0
        for name, (fn, fl) in cases.items():
            us = timeit(fn, args.iters)
            tot_t += us
            tot_f += fl
            print(f"[{variant}] {name:38s} {us:8.1f} us  {fl / us / 1e6:7.1f} TFLOP/s")
        print(f"[{variant}] layer GEMMs total {tot_t:8.1f} us  {tot_f / tot_t / 1e6:7.1f} TFLOP/s")
    os.environ["HSENET_GEMM_1CTA"] = "0"
    us = timeit(lambda: lib.hsenet_self_attention(qkv.data_ptr(), att.data_ptr(), B, 2049, 0, st), args.iters)
    fl = 4.0 * B * 12 * 2049 * 2049 * 64
    print(f"attention S=2049                       {us:8.1f} us  {fl / us / 1e6:7.1f} TFLOP/s")
    us = timeit(lambda: lib.hsenet_layernorm(x32.data_ptr(), g.data_ptr(), g.data_ptr(), M, xn.data_ptr(), 1, st),
                args.iters)
    print(f"layernorm fp32->bf16                   {us:8.1f} us  {M * 768 * 6 / us / 1e3:7.1f} GB/s")


if __name__ == "__main__":
    main()
