#!/usr/bin/env python
"""Timeline of CTA 0 of the split attention kernel (trace build): per 64-key step, clock64 stamps of the softmax warp's
p_full arrive and s_full observation and of the issuing warp's p_full observation / end of issue.  Prints the hop
latencies that make up the per-buffer cycle.  Usage (trace build): HSENET_LIB_PATH=.../libhsenet_sm100a_trace.so python tools/attn_timeline.py"""
import ctypes as C
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from hsenet_b200 import _lib  # noqa: E402

lib = _lib.load()
raw = C.CDLL(_lib.LIB_PATH)
fn = raw.hsenet_debug_att_times
fn.argtypes = [C.POINTER(C.c_longlong)]
B, S = int(sys.argv[1]) if len(sys.argv) > 1 else 8, 2049
dev = torch.device("cuda:0")
st = torch.cuda.current_stream(dev).cuda_stream
qkv = torch.randn(B * S, 2304, device=dev).to(torch.bfloat16)
out = torch.empty(B * S, 768, dtype=torch.bfloat16, device=dev)
os.environ["HSENET_ATT_KERNEL"] = "split"
for _ in range(3):
    lib.hsenet_self_attention(qkv.data_ptr(), out.data_ptr(), B, S, 0, st)
torch.cuda.synchronize()
buf = (C.c_longlong * 672)()
assert fn(buf) == 0
arr = [[buf[r * 48 + i] for i in range(48)] for r in range(14)]
t0 = arr[8][2]
print("CTA 0: issuing warp", arr[11][47], "(scheduler", arr[11][47] % 4, ")")
print("step  s_full_seen | arrive of softmax warps 4..11 (scheduler = column % 4; rel. s_full_seen) | last arrive -> issuer woken   issue   issue_end -> s_full(t+2) | issuer: prev issue_end -> operands ready -> p_full seen")
for t in range(2, 31):
    sf = arr[8][t]
    arrives = [arr[r][t] - sf for r in range(8)]
    last = max(arr[r][t] for r in range(8))
    iw, ie, nxt = arr[9][t], arr[10][t], arr[8][t + 2]
    print(f"{t:4d} {sf - t0:11d} | " + " ".join(f"{a:6d}" for a in arrives) + f" | {iw - last:10d} {ie - iw:14d} {nxt - ie:10d} | {arr[11][t] - arr[10][t - 1]:8d} {iw - arr[11][t]:8d} | top {arr[12][t] - arr[10][t - 1]:5d} v_full {arr[13][t] - arr[12][t]:5d} k_full {arr[11][t] - arr[13][t]:5d}")
