#!/usr/bin/env python
"""One ViT_stage1 forward (12 layers, direct launches) with the LayerNorm fold on or off -- meant to be run under
`ncu --metrics gpu__time_duration.sum` so that the launch list shows what each GEMM epilogue mode costs.
Usage: python tools/fold_probe.py <0|1> [batch]"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import hsenet_b200 as H  # noqa: E402

fold = sys.argv[1] == "1"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 8
dev = torch.device("cuda:0")
torch.manual_seed(0)
m = H.ViT_stage1(in_channels=1, img_size=(32, 256, 256), patch_size=(4, 16, 16), pos_embed="perceptron", spatial_dims=3,
                 classification=True).eval().requires_grad_(False).to(dev)
m.fold_layernorm = fold
m.use_cuda_graph = False
x = torch.rand(B, 1, 32, 256, 256, device=dev)
with torch.no_grad(), H.precision("bf16"):
    for _ in range(2):
        m(x)
torch.cuda.synchronize()
