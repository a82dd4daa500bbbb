"""Shared helpers for the parity tests (test infrastructure; may import oracle/)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from oracle import hsenet_oracle as O  # noqa: E402

GEOM = dict(in_channels=1, img_size=(32, 256, 256), patch_size=(4, 16, 16), pos_embed="perceptron",
            spatial_dims=3, classification=True)

# north-star tolerances
BF16_COS, BF16_MAXREL = 0.999, 2e-2
FP32_MAXREL = 1e-4


def randomize_params(module, seed=7):
    """Default inits leave cls_token = 0, LayerNorm = (1, 0), patch bias = 0: perturb them so that bugs in those
    paths cannot hide.  Deterministic given the seed."""
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for name, p in module.named_parameters():
            if name.endswith("cls_token"):
                p.copy_(0.02 * torch.randn(p.shape, generator=g))
            elif "norm" in name and name.endswith("weight"):
                p.copy_(1.0 + 0.1 * torch.randn(p.shape, generator=g))
            elif name.endswith("bias"):
                p.copy_(0.02 * torch.randn(p.shape, generator=g))
    return module


def cpu_state(module):
    return {k: v.detach().float().cpu() for k, v in module.state_dict().items()}


def synthetic_inputs(B, seed=1234):
    """uniform [0,1] volumes, N(0,1) slice features."""
    g = torch.Generator().manual_seed(seed)
    images = torch.rand(B, 1, 32, 256, 256, generator=g)
    images_2d = torch.randn(B, 32, 768, generator=g)
    return images, images_2d


def metrics(got, ref):
    return O.parity_metrics(got, ref)


def assert_bf16(got, ref, what=""):
    m = metrics(got, ref)
    assert m["cos"] >= BF16_COS and m["max_rel"] <= BF16_MAXREL, f"{what}: {m}"
    return m


def assert_fp32(got, ref, what="", tol=FP32_MAXREL):
    m = metrics(got, ref)
    assert m["max_rel"] <= tol, f"{what}: {m}"
    return m
