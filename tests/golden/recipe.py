"""Deterministic weight / input recipe shared by the fixture generator (make_golden.py, which runs the REAL reference
files) and the tests that replay the fixtures (oracle on CPU, CUDA path on the B200).  Independent of any module
constructor's RNG consumption: every tensor is drawn from its own generator seeded by (seed, key)."""
import zlib

import torch


def _gen(seed: int, key: str) -> torch.Generator:
    return torch.Generator().manual_seed((seed * 1000003 + zlib.crc32(key.encode())) % (2 ** 31))


def recipe_state_dict(shapes: dict, seed: int = 0) -> dict:
    """shapes: {state_dict key: shape}.  LayerNorm weights ~ 1 + 0.1 N, biases / cls / pos ~ 0.02 N,
    matrices ~ N(0, 1/(3 fan_in)) (the std of nn.Linear's default init)."""
    sd = {}
    for k, shp in shapes.items():
        g = _gen(seed, k)
        shp = tuple(shp)
        if "norm" in k and k.endswith("weight"):
            t = 1.0 + 0.1 * torch.randn(shp, generator=g)
        elif k.endswith("bias") or k.endswith("cls_token") or k.endswith("position_embeddings"):
            t = 0.02 * torch.randn(shp, generator=g)
        elif k.endswith("patch_score_proj.weight"):
            t = 0.5 * torch.randn(shp, generator=g)
        else:
            fan_in = shp[-1]
            t = torch.randn(shp, generator=g) * (1.0 / (3.0 * fan_in)) ** 0.5
        sd[k] = t
    return sd


def recipe_inputs(B: int, seed: int = 1234):
    g = _gen(seed, "inputs")
    images = torch.rand(B, 1, 32, 256, 256, generator=g)
    images_2d = torch.randn(B, 32, 768, generator=g)
    return images, images_2d


def recipe_tokens(B: int, seed: int = 99):
    return torch.randn(B, 2048, 768, generator=_gen(seed, "tokens"))


SAMPLE_ROWS = [0, 1, 2, 777, 2048]
PACKER_ROWS = [0, 5, 127]
