"""TEST INFRASTRUCTURE ONLY -- never imported by the product path (hsenet_b200/).

MONAI 1.3.0 block shim.

The reference builds its encoders out of two third-party classes that are NOT vendored under
/root/reference and are not installed in this image (requirements.txt:62 pins ``monai==1.3.0``):

    from monai.networks.blocks.patchembedding import PatchEmbeddingBlock     (vit.py:21)
    from monai.networks.blocks.transformerblock import TransformerBlock      (vit.py:22)

This file restates the published behaviour of those two classes (and the SABlock / MLPBlock they are
made of) for exactly the configuration the reference uses (vit.py:290-305, 428-443):
``pos_embed="perceptron"``, ``dropout_rate=0.0``, ``qkv_bias=False``, ``save_attn=False``,
``spatial_dims=3``.  ``install()`` registers them under the MONAI module names so that the
reference's own ``vit.py`` imports and runs unmodified.

Parity status: **unpinned at the MONAI boundary** -- the reference holds no tests or golden vectors.
What pins the shim indirectly (and is asserted in tests/test_oracle.py):
  * state-dict key counts 138 / 150 / 288 quoted in Preprint/Bench/eval/eval_HSENet_BIMCV_R_MRG.py:339-354;
  * the inline comment at vit.py:437
      "Rearrange('b c (h p1) (w p2) (d p3) -> b (h w d) (p1 p2 p3 c)', p1=4, p2=16, p3=16)
       + Linear(in_features=1024, out_features=768, bias=True)";
  * the hard-coded 2048 / 8x16x16 / 768 geometry in lamed_arch.py:125 and spatial_pooling_projector.py:140.

MONAI 1.3.0 facts restated here:
  PatchEmbeddingBlock (perceptron): ``patch_embeddings = Sequential(Rearrange(...), Linear(patch_dim, hidden))``,
      ``position_embeddings = Parameter(zeros(1, n_patches, hidden))`` trunc-normal(std=.02) initialised,
      ``forward = dropout(patch_embeddings(x) + position_embeddings)``; Linear weights trunc-normal(std=.02),
      bias 0, LayerNorm weight 1 / bias 0 via ``self.apply(self._init_weights)``.
  TransformerBlock: children registered in the order ``mlp, norm1, attn, norm2``;
      ``x = x + attn(norm1(x)); x = x + mlp(norm2(x))``.
  SABlock: children ``out_proj = Linear(h, h)``, ``qkv = Linear(h, 3h, bias=qkv_bias)``;
      ``Rearrange("b h (qkv l d) -> qkv b l h d", qkv=3, l=num_heads)``; ``softmax(q k^T * head_dim**-0.5)``;
      ``einsum("bhxy,bhyd->bhxd")`` then ``"b h l d -> b l (h d)"``; out_proj; dropouts (p = 0 here).
  MLPBlock: ``linear1``, ``linear2``, ``fn = nn.GELU()`` (exact erf), two dropouts.
"""
from __future__ import annotations

import sys
import types

import torch
import torch.nn as nn
from einops.layers.torch import Rearrange


class MLPBlock(nn.Module):
    def __init__(self, hidden_size: int, mlp_dim: int, dropout_rate: float = 0.0, act="GELU", dropout_mode="vit"):
        super().__init__()
        if not (0 <= dropout_rate <= 1):
            raise ValueError("dropout_rate should be between 0 and 1.")
        mlp_dim = mlp_dim or hidden_size
        self.linear1 = nn.Linear(hidden_size, mlp_dim)
        self.linear2 = nn.Linear(mlp_dim, hidden_size)
        self.fn = nn.GELU()
        self.drop1 = nn.Dropout(dropout_rate)
        self.drop2 = nn.Dropout(dropout_rate)

    def forward(self, x):
        x = self.fn(self.linear1(x))
        x = self.drop1(x)
        x = self.linear2(x)
        x = self.drop2(x)
        return x


class SABlock(nn.Module):
    def __init__(self, hidden_size: int, num_heads: int, dropout_rate: float = 0.0,
                 qkv_bias: bool = False, save_attn: bool = False):
        super().__init__()
        if not (0 <= dropout_rate <= 1):
            raise ValueError("dropout_rate should be between 0 and 1.")
        if hidden_size % num_heads != 0:
            raise ValueError("hidden size should be divisible by num_heads.")
        self.num_heads = num_heads
        self.out_proj = nn.Linear(hidden_size, hidden_size)
        self.qkv = nn.Linear(hidden_size, hidden_size * 3, bias=qkv_bias)
        self.input_rearrange = Rearrange("b h (qkv l d) -> qkv b l h d", qkv=3, l=num_heads)
        self.out_rearrange = Rearrange("b h l d -> b l (h d)")
        self.drop_output = nn.Dropout(dropout_rate)
        self.drop_weights = nn.Dropout(dropout_rate)
        self.head_dim = hidden_size // num_heads
        self.scale = self.head_dim ** -0.5
        self.save_attn = save_attn
        self.att_mat = torch.Tensor()

    def forward(self, x):
        output = self.input_rearrange(self.qkv(x))
        q, k, v = output[0], output[1], output[2]
        att_mat = (torch.einsum("blxd,blyd->blxy", q, k) * self.scale).softmax(dim=-1)
        if self.save_attn:
            self.att_mat = att_mat.detach()
        att_mat = self.drop_weights(att_mat)
        x = torch.einsum("bhxy,bhyd->bhxd", att_mat, v)
        x = self.out_rearrange(x)
        x = self.out_proj(x)
        x = self.drop_output(x)
        return x


class TransformerBlock(nn.Module):
    def __init__(self, hidden_size: int, mlp_dim: int, num_heads: int, dropout_rate: float = 0.0,
                 qkv_bias: bool = False, save_attn: bool = False):
        super().__init__()
        if not (0 <= dropout_rate <= 1):
            raise ValueError("dropout_rate should be between 0 and 1.")
        if hidden_size % num_heads != 0:
            raise ValueError("hidden_size should be divisible by num_heads.")
        self.mlp = MLPBlock(hidden_size, mlp_dim, dropout_rate)
        self.norm1 = nn.LayerNorm(hidden_size)
        self.attn = SABlock(hidden_size, num_heads, dropout_rate, qkv_bias, save_attn)
        self.norm2 = nn.LayerNorm(hidden_size)

    def forward(self, x):
        x = x + self.attn(self.norm1(x))
        x = x + self.mlp(self.norm2(x))
        return x


class PatchEmbeddingBlock(nn.Module):
    def __init__(self, in_channels: int, img_size, patch_size, hidden_size: int, num_heads: int,
                 pos_embed: str, dropout_rate: float = 0.0, spatial_dims: int = 3):
        super().__init__()
        if not (0 <= dropout_rate <= 1):
            raise ValueError("dropout_rate should be between 0 and 1.")
        if hidden_size % num_heads != 0:
            raise ValueError("hidden size should be divisible by num_heads.")
        if pos_embed != "perceptron":
            raise NotImplementedError("shim covers the reference's pos_embed='perceptron' only")
        img_size = tuple(img_size) if not isinstance(img_size, int) else (img_size,) * spatial_dims
        patch_size = tuple(patch_size) if not isinstance(patch_size, int) else (patch_size,) * spatial_dims
        for m, p in zip(img_size, patch_size):
            if m < p:
                raise ValueError("patch_size should be smaller than img_size.")
            if m % p != 0:
                raise ValueError("patch_size should be divisible by img_size for perceptron.")
        self.n_patches = 1
        for m, p in zip(img_size, patch_size):
            self.n_patches *= m // p
        self.patch_dim = int(in_channels)
        for p in patch_size:
            self.patch_dim *= p
        chars = (("h", "p1"), ("w", "p2"), ("d", "p3"))[:spatial_dims]
        from_chars = "b c " + " ".join(f"({k} {v})" for k, v in chars)
        to_chars = f"b ({' '.join([c[0] for c in chars])}) ({' '.join([c[1] for c in chars])} c)"
        axes_len = {f"p{i + 1}": p for i, p in enumerate(patch_size)}
        self.patch_embeddings = nn.Sequential(
            Rearrange(f"{from_chars} -> {to_chars}", **axes_len), nn.Linear(self.patch_dim, hidden_size)
        )
        self.position_embeddings = nn.Parameter(torch.zeros(1, self.n_patches, hidden_size))
        self.dropout = nn.Dropout(dropout_rate)
        nn.init.trunc_normal_(self.position_embeddings, mean=0.0, std=0.02, a=-2.0, b=2.0)
        self.apply(self._init_weights)

    def _init_weights(self, m):
        if isinstance(m, nn.Linear):
            nn.init.trunc_normal_(m.weight, mean=0.0, std=0.02, a=-2.0, b=2.0)
            if m.bias is not None:
                nn.init.constant_(m.bias, 0)
        elif isinstance(m, nn.LayerNorm):
            nn.init.constant_(m.bias, 0)
            nn.init.constant_(m.weight, 1.0)

    def forward(self, x):
        x = self.patch_embeddings(x)
        embeddings = x + self.position_embeddings
        embeddings = self.dropout(embeddings)
        return embeddings


def install() -> None:
    """Register the shim (and an empty ``open_clip`` stub, imported at vit.py:23 but unused on the live
    path) in ``sys.modules`` under the names the reference imports."""
    if "monai.networks.blocks.patchembedding" in sys.modules and getattr(
            sys.modules["monai"], "__hsenet_shim__", False):
        return
    def mod(name):
        m = types.ModuleType(name)
        sys.modules[name] = m
        return m
    monai = mod("monai")
    monai.__hsenet_shim__ = True
    monai.__version__ = "1.3.0+shim"
    networks = mod("monai.networks")
    blocks = mod("monai.networks.blocks")
    pe = mod("monai.networks.blocks.patchembedding")
    tb = mod("monai.networks.blocks.transformerblock")
    monai.networks = networks
    networks.blocks = blocks
    blocks.patchembedding = pe
    blocks.transformerblock = tb
    pe.PatchEmbeddingBlock = PatchEmbeddingBlock
    tb.TransformerBlock = TransformerBlock
    blocks.PatchEmbeddingBlock = PatchEmbeddingBlock
    blocks.TransformerBlock = TransformerBlock
    blocks.SABlock = SABlock
    blocks.MLPBlock = MLPBlock
    if "open_clip" not in sys.modules:
        oc = types.ModuleType("open_clip")
        oc.__hsenet_stub__ = True
        sys.modules["open_clip"] = oc
