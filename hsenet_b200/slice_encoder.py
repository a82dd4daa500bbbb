"""Online 2D-slice branch: ``SliceTrunkViTB16`` turns a CT volume into the [B,32,768] slice
features ``ViT_stage2`` consumes as ``image_2d`` -- the online counterpart of the reference's offline pipeline
(CT-RATE_nii_to_2D_slices.py -> JPEG -> BiomedCLIP trunk -> npy, Data/data_processing/CT-RATE/CT-RATE_2D_to_npy_file.py:75-98)
in the formulation of ``ViT4LLM_v3_med2e3.forward`` (vit.py:805-808, 816): trilinear resize to (32,224,224), three identical
channels, ``model.visual.trunk`` of BiomedCLIP.

The module has the parameter names, shapes and registration order of timm's ``vit_base_patch16_224`` (152 keys:
``cls_token``, ``pos_embed``, ``patch_embed.proj.*``, ``blocks.N.{norm1,attn.qkv,attn.proj,norm2,mlp.fc1,mlp.fc2}.*``,
``norm.*``), so ``load_state_dict(biomedclip.visual.trunk.state_dict())`` works.  The forward runs in
libhsenet_sm100a.so (hsenet_slice_trunk_forward) on the same tcgen05 GEMM / attention kernels as the 3D towers.
Parity: timm / open_clip are not installable offline, so the trunk is checked against a restatement of timm's forward
(oracle/hsenet_oracle.py::timm_vit_trunk) -- "parity unpinned" at that third-party boundary.  Unlike the offline pipeline
there is no uint8 / JPEG quantisation and no CLIP mean/std normalisation (the online formulation has neither).
Inference only (the reference freezes the trunk: vit.py:800 ``requires_grad_(False)``).
"""
from __future__ import annotations

import ctypes as C

import torch
import torch.nn as nn

from . import _lib
from . import runtime as rt


class _Attention(nn.Module):
    def __init__(self, dim):
        super().__init__()
        self.qkv = nn.Linear(dim, dim * 3, bias=True)
        self.proj = nn.Linear(dim, dim)


class _Mlp(nn.Module):
    def __init__(self, dim, hidden):
        super().__init__()
        self.fc1 = nn.Linear(dim, hidden)
        self.fc2 = nn.Linear(hidden, dim)


class _Block(nn.Module):
    def __init__(self, dim, hidden):
        super().__init__()
        self.norm1 = nn.LayerNorm(dim, eps=1e-6)
        self.attn = _Attention(dim)
        self.norm2 = nn.LayerNorm(dim, eps=1e-6)
        self.mlp = _Mlp(dim, hidden)


class _PatchEmbed(nn.Module):
    def __init__(self, dim):
        super().__init__()
        self.proj = nn.Conv2d(3, dim, kernel_size=16, stride=16)


class SliceTrunkViTB16(nn.Module):
    def __init__(self, num_layers: int = 12):
        super().__init__()
        dim = 768
        self.cls_token = nn.Parameter(torch.zeros(1, 1, dim))
        self.pos_embed = nn.Parameter(torch.randn(1, 197, dim) * 0.02)
        self.patch_embed = _PatchEmbed(dim)
        self.blocks = nn.Sequential(*[_Block(dim, 3072) for _ in range(num_layers)])
        self.norm = nn.LayerNorm(dim, eps=1e-6)
        self.requires_grad_(False)
        self._cache = rt.WeightCache()

    def refresh_weights(self):
        self._cache.invalidate()

    def _build_payload(self, prec: str):
        cw = lambda w: rt.cast_weight(w, prec)
        keep = []

        def k(t):
            keep.append(t)
            return t.data_ptr()

        n = len(self.blocks)
        blocks = (_lib.BlockWeights * max(n, 1))()
        for i, blk in enumerate(self.blocks):
            b = blocks[i]
            b.w_qkv = k(cw(blk.attn.qkv.weight)); b.b_qkv = k(rt.f32(blk.attn.qkv.bias))
            b.w_out = k(cw(blk.attn.proj.weight)); b.b_out = k(rt.f32(blk.attn.proj.bias))
            b.w_fc1 = k(cw(blk.mlp.fc1.weight)); b.b_fc1 = k(rt.f32(blk.mlp.fc1.bias))
            b.w_fc2 = k(cw(blk.mlp.fc2.weight)); b.b_fc2 = k(rt.f32(blk.mlp.fc2.bias))
            b.ln1_g = k(rt.f32(blk.norm1.weight)); b.ln1_b = k(rt.f32(blk.norm1.bias))
            b.ln2_g = k(rt.f32(blk.norm2.weight)); b.ln2_b = k(rt.f32(blk.norm2.bias))
        w = _lib.TrunkWeights()
        w.num_layers = n
        w.ln_eps = float(self.norm.eps)
        conv = self.patch_embed.proj
        # three identical input channels: fold the channel sum into the stem weight, [768,3,16,16] -> [768,256]
        w.w_patch_sum = k(cw(conv.weight.detach().float().sum(dim=1).reshape(768, 256)))
        w.b_patch = k(rt.f32(conv.bias))
        pos = rt.f32(self.pos_embed).reshape(197, 768)
        w.pos_patch = k(pos[1:].contiguous())
        w.cls_pos0 = k((rt.f32(self.cls_token).reshape(768) + pos[0]).contiguous())
        w.blocks_host = C.cast(blocks, C.c_void_p)
        w.norm_g = k(rt.f32(self.norm.weight)); w.norm_b = k(rt.f32(self.norm.bias))
        return {"struct": w, "blocks": blocks, "keep": keep}

    def forward(self, images: torch.Tensor) -> torch.Tensor:
        """images [B,1,32,256,256] -> slice features fp32 [B,32,768]."""
        rt.require_cuda(images, "images")
        rt.require_cuda(self.norm.weight, "SliceTrunkViTB16 parameters")
        rt.forbid_autograd(self.parameters(), "SliceTrunkViTB16")
        if images.dim() != 5 or tuple(images.shape[1:]) != (1, 32, 256, 256):
            raise ValueError(f"expected images of shape [B,1,32,256,256], got {tuple(images.shape)}")
        dev = images.device
        B = images.shape[0]
        prec = rt.get_precision()
        pc = rt.precision_code(prec)
        lib = _lib.load()
        x = images.detach().float().contiguous()
        out = torch.empty(B, 32, 768, dtype=torch.float32, device=dev)
        with torch.cuda.device(dev):
            payload = self._cache.get(self.parameters(), prec, self._build_payload)
            ws = rt.workspace(dev, lib.hsenet_slice_trunk_workspace_bytes(B, pc), "slice_trunk")
            rc = lib.hsenet_slice_trunk_forward(C.byref(payload["struct"]), x.data_ptr(), B, pc, out.data_ptr(),
                                                ws.data_ptr(), ws.numel(), rt.stream_ptr(dev))
        _lib.check(rc, "slice_trunk_forward")
        return out
