"""Spatial packer facade, drop-in for the live classes of the reference's
``Preprint/LaMed/src/model/multimodal_projector/spatial_pooling_projector.py``:

    resolution_attention_v3   (:48-83)    parameter container (Wq, Wk, Wv, output_linear, norm)
    VisualPacker_3d_phi_v3    (:121-153)  [B,2048,768] -> [B,128,out_dim]

Parameter names / order match the reference (14 state-dict keys: proj_mpls.{0,2}.{weight,bias},
resolution_attention.{Wq,Wk,Wv,output_linear,norm}.{weight,bias}).  ``forward`` runs hsenet_packer_forward.
"""
from __future__ import annotations

import ctypes as C

import torch
import torch.nn as nn

from . import _lib
from . import runtime as rt
from . import training as tr

HIDDEN = 768
N_PATCH = 2048


class resolution_attention_v3(nn.Module):
    def __init__(self, in_channels=16, out_channels=8, emb_dim=768, output_dim=768, dropout=0.1, aropout=0.0):
        super().__init__()
        self.emb_dim = emb_dim
        self.Wq = nn.Linear(emb_dim, emb_dim)
        self.Wk = nn.Linear(emb_dim, emb_dim)
        self.Wv = nn.Linear(emb_dim, emb_dim)
        self.attn = None
        self.output_linear = nn.Linear(emb_dim, emb_dim)
        self.dropout = nn.Dropout(p=dropout)
        self.dropout_2 = nn.Dropout(p=dropout)
        self.norm = nn.LayerNorm(emb_dim)


class VisualPacker_3d_phi_v3(nn.Module):
    """Same constructor as the reference (:122); ``layer_type``, ``layer_num``, ``pooling_type``, ``pooling_size``
    are accepted and ignored exactly as there."""

    def __init__(self, image_size, patch_size, in_dim, out_dim, layer_type, layer_num, pooling_type='spatial',
                 pooling_size=2):
        super().__init__()
        if tuple(image_size) != (32, 256, 256) or tuple(patch_size) != (4, 16, 16) or in_dim != HIDDEN:
            raise ValueError("hsenet_b200: the packer is hard-wired to the 8x16x16x768 token grid like the reference "
                             "(spatial_pooling_projector.py:140)")
        if out_dim % 256 != 0:
            raise ValueError("hsenet_b200: out_dim must be a multiple of 256 (3072 for Phi-4-mini)")
        self.in_dim = in_dim
        self.out_dim = out_dim
        self.proj_mpls = nn.Sequential(
            nn.Linear(self.in_dim, self.out_dim),
            nn.GELU(),
            nn.Linear(self.out_dim, self.out_dim),
        )
        self.kernel_size = (1, 4, 4)
        self.num_patches_pre = [img // pch for img, pch in zip(image_size, patch_size)]
        self.num_patches_post = [self.num_patches_pre[i] // self.kernel_size[i] for i in range(3)]
        self.resolution_attention = resolution_attention_v3()
        #: dtype of the returned tokens; None = activation dtype of the precision mode
        self.output_dtype = None
        self._cache = rt.WeightCache()
        self._train_cache = rt.WeightCache()

    def _dropout_active(self) -> bool:
        """True in ``.train()`` mode with p > 0 on resolution_attention's nn.Dropout members
        (spatial_pooling_projector.py:58-59): the forward then runs the training kernels, which apply dropout."""
        a = self.resolution_attention
        return bool(self.training and (a.dropout.p > 0 or a.dropout_2.p > 0))

    def disable_dropout(self):
        """Set p = 0 on the two Dropout members of resolution_attention (train without dropout)."""
        for m in self.modules():
            if isinstance(m, nn.Dropout):
                m.p = 0.0
        return self

    def refresh_weights(self):
        self._cache.invalidate()
        self._train_cache.invalidate()

    def _build_train_payload(self, prec: str):
        pl = self._build_payload(prec)
        keep = pl["keep"]

        def kt(w):
            t = tr.transpose_weight(w, prec)
            keep.append(t)
            return t.data_ptr()

        a = self.resolution_attention
        wt = _lib.PackerWeightsT()
        wt.w_q_t = kt(a.Wq.weight)
        wt.w_kv_t = kt(torch.cat([a.Wk.weight.detach(), a.Wv.weight.detach()], 0))
        wt.w_o_t = kt(a.output_linear.weight)
        wt.w_p0_t = kt(self.proj_mpls[0].weight)
        wt.w_p2_t = kt(self.proj_mpls[2].weight)
        pl["struct_t"] = wt
        return pl

    def _forward_train(self, visual_inputs):
        rt.require_cuda(visual_inputs, "visual_inputs")
        if visual_inputs.dim() != 3 or visual_inputs.shape[1] != N_PATCH or visual_inputs.shape[2] != HIDDEN:
            raise ValueError(f"expected visual_inputs [B,2048,768], got {tuple(visual_inputs.shape)}")
        out = tr.PackerTrainFn.apply(self, visual_inputs, *self.parameters())
        if self.output_dtype is not None and out.dtype != self.output_dtype:
            out = out.to(self.output_dtype)
        return out

    @property
    def proj_out_num(self):
        num = 1
        for n in self.num_patches_post:
            num *= n
        return num

    def _build_payload(self, prec: str):
        cw = lambda w: rt.cast_weight(w, prec)
        keep = []

        def k(t):
            keep.append(t)
            return t.data_ptr()

        a = self.resolution_attention
        w = _lib.PackerWeights()
        w.out_dim = self.out_dim
        w.w_q = k(cw(a.Wq.weight)); w.b_q = k(rt.f32(a.Wq.bias))
        w.w_kv = k(cw(torch.cat([a.Wk.weight.detach(), a.Wv.weight.detach()], 0)))
        w.b_kv = k(torch.cat([rt.f32(a.Wk.bias), rt.f32(a.Wv.bias)], 0))
        w.w_o = k(cw(a.output_linear.weight)); w.b_o = k(rt.f32(a.output_linear.bias))
        w.ln_g = k(rt.f32(a.norm.weight)); w.ln_b = k(rt.f32(a.norm.bias))
        w.w_p0 = k(cw(self.proj_mpls[0].weight)); w.b_p0 = k(rt.f32(self.proj_mpls[0].bias))
        w.w_p2 = k(cw(self.proj_mpls[2].weight)); w.b_p2 = k(rt.f32(self.proj_mpls[2].bias))
        return {"struct": w, "keep": keep}

    def forward_into(self, visual_inputs: torch.Tensor, out: torch.Tensor, token_offset: int = 0) -> torch.Tensor:
        """Pack ``visual_inputs [B,2048,768]`` into ``out[:, token_offset:token_offset+128, :]`` (``out`` is
        ``[B, T, out_dim]`` contiguous, fp32 or bf16).  Used by encode_images to skip the reference's torch.cat."""
        rt.require_cuda(visual_inputs, "visual_inputs")
        if tr.needs_grad(self, visual_inputs):
            raise RuntimeError("forward_into writes through raw pointers and is invisible to autograd; call the module "
                               "(forward) when gradients are required")
        if self._dropout_active():
            raise RuntimeError("forward_into is the inference path and applies no dropout; this packer is in train() "
                               "mode with p > 0 -- call the module (forward), .eval() or .disable_dropout()")
        if visual_inputs.dim() != 3 or visual_inputs.shape[1] != N_PATCH or visual_inputs.shape[2] != HIDDEN:
            raise ValueError(f"expected visual_inputs [B,2048,768], got {tuple(visual_inputs.shape)}")
        if visual_inputs.stride(2) != 1:
            visual_inputs = visual_inputs.contiguous()
        B = visual_inputs.shape[0]
        dev = visual_inputs.device
        prec = rt.get_precision()
        act = rt.act_dtype(prec)
        pc = rt.precision_code(prec)
        if not out.is_contiguous() or out.dim() != 3 or out.shape[0] != B or out.shape[2] != self.out_dim:
            raise ValueError("out must be a contiguous [B, T, out_dim] tensor")
        if prec == "fp32_verify" and out.dtype != torch.float32:
            raise ValueError("fp32_verify mode writes fp32 outputs")
        lib = _lib.load()
        st = rt.stream_ptr(dev)
        with torch.cuda.device(dev):
            payload = self._cache.get(self.parameters(), prec, self._build_payload)
            # tower features arrive as a [:, 1:] view (batch stride 2049*768) or contiguous; any float dtype
            if visual_inputs.dtype == act and visual_inputs.is_contiguous():
                hr = visual_inputs.detach()
            else:
                hr = torch.empty(B, N_PATCH, HIDDEN, dtype=act, device=dev)
                v = visual_inputs.detach()
                _lib.check(lib.hsenet_gather_rows(v.data_ptr(), rt.dtype_code(v.dtype), v.stride(0), v.stride(1), B,
                                                  N_PATCH, hr.data_ptr(), rt.dtype_code(act), st), "gather_rows")
            ws = rt.workspace(dev, lib.hsenet_packer_workspace_bytes(B, pc, self.out_dim), "packer")
            rc = lib.hsenet_packer_forward(C.byref(payload["struct"]), hr.data_ptr(), B, pc, out.data_ptr(),
                                           rt.dtype_code(out.dtype), out.shape[1], token_offset, ws.data_ptr(),
                                           ws.numel(), st)
        _lib.check(rc, "packer_forward")
        return out

    def forward(self, visual_inputs):
        if tr.needs_grad(self, visual_inputs) or self._dropout_active():
            return self._forward_train(visual_inputs)
        B = visual_inputs.shape[0]
        dt = self.output_dtype or rt.act_dtype()
        out = torch.empty(B, self.proj_out_num, self.out_dim, dtype=dt, device=visual_inputs.device)
        return self.forward_into(visual_inputs, out, 0)
