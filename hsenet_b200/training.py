"""torch.autograd.Function wrappers over the training entry points of libhsenet_sm100a.so.

The reference trains ``ViT_stage1`` / ``ViT_stage2`` in its CLIP stages (train/train_CLIP_stage1.py:231-257,
train_CLIP_stage2.py:239-267) and the two ``VisualPacker_3d_phi_v3`` packers in the VLM stage (train/train_VLM.py:406-414)
through ordinary autograd.  Here a module's forward under grad mode runs ``hsenet_*_forward_train`` (which keeps the
activations on a tape allocated by this file) and its backward runs ``hsenet_*_backward``, which writes one fp32 gradient
per parameter.  The Functions take the module's parameters as explicit inputs, in registration order, so that autograd
routes the returned gradients to them (DDP hooks, gradient accumulation and optimizers work unchanged).

Train-mode dropout of the two small attentions (``regular_attention`` / ``resolution_attention_v3``, p = 0.1 on the
attention probabilities and on the output projection in front of the residual add): applied inside the kernels from a
counter-based keep mask (``hsenet_dropout`` in include/hsenet_b200.h).  Two 63-bit seeds per call are drawn from torch's
CPU generator (``torch.manual_seed`` makes a run repeatable); the backward regenerates the masks from the same seeds.
The random stream differs from torch's own dropout kernels, so a run matches the reference in distribution, not bit for
bit; ``dropout_mask()`` returns the exact mask a seed stands for (the parity tests feed it to the oracle).
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib
from . import runtime as rt


def needs_grad(module, *tensors) -> bool:
    if not torch.is_grad_enabled():
        return False
    if any(p.requires_grad for p in module.parameters()):
        return True
    return any(isinstance(t, torch.Tensor) and t.requires_grad for t in tensors)


def transpose_weight(w: torch.Tensor, prec: str) -> torch.Tensor:
    """[out,in] fp32 parameter -> [in,out] contiguous in the activation dtype (operand of the input-gradient GEMMs)."""
    src = w.detach().float().contiguous()
    rows, cols = src.shape
    act = rt.act_dtype(prec)
    out = torch.empty(cols, rows, dtype=act, device=src.device)
    with torch.cuda.device(src.device):
        rc = _lib.load().hsenet_transpose_weight(src.data_ptr(), rows, cols, out.data_ptr(), rt.dtype_code(act),
                                                 rt.stream_ptr(src.device))
    _lib.check(rc, "transpose_weight")
    return out


def draw_dropout(attn_module, training: bool):
    """``_lib.Dropout`` for one forward/backward pair of a small-attention module (members ``dropout`` on the
    probabilities and ``dropout_2`` on the output projection), or None when both are inactive (eval / p == 0)."""
    pa, po = float(attn_module.dropout.p), float(attn_module.dropout_2.p)
    if not training or (pa <= 0.0 and po <= 0.0):
        return None
    if pa >= 1.0 or po >= 1.0:
        raise ValueError("hsenet_b200: dropout p must be < 1")
    seeds = torch.randint(0, 2 ** 62, (2,), dtype=torch.int64)      # CPU generator: follows torch.manual_seed
    return _lib.Dropout(max(pa, 0.0), max(po, 0.0), int(seeds[0]), int(seeds[1]))


def dropout_mask(p: float, seed: int, shape, device) -> torch.Tensor:
    """The keep mask (0 or 1/(1-p), fp32) the kernels apply for (p, seed) over a row-major tensor of ``shape``."""
    out = torch.empty(shape, dtype=torch.float32, device=device)
    with torch.cuda.device(device):
        rc = _lib.load().hsenet_dropout_mask(C.c_float(p), C.c_ulonglong(seed), out.numel(), out.data_ptr(),
                                             rt.stream_ptr(device))
    _lib.check(rc, "dropout_mask")
    return out


def _drop_ref(d):
    return C.byref(d) if d is not None else None


def _grad_like(p: torch.Tensor, want: bool):
    return torch.empty(p.shape, dtype=torch.float32, device=p.device) if want else None


class VitTrainFn(torch.autograd.Function):
    """(module, images, image_2d, *parameters) -> (tokens [B,2049,768], patch tokens [B,2048,768])."""

    @staticmethod
    def forward(ctx, module, x, image_2d, *params):
        ctx.set_materialize_grads(False)      # an unused output arrives as None instead of a zero tensor
        lib = _lib.load()
        dev = x.device
        prec = rt.get_precision()
        pc = rt.precision_code(prec)
        act = rt.act_dtype(prec)
        B = x.shape[0]
        L = len(module.blocks)
        stage = module._stage
        with torch.cuda.device(dev):
            payload = module._train_cache.get(module.parameters(), prec, module._build_train_payload)
            xin = x.detach().float().contiguous()
            s2d = None
            if stage == 2:
                s2d = image_2d.detach().to(dev).reshape(B, 32, -1).float().contiguous()
            tape = torch.empty(lib.hsenet_vit_tape_bytes(B, pc, stage, L), dtype=torch.uint8, device=dev)
            ws = rt.workspace(dev, lib.hsenet_vit_train_workspace_bytes(B, pc, stage), "vit_train")
            tokens = torch.empty(B, 2049, 768, dtype=act, device=dev)
            patch = torch.empty(B, 2048, 768, dtype=act, device=dev)
            scores = torch.empty(B, 2048, dtype=torch.float32, device=dev) if stage == 2 else None
            drop = draw_dropout(module.slice_guided_attention, module.training) if stage == 2 else None
            rc = lib.hsenet_vit_forward_train(C.byref(payload["struct"]), xin.data_ptr(), rt.ptr(s2d), B, pc,
                                              tokens.data_ptr(), patch.data_ptr(), rt.ptr(scores), tape.data_ptr(),
                                              tape.numel(), ws.data_ptr(), ws.numel(), _drop_ref(drop),
                                              rt.stream_ptr(dev))
        _lib.check(rc, f"vit_forward_train(stage={stage})")
        ctx.drop = drop
        module.last_dropout = drop
        ctx.module, ctx.prec, ctx.payload, ctx.tape, ctx.xin, ctx.B = module, prec, payload, tape, xin, B
        ctx.param_needs = [p.requires_grad for p in params]
        module.last_scores = scores
        return tokens, patch

    @staticmethod
    def backward(ctx, d_tokens, d_patch):
        if d_tokens is None and d_patch is None:
            return (None,) * (3 + len(ctx.param_needs))
        module, prec, payload, B = ctx.module, ctx.prec, ctx.payload, ctx.B
        lib = _lib.load()
        dev = ctx.xin.device
        pc = rt.precision_code(prec)
        act = rt.act_dtype(prec)
        L = len(module.blocks)
        stage = module._stage
        params = list(module.parameters())
        needs = ctx.param_needs
        grads = {id(p): _grad_like(p, n) for p, n in zip(params, needs)}
        g = lambda p: rt.ptr(grads[id(p)])
        bg = (_lib.BlockGrads * max(L, 1))()
        for i, blk in enumerate(module.blocks):
            e = bg[i]
            e.w_qkv = g(blk.attn.qkv.weight)
            e.w_out, e.b_out = g(blk.attn.out_proj.weight), g(blk.attn.out_proj.bias)
            e.w_fc1, e.b_fc1 = g(blk.mlp.linear1.weight), g(blk.mlp.linear1.bias)
            e.w_fc2, e.b_fc2 = g(blk.mlp.linear2.weight), g(blk.mlp.linear2.bias)
            e.ln1_g, e.ln1_b = g(blk.norm1.weight), g(blk.norm1.bias)
            e.ln2_g, e.ln2_b = g(blk.norm2.weight), g(blk.norm2.bias)
        vg = _lib.VitGrads()
        vg.blocks_host = C.cast(bg, C.c_void_p)
        vg.cls_token = g(module.cls_token)
        vg.pos_embed = g(module.patch_embedding.position_embeddings)
        lin = module.patch_embedding.patch_embeddings[1]
        vg.w_patch, vg.b_patch = g(lin.weight), g(lin.bias)
        vg.norm_g, vg.norm_b = g(module.norm.weight), g(module.norm.bias)
        wkv = bkv = None
        if stage == 2:
            a = module.slice_guided_attention
            vg.w_sq, vg.b_sq = g(a.Wq.weight), g(a.Wq.bias)
            if a.Wk.weight.requires_grad or a.Wv.weight.requires_grad:
                wkv = torch.empty(1536, 768, dtype=torch.float32, device=dev)
            if a.Wk.bias.requires_grad or a.Wv.bias.requires_grad:
                bkv = torch.empty(1536, dtype=torch.float32, device=dev)
            vg.w_skv, vg.b_skv = rt.ptr(wkv), rt.ptr(bkv)
            vg.w_so, vg.b_so = g(a.output_linear.weight), g(a.output_linear.bias)
            vg.sn_g, vg.sn_b = g(a.norm.weight), g(a.norm.bias)
            vg.w_score, vg.b_score = g(module.patch_score_proj.weight), g(module.patch_score_proj.bias)

        def as_act(t, shape):
            if t is None:
                return None
            t = t.detach()
            if t.dtype != act or not t.is_contiguous():
                t = t.to(act).contiguous()
            assert tuple(t.shape) == shape
            return t

        dt = as_act(d_tokens, (B, 2049, 768))
        dp = as_act(d_patch, (B, 2048, 768))
        with torch.cuda.device(dev):
            ws = rt.workspace(dev, lib.hsenet_vit_train_workspace_bytes(B, pc, stage), "vit_train")
            rc = lib.hsenet_vit_backward(C.byref(payload["struct"]), C.byref(payload["struct_t"]), ctx.xin.data_ptr(), B,
                                         pc, rt.ptr(dt), rt.ptr(dp), ctx.tape.data_ptr(), ctx.tape.numel(), C.byref(vg),
                                         ws.data_ptr(), ws.numel(), _drop_ref(ctx.drop), rt.stream_ptr(dev))
        _lib.check(rc, f"vit_backward(stage={stage})")
        if stage == 2:
            a = module.slice_guided_attention
            if wkv is not None:
                if grads[id(a.Wk.weight)] is not None:
                    grads[id(a.Wk.weight)].copy_(wkv[:768])
                if grads[id(a.Wv.weight)] is not None:
                    grads[id(a.Wv.weight)].copy_(wkv[768:])
            if bkv is not None:
                if grads[id(a.Wk.bias)] is not None:
                    grads[id(a.Wk.bias)].copy_(bkv[:768])
                if grads[id(a.Wv.bias)] is not None:
                    grads[id(a.Wv.bias)].copy_(bkv[768:])
        ctx.tape = None
        out = []
        for p in params:
            gr = grads[id(p)]
            out.append(None if gr is None else gr.to(p.dtype))
        return (None, None, None) + tuple(out)


class PackerTrainFn(torch.autograd.Function):
    """(module, visual_inputs [B,2048,768], *parameters) -> [B,128,out_dim]."""

    @staticmethod
    def forward(ctx, module, hr_in, *params):
        ctx.set_materialize_grads(False)
        lib = _lib.load()
        dev = hr_in.device
        prec = rt.get_precision()
        pc = rt.precision_code(prec)
        act = rt.act_dtype(prec)
        B = hr_in.shape[0]
        D = module.out_dim
        with torch.cuda.device(dev):
            payload = module._train_cache.get(module.parameters(), prec, module._build_train_payload)
            hr = hr_in.detach()
            if hr.dtype != act or not hr.is_contiguous():
                hr = hr.to(act).contiguous()
            tape = torch.empty(lib.hsenet_packer_tape_bytes(B, pc, D), dtype=torch.uint8, device=dev)
            ws = rt.workspace(dev, lib.hsenet_packer_train_workspace_bytes(B, pc, D), "packer_train")
            out = torch.empty(B, 128, D, dtype=act, device=dev)
            drop = draw_dropout(module.resolution_attention, module.training)
            rc = lib.hsenet_packer_forward_train(C.byref(payload["struct"]), hr.data_ptr(), B, pc, out.data_ptr(),
                                                 tape.data_ptr(), tape.numel(), ws.data_ptr(), ws.numel(),
                                                 _drop_ref(drop), rt.stream_ptr(dev))
        _lib.check(rc, "packer_forward_train")
        ctx.drop = drop
        module.last_dropout = drop
        ctx.module, ctx.prec, ctx.payload, ctx.tape, ctx.hr, ctx.B = module, prec, payload, tape, hr, B
        ctx.param_needs = [p.requires_grad for p in params]
        ctx.hr_needs = hr_in.requires_grad
        ctx.hr_dtype = hr_in.dtype
        return out

    @staticmethod
    def backward(ctx, d_out):
        if d_out is None:
            return (None,) * (2 + len(ctx.param_needs))
        module, prec, payload, B = ctx.module, ctx.prec, ctx.payload, ctx.B
        lib = _lib.load()
        dev = ctx.hr.device
        pc = rt.precision_code(prec)
        act = rt.act_dtype(prec)
        D = module.out_dim
        params = list(module.parameters())
        grads = {id(p): _grad_like(p, n) for p, n in zip(params, ctx.param_needs)}
        g = lambda p: rt.ptr(grads[id(p)])
        a = module.resolution_attention
        pg = _lib.PackerGrads()
        pg.w_q, pg.b_q = g(a.Wq.weight), g(a.Wq.bias)
        wkv = torch.empty(1536, 768, dtype=torch.float32, device=dev) \
            if (a.Wk.weight.requires_grad or a.Wv.weight.requires_grad) else None
        bkv = torch.empty(1536, dtype=torch.float32, device=dev) \
            if (a.Wk.bias.requires_grad or a.Wv.bias.requires_grad) else None
        pg.w_kv, pg.b_kv = rt.ptr(wkv), rt.ptr(bkv)
        pg.w_o, pg.b_o = g(a.output_linear.weight), g(a.output_linear.bias)
        pg.ln_g, pg.ln_b = g(a.norm.weight), g(a.norm.bias)
        pg.w_p0, pg.b_p0 = g(module.proj_mpls[0].weight), g(module.proj_mpls[0].bias)
        pg.w_p2, pg.b_p2 = g(module.proj_mpls[2].weight), g(module.proj_mpls[2].bias)
        do = d_out.detach()
        if do.dtype != act or not do.is_contiguous():
            do = do.to(act).contiguous()
        d_hr = torch.empty(B, 2048, 768, dtype=torch.float32, device=dev) if ctx.hr_needs else None
        with torch.cuda.device(dev):
            ws = rt.workspace(dev, lib.hsenet_packer_train_workspace_bytes(B, pc, D), "packer_train")
            rc = lib.hsenet_packer_backward(C.byref(payload["struct"]), C.byref(payload["struct_t"]), ctx.hr.data_ptr(), B,
                                            pc, do.data_ptr(), ctx.tape.data_ptr(), ctx.tape.numel(), C.byref(pg),
                                            rt.ptr(d_hr), ws.data_ptr(), ws.numel(), _drop_ref(ctx.drop),
                                            rt.stream_ptr(dev))
        _lib.check(rc, "packer_backward")
        for (wp, src, lo) in ((a.Wk.weight, wkv, 0), (a.Wv.weight, wkv, 768), (a.Wk.bias, bkv, 0), (a.Wv.bias, bkv, 768)):
            if src is not None and grads[id(wp)] is not None:
                grads[id(wp)].copy_(src[lo:lo + 768])
        ctx.tape = None
        out = [None if grads[id(p)] is None else grads[id(p)].to(p.dtype) for p in params]
        return (None, None if d_hr is None else d_hr.to(ctx.hr_dtype)) + tuple(out)
