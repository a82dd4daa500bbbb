"""Gradient parity of the training path (-m gpu): the autograd Functions over
hsenet_*_forward_train / hsenet_*_backward against torch.autograd through the CPU oracle on identical inputs and weights.
fp32 verification mode <= 1e-4 (max|err| / max|ref| per gradient tensor); bf16: cosine >= 0.999."""
import pytest
import torch
import torch.nn.functional as F

from util import GEOM, O, cpu_state, metrics, randomize_params, synthetic_inputs

pytestmark = pytest.mark.gpu

FP32_TOL = 1e-4
BF16_COS = 0.999


def _st():
    return torch.cuda.current_stream().cuda_stream


@pytest.fixture(scope="module")
def lib():
    from hsenet_b200 import _lib
    return _lib.load()


def _attn_ref(qkv, B, S):
    q, k, v = qkv.reshape(B, S, 3, 12, 64).permute(2, 0, 3, 1, 4)
    att = (q @ k.transpose(-1, -2) * 0.125).softmax(-1)
    return (att @ v).permute(0, 2, 1, 3).reshape(B * S, 768)


@pytest.mark.parametrize("prec", [1, 0])
@pytest.mark.parametrize("B,S", [(2, 2049), (2, 130), (3, 197), (1, 64), (1, 1), (2, 129)])
def test_attention_backward(lib, cuda, B, S, prec):
    from hsenet_b200 import _lib
    g = torch.Generator().manual_seed(S * 7 + B)
    dt = torch.float32 if prec == 1 else torch.bfloat16
    qkv = (torch.randn(B * S, 2304, generator=g) * 1.5).to(dt)
    dout = torch.randn(B * S, 768, generator=g).to(dt)
    ref_in = qkv.float().requires_grad_(True)
    ref_out = _attn_ref(ref_in, B, S)
    (ref_grad,) = torch.autograd.grad(ref_out, ref_in, dout.float())
    sp = (S + 127) // 128 * 128
    qd, dd = qkv.to(cuda), dout.to(cuda)
    out = torch.empty(B * S, 768, dtype=dt, device=cuda)
    lse = torch.full((B, 12, sp), float("nan"), device=cuda)
    dvec = torch.empty(B, 12, sp, device=cuda)
    dq = torch.full((B * S, 2304), float("nan"), dtype=dt, device=cuda)
    _lib.check(lib.hsenet_self_attention_train(qd.data_ptr(), out.data_ptr(), lse.data_ptr(), B, S, prec, _st()), "fwd")
    # lse: log2-domain log-sum-exp of the scaled scores; +inf on the padding
    q, k, _ = qkv.float().reshape(B, S, 3, 12, 64).permute(2, 0, 3, 1, 4)
    ref_lse = torch.logsumexp(q @ k.transpose(-1, -2) * 0.125, dim=-1) * 1.4426950408889634
    assert torch.allclose(lse[:, :, :S].cpu(), ref_lse, atol=2e-2 if prec == 0 else 1e-4, rtol=0)
    assert torch.isinf(lse[:, :, S:]).all()
    _lib.check(lib.hsenet_self_attention_backward(qd.data_ptr(), out.data_ptr(), dd.data_ptr(), lse.data_ptr(),
                                                  dvec.data_ptr(), dq.data_ptr(), B, S, prec, _st()), "bwd")
    torch.cuda.synchronize()
    for name, lo in (("dq", 0), ("dk", 768), ("dv", 1536)):
        if float(ref_grad[:, lo:lo + 768].abs().max()) < 1e-6:      # S = 1: softmax over one key, dq = dk = 0 exactly
            assert float(dq[:, lo:lo + 768].float().abs().max()) < 1e-4, name
            continue
        m = metrics(dq[:, lo:lo + 768], ref_grad[:, lo:lo + 768])
        if prec == 1:
            assert m["max_rel"] <= FP32_TOL, (name, m)
        else:
            assert m["cos"] >= BF16_COS and m["max_rel"] <= 3e-2, (name, m)


def _oracle_grads(fn, sd, *inputs, seed=0):
    """Run `fn(sd, *inputs)` (a CPU oracle forward) under autograd with a fixed random cotangent; returns
    (cotangent(s), {name: grad}, input grads)."""
    leaves = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    outs = fn(leaves, *inputs)
    outs = outs if isinstance(outs, tuple) else (outs,)
    g = torch.Generator().manual_seed(seed)
    cots = [torch.randn(o.shape, generator=g) for o in outs]
    loss = sum((o * c).sum() for o, c in zip(outs, cots))
    ins = [t for t in inputs if isinstance(t, torch.Tensor) and t.requires_grad]
    grads = torch.autograd.grad(loss, list(leaves.values()) + ins, allow_unused=True)
    named = dict(zip(leaves.keys(), grads[:len(leaves)]))
    return cots, named, grads[len(leaves):]


def _check_grads(module, ref, prec, skip=()):
    worst = {"max_rel": 0.0, "cos": 1.0}
    for name, p in module.named_parameters():
        if name in skip:
            continue
        r = ref[name]
        assert p.grad is not None, name
        if name.endswith("Wk.bias"):
            # key bias of a softmax attention: a constant added to every key cancels, the true gradient is 0 and both
            # sides only hold rounding noise -- judge the error against the scale of the sibling value-bias gradient
            scale = float(ref[name.replace("Wk.bias", "Wv.bias")].abs().max())
            err = float((p.grad.detach().cpu().float() - r).abs().max())
            assert err <= (FP32_TOL if prec == "fp32_verify" else 2e-2) * scale, (name, err, scale)
            continue
        if r is None or float(r.abs().max()) == 0.0:
            assert float(p.grad.abs().max()) < 1e-6, name
            continue
        m = metrics(p.grad, r)
        if prec == "fp32_verify":
            assert m["max_rel"] <= FP32_TOL, (name, m)
        else:
            assert m["cos"] >= BF16_COS, (name, m)
        worst["max_rel"] = max(worst["max_rel"], m["max_rel"])
        worst["cos"] = min(worst["cos"], m["cos"])
    return worst


@pytest.mark.parametrize("prec", ["fp32_verify", "bf16"])
def test_packer_backward(cuda, prec):
    import hsenet_b200 as H
    torch.manual_seed(3)
    p = randomize_params(H.VisualPacker_3d_phi_v3((32, 256, 256), (4, 16, 16), 768, 3072, "mlp", 2)).eval()
    sd = cpu_state(p)
    g = torch.Generator().manual_seed(9)
    feats = torch.randn(2, 2048, 768, generator=g)
    fr = feats.clone().requires_grad_(True)
    (cot,), ref, (ref_dhr,) = _oracle_grads(lambda s, f: O.visual_packer(s, f), sd, fr, seed=5)
    p = p.to(cuda)
    act = torch.float32 if prec == "fp32_verify" else torch.bfloat16
    x = feats.to(cuda).to(act).requires_grad_(True)
    with H.precision(prec):
        y = p(x)
        assert y.requires_grad and y.shape == (2, 128, 3072)
        (y.float() * cot.to(cuda)).sum().backward()
    print(prec, "packer grads", _check_grads(p, ref, prec))
    m = metrics(x.grad, ref_dhr)
    assert (m["max_rel"] <= FP32_TOL) if prec == "fp32_verify" else (m["cos"] >= BF16_COS), m


@pytest.mark.parametrize("prec", ["fp32_verify", "bf16"])
@pytest.mark.parametrize("stage", [1, 2])
def test_vit_backward(cuda, stage, prec):
    import hsenet_b200 as H
    torch.manual_seed(stage)
    cls = H.ViT_stage1 if stage == 1 else H.ViT_stage2
    m = randomize_params(cls(num_layers=2, **GEOM)).eval()
    sd = cpu_state(m)
    B = 2
    x, s = synthetic_inputs(B, seed=21)
    if stage == 1:
        fn = lambda sd_, x_: O.vit_stage1(sd_, x_)[0]
        cots, ref, _ = _oracle_grads(fn, sd, x, seed=11)
    else:
        fn = lambda sd_, x_, s_: O.vit_stage2(sd_, x_, s_)[0]
        cots, ref, _ = _oracle_grads(fn, sd, x, s, seed=11)
    m = m.to(cuda)
    with H.precision(prec):
        y, hs = (m(x.to(cuda)) if stage == 1 else m(x.to(cuda), s.to(cuda)))
        assert y.requires_grad and hs == []
        (y.float() * cots[0].to(cuda)).sum().backward()
    print(prec, f"stage{stage} grads", _check_grads(m, ref, prec))


def test_patch_output_gradient_and_repeatability(cuda):
    """The tower hands out last_patch_tokens ([:, 1:] of the final norm): gradients through that output alone must match the
    oracle's, and two backward passes must agree bit for bit (no atomics anywhere on the training path)."""
    import hsenet_b200 as H
    torch.manual_seed(5)
    m = randomize_params(H.ViT_stage1(num_layers=1, **GEOM)).eval()
    sd = cpu_state(m)
    x, _ = synthetic_inputs(1, seed=4)
    cots, ref, _ = _oracle_grads(lambda sd_, x_: O.vit_stage1(sd_, x_)[0][:, 1:], sd, x, seed=2)
    m = m.to(cuda)
    runs = []
    for _ in range(2):
        m.zero_grad(set_to_none=True)
        with H.precision("fp32_verify"):
            m(x.to(cuda))
            (m.last_patch_tokens * cots[0].to(cuda)).sum().backward()
        runs.append({n: p.grad.clone() for n, p in m.named_parameters()})
    _check_grads(m, ref, "fp32_verify")
    assert all(torch.equal(runs[0][n], runs[1][n]) for n in runs[0])
    with H.precision("bf16"):
        a = []
        for _ in range(2):
            m.zero_grad(set_to_none=True)
            m(x.to(cuda))
            (m.last_patch_tokens.float() * cots[0].to(cuda)).sum().backward()
            a.append({n: p.grad.clone() for n, p in m.named_parameters()})
    assert all(torch.equal(a[0][n], a[1][n]) for n in a[0])


def test_clip_stage1_training_step_smoke(cuda):
    """One step of the M3DCLIP_stage1.forward formulation (CLIP_stage1.py:104-139): ViT_stage1 -> cls head -> contrastive
    loss -> backward -> optimizer step; all gradients finite, loss decreases on the same batch."""
    import hsenet_b200 as H
    torch.manual_seed(0)
    vit = H.ViT_stage1(num_layers=2, **GEOM).to(cuda).train()
    head = H.ClipImageHead().to(cuda).train()
    x, _ = synthetic_inputs(4, seed=8)
    x = x.to(cuda)
    text = F.normalize(torch.randn(4, 768, generator=torch.Generator().manual_seed(1))).to(cuda)
    scale = torch.tensor(1.0 / 0.07, device=cuda)
    labels = torch.arange(4, device=cuda)
    params = list(vit.parameters()) + list(head.parameters())
    opt = torch.optim.SGD(params, lr=1e-3)
    losses = []
    for _ in range(3):
        opt.zero_grad(set_to_none=True)
        tokens, _ = vit(x)
        emb = head(tokens)
        loss, _, _ = H.contrastive_logits(emb, text, scale, labels)
        loss.backward()
        assert all(p.grad is not None and torch.isfinite(p.grad).all() for p in params)
        opt.step()
        losses.append(float(loss.detach()))
    assert losses[-1] < losses[0], losses


def _masks(drop, shapes, cuda):
    """The keep masks the kernels applied for this forward (hsenet_dropout_mask), on the CPU for the oracle."""
    from hsenet_b200 import training as tr
    assert drop is not None
    return (tr.dropout_mask(drop.p_attn, drop.seed_attn, shapes[0], cuda).cpu(),
            tr.dropout_mask(drop.p_out, drop.seed_out, shapes[1], cuda).cpu())


def test_dropout_mask_statistics(cuda):
    """Counter-based keep mask: values in {0, 1/(1-p)}, keep rate 1-p, different seeds decorrelated, same seed repeatable,
    p = 0 -> all ones."""
    from hsenet_b200 import training as tr
    n = 1 << 20
    for p in (0.1, 0.5):
        m = tr.dropout_mask(p, 1234, (n,), cuda)
        vals = torch.unique(m)
        assert vals.numel() == 2 and float(vals[0]) == 0.0 and abs(float(vals[1]) - 1.0 / (1.0 - p)) < 1e-6
        keep = float((m > 0).float().mean())
        assert abs(keep - (1.0 - p)) < 4.0 * (p * (1 - p) / n) ** 0.5 + 1e-4, (p, keep)
        assert torch.equal(m, tr.dropout_mask(p, 1234, (n,), cuda))
        m2 = tr.dropout_mask(p, 1235, (n,), cuda)
        both = float(((m > 0) & (m2 > 0)).float().mean())
        assert abs(both - (1.0 - p) ** 2) < 5e-3, (p, both)
        # no structure along rows of 32 / 768 (the shapes the kernels index)
        assert abs(float((m.view(-1, 32)[:, 0] > 0).float().mean()) - (1.0 - p)) < 2e-2
    assert bool((tr.dropout_mask(0.0, 7, (1000,), cuda) == 1.0).all())


@pytest.mark.parametrize("prec", ["fp32_verify", "bf16"])
def test_packer_train_mode_dropout(cuda, prec):
    """train() mode: Dropout(p=0.1) on the window-attention probabilities and on output_linear's result
    (spatial_pooling_projector.py:58-59, 76-78) is applied by the kernels; forward and every gradient match the oracle run
    with the SAME keep masks.  A second forward draws new masks; eval() is unchanged."""
    import hsenet_b200 as H
    torch.manual_seed(3)
    p = randomize_params(H.VisualPacker_3d_phi_v3((32, 256, 256), (4, 16, 16), 768, 3072, "mlp", 2))
    sd = cpu_state(p)
    feats = torch.randn(2, 2048, 768, generator=torch.Generator().manual_seed(9))
    p = p.to(cuda).train()
    assert p._dropout_active()
    act = torch.float32 if prec == "fp32_verify" else torch.bfloat16
    x = feats.to(cuda).to(act).requires_grad_(True)
    torch.manual_seed(77)
    with H.precision(prec):
        y = p(x)
    drop = p.last_dropout
    assert abs(drop.p_attn - 0.1) < 1e-7 and abs(drop.p_out - 0.1) < 1e-7 and drop.seed_attn != drop.seed_out
    masks = _masks(drop, ((2, 128, 16), (2, 128, 768)), cuda)
    assert 0.8 < float((masks[0] > 0).float().mean()) < 0.97
    fr = feats.clone().requires_grad_(True)
    (cot,), ref, (ref_dhr,) = _oracle_grads(lambda s, f: O.visual_packer(s, f, drop=masks), sd, fr, seed=5)
    ref_y = O.visual_packer(sd, feats, drop=masks)
    m = metrics(y, ref_y)
    assert (m["max_rel"] <= FP32_TOL) if prec == "fp32_verify" else (m["cos"] >= BF16_COS), m
    no_drop = metrics(y, O.visual_packer(sd, feats))
    assert no_drop["max_rel"] > 1e-2, no_drop            # the masks really changed the result
    (y.float() * cot.to(cuda)).sum().backward()
    print(prec, "packer train-mode grads", _check_grads(p, ref, prec))
    m = metrics(x.grad, ref_dhr)
    assert (m["max_rel"] <= FP32_TOL) if prec == "fp32_verify" else (m["cos"] >= BF16_COS), m
    with H.precision(prec), torch.no_grad():
        y2 = p(x)                                          # train() without grad: still dropout, new masks
        assert p.last_dropout.seed_attn != drop.seed_attn
        assert not torch.equal(y2, y.detach())
        out = torch.empty(2, 128, 3072, dtype=act, device=cuda)
        with pytest.raises(RuntimeError):
            p.forward_into(x.detach(), out, 0)             # the raw-pointer inference path refuses to skip dropout
        p.eval()
        ye = p(x)
    me = metrics(ye, O.visual_packer(sd, feats))
    assert (me["max_rel"] <= FP32_TOL) if prec == "fp32_verify" else (me["cos"] >= BF16_COS), me
    # same torch seed -> same masks -> same bits
    p.train()
    outs = []
    for _ in range(2):
        torch.manual_seed(5)
        with H.precision(prec), torch.no_grad():
            outs.append(p(x).clone())
    assert torch.equal(outs[0], outs[1])
    assert p.disable_dropout()._dropout_active() is False


@pytest.mark.parametrize("prec", ["fp32_verify", "bf16"])
def test_vit_stage2_train_mode_dropout(cuda, prec):
    """ViT_stage2.train(): the two Dropout(p=0.1) of slice_guided_attention (vit.py:46-47, 60-62) inside the kernels; output
    and gradients against the oracle with the same masks."""
    import hsenet_b200 as H
    torch.manual_seed(2)
    m = randomize_params(H.ViT_stage2(num_layers=1, **GEOM))
    with torch.no_grad():
        m.patch_score_proj.weight.mul_(8.0)              # informative gate (default init gives scores ~0.5)
    sd = cpu_state(m)
    B = 2
    x, s = synthetic_inputs(B, seed=21)
    m = m.to(cuda).train()
    torch.manual_seed(123)
    with H.precision(prec):
        y, _ = m(x.to(cuda), s.to(cuda))
    masks = _masks(m.last_dropout, ((B, 2048, 32), (B, 2048, 768)), cuda)
    fn = lambda sd_, x_, s_: O.vit_stage2(sd_, x_, s_, drop=masks)[0]
    cots, ref, _ = _oracle_grads(fn, sd, x, s, seed=11)
    ref_y = O.vit_stage2(sd, x, s, drop=masks)[0]
    mm = metrics(y, ref_y)
    assert (mm["max_rel"] <= FP32_TOL) if prec == "fp32_verify" else (mm["cos"] >= BF16_COS), mm
    ref_scores, _ = O.patch_scores(sd, O.patch_embedding(sd, x), s, drop=masks)
    plain_scores, _ = O.patch_scores(sd, O.patch_embedding(sd, x), s)
    ms = metrics(m.last_scores, ref_scores)
    assert ms["max_rel"] <= (FP32_TOL if prec == "fp32_verify" else 2e-2), ms
    assert metrics(m.last_scores, plain_scores)["max_rel"] > 10 * ms["max_rel"]      # dropout moved the scores
    (y.float() * cots[0].to(cuda)).sum().backward()
    print(prec, "stage2 train-mode grads", _check_grads(m, ref, prec))
