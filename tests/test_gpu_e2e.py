"""End-to-end parity (-m gpu): the module facades (through the C ABI) against the CPU oracle on identical synthetic
volumes and shared random-init weights.  Tolerances are BASELINE.json's: bf16 cos >= 0.999 and
max|err|/max|ref| <= 2e-2; fp32 verification mode <= 1e-4."""
import pytest
import torch

from util import GEOM, O, assert_bf16, assert_fp32, cpu_state, metrics, randomize_params, synthetic_inputs

pytestmark = pytest.mark.gpu


def _build(cls, layers, seed=0, **kw):
    torch.manual_seed(seed)
    m = cls(num_layers=layers, **GEOM, **kw) if layers is not None else cls(**kw)
    return randomize_params(m).eval()


@pytest.mark.parametrize("layers,B", [(2, 2), (12, 1)])
def test_vit_stage1(cuda, layers, B):
    import hsenet_b200 as H
    m = _build(H.ViT_stage1, layers)
    sd = cpu_state(m)
    x, _ = synthetic_inputs(B)
    ref, ref_h = O.vit_stage1(sd, x)
    m = m.to(cuda)
    m.return_hidden_states = True
    with torch.no_grad():
        with H.precision("fp32_verify"):
            got, hs = m(x.to(cuda))
        assert got.dtype == torch.float32 and got.shape == (B, 2049, 768) and len(hs) == layers
        print("fp32_verify", assert_fp32(got, ref, "vit_stage1 fp32"))
        assert_fp32(hs[0], ref_h[0], "hidden[0] fp32")
        assert_fp32(m.last_patch_tokens, ref[:, 1:], "patch tokens fp32")
        with H.precision("bf16"):
            got, hs = m(x.to(cuda))
        assert got.dtype == torch.bfloat16
        print("bf16", assert_bf16(got, ref, "vit_stage1 bf16"))
        assert_bf16(hs[-1], ref_h[-1], "hidden[-1] bf16")
        assert torch.equal(m.last_patch_tokens, got[:, 1:])


def test_layernorm_folded_into_gemms_matches_separate_kernels(cuda):
    """bf16 path: norm1 / norm2 folded into the GEMM epilogues (default) against the same tower with every LayerNorm
    as its own kernel, and both against the oracle.  LayerNorm gains / biases are randomised (util.randomize_params)
    and given a large common offset here so that mean subtraction and the beta term are really exercised."""
    import hsenet_b200 as H
    m = _build(H.ViT_stage1, 3)
    with torch.no_grad():
        for blk in m.blocks:
            blk.norm1.bias.add_(0.3)
            blk.norm2.weight.mul_(1.7)
    sd = cpu_state(m)
    x, _ = synthetic_inputs(2)
    ref, ref_h = O.vit_stage1(sd, x)
    m = m.to(cuda)
    m.return_hidden_states = True
    m.use_cuda_graph = False            # direct launches: the kernel counter then counts exactly one forward
    outs = {}
    with torch.no_grad(), H.precision("bf16"):
        for fold in (True, False):
            m.fold_layernorm = fold
            m(x.to(cuda))                   # builds the weight cache (the fold kernels run here)
            n0 = H.runtime.kernel_launch_count()
            got, hs = m(x.to(cuda))
            outs[fold] = (got.float().cpu(), hs[-1].float().cpu(), H.runtime.kernel_launch_count() - n0)
            print("fold", fold, assert_bf16(got, ref, f"vit_stage1 bf16 fold={fold}"))
            assert_bf16(hs[-1], ref_h[-1], f"hidden[-1] bf16 fold={fold}")
    mm = metrics(outs[True][0], outs[False][0])
    assert mm["cos"] > 0.9999 and mm["max_rel"] < 1e-2, mm
    # 3 blocks: norm2 of each + norm1 of blocks 1, 2 lose their kernel (5 launches)
    assert outs[False][2] - outs[True][2] == 5, (outs[True][2], outs[False][2])
    # per-slab statistics slots, fixed summation order: the folded path repeats bit for bit
    with torch.no_grad(), H.precision("bf16"):
        m.fold_layernorm = True
        again, _ = m(x.to(cuda))
    assert torch.equal(again.float().cpu(), outs[True][0])


@pytest.mark.parametrize("layers,B", [(2, 2), (12, 1)])
def test_vit_stage2(cuda, layers, B):
    import hsenet_b200 as H
    m = _build(H.ViT_stage2, layers, seed=1)
    with torch.no_grad():    # make the gate informative: default init gives scores ~0.5 everywhere
        m.patch_score_proj.weight.mul_(8.0)
    sd = cpu_state(m)
    x, s = synthetic_inputs(B, seed=77)
    ref, _ = O.vit_stage2(sd, x, s)
    ref_scores, _ = O.patch_scores(sd, O.patch_embedding(sd, x), s)
    m = m.to(cuda)
    with torch.no_grad():
        with H.precision("fp32_verify"):
            got, _ = m(x.to(cuda), s.to(cuda))
        print("fp32_verify", assert_fp32(got, ref, "vit_stage2 fp32"))
        assert_fp32(m.last_scores, ref_scores, "scores fp32")
        with H.precision("bf16"):
            got, _ = m(x.to(cuda), s.to(cuda))
        print("bf16", assert_bf16(got, ref, "vit_stage2 bf16"))
        assert metrics(m.last_scores, ref_scores)["max_rel"] < 2e-2


def test_packer(cuda):
    import hsenet_b200 as H
    torch.manual_seed(3)
    p = randomize_params(H.VisualPacker_3d_phi_v3((32, 256, 256), (4, 16, 16), 768, 3072, "mlp", 2)).eval()
    sd = cpu_state(p)
    g = torch.Generator().manual_seed(5)
    tok = torch.randn(2, 2049, 768, generator=g)
    feats = tok[:, 1:]                                   # the non-contiguous view the tower hands over
    ref = O.visual_packer(sd, feats)
    p = p.to(cuda)
    with torch.no_grad():
        with H.precision("fp32_verify"):
            got = p(tok.to(cuda)[:, 1:])
        assert got.shape == (2, 128, 3072) and p.proj_out_num == 128
        print("fp32_verify", assert_fp32(got, ref, "packer fp32"))
        with H.precision("bf16"):
            got = p(tok.to(cuda)[:, 1:])
        assert got.dtype == torch.bfloat16
        print("bf16", assert_bf16(got, ref, "packer bf16"))


def test_encode_images_full_path(cuda):
    """BASELINE config 3 at B=2: dual tower + two packers -> [B,256,3072]."""
    import hsenet_b200 as H
    torch.manual_seed(0)
    enc = randomize_params(H.HSENetVisualEncoder(H.VisionConfig())).eval()
    tsd = {k: v for k, v in cpu_state(enc.vision_tower).items()}
    p1, p2 = cpu_state(enc.mm_projector), cpu_state(enc.mm_projector2)
    x, s = synthetic_inputs(2)
    ref = O.encode_images(tsd, p1, p2, x, s)
    enc = enc.to(cuda)
    with torch.no_grad():
        with H.precision("fp32_verify"):
            enc.mm_projector.output_dtype = torch.float32
            got = enc(x.to(cuda), s.to(cuda))
        assert got.shape == (2, 256, 3072)
        print("fp32_verify", assert_fp32(got, ref, "encode_images fp32"))
        enc.mm_projector.output_dtype = None
        with H.precision("bf16"):
            got = enc(x.to(cuda), s.to(cuda))
        print("bf16", assert_bf16(got, ref, "encode_images bf16"))


def test_clip_image_head(cuda):
    import hsenet_b200 as H
    torch.manual_seed(2)
    head = randomize_params(H.ClipImageHead()).eval()
    sd = cpu_state(head)
    tok = torch.randn(5, 2049, 768, generator=torch.Generator().manual_seed(1))
    ref = O.clip_encode_image(sd, tok)
    head = head.to(cuda)
    with torch.no_grad():
        with H.precision("fp32_verify"):
            got = head(tok.to(cuda))
        assert_fp32(got, ref, "clip head fp32", tol=1e-5)
        with H.precision("bf16"):
            got = head(tok.to(cuda).to(torch.bfloat16))
        assert_bf16(got, ref, "clip head bf16")


def test_no_cpu_path_and_errors(cuda):
    import hsenet_b200 as H
    m = _build(H.ViT_stage1, 1)
    with torch.no_grad(), pytest.raises(RuntimeError):
        m(torch.zeros(1, 1, 32, 256, 256))                 # CPU tensors: must fail loudly, never fall back
    m = m.to(cuda)
    with torch.no_grad(), pytest.raises(ValueError):
        m(torch.zeros(1, 1, 16, 256, 256, device=cuda))
    # grad enabled + trainable parameters: the training path (row f-1) -- never a silent detach
    y, _ = m(torch.zeros(1, 1, 32, 256, 256, device=cuda))
    assert y.requires_grad and y.grad_fn is not None
    # train-mode dropout of the slice-guided attention: applied by the training kernels, also without grad
    m2 = _build(H.ViT_stage2, 1).to(cuda).train()
    with torch.no_grad():
        m2(torch.zeros(1, 1, 32, 256, 256, device=cuda), torch.zeros(1, 32, 768, device=cuda))
    assert m2.last_dropout is not None and abs(m2.last_dropout.p_attn - 0.1) < 1e-7


def test_splice_into_llm_embeddings(cuda):
    """prepare_inputs_for_multimodal (lamed_arch.py:143-155).  The packers write straight into
    positions 1..256 of the embedding sequence; result must equal the reference's cat(embeds[:, :1], feats, embeds[:, 257:])."""
    import hsenet_b200 as H
    torch.manual_seed(0)
    enc = randomize_params(H.HSENetVisualEncoder(H.VisionConfig())).eval().requires_grad_(False).to(cuda)
    x, s = synthetic_inputs(2, seed=3)
    x, s = x.to(cuda), s.to(cuda)
    L = 300

    class FakeLM:
        def __init__(self, dtype):
            self.emb = torch.nn.Embedding(1000, 3072).to(cuda).to(dtype).requires_grad_(False)
            self.mm_projector, self.mm_projector2 = enc.mm_projector, enc.mm_projector2
        def get_model(self): return self
        def get_vision_tower(self): return enc.vision_tower
        def embed_tokens(self, ids): return self.emb(ids)

    ids = torch.randint(0, 1000, (2, L), device=cuda)
    for dtype in (torch.bfloat16, torch.float32, torch.float16):
        lm = FakeLM(dtype)
        with torch.no_grad(), H.precision("bf16"):
            feats = H.encode_images(lm, x, None, s)                               # [2,256,3072] bf16
            r = H.prepare_inputs_for_multimodal(lm, ids, None, None, None, None, x, s)
            assert r[0] is None and r[4].shape == (2, L, 3072) and r[4].dtype == dtype
            e = lm.embed_tokens(ids)
            ref = torch.cat((e[:, :1], feats.to(dtype), e[:, 257:]), dim=1)        # the reference's formulation
            if dtype == torch.float32:
                # fp32 slot: the GEMM epilogue stores its fp32 accumulator instead of a bf16-rounded value
                assert metrics(r[4][:, 1:257], feats)["max_rel"] < 5e-3
                assert torch.equal(r[4][:, :1], e[:, :1]) and torch.equal(r[4][:, 257:], e[:, 257:])
            else:
                assert torch.equal(r[4], ref)
            # decode step (one token) and missing image: passthrough like lamed_arch.py:148
            r1 = H.prepare_inputs_for_multimodal(lm, ids[:, :1], None, None, None, None, x, s)
            assert r1[0] is ids[:, :1] or torch.equal(r1[0], ids[:, :1])
            assert r1[4] is None
            assert H.prepare_inputs_for_multimodal(lm, ids, None, None, None, None, None, None)[4] is None


@pytest.mark.parametrize("remain,select", [("dual_vits", "patch"), ("dual_vits", "cls_patch"), ("3d_vit", "patch"),
                                           ("2e3_vit", "cls_patch")])
def test_tower_config_variants(cuda, remain, select):
    """ViT3DTower_dual_encoders dispatch (vit.py:934-948): tuple vs single tensor, with / without the cls row."""
    import hsenet_b200 as H
    torch.manual_seed(0)
    cfg = H.VisionConfig(select_feature=select, remain=remain)
    tower = H.build_vision_tower(cfg)
    for t in (tower.vision_tower_stage1, tower.vision_tower_stage2):      # keep the CPU oracle cheap: 1 block each
        del t.blocks[1:]
    tower = randomize_params(tower).eval().requires_grad_(False)
    sd = cpu_state(tower)
    x, s = synthetic_inputs(1, seed=11)
    ref = O.dual_tower(sd, x, s, select_feature=select, remain=remain)
    tower = tower.to(cuda)
    with torch.no_grad(), H.precision("bf16"):
        got = tower(x.to(cuda), s.to(cuda))
    n = 2048 if select == "patch" else 2049
    if remain == "dual_vits":
        assert isinstance(got, tuple) and len(got) == 2
        for g, r in zip(got, ref):
            assert g.shape == (1, n, 768)
            assert_bf16(g, r, f"{remain}/{select}")
    else:
        assert torch.is_tensor(got) and got.shape == (1, n, 768)
        assert_bf16(got, ref, f"{remain}/{select}")
    assert tower.hidden_size == 768 and tower.device.type == "cuda"


def test_concurrent_towers_match_serial_execution(cuda):
    """The dual tower runs its two encoders on two streams (own workspaces, joined before returning): results must be
    bit-identical to running them one after the other, also when the caller itself is on a non-default stream and
    consumes the features right away."""
    import hsenet_b200 as H
    torch.manual_seed(0)
    enc = randomize_params(H.HSENetVisualEncoder(H.VisionConfig())).eval().requires_grad_(False).to(cuda)
    tower = enc.vision_tower
    x, s = synthetic_inputs(2, seed=5)
    x, s = x.to(cuda), s.to(cuda)
    with torch.no_grad(), H.precision("bf16"):
        tower.concurrent_towers = False
        a1, a2 = tower(x, s)
        tower.concurrent_towers = True
        b1, b2 = tower(x, s)
        assert torch.equal(a1, b1) and torch.equal(a2, b2)
        side = torch.cuda.Stream(device=cuda)
        side.wait_stream(torch.cuda.current_stream(cuda))
        with torch.cuda.stream(side):
            for _ in range(3):                      # back-to-back calls reuse the graphs' static buffers
                c1, c2 = tower(x * 1.0, s)
                tot = c1.float().sum() + c2.float().sum()
        torch.cuda.current_stream(cuda).wait_stream(side)
        assert torch.equal(c1, a1) and torch.equal(c2, a2)
        assert torch.isfinite(tot)


def test_input_dtypes_layouts_and_output_dtype(cuda):
    """Inputs arrive as fp32 in the reference even in bf16 runs, but fp16 / bf16 / non-contiguous tensors must work too
    (SURVEY 8b 'Threading / devices'); output_dtype controls the returned dtype."""
    import hsenet_b200 as H
    torch.manual_seed(2)
    m = randomize_params(H.ViT_stage2(num_layers=1, **GEOM)).eval().requires_grad_(False).to(cuda)
    x, s = synthetic_inputs(2, seed=13)
    x, s = x.to(cuda), s.to(cuda)
    with torch.no_grad(), H.precision("bf16"):
        base, _ = m(x, s)
        # (a) non-contiguous volume (a strided view of a larger tensor) and a flattened [B, 32*768] slice tensor
        big = torch.zeros(2, 1, 32, 256, 512, device=cuda)
        big[..., ::2] = x
        y, _ = m(big[..., ::2], s.reshape(2, -1))
        assert torch.equal(y, base)
        # (b) half-precision inputs: same as feeding their fp32 up-casts
        y16, _ = m(x.half(), s.half())
        yref, _ = m(x.half().float(), s.half().float())
        assert torch.equal(y16, yref)
        # (c) output dtype
        m.output_dtype = torch.float32
        y32, _ = m(x, s)
        assert y32.dtype == torch.float32 and torch.equal(y32, base.float())
        m.output_dtype = None
        # (d) graphs off == graphs on (bitwise)
        m.use_cuda_graph = False
        yd, _ = m(x, s)
        m.use_cuda_graph = True
        assert torch.equal(yd, base)
    # (e) weights updated in place (optimizer-style): the bf16 weight cache and the captured graph must follow
    with torch.no_grad(), H.precision("bf16"):
        m.blocks[0].mlp.linear2.weight.mul_(0.5)
        y2, _ = m(x, s)
        assert not torch.equal(y2, base)
        m.blocks[0].mlp.linear2.weight.mul_(2.0)
        y3, _ = m(x, s)
        assert torch.equal(y3, base)


@pytest.mark.parametrize("layers", [2, 12])
def test_slice_trunk_online_branch(cuda, layers):
    """Row f-2: volume -> 32 resized slices -> ViT-B/16 trunk -> [B,32,768], against the restated timm forward (parity
    unpinned at the timm / open_clip boundary: neither is installable offline).  The result feeds ViT_stage2 as image_2d."""
    import hsenet_b200 as H
    torch.manual_seed(3)
    m = H.SliceTrunkViTB16(num_layers=layers).eval()
    g = torch.Generator().manual_seed(17)
    with torch.no_grad():
        for name, p in m.named_parameters():          # timm-like scales; norms / biases perturbed so they cannot hide
            if name.endswith("norm1.weight") or name.endswith("norm2.weight") or name == "norm.weight":
                p.copy_(1.0 + 0.1 * torch.randn(p.shape, generator=g))
            elif name.endswith("bias"):
                p.copy_(0.02 * torch.randn(p.shape, generator=g))
            elif name == "cls_token":
                p.copy_(0.02 * torch.randn(p.shape, generator=g))
    assert len(m.state_dict()) == 4 + 12 * layers + 2
    sd = cpu_state(m)
    x, _ = synthetic_inputs(1, seed=9)
    ref = O.slice_branch(sd, x)
    m = m.to(cuda)
    with torch.no_grad():
        with H.precision("fp32_verify"):
            got = m(x.to(cuda))
        assert got.shape == (1, 32, 768) and got.dtype == torch.float32
        print("slice trunk fp32_verify", assert_fp32(got, ref, "slice trunk fp32"))
        with H.precision("bf16"):
            got = m(x.to(cuda))
        print("slice trunk bf16", assert_bf16(got, ref, "slice trunk bf16"))
        # feeds the 2E3 encoder in place of the offline npy features
        t2 = _build(H.ViT_stage2, 1).to(cuda).requires_grad_(False)
        with H.precision("bf16"):
            y, _ = t2(x.to(cuda), got)
        assert y.shape == (1, 2049, 768) and torch.isfinite(y.float()).all()
