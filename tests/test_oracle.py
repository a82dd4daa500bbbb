"""CPU tests of the oracle: (a) the restatement against the committed golden fixtures generated from the reference's
own files (tests/golden/make_golden.py); (b) when /root/reference is present (build container), the restatement and
the integer maps against the reference files executed live through the MONAI shim; (c) the shim's known-answer pins
(state-dict key counts 138 / 150 / 288 quoted in the reference's eval script)."""
import hashlib
import os
import sys

import numpy as np
import pytest
import torch

from util import GEOM, O, metrics

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
from recipe import PACKER_ROWS, SAMPLE_ROWS, recipe_inputs, recipe_state_dict, recipe_tokens  # noqa: E402

from oracle import reference_loader as RL  # noqa: E402

needs_ref = pytest.mark.skipif(not RL.available(), reason="/root/reference not present (GPU box)")


def _facade_shapes(kind, layers):
    import hsenet_b200 as H
    if kind == "packer":
        m = H.VisualPacker_3d_phi_v3((32, 256, 256), (4, 16, 16), 768, 3072, "mlp", 2)
    else:
        m = (H.ViT_stage1 if kind == 1 else H.ViT_stage2)(num_layers=layers, **GEOM)
    return {k: tuple(v.shape) for k, v in m.state_dict().items()}


def _golden(name):
    return np.load(os.path.join(HERE, "golden", name))


# ---------------------------------------------------------------------------------------------- fixtures (travel)
def test_maps_against_reference_fixture():
    g = _golden("maps.npz")
    pm = O.patch_gather_map()
    assert hashlib.sha256(pm.tobytes()).digest() == bytes(g["patch_map_sha256"])       # bit exact, whole map
    assert np.array_equal(pm[g["patch_map_row_ids"]], g["patch_map_rows"])
    assert np.array_equal(O.packer_window_map(), g["window_map"])
    assert int(g["n_stage1"]) == 138 and int(g["n_stage2"]) == 150                     # reference's own KAT
    assert len(g["tower_keys"]) == 288 and len(g["packer_keys"]) == 14


def test_facade_state_dict_matches_reference_keys():
    """Names AND order are contractual: train_VLM.py:477-503 copies CLIP weights by position."""
    import hsenet_b200 as H
    g = _golden("maps.npz")
    tower = H.ViT3DTower_dual_encoders(H.VisionConfig())
    assert list(tower.state_dict().keys()) == [str(k) for k in g["tower_keys"]]
    p = H.VisualPacker_3d_phi_v3((32, 256, 256), (4, 16, 16), 768, 3072, "mlp", 2)
    assert list(p.state_dict().keys()) == [str(k) for k in g["packer_keys"]]
    assert p.proj_out_num == 128 and tower.hidden_size == 768


@pytest.mark.parametrize("layers", [2])
def test_restatement_vs_golden_vit(layers):
    x, s = recipe_inputs(1)
    g1 = _golden(f"vit_stage1_L{layers}.npz")
    sd = recipe_state_dict(_facade_shapes(1, layers), seed=layers)
    with torch.no_grad():
        y, h = O.vit_stage1(sd, x)
    assert metrics(y[:, SAMPLE_ROWS], torch.from_numpy(g1["rows"]))["max_rel"] < 1e-5
    assert metrics(y.norm(dim=-1), torch.from_numpy(g1["row_norms"]))["max_rel"] < 1e-5
    assert metrics(h[0][:, SAMPLE_ROWS], torch.from_numpy(g1["hidden0_rows"]))["max_rel"] < 1e-5
    g2 = _golden(f"vit_stage2_L{layers}.npz")
    sd = recipe_state_dict(_facade_shapes(2, layers), seed=100 + layers)
    with torch.no_grad():
        y, _ = O.vit_stage2(sd, x, s)
        sc, att = O.patch_scores(sd, O.patch_embedding(sd, x), s)
    assert metrics(y[:, SAMPLE_ROWS], torch.from_numpy(g2["rows"]))["max_rel"] < 1e-5
    assert metrics(sc, torch.from_numpy(g2["scores"]))["max_rel"] < 1e-5
    assert metrics(att[:, SAMPLE_ROWS[:4]], torch.from_numpy(g2["attn_rows"]))["max_rel"] < 1e-5


def test_restatement_vs_golden_packer():
    g = _golden("packer.npz")
    sd = recipe_state_dict(_facade_shapes("packer", 0), seed=7)
    with torch.no_grad():
        y = O.visual_packer(sd, recipe_tokens(2))
    assert y.shape == (2, 128, 3072)
    assert metrics(y[:, PACKER_ROWS], torch.from_numpy(g["rows"]))["max_rel"] < 1e-5
    assert metrics(y.norm(dim=-1), torch.from_numpy(g["row_norms"]))["max_rel"] < 1e-5


def test_slice_extract_is_per_slice_bilinear():
    """Depth 32 -> 32 makes the trilinear resize a per-slice bilinear one."""
    import torch.nn.functional as F
    x = torch.rand(1, 1, 32, 256, 256, generator=torch.Generator().manual_seed(0))
    a = O.slice_extract(x)
    b = F.interpolate(x[0].permute(1, 0, 2, 3), size=(224, 224), mode="bilinear", align_corners=False)
    assert a.shape == (32, 3, 224, 224)
    assert torch.allclose(a[:, 0], b[:, 0], atol=1e-6) and torch.equal(a[:, 0], a[:, 2])


# ---------------------------------------------------------------------------------------------- live reference
@needs_ref
def test_shim_key_counts_live():
    with RL.quiet():
        vit, pk = RL.vit(), RL.packer()
        t = vit.ViT3DTower_dual_encoders(RL.TowerConfig())
        p = pk.VisualPacker_3d_phi_v3((32, 256, 256), (4, 16, 16), 768, 3072, "mlp", 2)
    keys = list(t.state_dict().keys())
    assert sum("vision_tower_stage1" in k for k in keys) == 138      # eval_HSENet_BIMCV_R_MRG.py:342-344
    assert sum("vision_tower_stage2" in k for k in keys) == 150      # :350-352 ("12 more")
    assert len(keys) == 288 and len(p.state_dict()) == 14


@needs_ref
def test_restatement_vs_reference_live():
    with RL.quiet():
        vit, pk = RL.vit(), RL.packer()
        torch.manual_seed(0)
        m1 = vit.ViT_stage1(num_layers=1, **GEOM).eval()
        m2 = vit.ViT_stage2(num_layers=1, **GEOM).eval()
        p = pk.VisualPacker_3d_phi_v3((32, 256, 256), (4, 16, 16), 768, 3072, "mlp", 2).eval()
        tower = vit.ViT3DTower_dual_encoders(RL.TowerConfig())
    x, s = recipe_inputs(1, seed=5)
    with torch.no_grad():
        r1, _ = m1(x)
        r2, _ = m2(x, s)
        o1, _ = O.vit_stage1(m1.state_dict(), x)
        o2, _ = O.vit_stage2(m2.state_dict(), x, s)
        assert metrics(o1, r1)["max_rel"] < 1e-6 and metrics(o2, r2)["max_rel"] < 1e-6
        assert metrics(O.visual_packer(p.state_dict(), r1[:, 1:]), p(r1[:, 1:]))["max_rel"] < 1e-5
        assert torch.equal(m1.patch_embedding.patch_embeddings[0](x), O.patchify(x))     # integer map, bit exact
    assert tower.hidden_size == 768


class _MaskDropout(torch.nn.Module):
    """Stand-in for an nn.Dropout member of the reference module: multiplies by a preset keep mask."""

    def __init__(self, mask):
        super().__init__()
        self.mask = mask

    def forward(self, x):
        return x * self.mask.reshape(x.shape)


@needs_ref
def test_train_mode_restatement_vs_reference_live():
    """The oracle's ``drop=`` restatement of train mode (dropout on the probabilities, vit.py:31-32 /
    spatial_pooling_projector.py:14-15, and dropout_2 in front of the residual, vit.py:62 / :78) against the reference's own
    modules in .train() with their two nn.Dropout members replaced by the same explicit masks -- outputs and gradients."""
    with RL.quiet():
        vit, pk = RL.vit(), RL.packer()
        torch.manual_seed(0)
        m2 = vit.ViT_stage2(num_layers=1, **GEOM).train()
        p = pk.VisualPacker_3d_phi_v3((32, 256, 256), (4, 16, 16), 768, 3072, "mlp", 2).train()
    g = torch.Generator().manual_seed(3)
    keep = lambda shape: (torch.rand(shape, generator=g) < 0.9).float() / 0.9
    x, s = recipe_inputs(1, seed=5)
    masks = (keep((1, 2048, 32)), keep((1, 2048, 768)))
    a = m2.slice_guided_attention
    a.dropout, a.dropout_2 = _MaskDropout(masks[0]), _MaskDropout(masks[1])
    r2, _ = m2(x, s)
    sd = {k: v.detach().clone().requires_grad_(True) for k, v in m2.state_dict().items()}
    o2, _ = O.vit_stage2(sd, x, s, drop=masks)
    assert metrics(o2.detach(), r2.detach())["max_rel"] < 1e-6
    assert metrics(O.vit_stage2(sd, x, s)[0].detach(), r2.detach())["max_rel"] > 1e-3     # the masks matter
    cot = torch.randn(r2.shape, generator=g)
    (r2 * cot).sum().backward()
    (o2 * cot).sum().backward()
    for name, prm in m2.named_parameters():
        ref_g, got = prm.grad, sd[name].grad
        if ref_g is None or float(ref_g.abs().max()) < 1e-7:      # e.g. the key bias (cancels in the softmax)
            assert got is None or float(got.abs().max()) < 1e-5, name
            continue
        assert metrics(got, ref_g)["max_rel"] < 1e-4, name
    feats = torch.randn(1, 2048, 768, generator=g)
    pm = (keep((1, 128, 16)), keep((1, 128, 768)))
    ra = p.resolution_attention
    ra.dropout, ra.dropout_2 = _MaskDropout(pm[0]), _MaskDropout(pm[1])
    assert metrics(O.visual_packer(p.state_dict(), feats, drop=pm).detach(), p(feats).detach())["max_rel"] < 1e-5


@needs_ref
def test_gather_features_reference_single_process():
    """The reference's gather_features on a 1-process gloo group returns its inputs (dist_utils.py:292-293)."""
    import torch.distributed as dist
    du = RL.dist_utils()
    from hsenet_b200 import gather_features
    if not dist.is_initialized():
        dist.init_process_group("gloo", init_method="tcp://127.0.0.1:29541", rank=0, world_size=1)
    try:
        a, b = torch.randn(4, 768), torch.randn(4, 768)
        ra, rb = du.gather_features(a, b)
        ga, gb = gather_features(a, b)
        assert torch.equal(ra, ga) and torch.equal(rb, gb)
    finally:
        dist.destroy_process_group()


def test_ingest_oracle_restatement():
    """Row f-4 (parity unpinned against the script, see oracle header): shapes, range, and the CropForeground box on a
    volume whose air border survives an identity resample exactly."""
    g = torch.Generator().manual_seed(9)
    raw = torch.randint(-1200, 600, (96, 80, 40), generator=g).float()
    raw[:8] = -2000; raw[-8:] = -2000; raw[:, :8] = -2000; raw[:, -8:] = -2000
    out, res, (mn, mx), box = O.preprocess_volume(raw, 1.0, 0.0, 0.75, 1.5, return_intermediates=True)
    assert out.shape == (1, 32, 256, 256) and res.shape == (40, 96, 80)
    assert (mn, mx) == (-1000.0, 200.0) and box == (0, 8, 8, 40, 88, 72)
    assert float(out.min()) >= 0.0 and float(out.max()) <= 1.0
    assert O.preprocess_resampled_shape((512, 512, 303), 0.7, 1.0) == (int(303 * (1.0 / 1.5)), int(512 * (0.7 / 0.75)),
                                                                      int(512 * (0.7 / 0.75)))
    x = torch.zeros(5, 6, 7)
    assert O.foreground_bbox(x) == ((0, 0, 0), (5, 6, 7))
    x[1, 2, 3] = 1.0; x[3, 4, 6] = 2.0
    assert O.foreground_bbox(x) == ((1, 2, 3), (4, 5, 7))
