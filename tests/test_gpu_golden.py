"""CUDA path against the committed golden fixtures (-m gpu).  The fixtures were produced by running the reference's OWN
files (tests/golden/make_golden.py: vit.py / spatial_pooling_projector.py through the MONAI shim) on the deterministic
recipe of tests/golden/recipe.py, so these tests pin the kernels to the reference itself, not only to the oracle port.
Also: size-independent properties of the full path at BASELINE.json's batch sizes."""
import os
import sys

import numpy as np
import pytest
import torch

from util import GEOM, assert_bf16, assert_fp32, metrics

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "golden"))
from recipe import PACKER_ROWS, SAMPLE_ROWS, recipe_inputs, recipe_state_dict, recipe_tokens  # noqa: E402

pytestmark = pytest.mark.gpu


def _golden(name):
    return np.load(os.path.join(HERE, "golden", name))


def _load_recipe(m, seed):
    shapes = {k: tuple(v.shape) for k, v in m.state_dict().items()}
    m.load_state_dict(recipe_state_dict(shapes, seed), strict=True)
    return m.eval().requires_grad_(False)


@pytest.mark.parametrize("layers", [2, 12])
def test_vit_stage1_vs_reference_fixture(cuda, layers):
    import hsenet_b200 as H
    g = _golden(f"vit_stage1_L{layers}.npz")
    m = _load_recipe(H.ViT_stage1(num_layers=layers, **GEOM), seed=layers).to(cuda)
    m.return_hidden_states = True
    x, _ = recipe_inputs(1)
    ref_rows, ref_norms = torch.from_numpy(g["rows"]), torch.from_numpy(g["row_norms"])
    with torch.no_grad():
        with H.precision("fp32_verify"):
            y, hs = m(x.to(cuda))
        assert_fp32(y[:, SAMPLE_ROWS], ref_rows, "stage1 rows fp32")
        assert_fp32(y.float().norm(dim=-1), ref_norms, "stage1 row norms fp32")
        assert_fp32(hs[0][:, SAMPLE_ROWS], torch.from_numpy(g["hidden0_rows"]), "stage1 hidden[0] fp32")
        with H.precision("bf16"):
            y, _ = m(x.to(cuda))
        assert_bf16(y[:, SAMPLE_ROWS], ref_rows, "stage1 rows bf16")
        assert metrics(y.float().norm(dim=-1), ref_norms)["max_rel"] < 2e-2


@pytest.mark.parametrize("layers", [2, 12])
def test_vit_stage2_vs_reference_fixture(cuda, layers):
    import hsenet_b200 as H
    g = _golden(f"vit_stage2_L{layers}.npz")
    m = _load_recipe(H.ViT_stage2(num_layers=layers, **GEOM), seed=100 + layers).to(cuda)
    x, s = recipe_inputs(1)
    ref_rows = torch.from_numpy(g["rows"])
    with torch.no_grad():
        with H.precision("fp32_verify"):
            y, _ = m(x.to(cuda), s.to(cuda))
        assert_fp32(y[:, SAMPLE_ROWS], ref_rows, "stage2 rows fp32")
        assert_fp32(m.last_scores, torch.from_numpy(g["scores"]), "stage2 scores fp32")
        with H.precision("bf16"):
            y, _ = m(x.to(cuda), s.to(cuda))
        assert_bf16(y[:, SAMPLE_ROWS], ref_rows, "stage2 rows bf16")
        # the sigmoid gate is an intermediate: the recipe draws patch_score_proj ~ 0.5 N(0,1) (14x the default init), which
        # amplifies bf16 noise in the logit; the north-star tolerance (2e-2) applies to the features checked above
        assert metrics(m.last_scores, torch.from_numpy(g["scores"]))["max_rel"] < 5e-2


def test_packer_vs_reference_fixture(cuda):
    import hsenet_b200 as H
    g = _golden("packer.npz")
    p = _load_recipe(H.VisualPacker_3d_phi_v3((32, 256, 256), (4, 16, 16), 768, 3072, "mlp", 2), seed=7).to(cuda)
    t = recipe_tokens(2).to(cuda)
    ref_rows = torch.from_numpy(g["rows"])
    with torch.no_grad():
        with H.precision("fp32_verify"):
            y = p(t)
        assert_fp32(y[:, PACKER_ROWS], ref_rows, "packer rows fp32", tol=1e-4)
        assert_fp32(y.norm(dim=-1), torch.from_numpy(g["row_norms"]), "packer norms fp32")
        with H.precision("bf16"):
            y = p(t)
        assert_bf16(y[:, PACKER_ROWS], ref_rows, "packer rows bf16")


def test_integer_maps_vs_reference_fixture(cuda):
    """Device-side gather maps against the maps obtained by executing the reference's einops / reshape-permute code."""
    import hashlib
    from hsenet_b200 import _lib
    lib = _lib.load()
    g = _golden("maps.npz")
    st = torch.cuda.current_stream().cuda_stream
    pm = torch.empty(2048, 1024, dtype=torch.int32, device=cuda)
    wm = torch.empty(128, 16, dtype=torch.int32, device=cuda)
    _lib.check(lib.hsenet_patch_gather_map(pm.data_ptr(), st), "map")
    _lib.check(lib.hsenet_packer_window_map(wm.data_ptr(), st), "map")
    assert hashlib.sha256(pm.cpu().numpy().tobytes()).digest() == bytes(g["patch_map_sha256"])
    assert np.array_equal(wm.cpu().numpy(), g["window_map"])


# ---- size-independent properties at full batch sizes (BASELINE configs 2 / 3) -------------------------------------
def test_full_path_batch_properties(cuda):
    """C3-sized run (B = 32): (i) batch-permutation equivariance, (ii) a batch equals the concatenation of its
    halves, (iii) repeatability, (iv) every volume's tokens are finite and the right shape."""
    import hsenet_b200 as H
    torch.manual_seed(0)
    enc = H.HSENetVisualEncoder(H.VisionConfig()).eval().requires_grad_(False).to(cuda)
    g = torch.Generator().manual_seed(5)
    B = 32
    x = torch.rand(B, 1, 32, 256, 256, generator=g).to(cuda)
    s = torch.randn(B, 32, 768, generator=g).to(cuda)
    with torch.no_grad(), H.precision("bf16"):
        y = enc(x, s)
        assert y.shape == (B, 256, 3072) and torch.isfinite(y.float()).all()
        y2 = enc(x, s)
        assert torch.equal(y, y2)                                            # deterministic: no atomics on the path
        perm = torch.randperm(B, generator=g).to(cuda)
        yp = enc(x[perm], s[perm])
        assert torch.equal(yp, y[perm])                                      # volumes are independent units
        ya = enc(x[:8], s[:8])
        assert torch.equal(ya, y[:8])                                        # B = 8 (config 2 size) == slice of B = 32
    # the two 128-token halves come from different towers/packers: they must differ
    assert (y[:, :128].float() - y[:, 128:].float()).abs().max() > 0


@pytest.mark.parametrize("B", [1, 3])
def test_odd_batches_match_oracle(cuda, B):
    import hsenet_b200 as H
    from util import O, cpu_state, randomize_params, synthetic_inputs
    torch.manual_seed(4)
    m = randomize_params(H.ViT_stage1(num_layers=1, **GEOM)).eval()
    sd = cpu_state(m)
    x, _ = synthetic_inputs(B, seed=B)
    ref, _ = O.vit_stage1(sd, x)
    m = m.to(cuda)
    with torch.no_grad():
        with H.precision("fp32_verify"):
            got, _ = m(x.to(cuda))
        assert_fp32(got, ref, f"B={B} fp32")
        with H.precision("bf16"):
            got, _ = m(x.to(cuda))
        assert_bf16(got, ref, f"B={B} bf16")


def test_graph_cache_survives_weight_reloads_and_precision_toggles(cuda):
    """VERDICT r1 / ADVICE: bf16 -> fp32_verify -> bf16 with weight reloads in between used to be able to replay a stale
    graph (entries were validated by id(payload), which CPython re-uses).  Every graph-replayed result must equal a
    graph-free run with the weights that are loaded at that moment."""
    import hsenet_b200 as H
    torch.manual_seed(11)
    m = H.ViT_stage1(num_layers=1, **GEOM).eval().requires_grad_(False).to(cuda)
    ref = H.ViT_stage1(num_layers=1, **GEOM).eval().requires_grad_(False).to(cuda)
    ref.use_cuda_graph = False
    g = torch.Generator().manual_seed(3)
    x = torch.rand(2, 1, 32, 256, 256, generator=g).to(cuda)
    shapes = {k: tuple(v.shape) for k, v in m.state_dict().items()}

    def check(tag):
        ref.load_state_dict(m.state_dict())
        for prec in ("bf16", "fp32_verify", "bf16"):
            with torch.no_grad(), H.precision(prec):
                a, _ = m(x)
                b, _ = ref(x)
            assert torch.equal(a, b), f"{tag}/{prec}: graph replay differs from direct launches"

    check("initial")
    for it in range(3):                                   # reload twice (and once more), toggling precision in between
        m.load_state_dict(recipe_state_dict(shapes, seed=40 + it), strict=True)
        check(f"reload{it}")
        with torch.no_grad():
            m.norm.weight.mul_(1.5)                       # optimizer-style in-place update (bumps _version)
        check(f"inplace{it}")
    # a write through .data is invisible to the signature: refresh_weights() is the documented escape hatch
    m.norm.weight.data.mul_(0.5)
    m.refresh_weights()
    check("data-write + refresh")


def test_report_worst_case_bf16_margin(cuda, capsys):
    """Prints the worst bf16 max|err|/max|ref| and cosine over the 12-layer fixtures so the margin against the 2e-2 gate
    is visible in the driver's GPU test log."""
    import hsenet_b200 as H
    worst = {"max_rel": 0.0, "cos": 1.0}
    for stage, cls, seed in ((1, H.ViT_stage1, 12), (2, H.ViT_stage2, 112)):
        g = _golden(f"vit_stage{stage}_L12.npz")
        m = _load_recipe(cls(num_layers=12, **GEOM), seed=seed).to(cuda)
        x, s = recipe_inputs(1)
        with torch.no_grad(), H.precision("bf16"):
            y, _ = m(x.to(cuda)) if stage == 1 else m(x.to(cuda), s.to(cuda))
        mm = metrics(y[:, SAMPLE_ROWS], torch.from_numpy(g["rows"]))
        worst["max_rel"] = max(worst["max_rel"], mm["max_rel"])
        worst["cos"] = min(worst["cos"], mm["cos"])
    with capsys.disabled():
        print(f"\n[hsenet_b200] worst-case 12-layer bf16 parity vs reference fixtures: max_rel={worst['max_rel']:.4e} "
              f"(gate 2e-2), cos={worst['cos']:.6f} (gate 0.999)")
    assert worst["max_rel"] <= 2e-2 and worst["cos"] >= 0.999
