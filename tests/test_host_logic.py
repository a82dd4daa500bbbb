"""Host-side logic on CPU: facade contracts, error behaviour, weight cache, world_size-2 gloo gather_features."""
import copy
import os
import sys

import pytest
import torch
import torch.multiprocessing as mp

from util import GEOM

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_constructor_contract_and_errors():
    import hsenet_b200 as H
    with pytest.raises(ValueError):
        H.ViT_stage1(1, (32, 256, 256), (4, 16, 16), hidden_size=770, num_heads=12, pos_embed="perceptron",
                     classification=True)
    with pytest.raises(ValueError):
        H.ViT_stage1(1, (64, 256, 256), (4, 16, 16), pos_embed="perceptron", classification=True)
    with pytest.raises(ValueError):
        H.ViT_stage1(dropout_rate=1.5, **GEOM)
    cfg = H.VisionConfig()
    cfg.vision_tower = "something_else"
    with pytest.raises(ValueError, match="Unknown vision tower"):
        H.build_vision_tower(cfg)
    cfg = H.VisionConfig()
    cfg.mm_projector_type = "nope"
    with pytest.raises(ValueError, match="Unknown projector type"):
        H.build_mm_projector(cfg)
    cfg = H.VisionConfig(select_feature="bogus")
    t = H.build_vision_tower(cfg)
    with pytest.raises(ValueError, match="Unexpected select feature"):
        t(torch.zeros(1, 1, 32, 256, 256), torch.zeros(1, 32, 768))


def test_cpu_inputs_fail_loudly_no_fallback():
    import hsenet_b200 as H
    m = H.ViT_stage1(num_layers=1, **GEOM).eval()
    with torch.no_grad(), pytest.raises(RuntimeError, match="no CPU path"):
        m(torch.zeros(1, 1, 32, 256, 256))
    p = H.VisualPacker_3d_phi_v3((32, 256, 256), (4, 16, 16), 768, 3072, "mlp", 2).eval()
    with torch.no_grad(), pytest.raises(RuntimeError, match="no CPU path"):
        p(torch.zeros(1, 2048, 768))


def test_state_dict_roundtrip_and_deepcopy():
    import hsenet_b200 as H
    a = H.ViT_stage2(num_layers=2, **GEOM)
    b = H.ViT_stage2(num_layers=2, **GEOM)
    b.load_state_dict(a.state_dict(), strict=True)
    for (ka, va), (kb, vb) in zip(a.state_dict().items(), b.state_dict().items()):
        assert ka == kb and torch.equal(va, vb)
    c = copy.deepcopy(a)
    assert list(c.state_dict().keys()) == list(a.state_dict().keys())
    assert a.patch_embedding.patch_embeddings[1].weight.shape == (768, 1024)
    assert a.blocks[0].attn.qkv.bias is None and a.blocks[0].attn.qkv.weight.shape == (2304, 768)


def test_weight_cache_invalidation():
    from hsenet_b200 import runtime as rt
    lin = torch.nn.Linear(4, 4)
    cache, calls = rt.WeightCache(), []
    build = lambda p: calls.append(p) or len(calls)
    assert cache.get(lin.parameters(), "bf16", build) == 1
    assert cache.get(lin.parameters(), "bf16", build) == 1            # unchanged -> cached
    with torch.no_grad():
        lin.weight.add_(1.0)                                          # optimizer-style in-place update
    assert cache.get(lin.parameters(), "bf16", build) == 2
    assert cache.get(lin.parameters(), "fp32_verify", build) == 3     # precision switch
    lin.load_state_dict({k: v.clone() for k, v in lin.state_dict().items()})
    assert cache.get(lin.parameters(), "fp32_verify", build) == 4


def test_precision_context():
    import hsenet_b200 as H
    assert H.get_precision() in ("bf16", "fp32_verify")
    with H.precision("fp32_verify"):
        assert H.get_precision() == "fp32_verify"
        with H.precision("bf16"):
            assert H.get_precision() == "bf16"
        assert H.get_precision() == "fp32_verify"
    with pytest.raises(ValueError):
        H.set_precision("fp8")


# ---- world_size 2 over gloo: packed single all-gather == the reference's two all-gathers, incl. gradients ----------
def _gather_worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    import torch.distributed.nn
    from hsenet_b200 import gather_features
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    try:
        g = torch.Generator().manual_seed(100 + rank)
        img = torch.randn(3, 768, generator=g, requires_grad=True)
        txt = torch.randn(3, 768, generator=g, requires_grad=True)
        ai, at = gather_features(img, txt, rank=rank, world_size=world)
        # the reference's formulation (utils/dist_utils.py:292-293)
        ri = torch.cat(torch.distributed.nn.all_gather(img), dim=0)
        rt_ = torch.cat(torch.distributed.nn.all_gather(txt), dim=0)
        ok = torch.equal(ai, ri) and torch.equal(at, rt_) and ai.shape == (3 * world, 768)
        w = torch.arange(1, 3 * world + 1, dtype=torch.float32).unsqueeze(1)
        (gi, gt) = torch.autograd.grad(((ai * w).sum() + 2 * (at * w).sum()), (img, txt))
        (hi, ht) = torch.autograd.grad(((ri * w).sum() + 2 * (rt_ * w).sum()), (img, txt))
        ok = ok and torch.allclose(gi, hi) and torch.allclose(gt, ht)
        # no-grad variant keeps the local block differentiable (dist_utils.py:300-303)
        bi, bt = gather_features(img, txt, gather_with_grad=False, rank=rank, world_size=world)
        ok = ok and torch.equal(bi.detach(), ri.detach()) and bi.requires_grad
        # mixed dtypes (fp32 image head output, bf16 text features under autocast): each output keeps its input's dtype
        ci, ct = gather_features(img.detach(), txt.detach().to(torch.bfloat16), rank=rank, world_size=world)
        ok = ok and ci.dtype == torch.float32 and ct.dtype == torch.bfloat16 and torch.equal(ci, ri.detach())
        ok = ok and torch.equal(ct, rt_.detach().to(torch.bfloat16))
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


def test_gather_features_gloo_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 400)
    procs = [ctx.Process(target=_gather_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    assert res == [(0, True), (1, True)]


def test_graph_cache_keys_on_generation_and_workspace():
    from hsenet_b200 import runtime as rt
    gc = rt.GraphCache(max_entries=2)
    assert gc.get(("k", 1), 10, 100) is None
    payload = {"keep": []}
    ent = gc.put(("k", 1), 10, 100, {"graph": object()}, keep=(payload,))
    assert gc.get(("k", 1), 10, 100) is ent and ent["_keep"][0] is payload     # entry pins what the graph points at
    assert gc.get(("k", 1), 11, 100) is None          # weights re-packed -> captured pointers are stale
    assert ("k", 1) not in gc.entries                 # ... and a stale entry is dropped, not kept for a later id clash
    ent = gc.put(("k", 1), 10, 100, {"graph": object()})
    assert gc.get(("k", 1), 10, 101) is None          # workspace re-allocated
    gc.put(("k", 1), 10, 100, {})
    gc.put(("k", 2), 10, 100, {})
    gc.put(("k", 3), 10, 100, {})                      # evicts the oldest entry
    assert len(gc.entries) == 2 and ("k", 1) not in gc.entries
    assert isinstance(copy.deepcopy(gc), rt.GraphCache) and not copy.deepcopy(gc).entries


def test_weight_cache_generation_is_monotonic():
    """id(payload) is reused by CPython after a rebuild (A, B, A, B ...); the generation counter never is."""
    from hsenet_b200 import runtime as rt
    wc = rt.WeightCache()
    p = torch.nn.Parameter(torch.zeros(4))
    gens = []
    for prec in ("bf16", "fp32_verify", "bf16", "fp32_verify", "bf16"):
        wc.get([p], prec, lambda key: {"k": key})
        gens.append(wc.generation)
    assert gens == sorted(gens) and len(set(gens)) == len(gens)
    g0 = wc.generation
    wc.get([p], "bf16", lambda key: {"k": key})
    assert wc.generation == g0                          # unchanged weights and precision: no rebuild
    with torch.no_grad():
        p.add_(1.0)                                     # optimizer-style in-place update bumps _version
    wc.get([p], "bf16", lambda key: {"k": key})
    assert wc.generation > g0
    g1 = wc.generation
    p.data.mul_(2.0)                                    # invisible to (data_ptr, _version): needs invalidate()
    wc.invalidate()
    wc.get([p], "bf16", lambda key: {"k": key})
    assert wc.generation > g1


def test_gather_features_keeps_each_dtype():
    """world == 1 returns the inputs untouched; the packing path's dtype handling is covered by the gloo worker below."""
    import hsenet_b200 as H
    a, b = torch.randn(2, 8), torch.randn(2, 8).to(torch.bfloat16)
    x, y = H.gather_features(a, b)
    assert x is a and y is b


def test_bench_reference_arm_contract():
    """`bench.py --impl reference` prints exactly ONE JSON line with the keys the driver reads."""
    import json
    import subprocess
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--workload", "c4",
                          "--steps", "1", "--warmup", "1"], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "volumes/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1
    assert d["e2e"] == {"value": d["value"], "unit": "volumes/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"] and "model" not in d["config"]


def test_dual_tower_two_stream_path_plumbing(monkeypatch):
    """Host logic of ViT3DTower_dual_encoders._forward_concurrent with the CUDA side stubbed out: tower 1 is launched
    inside the side-stream context, tower 2 on the caller's stream, both are finished (cloned) only after the join, and
    .train() mode of the 2E3 encoder is still rejected on this path."""
    import contextlib
    import hsenet_b200 as H
    from hsenet_b200 import vit
    tw = H.ViT3DTower_dual_encoders(H.VisionConfig()).eval()
    tw.shared_patch_embedding = False        # the shared patch-embedding kernel needs a device; this test is host plumbing
    log = []

    class FakeStream:
        def __init__(self, name):
            self.name = name

        def wait_stream(self, other):
            log.append(("wait", self.name, other.name))

    cur, side = FakeStream("cur"), FakeStream("side")
    active = ["cur"]

    @contextlib.contextmanager
    def fake_stream_ctx(s):
        active.append(s.name)
        try:
            yield
        finally:
            active.pop()

    monkeypatch.setattr(vit.rt, "side_stream", lambda dev: side)
    monkeypatch.setattr(torch.cuda, "current_stream", lambda dev=None: cur)
    monkeypatch.setattr(torch.cuda, "stream", fake_stream_ctx)
    for t in (tw.vision_tower_stage1, tw.vision_tower_stage2):
        def prepare(x, s, patch_done=False, _t=t):
            log.append(("prepare", _t._stage, active[-1], s is not None))
            return {"stage": _t._stage}

        def replay(call, _t=t):
            log.append(("replay", call["stage"], active[-1]))
            return (torch.zeros(1, 2049, 768), torch.zeros(1, 2048, 768), None, None), True, torch.float32

        def finish(launched, _t=t, _orig=t._finish):
            log.append(("finish", _t._stage, active[-1]))
            return _orig(launched)
        monkeypatch.setattr(t, "_prepare", prepare)
        monkeypatch.setattr(t, "_replay", replay)
        monkeypatch.setattr(t, "_finish", finish)
    f1, f2 = tw._forward_concurrent(torch.zeros(1, 1, 32, 256, 256), torch.zeros(1, 32, 768))
    assert f1.shape == f2.shape == (1, 2048, 768)
    # graphs are prepared first (tower 1 on the side stream), the streams join (the shared patch embedding would run here),
    # then both replays, then the outputs are cloned on the caller's stream after the final join
    assert log == [("wait", "side", "cur"), ("prepare", 1, "side", False), ("prepare", 2, "cur", True),
                   ("wait", "cur", "side"), ("wait", "side", "cur"), ("replay", 1, "side"), ("replay", 2, "cur"),
                   ("wait", "cur", "side"), ("finish", 1, "cur"), ("finish", 2, "cur")]
    # train() mode with dropout: the tower leaves the graph path (the training kernels apply dropout)
    assert not tw.vision_tower_stage2._dropout_active()
    tw.vision_tower_stage2.train()
    assert tw.vision_tower_stage2._dropout_active() and not tw.vision_tower_stage1._dropout_active()
    assert not tw.vision_tower_stage2.disable_dropout()._dropout_active()


def test_preprocess_host_logic():
    """Row f-4 host side: the resampled shape follows the script's int(orig * cur / target) per axis of the transposed
    volume (and agrees with the oracle), CPU tensors fail loudly, malformed inputs are rejected."""
    import hsenet_b200 as H
    from hsenet_b200 import preprocess as P
    from util import O
    for shape, xy, z in (((512, 512, 303), 0.7, 1.0), ((96, 80, 40), 0.9, 2.0), ((64, 64, 33), 0.75, 1.5)):
        assert P.resampled_shape(shape, xy, z) == O.preprocess_resampled_shape(shape, xy, z)
    assert P.resampled_shape((64, 64, 33), 0.75, 1.5) == (33, 64, 64)          # identity spacing: pure transpose
    with pytest.raises(RuntimeError):
        H.preprocess_ct_volume(torch.zeros(8, 8, 8), 1.0, 0.0, 0.75, 1.5)      # CPU tensor: no fallback
