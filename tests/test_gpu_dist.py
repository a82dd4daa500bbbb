"""gather_features over NCCL: needs two GPUs; skipped on a one-GPU box.

Run with `gpurun --gpus 2 -- python -m pytest tests/test_gpu_dist.py -m gpu -q`.  bench.py's `clip_step` record performs
the same check inside every N > 1 bench run (collective_check)."""
import os

import pytest
import torch
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _nccl_worker(rank, world, port, q):
    import torch.distributed as dist
    import torch.distributed.nn
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        import hsenet_b200 as H
        g = torch.Generator().manual_seed(100 + rank)
        img = torch.randn(32, 768, generator=g).to(dev).requires_grad_(True)
        txt = torch.randn(32, 768, generator=g).to(dev).requires_grad_(True)
        ai, at = H.gather_features(img, txt, rank=rank, world_size=world)
        ri = torch.cat(torch.distributed.nn.all_gather(img), dim=0)         # utils/dist_utils.py:292-293
        rt_ = torch.cat(torch.distributed.nn.all_gather(txt), dim=0)
        ok = torch.equal(ai, ri) and torch.equal(at, rt_) and ai.shape == (32 * world, 768)
        w = torch.arange(1, 32 * world + 1, dtype=torch.float32, device=dev).unsqueeze(1)
        gi, gt = torch.autograd.grad(((ai * w).sum() + 2 * (at * w).sum()), (img, txt))    # reduce-scatter backward
        hi, ht = torch.autograd.grad(((ri * w).sum() + 2 * (rt_ * w).sum()), (img, txt))
        ok = ok and torch.allclose(gi, hi) and torch.allclose(gt, ht)
        # the contrastive step on top (CLIP_stage1.py:141-155): identical loss on every rank
        labels = torch.arange(32 * world, device=dev)
        loss, lpi, _ = H.contrastive_logits(img, txt, torch.tensor(10.0, device=dev), labels)
        ref = torch.nn.functional.cross_entropy(10.0 * ri @ rt_.T, labels)
        ref = (ref + torch.nn.functional.cross_entropy((10.0 * ri @ rt_.T).T, labels)) / 2
        ok = ok and torch.allclose(loss, ref, rtol=1e-5, atol=1e-5) and lpi.shape == (32 * world, 32 * world)
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


def test_gather_features_nccl_world2(cuda):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2)")
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29900 + (os.getpid() % 90)
    procs = [ctx.Process(target=_nccl_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=300) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    assert res == [(0, True), (1, True)]


def _ddp_worker(rank, world, port, q):
    """DistributedDataParallel over the training path (train_CLIP_stage1.py:231-257 runs the
    encoder under accelerate's DDP): the autograd Function returns ordinary parameter gradients, so DDP's bucketed NCCL
    all-reduce (overlapped with the rest of backward by its hooks) must give every rank the mean of the per-rank gradients."""
    import torch.distributed as dist
    from torch.nn.parallel import DistributedDataParallel as DDP
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        import hsenet_b200 as H
        geom = dict(in_channels=1, img_size=(32, 256, 256), patch_size=(4, 16, 16), pos_embed="perceptron",
                    spatial_dims=3, classification=True)
        torch.manual_seed(0)                                   # same initial weights on every rank
        m = H.ViT_stage1(num_layers=1, **geom).to(dev).train()
        xs = [torch.rand(1, 1, 32, 256, 256, generator=torch.Generator().manual_seed(50 + r)) for r in range(world)]
        cot = torch.randn(1, 2049, 768, generator=torch.Generator().manual_seed(9)).to(dev)
        # reference: every rank computes all per-rank gradients locally (no DDP) and averages them
        mean = None
        for r in range(world):
            m.zero_grad(set_to_none=True)
            with H.precision("fp32_verify"):
                y, _ = m(xs[r].to(dev))
            (y * cot).sum().backward()
            g = [p.grad.clone() for p in m.parameters()]
            mean = g if mean is None else [a + b for a, b in zip(mean, g)]
        mean = [a / world for a in mean]
        m.zero_grad(set_to_none=True)
        ddp = DDP(m, device_ids=[rank])
        with H.precision("fp32_verify"):
            y, _ = ddp(xs[rank].to(dev))
        (y * cot).sum().backward()
        ok = True
        for p, ref in zip(m.parameters(), mean):
            scale = float(ref.abs().max()) + 1e-12
            ok = ok and p.grad is not None and float((p.grad - ref).abs().max()) <= 1e-5 * scale + 1e-7
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


def test_ddp_backward_nccl_world2(cuda):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2)")
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29800 + (os.getpid() % 90)
    procs = [ctx.Process(target=_ddp_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=600) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    assert res == [(0, True), (1, True)]
