#!/usr/bin/env python
"""bench.py -- CT volumes/s of the HSENet visual-encoding hot path on B200 (see BASELINE.json / SURVEY.md section 8d).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload c3|c2|c4] [--batch B] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W

One "step" = one pass of the hot path over one batch of synthetic volumes per GPU.
  workload c3 (default, BASELINE.json configs[2] = the full north-star path, lamed_arch.py:122-141): dual encoder + the two
      spatial packers -> [B,256,3072], batch 32 per GPU, bf16.  1033.54 algorithmic GFLOP per volume.
  workload c2 (configs[1]): dual encoder forward only, batch 8 per GPU. 1017.22 GFLOP/volume.  Also measured (short) inside
      the default run and reported under the extra key "c2".
  workload c4 (configs[3]): stage-1 CLIP image side with the packed NCCL all-gather; at N > 1 the default run also reports
      it under "clip_step" together with an in-bench check of the collective's result.
Extra keys of the default line: "c2", "clip_step" (N > 1), "gpu_eager_baseline" (the reference algorithm as eager PyTorch
under bf16 autocast on the same B200, N = 1), "roofline_attention" (tensor and MUFU bounds), "roofline_packer" and
"roofline_layernorm" (HBM), "cpu_baseline".
Prints ONE JSON line (rank 0).  `value` = volumes/s with inputs resident in HBM (CUDA events, max over ranks);
`e2e` = the same through the public module API with pinned HOST inputs, H2D and D2H inside the timed region.
`--impl reference` times the reference algorithm's CPU path (the oracle port, fp32 eager PyTorch, all host threads) on a
bounded sample of the same workload; it is the one place outside tests/ and smoke() where oracle/ is executed.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

GFLOP_PER_VOLUME = {"c2": 1017.22, "c3": 1033.54, "c4": 506.05}
DEFAULT_BATCH = {"c2": 8, "c3": 32, "c4": 32}
WORKLOAD_NAME = {
    "c2": "C2 dual encoder (ViT_stage1 + ViT_stage2/2E3) forward, bf16",
    "c3": "C3 dual encoder + two spatial packers -> [B,256,3072], bf16",
    "c4": "C4 stage-1 CLIP image side: ViT_stage1 -> cls head -> packed all-gather of image/text embeddings -> logits",
}


def load_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        d = json.load(open(p))
        return {"bf16_sustained": d.get("bf16_tflops_sustained", 1402.0), "bf16_burst": d.get("bf16_tflops", 1673.1),
                "hbm_gbs": d.get("hbm_gbs", 6555.8), "source": "measured"}
    return {"bf16_sustained": 1400.0, "bf16_burst": 1590.0, "hbm_gbs": 6650.0, "source": "fallback"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    FIELDS = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                 "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), line.strip()))

    def stop(self, t0=None, t1=None):
        """Summary of the samples taken inside the host-time window [t0, t1] (the timed region; all samples if that window
        caught none).  The sampler is started BEFORE the warm-up steps so that nvidia-smi's start-up (NVML init, ~0.1 s)
        neither overlaps the timed steps nor contributes idle-clock samples."""
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        rows = [r for (t, r) in self.rows if t0 is None or (t0 <= t <= (t1 if t1 is not None else t))]
        if not rows:
            rows = [r for (_, r) in self.rows]
        for r in rows:
            parts = [p.strip() for p in r.split(",")]
            if len(parts) < 6:
                continue
            try:
                sm.append(float(parts[0]))
                mx = float(parts[1])
            except ValueError:
                continue
            for n, v in zip(names, parts[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


def build_modules(workload, device):
    import hsenet_b200 as H
    torch.manual_seed(0)
    enc = H.HSENetVisualEncoder(H.VisionConfig(), use_parallel_projector=True).eval().requires_grad_(False)
    return enc.to(device)


def make_inputs(B, rank, n_sets, pinned):
    g = torch.Generator().manual_seed(1234 + rank)
    sets = []
    for _ in range(n_sets):
        x = torch.rand(B, 1, 32, 256, 256, generator=g)
        s = torch.randn(B, 32, 768, generator=g)
        if pinned:
            x, s = x.pin_memory(), s.pin_memory()
        sets.append((x, s))
    return sets


_c4_state = {}


def run_step(enc, workload, x, s):
    if workload == "c2":
        return enc.vision_tower(x, s)                 # (feat_stage1 [B,2048,768], feat_stage2 [B,2048,768])
    if workload == "c4":
        import hsenet_b200 as H
        if "head" not in _c4_state:
            torch.manual_seed(1)
            _c4_state["head"] = H.ClipImageHead().eval().requires_grad_(False).to(x.device)
            g = torch.Generator().manual_seed(99 + int(os.environ.get("RANK", "0")))      # every rank its own captions
            _c4_state["text"] = torch.nn.functional.normalize(torch.randn(x.shape[0], 768, generator=g)).to(x.device)
            _c4_state["scale"] = torch.tensor(1.0 / 0.07, device=x.device)
        tokens, _ = enc.vision_tower.vision_tower_stage1(x)
        emb = _c4_state["head"](tokens)                                    # [B_loc,768] unit-norm, fp32
        _c4_state["emb"] = emb
        _, lpi, _ = H.contrastive_logits(emb, _c4_state["text"], _c4_state["scale"])   # NCCL all-gather inside
        return lpi                                                          # [B_global, B_global]
    return enc(x, s)                                  # [B,256,3072]


def cpu_reference_run(workload, threads, reps):
    """The reference algorithm on host cores: oracle port (fp32 eager PyTorch).  Bounded sample: 1 volume per rep."""
    from oracle import hsenet_oracle as O
    import hsenet_b200 as H
    torch.set_num_threads(threads)
    torch.manual_seed(0)
    enc = H.HSENetVisualEncoder(H.VisionConfig(), use_parallel_projector=True).eval()
    tsd = {k: v.detach().float() for k, v in enc.vision_tower.state_dict().items()}
    p1 = {k: v.detach().float() for k, v in enc.mm_projector.state_dict().items()}
    p2 = {k: v.detach().float() for k, v in enc.mm_projector2.state_dict().items()}
    g = torch.Generator().manual_seed(1234)
    x = torch.rand(1, 1, 32, 256, 256, generator=g)
    s = torch.randn(1, 32, 768, generator=g)
    times = []
    with torch.no_grad():
        for i in range(reps + 1):
            t0 = time.perf_counter()
            if workload == "c2":
                O.dual_tower(tsd, x, s)
            elif workload == "c4":
                O.vit_stage1({k[len("vision_tower_stage1."):]: v for k, v in tsd.items()
                              if k.startswith("vision_tower_stage1.")}, x)
            else:
                O.encode_images(tsd, p1, p2, x, s)
            dt = time.perf_counter() - t0
            if i > 0:
                times.append(dt)
    return times


def _event_ms(fn, reps, dev):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(dev)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize(dev)
    return e0.elapsed_time(e1) / reps


def measure_c2(enc, dev, rank, world, barrier, max_over_ranks, steps=10):
    """BASELINE configs[1] (dual encoder forward, batch 8 per GPU) as a short extra measurement."""
    B2 = DEFAULT_BATCH["c2"]
    sets = [(x.to(dev), s.to(dev)) for x, s in make_inputs(B2, rank, 4, pinned=False)]
    for i in range(3):
        run_step(enc, "c2", *sets[i % 4])
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for i in range(steps):
        run_step(enc, "c2", *sets[i % 4])
    e1.record()
    barrier()
    ms = max_over_ranks(e0.elapsed_time(e1))
    v = B2 * world * steps / (ms * 1e-3)
    return {"workload": WORKLOAD_NAME["c2"], "value": v, "unit": "volumes/s", "volumes_per_gpu_per_step": B2,
            "steps": steps, "warmup": 3, "ms_per_step": ms / steps,
            "step_tflops_per_gpu": v / world * GFLOP_PER_VOLUME["c2"] / 1e3}


def measure_clip_step(enc, dev, rank, world, x, s, barrier, max_over_ranks, steps=5):
    """BASELINE configs[3]: stage-1 ViT -> cls head -> gather_features (ONE packed NCCL all-gather) -> [B,B] logits, with the
    all-gather timed on its own and its result checked on every rank against the reference formulation
    (utils/dist_utils.py:292-293: rank-ordered concatenation of per-rank all_gather lists)."""
    import torch.distributed as dist
    import hsenet_b200 as H
    B = x.shape[0]
    for _ in range(2):
        run_step(enc, "c4", x, s)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(steps):
        lpi = run_step(enc, "c4", x, s)
    e1.record()
    barrier()
    ms = max_over_ranks(e0.elapsed_time(e1))
    emb, text = _c4_state["emb"], _c4_state["text"]
    ag_ms = max_over_ranks(_event_ms(lambda: H.gather_features(emb, text), 20, dev))
    all_i, all_t = H.gather_features(emb, text)
    li = [torch.empty_like(emb) for _ in range(world)]
    lt = [torch.empty_like(text) for _ in range(world)]
    dist.all_gather(li, emb)
    dist.all_gather(lt, text)
    ok = (torch.equal(all_i, torch.cat(li, 0)) and torch.equal(all_t, torch.cat(lt, 0))
          and torch.equal(all_i[rank * B:(rank + 1) * B], emb) and tuple(lpi.shape) == (B * world, B * world)
          and bool(torch.isfinite(lpi).all()))
    flag = torch.tensor([1 if ok else 0], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    v = B * world * steps / (ms * 1e-3)
    return {"workload": WORKLOAD_NAME["c4"], "value": v, "unit": "volumes/s", "ranks": world,
            "volumes_per_gpu_per_step": B, "global_batch": B * world, "steps": steps, "ms_per_step": ms / steps,
            "all_gather_us": ag_ms * 1e3, "all_gather_bytes_per_rank": int(emb.numel() * 4 + text.numel() * 4),
            "logits_shape": list(lpi.shape), "backend": dist.get_backend(),
            "collective_check": "ok" if int(flag.item()) == 1 else "MISMATCH",
            "check": "every rank: packed all_gather_into_tensor result == rank-ordered cat of dist.all_gather lists "
                     "(bit-exact), own block == local embeddings, logits finite; MIN-reduced over ranks"}


def measure_train_step(dev, B=8, steps=3):
    """forward + backward of a trainable 12-layer ViT_stage1 through the autograd Function
    (stage-1 CLIP image side, train_CLIP_stage1.py:231-257), batch 8, bf16.  An extra record, not the headline."""
    import hsenet_b200 as H
    torch.manual_seed(0)
    vit = H.ViT_stage1(1, (32, 256, 256), (4, 16, 16), pos_embed="perceptron", spatial_dims=3, classification=True).to(dev)
    x = torch.rand(B, 1, 32, 256, 256, generator=torch.Generator().manual_seed(5)).to(dev)

    def step():
        vit.zero_grad(set_to_none=True)
        y, _ = vit(x)
        y.float().square().mean().backward()

    step()
    ms = _event_ms(step, steps, dev)
    finite = all(p.grad is not None and bool(torch.isfinite(p.grad).all()) for p in vit.parameters())
    del vit
    H.release_workspaces()
    torch.cuda.empty_cache()
    return {"workload": "ViT_stage1 (12 layers) forward + backward through hsenet_vit_forward_train / hsenet_vit_backward",
            "value": B / (ms * 1e-3), "unit": "volumes/s", "ms_per_step": ms, "volumes_per_step": B, "steps": steps,
            "tflops_at_3x_forward": 3 * 506.05 * B / ms, "grads_finite": finite}


def gpu_eager_baseline(enc, dev, B=8, reps=3):
    """The reference algorithm (oracle port of vit.py / packer, materialised scores, eager PyTorch kernels) on the SAME
    B200 under bf16 autocast: BASELINE.md section 4's 'reference on the same box' number.  A baseline leg, like cpu_baseline."""
    from oracle import hsenet_oracle as O
    tsd = {k: v.detach().float() for k, v in enc.vision_tower.state_dict().items()}
    p1 = {k: v.detach().float() for k, v in enc.mm_projector.state_dict().items()}
    p2 = {k: v.detach().float() for k, v in enc.mm_projector2.state_dict().items()}
    g = torch.Generator().manual_seed(4321)
    x = torch.rand(B, 1, 32, 256, 256, generator=g).to(dev)
    s = torch.randn(B, 32, 768, generator=g).to(dev)

    def step():
        with torch.autocast("cuda", dtype=torch.bfloat16):
            return O.encode_images(tsd, p1, p2, x, s)

    step()
    ms = _event_ms(step, reps, dev)
    del tsd, p1, p2
    torch.cuda.empty_cache()
    return {"value": B / (ms * 1e-3), "unit": "volumes/s", "ms_per_step": ms, "volumes_per_step": B, "steps": reps,
            "kind": "port", "dtype": "bf16 autocast (fp32 softmax / LayerNorm)",
            "how": "oracle port of the reference modules run as eager PyTorch (cuBLAS GEMMs, materialised "
                   "[B,12,2049,2049] scores) on the same GPU, full C3 path"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--workload", default="c3", choices=["c2", "c3", "c4"])
    ap.add_argument("--batch", type=int, default=0, help="volumes per GPU per step (default: 8 for c2, 32 for c3)")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the c2 / clip_step / gpu_eager_baseline records")
    args = ap.parse_args()

    # The contract is ONE JSON line on stdout.  Libraries (NCCL prints its version banner from C) write to fd 1 too, so
    # everything but the final line is routed to stderr.
    sys.stdout.flush()
    saved_stdout_fd = os.dup(1)
    os.dup2(2, 1)

    def emit(obj):
        sys.stdout.flush()
        os.dup2(saved_stdout_fd, 1)
        print(json.dumps(obj), flush=True)

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    B = args.batch or DEFAULT_BATCH[args.workload]
    W = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    K = max(args.steps, 1)
    gflop = GFLOP_PER_VOLUME[args.workload]
    config = {"workload": WORKLOAD_NAME[args.workload], "volumes_per_gpu_per_step": B, "global_batch": B * world,
              "volume": "1x32x256x256 fp32 in [0,1]", "tokens": 2049, "hidden": 768, "layers": 12,
              "parallelism": f"dp{world} (independent volumes per rank, no data-path collective)"}

    # ------------------------------------------------------------------ reference arm (CPU)
    if args.impl == "reference":
        if rank != 0:
            return
        threads = os.cpu_count() or 1
        reps = max(1, min(K, 3))
        times = cpu_reference_run(args.workload, threads, reps)
        t = statistics.median(times)
        v = 1.0 / t
        sample = f"1 volume per step, {len(times)} timed steps after 1 warm-up (median), fp32 eager PyTorch oracle port"
        emit({
            "impl": "reference", "metric": "CT volumes/s encoded", "value": v, "unit": "volumes/s", "n_gpus": world,
            "steps": len(times), "warmup": 1, "ms_per_step": t * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": config,
            "cpu_baseline": {"value": v, "unit": "volumes/s", "cores": threads, "kind": "port", "sample": sample},
            "e2e": {"value": v, "unit": "volumes/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        })
        return

    # ------------------------------------------------------------------ our arm (B200)
    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl ours needs a CUDA device (no CPU fallback exists)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
    import hsenet_b200 as H
    from hsenet_b200 import _lib
    lib = _lib.load()
    H.set_precision("bf16")
    enc = build_modules(args.workload, dev)
    n_sets = 4 if B <= 8 else 2                           # >= 268 MB of rotating inputs > 126 MB L2: inputs come from HBM
    host_sets = make_inputs(B, rank, n_sets, pinned=True)
    dev_sets = [(x.to(dev), s.to(dev)) for x, s in host_sets]
    _t1 = getattr(getattr(enc, "vision_tower", None), "vision_tower_stage1", None)
    if _t1 is not None:
        config["layernorm_fold"] = bool(_t1.fold_layernorm)                                # HSENET_LN_FOLD (default on)
    config["l2_policy"] = (f"{n_sets} rotating input sets ({n_sets * B * 8.39:.0f} MB) + ~{B * 0.28:.1f} GB of "
                           "activations per step, both larger than the 126 MB L2; no explicit flush")

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def max_over_ranks(ms):
        if world > 1:
            t = torch.tensor([ms], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            return float(t.item())
        return ms

    with torch.no_grad():
        sampler = ClockSampler(local_rank)
        if rank == 0:
            sampler.start()
        for i in range(W):
            run_step(enc, args.workload, *dev_sets[i % n_sets])
        barrier()
        from hsenet_b200 import runtime as hrt
        launches0 = hrt.kernel_launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        t_host0 = time.perf_counter()
        e0.record()
        for i in range(K):
            out = run_step(enc, args.workload, *dev_sets[i % n_sets])
        e1.record()
        host_enqueue_ms = (time.perf_counter() - t_host0) * 1e3 / K
        barrier()
        t_host1 = time.perf_counter()
        ms_total = max_over_ranks(e0.elapsed_time(e1))
        launches = hrt.kernel_launch_count() - launches0
        clocks = sampler.stop(t_host0, t_host1 + 0.1) if rank == 0 else None

        # ---- roofline pass: the same K steps, instrumented with CUDA events around every launch of the library ----
        # (direct launches: the per-launch event hooks live in the library's launchers, which graph replays bypass)
        for tw in (enc.vision_tower.vision_tower_stage1, enc.vision_tower.vision_tower_stage2):
            tw.use_cuda_graph = False
        concurrent = enc.vision_tower.concurrent_towers
        enc.vision_tower.concurrent_towers = False      # one stream: per-launch event times must not overlap
        lib.hsenet_profile_start()
        for i in range(K):
            run_step(enc, args.workload, *dev_sets[i % n_sets])
        NC = _lib.PROFILE_CLASSES
        ms = (C.c_double * NC)(); fl = (C.c_double * NC)(); by = (C.c_double * NC)(); ln = (C.c_uint64 * NC)()
        _lib.check(lib.hsenet_profile_stop(ms, fl, by, ln), "profile_stop")
        # With the LayerNorm fold (default) the GEMM launches also carry the LayerNorm work in their epilogues, which
        # lowers the GEMM class's FLOP rate although the step is faster.  Second instrumented pass with every LayerNorm
        # as its own kernel, so that both views are on record (rank 0 of the default run only).
        unfolded = None
        towers = (enc.vision_tower.vision_tower_stage1, enc.vision_tower.vision_tower_stage2)
        if rank == 0 and not args.no_extras and all(t.fold_layernorm for t in towers):
            try:                                         # an extra record: it must never take the headline line down
                for tw in towers:
                    tw.fold_layernorm = False
                KU = min(K, 5)
                for i in range(2):
                    run_step(enc, args.workload, *dev_sets[i % n_sets])
                lib.hsenet_profile_start()
                for i in range(KU):
                    run_step(enc, args.workload, *dev_sets[i % n_sets])
                ms_u = (C.c_double * NC)(); fl_u = (C.c_double * NC)(); by_u = (C.c_double * NC)(); ln_u = (C.c_uint64 * NC)()
                _lib.check(lib.hsenet_profile_stop(ms_u, fl_u, by_u, ln_u), "profile_stop")
                g_tf = fl_u[0] / (ms_u[0] * 1e-3) / 1e12 if ms_u[0] > 0 else 0.0
                unfolded = {"gemm_tflops": g_tf, "steps": KU,
                            "share_ms": {"gemm": ms_u[0] / KU, "attention": ms_u[1] / KU, "layernorm": ms_u[2] / KU},
                            "note": "same steps with HSENET_LN_FOLD=0 (every LayerNorm its own kernel): GEMM launches "
                                    "without the folded LayerNorm epilogue work"}
            except Exception as exc:                     # noqa: BLE001
                unfolded = None
                sys.stderr.write(f"[bench] unfolded roofline pass skipped: {type(exc).__name__}: {exc}\n")
            finally:
                for tw in towers:
                    tw.fold_layernorm = True
        for tw in towers:
            tw.use_cuda_graph = True
        enc.vision_tower.concurrent_towers = concurrent

        # ---- e2e: public API, pinned host inputs, H2D + D2H inside the timed region ------------------------------------
        e2e = None
        if not args.no_e2e:
            # Double-buffered ingest loop: H2D of step i+1 and D2H of step i-1 run on side streams while step i
            # computes.  Every step still moves its own inputs host->device and its own result device->host inside
            # the timed region; they are overlapped, not skipped.
            cur = torch.cuda.current_stream(dev)
            s_h2d, s_d2h = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
            xd = [torch.empty_like(dev_sets[0][0]) for _ in range(2)]
            sd = [torch.empty_like(dev_sets[0][1]) for _ in range(2)]
            ev_in = [torch.cuda.Event() for _ in range(2)]
            ev_done = [torch.cuda.Event() for _ in range(2)]
            ev_out = [torch.cuda.Event() for _ in range(2)]
            host_out = [None, None]
            live_out = [None, None]
            used = [False, False]

            def e2e_step(i):
                slot = i & 1
                x, s = host_sets[i % n_sets]
                with torch.cuda.stream(s_h2d):
                    if used[slot]:
                        s_h2d.wait_event(ev_done[slot])          # step i-2 has finished reading this input slot
                    xd[slot].copy_(x, non_blocking=True)
                    sd[slot].copy_(s, non_blocking=True)
                    ev_in[slot].record(s_h2d)
                cur.wait_event(ev_in[slot])
                o = run_step(enc, args.workload, xd[slot], sd[slot])
                ev_done[slot].record(cur)
                outs = o if isinstance(o, tuple) else (o,)
                if host_out[slot] is None:
                    host_out[slot] = [torch.empty(t.shape, dtype=t.dtype, pin_memory=True) for t in outs]
                if used[slot]:
                    # The device tensors of step i-2 are released below.  No record_stream(): a block that another
                    # stream has touched cannot be recycled until cross-stream events are polled complete, and the
                    # caching allocator then falls back to a synchronising cudaMalloc every few steps (this made the
                    # end-to-end number swing between 280 and 750 volumes/s).  Instead the compute stream waits for the
                    # copy that read them, after which freeing them in stream order is safe.
                    cur.wait_event(ev_out[slot])
                live_out[slot] = outs
                with torch.cuda.stream(s_d2h):
                    s_d2h.wait_event(ev_done[slot])
                    for hbuf, t in zip(host_out[slot], outs):      # host buffers are reused in order on one stream
                        hbuf.copy_(t, non_blocking=True)
                    ev_out[slot].record(s_d2h)
                used[slot] = True
                return host_out[slot]

            for i in range(4):
                e2e_step(i)
            barrier()
            s_d2h.synchronize()
            e0.record()
            for i in range(K):
                res = e2e_step(i)
            cur.wait_stream(s_d2h)                               # the last results must be on the host before we stop
            e1.record()
            barrier()
            ms_e2e = max_over_ranks(e0.elapsed_time(e1))
            h2d = host_sets[0][0].numel() * 4 + host_sets[0][1].numel() * 4
            d2h = sum(t.numel() * t.element_size() for t in res)
            e2e = {"value": B * world * K / (ms_e2e * 1e-3), "unit": "volumes/s", "h2d_bytes_per_step": h2d,
                   "d2h_bytes_per_step": d2h, "ms_per_step": ms_e2e / K,
                   "how": "public module API, pinned host inputs/outputs, H2D and D2H of every step inside the timed "
                          "region on side streams (double buffered) overlapping the previous/next step's compute"}

        # ---- extra records (short): C2, the C4 CLIP step with its collective (N > 1), eager PyTorch on the same GPU (N = 1)
        extras = {}
        if not args.no_extras and args.workload == "c3":
            extras["c2"] = measure_c2(enc, dev, rank, world, barrier, max_over_ranks)
            if world > 1:
                extras["clip_step"] = measure_clip_step(enc, dev, rank, world, *dev_sets[0], barrier, max_over_ranks)
            elif rank == 0:
                try:
                    extras["gpu_eager_baseline"] = gpu_eager_baseline(enc, dev)
                except Exception as exc:                       # a baseline leg must never take the bench line down
                    extras["gpu_eager_baseline"] = {"unavailable": f"{type(exc).__name__}: {exc}"[:200]}
                try:
                    with torch.enable_grad():
                        extras["train_step"] = measure_train_step(dev)
                except Exception as exc:
                    extras["train_step"] = {"unavailable": f"{type(exc).__name__}: {exc}"[:200]}

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    peaks = load_peaks()
    value = B * world * K / (ms_total * 1e-3)
    gemm_tf = fl[0] / (ms[0] * 1e-3) / 1e12 if ms[0] > 0 else 0.0
    att_tf = fl[1] / (ms[1] * 1e-3) / 1e12 if ms[1] > 0 else 0.0
    ln_gbs = by[2] / (ms[2] * 1e-3) / 1e9 if ms[2] > 0 else 0.0
    step_tf = value / world * gflop / 1e3
    traffic = None
    tp = os.path.join(ROOT, "profiles", "roofline_traffic.json")
    if os.path.isfile(tp):
        try:
            traffic = json.load(open(tp)).get("gemm_bf16_kernel_dram_bytes_per_launch")
        except Exception:
            traffic = None

    def hbm_record(cls, kernel):
        gbs = by[cls] / (ms[cls] * 1e-3) / 1e9 if ms[cls] > 0 else 0.0
        return {"kernel": kernel, "bound": "hbm", "achieved": gbs, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                "frac": gbs / peaks["hbm_gbs"], "launches": int(ln[cls]),
                "avg_launch_us": ms[cls] / ln[cls] * 1e3 if ln[cls] else None,
                "algorithmic_mb_per_launch": by[cls] / ln[cls] / 1e6 if ln[cls] else None}

    # attention: d = 64 makes the exponentials, not the MMAs, the longer pole: 4*64 = 256 tensor FLOPs per score against
    # one ex2, i.e. 32 tensor-pipe cycles vs 64 MUFU cycles per 128x64 tile row at 8192 FLOP/clk and 16 ex2/clk per SM
    sm_max = (clocks or {}).get("sm_max_mhz") or 1965.0
    mufu_peak = 16.0 * 148 * sm_max * 1e6                       # ex2 per second per GPU at the maximum SM clock
    att_exps = fl[1] / (4.0 * 64)                               # one exponential per score
    att_exp_rate = att_exps / (ms[1] * 1e-3) if ms[1] > 0 else 0.0
    att_bounds = {
        "tensor": {"achieved": att_tf, "peak": peaks["bf16_sustained"], "unit": "TFLOP/s",
                   "frac": att_tf / peaks["bf16_sustained"]},
        "mufu": {"achieved": att_exp_rate / 1e12, "peak": mufu_peak / 1e12, "unit": "Tex2/s",
                 "frac": att_exp_rate / mufu_peak,
                 "note": "algorithmic exponentials (one per score) against 16 ex2/clk/SM x 148 SMs at the maximum SM clock; "
                         "the kernel evaluates 2 of every 8 on the FMA pipe instead"},
    }
    line = {
        "metric": "CT volumes/s encoded", "value": value, "unit": "volumes/s", "n_gpus": world, "steps": K,
        "warmup": W, "ms_per_step": ms_total / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "bf16", "data": "synthetic", "config": config, "clocks": clocks, "gpu_launches": int(launches),
        "host_enqueue_ms_per_step": host_enqueue_ms,
        "roofline": {
            "kernel": "gemm_bf16_2cta_kernel (tcgen05 cta_group::2; 69% of the path's FLOPs)", "bound": "tensor",
            "achieved": gemm_tf, "peak": peaks["bf16_sustained"], "unit": "TFLOP/s",
            "frac": gemm_tf / peaks["bf16_sustained"], "traffic": traffic,
            "frac_of_burst_peak": gemm_tf / peaks["bf16_burst"],
            "peak_source": f"MEASURED_PEAKS.json bf16_tflops_sustained ({peaks['source']}); kernel timed inside a long step",
            "launches": int(ln[0]), "avg_launch_ms": ms[0] / ln[0] if ln[0] else None,
            "how": "algorithmic 2*M*N*K per launch / CUDA-event duration per launch, instrumented re-run of the same steps",
        },
        "roofline_attention": {"kernel": {"s": "attention_split_kernel", "t": "attention_kernel<POLY,3>", "r": "attention_kernel<POLY,2>"}[
                                   (os.environ.get("HSENET_ATT_KERNEL") or "split")[0]] + " (tcgen05 flash attention; 31% of FLOPs)",
                               "bound": "mufu" if att_bounds["mufu"]["frac"] > att_bounds["tensor"]["frac"] else "tensor",
                               "achieved": att_tf, "peak": peaks["bf16_sustained"], "unit": "TFLOP/s",
                               "frac": att_tf / peaks["bf16_sustained"], "bounds": att_bounds, "launches": int(ln[1]),
                               "avg_launch_us": ms[1] / ln[1] * 1e3 if ln[1] else None},
        "roofline_layernorm": hbm_record(2, "layernorm_kernel"),
        "roofline_step": {"achieved": step_tf, "peak": peaks["bf16_sustained"], "unit": "TFLOP/s",
                          "frac": step_tf / peaks["bf16_sustained"], "frac_of_burst_peak": step_tf / peaks["bf16_burst"],
                          "gflop_per_volume": gflop,
                          "share_ms": {"gemm": ms[0] / K, "attention": ms[1] / K, "layernorm": ms[2] / K,
                                       "packer_pool": ms[4] / K, "packer_window_attn": ms[5] / K, "im2col": ms[6] / K,
                                       "slice_xattn": ms[7] / K, "score_scale": ms[8] / K}},
    }
    if ln[4] or ln[5]:
        pk_ms, pk_by = ms[4] + ms[5], by[4] + by[5]
        gbs = pk_by / (pk_ms * 1e-3) / 1e9 if pk_ms > 0 else 0.0
        line["roofline_packer"] = {
            "kernel": "packer_pool_kernel + packer_window_attn_kernel", "bound": "hbm", "achieved": gbs,
            "peak": peaks["hbm_gbs"], "unit": "GB/s", "frac": gbs / peaks["hbm_gbs"],
            "pool": hbm_record(4, "packer_pool_kernel"), "window_attention": hbm_record(5, "packer_window_attn_kernel"),
            "algorithmic_bytes": "pool: 2048x768 in + 128x768 out (bf16) per volume; window attention: 2048x1536 bf16 K|V "
                                 "+ 128x768 fp32 Q + 128x768 bf16 out per volume"}
    line["roofline_rowops"] = {"im2col": hbm_record(6, "im2col_kernel"), "slice_xattn": hbm_record(7, "slice_xattn_kernel"),
                               "score_scale": hbm_record(8, "score_scale_kernel")}
    if unfolded is not None:
        unfolded["gemm_frac"] = unfolded["gemm_tflops"] / peaks["bf16_sustained"]
        unfolded["gemm_frac_of_burst_peak"] = unfolded["gemm_tflops"] / peaks["bf16_burst"]
        line["roofline_unfolded"] = unfolded
        line["roofline"]["note"] = ("LayerNorm fold on: these launches also compute the LayerNorm statistics / scale+shift in "
                                    "their epilogues (23 of 25 layernorm launches per tower removed); see roofline_unfolded")
    line.update(extras)
    if e2e is not None:
        line["e2e"] = e2e
    if not args.no_cpu_baseline and world == 1:
        threads = os.cpu_count() or 1
        times = cpu_reference_run(args.workload, threads, 2)
        t = statistics.median(times)
        line["cpu_baseline"] = {"value": 1.0 / t, "unit": "volumes/s", "cores": threads, "kind": "port",
                                "sample": "1 volume per step, 2 timed steps after 1 warm-up (median), fp32 eager "
                                          "PyTorch oracle port of the reference's vit.py / packer"}
    if world > 1:
        dist.destroy_process_group()
    emit(line)


if __name__ == "__main__":
    main()
