#!/usr/bin/env python
"""BASELINE.json configs[4] (C5): throughput sweep of the FULL encoding path (dual tower + two packers -> [B,256,3072], bf16)
over the batch size on one GPU, CUDA-event timed with inputs resident in HBM.  Writes one JSON document.

    python tools/sweep.py [--batches 1,2,4,8,14,16,32,64,128,256,512] [--out profiles/r2/sweep.json] [--cpu-baseline]

Batch 14 is the reference's evaluation batch (Bench/eval/eval_HSENet_CT_Rate_MRG.py:388).  The multi-GPU points of the
sweep (32 volumes per GPU on 2 / 4 / 8 GPUs) are ordinary `bench.py --gpus N` lines.
"""
import argparse
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

GFLOP = 1033.54


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batches", default="1,2,4,8,14,16,32,64,128,256,512")
    ap.add_argument("--out", default="")
    ap.add_argument("--cpu-baseline", action="store_true")
    ap.add_argument("--min-ms", type=float, default=400.0, help="timed region per point")
    a = ap.parse_args()
    import hsenet_b200 as H
    dev = torch.device("cuda:0")
    H.set_precision("bf16")
    torch.manual_seed(0)
    enc = H.HSENetVisualEncoder(H.VisionConfig(), use_parallel_projector=True).eval().requires_grad_(False).to(dev)
    peaks = {"bf16_sustained": 1402.0, "bf16_burst": 1673.1}
    pp = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(pp):
        d = json.load(open(pp))
        peaks = {"bf16_sustained": d.get("bf16_tflops_sustained", 1402.0), "bf16_burst": d.get("bf16_tflops", 1673.1)}
    points = []
    g = torch.Generator().manual_seed(1234)
    for B in [int(b) for b in a.batches.split(",")]:
        nset = max(2, min(8, 300 // max(B * 8, 1) + 1))          # rotating inputs > 126 MB L2
        sets = [(torch.rand(B, 1, 32, 256, 256, generator=g).to(dev), torch.randn(B, 32, 768, generator=g).to(dev))
                for _ in range(nset)]
        with torch.no_grad():
            for i in range(3):
                enc(*sets[i % nset])
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            enc(*sets[0])
            e1.record()
            torch.cuda.synchronize()
            steps = max(3, int(a.min_ms / max(e0.elapsed_time(e1), 1e-3)))
            e0.record()
            t0 = time.perf_counter()
            for i in range(steps):
                y = enc(*sets[i % nset])
            host_ms = (time.perf_counter() - t0) * 1e3 / steps
            e1.record()
            torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / steps
        v = B / (ms * 1e-3)
        tf = v * GFLOP / 1e3
        points.append({"batch": B, "volumes_per_s": v, "ms_per_step": ms, "steps": steps, "step_tflops": tf,
                       "frac_of_sustained_peak": tf / peaks["bf16_sustained"], "frac_of_burst_peak": tf / peaks["bf16_burst"],
                       "host_enqueue_ms_per_step": host_ms, "out_shape": list(y.shape)})
        print(f"B={B:4d}  {v:8.1f} volumes/s  {ms:9.3f} ms/step  {tf:7.1f} TFLOP/s  host enqueue {host_ms:6.2f} ms",
              flush=True)
        del sets, y
        H.release_workspaces()
        torch.cuda.empty_cache()
    doc = {"workload": "C5: full encoding path (ViT_stage1 + ViT_stage2 + two VisualPacker_3d_phi_v3) forward, bf16, 1 GPU",
           "gflop_per_volume": GFLOP, "peaks": peaks, "gpu": torch.cuda.get_device_name(0), "points": points}
    if a.cpu_baseline:
        import bench
        import statistics
        times = bench.cpu_reference_run("c3", os.cpu_count() or 1, 2)
        doc["cpu_baseline"] = {"value": 1.0 / statistics.median(times), "unit": "volumes/s", "cores": os.cpu_count(),
                               "kind": "port", "sample": "1 volume per step, 2 timed steps after 1 warm-up"}
    s = json.dumps(doc, indent=1)
    if a.out:
        os.makedirs(os.path.dirname(os.path.abspath(a.out)), exist_ok=True)
        open(a.out, "w").write(s + "\n")
    else:
        print(s)


if __name__ == "__main__":
    main()
