#!/usr/bin/env python
"""Summarise an .ncu-rep (read on the CPU box): headline metrics per launch + top stall-sample SASS lines.
usage: python tools/ncu_summary.py gpurun_out/prof.ncu-rep [ntop]"""
import csv
import io
import subprocess
import sys

KEYS = ["gpu__time_duration.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "launch__registers_per_thread", "launch__grid_size", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed"]


def main():
    rep = sys.argv[1]
    ntop = int(sys.argv[2]) if len(sys.argv) > 2 else 25
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    print("== launches:", len(data))
    ik = hdr.index("Kernel Name")
    for r in data:
        print("--", r[ik][:90])
        for k in KEYS:
            if k in hdr:
                i = hdr.index(k)
                print(f"   {k:80s} {r[i]:>14s} {units[i]}")
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(src)))
    blocks, cur = [], None
    for r in rows:
        if r and r[0] == "Kernel Name":
            cur = []
            blocks.append((r[1], cur))
            continue
        if cur is not None:
            cur.append(r)
    for name, b in blocks[:1] if len(sys.argv) < 4 else blocks:
        h, d = b[0], b[1:]
        ia, isamp, iex = h.index("Source"), h.index("Warp Stall Sampling (All Samples)"), h.index("Instructions Executed")
        tot = sum(int(r[isamp]) for r in d if r[isamp].isdigit())
        print(f"== {name[:80]}: {tot} samples, {len(d)} SASS instructions")
        top = sorted(range(len(d)), key=lambda i: -int(d[i][isamp]) if d[i][isamp].isdigit() else 0)[:ntop]
        for i in top:
            r = d[i]
            print(f"   {i:5d} {int(r[isamp]):6d} {100 * int(r[isamp]) / max(tot, 1):5.1f}%  ex={r[iex]:>9s}  {r[ia].strip()[:100]}")


if __name__ == "__main__":
    main()
