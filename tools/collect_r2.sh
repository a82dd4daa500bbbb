#!/bin/bash
# Collects the round-2 evidence set on one B200 (run through gpurun); outputs under gpurun_out/r2/
set -u
O=gpurun_out/r2
mkdir -p $O
python bench.py --steps 20 --warmup 5 > $O/bench_c3.json 2> $O/bench_c3.err
python bench.py --workload c2 --steps 20 --warmup 5 --no-extras > $O/bench_c2.json 2>/dev/null
python tools/kernel_bench.py --batch 8 > $O/kernel_bench.txt 2>&1
python tools/kernel_bench.py --batch 32 >> $O/kernel_bench.txt 2>&1
python tools/attn_sweep.py --batches 8,32 --env HSENET_ATT_KERNEL --modes split,tri,rowwarp --reps 3 > $O/attn_sweep.txt 2>&1
python tools/rowops_bench.py --batch 32 > $O/rowops_bench.txt 2>&1
python tools/ingest_bench.py > $O/ingest_bench.txt 2>&1
# launch list of the default bench command (cold-cache, serialised: shares, not absolutes)
ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file $O/launches_bench_c3.csv \
    python bench.py --steps 1 --warmup 1 --no-extras --no-e2e --no-cpu-baseline > $O/launches_bench.log 2>&1
# full captures
HSENET_ATT_KERNEL=split ncu --set full --import-source on --clock-control none -k regex:attention_split --launch-skip 3 -c 1 -f \
    -o $O/ncu_attention python tools/attn_sweep.py --batches 8 --modes split --env HSENET_ATT_KERNEL --reps 1 --iters 2 > $O/ncu_attention.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:gemm_bf16_2cta -c 16 -f \
    -o $O/ncu_gemm python tools/kernel_bench.py --batch 8 --iters 1 > $O/ncu_gemm.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:"packer_pool|packer_window|slice_extract|slice_xattn|layernorm" -c 10 -f \
    -o $O/ncu_rowops python tools/rowops_bench.py --batch 32 --once > $O/ncu_rowops.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:"hu_transpose_rows|resample_kernel|minmax_kernel|bbox_kernel|crop_normalize" --launch-skip 10 -c 6 -f \
    -o $O/ncu_ingest python tools/ingest_bench.py > $O/ncu_ingest.log 2>&1
ls -la $O
