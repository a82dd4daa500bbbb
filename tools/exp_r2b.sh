#!/bin/bash
mkdir -p gpurun_out/exp
python -m pytest tests/test_gpu_ops.py tests/test_gpu_backward.py tests/test_gpu_golden.py -q -x -k "attention or golden or repeat or permut or vit" 2>&1 | tail -3 > gpurun_out/exp/orphan_tests.txt
python tools/attn_sweep.py --batches 8,32 --env HSENET_ATT_KERNEL --modes split --reps 3 > gpurun_out/exp/attn_sweep_orphan.txt 2>&1
python tools/attn_sweep.py --batches 8 --seq 2048 --env HSENET_ATT_KERNEL --modes split --reps 3 >> gpurun_out/exp/attn_sweep_orphan.txt 2>&1
for i in 1 2; do python bench.py --steps 10 --warmup 4 --no-extras > gpurun_out/exp/bench_orphan_$i.json 2>/dev/null; done
