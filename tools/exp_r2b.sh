#!/bin/bash
mkdir -p gpurun_out/exp
python -m pytest tests/test_gpu_e2e.py tests/test_gpu_ops.py -q -x -k "layernorm_folded or linear or gemm or fold" 2>&1 | tail -3 > gpurun_out/exp/fold_test2.txt
python tools/kernel_bench.py --batch 8 > gpurun_out/exp/kernel_bench_b8.txt 2>&1
for f in 0 1; do
  ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/exp/fold${f}_launches.csv python tools/fold_probe.py $f 8 > gpurun_out/exp/fold${f}.log 2>&1
done
for f in 0 1 0 1; do
  HSENET_LN_FOLD=$f python bench.py --steps 10 --warmup 4 --no-extras > gpurun_out/exp/bench2_fold${f}_$RANDOM.json 2> gpurun_out/exp/bench_err.txt
done
