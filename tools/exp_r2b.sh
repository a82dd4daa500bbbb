#!/bin/bash
mkdir -p gpurun_out/exp
L=hsenet_b200/libhsenet_sm100a
python -m pytest tests/test_gpu_ops.py tests/test_gpu_e2e.py tests/test_gpu_golden.py -q -x 2>&1 | tail -3 > gpurun_out/exp/pf_tests.txt
rm -f gpurun_out/exp/pf_kernel_bench.txt
for v in "" _nopf "" _nopf; do
  echo "== variant '$v'" >> gpurun_out/exp/pf_kernel_bench.txt
  HSENET_LIB_PATH=$PWD/${L}$v.so python tools/kernel_bench.py --batch 8 2>&1 | grep "2cta" >> gpurun_out/exp/pf_kernel_bench.txt
done
for v in "" _nopf "" _nopf; do
  HSENET_LIB_PATH=$PWD/${L}$v.so python bench.py --steps 10 --warmup 4 --no-extras > gpurun_out/exp/bench_pf${v}_$RANDOM.json 2>/dev/null
done
