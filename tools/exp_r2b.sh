#!/bin/bash
mkdir -p gpurun_out/exp
L=hsenet_b200/libhsenet_sm100a
python -m pytest tests/test_gpu_ops.py tests/test_gpu_golden.py -q -x -k "attention or repeat or permut" 2>&1 | tail -2 > gpurun_out/exp/att_test.txt
for v in "" _nolean; do
  echo "== variant '$v'" >> gpurun_out/exp/attn_sweep4.txt
  HSENET_LIB_PATH=$PWD/${L}$v.so python tools/attn_sweep.py --batches 8,32 --env HSENET_ATT_KERNEL --modes split --reps 3 >> gpurun_out/exp/attn_sweep4.txt 2>&1
done
HSENET_LIB_PATH=$PWD/${L}_trace.so python tools/attn_timeline.py 8 > gpurun_out/exp/timeline4.txt 2>&1
