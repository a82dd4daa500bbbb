#!/bin/bash
mkdir -p gpurun_out/exp
L=hsenet_b200/libhsenet_sm100a
for v in _pace4 _pace8 _pace12; do
  echo "== variant '$v'" >> gpurun_out/exp/attn_sweep6.txt
  HSENET_LIB_PATH=$PWD/${L}$v.so python tools/attn_sweep.py --batches 8,32 --env HSENET_ATT_KERNEL --modes split --reps 3 >> gpurun_out/exp/attn_sweep6.txt 2>&1
done
HSENET_LIB_PATH=$PWD/${L}_trace.so python tools/attn_timeline.py 8 > gpurun_out/exp/timeline6.txt 2>&1
