#!/bin/bash
mkdir -p gpurun_out/exp
L=hsenet_b200/libhsenet_sm100a
rm -f gpurun_out/exp/attn_sweep_tail2.txt
for v in "" _notail; do
  for seq in 2048 2049; do
    echo "== variant '$v' seq $seq" >> gpurun_out/exp/attn_sweep_tail2.txt
    HSENET_LIB_PATH=$PWD/${L}$v.so python tools/attn_sweep.py --batches 8,32 --seq $seq --env HSENET_ATT_KERNEL --modes split --reps 3 >> gpurun_out/exp/attn_sweep_tail2.txt 2>&1
  done
done
