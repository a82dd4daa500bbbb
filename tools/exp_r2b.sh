#!/bin/bash
mkdir -p gpurun_out/exp
HSENET_LIB_PATH=$PWD/hsenet_b200/libhsenet_sm100a_trace.so python tools/attn_timeline.py 8 > gpurun_out/exp/timeline7.txt 2>&1
