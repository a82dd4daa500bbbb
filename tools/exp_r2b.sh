#!/bin/bash
mkdir -p gpurun_out/exp
timeout 900 compute-sanitizer --tool memcheck --print-limit 20 python tools/sanitize_small.py > gpurun_out/exp/sanitizer_memcheck.log 2>&1
timeout 900 compute-sanitizer --tool synccheck --print-limit 20 python tools/sanitize_small.py > gpurun_out/exp/sanitizer_synccheck.log 2>&1
tail -2 gpurun_out/exp/sanitizer_memcheck.log; tail -2 gpurun_out/exp/sanitizer_synccheck.log
