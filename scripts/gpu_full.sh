#!/bin/bash
# one GPU call: the whole -m gpu suite (not fail-fast), then the timing breakdown at batch 512 and for a single pair
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -15 gpurun_out/pytest_gpu.log
for B in 512 1; do
  DVO_B200_LIB=$PWD/dvo_slam_b200/variants/timing.so DVO_B200_TIMING=1 timeout 600 python scripts/dev_bench.py $B $([ $B = 1 ] && echo 20 || echo 2) 0 > gpurun_out/timing_b$B.log 2>&1
  grep -vE "^\[dvo_b200 timing\]   consumer" gpurun_out/timing_b$B.log | tail -22
done
