#!/bin/bash
# session-3 GPU call H: pipeline depth 2 vs 3 at the same window capacity (17 x 144), with and without the pipe timers
mkdir -p gpurun_out; L=gpurun_out/s3h.log; : > $L
for v in timing2 timing3; do
  echo "=== $v" >> $L
  DVO_B200_LIB=$PWD/dvo_slam_b200/variants/$v.so DVO_B200_TIMING=1 timeout 300 python scripts/dev_bench.py 512 1 0 2>&1 | grep -E "timing\] level-slot|consumer warp|tiles [0-9]|^spc" | tail -13 >> $L
done
run() { echo "=== $1" >> $L; shift; env "$@" timeout 300 python scripts/dev_bench.py 512 3 0 2>&1 | tail -1 >> $L; }
run s2 DVO_B200_LIB=$PWD/dvo_slam_b200/variants/s2.so
run s3 DVO_B200_LIB=$PWD/dvo_slam_b200/variants/s3.so
run s2 DVO_B200_LIB=$PWD/dvo_slam_b200/variants/s2.so
run s3 DVO_B200_LIB=$PWD/dvo_slam_b200/variants/s3.so
cat $L
