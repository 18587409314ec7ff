#!/bin/bash
# one GPU call: the developer benchmark for several library variants (dvo_slam_b200/variants/*.so)
mkdir -p gpurun_out; : > gpurun_out/variants.log
for v in "$@"; do
  echo "=== variant $v" >> gpurun_out/variants.log
  DVO_B200_LIB=$PWD/dvo_slam_b200/variants/$v.so timeout 300 python scripts/dev_bench.py 512 3 ${SPCS:-0} >> gpurun_out/variants.log 2>&1
done
grep -E "^===|^spc|ms per launch: [0-9]" gpurun_out/variants.log
