#!/bin/bash
mkdir -p gpurun_out
DVO_B200_LIB=$PWD/dvo_slam_b200/variants/timing.so DVO_B200_TIMING=1 timeout 600 python scripts/dev_bench.py ${B:-512} 2 ${SPCS:-0} > gpurun_out/dev_bench_timing.log 2>&1; grep -vE "^\[dvo_b200 timing\]   consumer" gpurun_out/dev_bench_timing.log | tail -${TAIL:-30}
