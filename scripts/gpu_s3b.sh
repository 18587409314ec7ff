#!/bin/bash
# session-3 GPU call B: quick parity subset on the default build, then dev bench: default / variants, timing breakdown
mkdir -p gpurun_out; L=gpurun_out/s3b.log; : > $L
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "records or linearisation or batch_equals or pose_within or golden or odd or degenerate" > gpurun_out/s3b_pytest.log 2>&1; echo "pytest rc=$?" >> $L; tail -3 gpurun_out/s3b_pytest.log >> $L
echo "=== default" >> $L; timeout 300 python scripts/dev_bench.py 512 3 0 2>&1 | tail -1 >> $L
for v in "$@"; do echo "=== variant $v" >> $L; DVO_B200_LIB=$PWD/dvo_slam_b200/variants/$v.so timeout 300 python scripts/dev_bench.py 512 3 0 2>&1 | tail -1 >> $L; done
echo "=== timing" >> $L
DVO_B200_LIB=$PWD/dvo_slam_b200/variants/timing.so DVO_B200_TIMING=1 timeout 300 python scripts/dev_bench.py 512 1 0 2>&1 | grep -E "timing\]" | head -8 >> $L
cat $L
