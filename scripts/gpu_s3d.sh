#!/bin/bash
# session-3 GPU call D: reference tile records: parity (whole gpu parity file), dev bench, timing
mkdir -p gpurun_out; L=gpurun_out/s3d.log; : > $L
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q > gpurun_out/s3d_pytest.log 2>&1; echo "pytest rc=$?" >> $L; tail -5 gpurun_out/s3d_pytest.log >> $L
echo "=== default" >> $L; timeout 300 python scripts/dev_bench.py 512 3 0 2>&1 | tail -3 >> $L
for v in "$@"; do echo "=== variant $v" >> $L; DVO_B200_LIB=$PWD/dvo_slam_b200/variants/$v.so timeout 300 python scripts/dev_bench.py 512 3 0 2>&1 | tail -3 >> $L; done
echo "=== timing" >> $L
DVO_B200_LIB=$PWD/dvo_slam_b200/variants/timing.so DVO_B200_TIMING=1 timeout 300 python scripts/dev_bench.py 512 1 0 2>&1 | grep -E "timing\] level-slot|consumer warp" | tail -8 >> $L
cat $L
