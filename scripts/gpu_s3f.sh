#!/bin/bash
# session-3 GPU call F: the whole -m gpu suite on the default build, then window-capacity variants
mkdir -p gpurun_out; L=gpurun_out/s3f.log; : > $L
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/s3f_pytest.log 2>&1; echo "pytest rc=$?" >> $L; tail -5 gpurun_out/s3f_pytest.log >> $L
run() { echo "=== $1" >> $L; shift; env "$@" timeout 300 python scripts/dev_bench.py 512 3 0 2>&1 | tail -3 | grep -v iterations >> $L; }
run default A=1
for v in "$@"; do run $v DVO_B200_LIB=$PWD/dvo_slam_b200/variants/$v.so; done
run default_again A=1
cat $L
