#!/bin/bash
# session-3 GPU call I: canonical sums: whole -m gpu suite, then dev bench (default vs the build before: variants/base2.so)
mkdir -p gpurun_out; L=gpurun_out/s3i.log; : > $L
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/s3i_pytest.log 2>&1; echo "pytest rc=$?" >> $L; tail -25 gpurun_out/s3i_pytest.log >> $L
run() { echo "=== $1" >> $L; shift; env "$@" timeout 300 python scripts/dev_bench.py 512 3 0 2>&1 | tail -3 | grep -v iterations >> $L; }
run default A=1
run before DVO_B200_LIB=$PWD/dvo_slam_b200/variants/base2.so
run default A=1
cat $L
