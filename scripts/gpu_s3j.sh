#!/bin/bash
# session-3 GPU call J: cyclic vs contiguous strip assignment (identical results by construction): parity, speed, squad waits
mkdir -p gpurun_out; L=gpurun_out/s3j.log; : > $L
timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_batch_parity.py -q -x > gpurun_out/s3j_pytest.log 2>&1; echo "pytest rc=$?" >> $L; tail -6 gpurun_out/s3j_pytest.log >> $L
run() { echo "=== $1" >> $L; shift; env "$@" timeout 300 python scripts/dev_bench.py 512 3 0 2>&1 | tail -3 | grep -v iterations >> $L; }
run cyclic A=1
run contiguous DVO_B200_CONTIGUOUS=1
run cyclic A=1
run contiguous DVO_B200_CONTIGUOUS=1
run cyclic_tail120,60 DVO_B200_TAIL=120,60
run cyclic_g4 DVO_B200_FINE_G=4 DVO_B200_TAIL=60,30
for m in 0 1; do
  echo "=== timing contiguous=$m" >> $L
  if [ $m = 1 ]; then export DVO_B200_CONTIGUOUS=1; else unset DVO_B200_CONTIGUOUS; fi
  DVO_B200_LIB=$PWD/dvo_slam_b200/variants/timing.so DVO_B200_TIMING=1 timeout 300 python scripts/dev_bench.py 512 1 0 2>&1 | grep -E "timing\] level-slot" | tail -4 >> $L
done
cat $L
