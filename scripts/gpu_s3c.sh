#!/bin/bash
# session-3 GPU call C: tail slices (squads of g / 2g / 4g over fixed pair ranges): slice sizes, determinism, batch-512 parity
mkdir -p gpurun_out; L=gpurun_out/s3c.log; : > $L
for t in "" "0,0" "130,0" "60,30" "120,60"; do
  echo "=== TAIL='$t'" >> $L
  if [ -z "$t" ]; then unset DVO_B200_TAIL; else export DVO_B200_TAIL=$t; fi
  timeout 300 python scripts/dev_bench.py 512 3 0 2>&1 | tail -2 >> $L
done
unset DVO_B200_TAIL
for v in "$@"; do echo "=== variant $v" >> $L; DVO_B200_LIB=$PWD/dvo_slam_b200/variants/$v.so timeout 300 python scripts/dev_bench.py 512 3 0 2>&1 | tail -2 >> $L; done
echo "=== timing" >> $L
DVO_B200_LIB=$PWD/dvo_slam_b200/variants/timing.so DVO_B200_TIMING=1 timeout 300 python scripts/dev_bench.py 512 1 0 2>&1 | grep -E "timing\] level-slot|lifetime" | tail -8 >> $L
timeout 900 python -m pytest tests/test_gpu_batch_parity.py -x -q > gpurun_out/s3c_pytest.log 2>&1; echo "pytest batch parity rc=$?" >> $L; tail -3 gpurun_out/s3c_pytest.log >> $L
cat $L
