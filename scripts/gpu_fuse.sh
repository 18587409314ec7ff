#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
for NF in 0 1; do
  if [ $NF = 1 ]; then export DVO_B200_NO_FUSE=1; else unset DVO_B200_NO_FUSE; fi
  echo "=== NO_FUSE=$NF"
  timeout 300 python scripts/dev_bench.py 512 4 0 2>&1 | grep -E "batch=|kernels"
done
unset DVO_B200_NO_FUSE
bash scripts/gpu_timing.sh
