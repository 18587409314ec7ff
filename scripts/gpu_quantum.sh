#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
for Q in ${QS:-0 1 2 3 4}; do
  export DVO_B200_QUANTUM=$Q
  echo "=== quantum $Q"
  timeout 300 python scripts/dev_bench.py 512 4 0 2>&1 | grep -E "batch=|kernels"
done
unset DVO_B200_QUANTUM
bash scripts/gpu_timing.sh | grep -E "lifetime|level-slot|batch="
