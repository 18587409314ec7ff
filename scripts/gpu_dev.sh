#!/bin/bash
# GPU call for kernel development: fail-fast parity tests, then the developer benchmark with plan overrides
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
DVO_B200_TIMING=1 timeout 600 python scripts/dev_bench.py 512 3 ${1:-0} > gpurun_out/dev_bench.log 2>&1
cat gpurun_out/dev_bench.log
