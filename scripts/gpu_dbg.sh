#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; tail -5 gpurun_out/pytest_gpu.log
DVO_B200_TIMING=1 timeout 600 python scripts/dev_bench.py 512 3 0 > gpurun_out/dev_bench.log 2>&1; grep -E "^spc|ms per launch: [0-9]|iterations per" gpurun_out/dev_bench.log | tail -3
DVO_B200_NO_WALK=1 DVO_B200_TIMING=1 timeout 600 python scripts/dev_bench.py 512 3 0 > gpurun_out/dev_bench_nowalk.log 2>&1; grep -E "^spc|ms per launch: [0-9]" gpurun_out/dev_bench_nowalk.log | tail -2
