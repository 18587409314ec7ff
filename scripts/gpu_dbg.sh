#!/bin/bash
mkdir -p gpurun_out
timeout 600 python scripts/dbg_determinism.py > gpurun_out/dbg.log 2>&1; cat gpurun_out/dbg.log | tail -12
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; tail -5 gpurun_out/pytest_gpu.log
DVO_B200_TIMING=1 timeout 600 python scripts/dev_bench.py 512 3 0 > gpurun_out/dev_bench.log 2>&1; grep -E "^spc|ms per launch|iterations per" gpurun_out/dev_bench.log | tail -4
DVO_B200_LIB=$PWD/dvo_slam_b200/variants/timing.so DVO_B200_TIMING=1 timeout 600 python scripts/dev_bench.py 512 2 0 > gpurun_out/dev_bench_timing.log 2>&1; tail -14 gpurun_out/dev_bench_timing.log
