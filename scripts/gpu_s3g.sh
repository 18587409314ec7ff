#!/bin/bash
# session-3 GPU call G: whole -m gpu suite, the default bench line, config 5, then the profile captures of this build
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu.log
timeout 900 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench rc=$?"; cut -c1-700 gpurun_out/bench_n1.json; tail -3 gpurun_out/bench_n1.err
timeout 600 python bench.py --config 5 --steps 5 --no-cpu-baseline > gpurun_out/bench_c5.json 2> gpurun_out/bench_c5.err; echo "c5 rc=$?"; cut -c1-400 gpurun_out/bench_c5.json
NAME=r02 COUNT=1 bash scripts/gpu_profile.sh
