#!/bin/bash
# one GPU call: parity tests (fail fast), then a short bench; everything lands in gpurun_out/
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -30 gpurun_out/pytest_gpu.log
if [ "$1" = "bench" ]; then
  timeout 600 python bench.py > gpurun_out/bench_dev.json 2> gpurun_out/bench_dev.err
  echo "bench exit $?"; cat gpurun_out/bench_dev.json; tail -5 gpurun_out/bench_dev.err
fi
if [ "$2" = "c5" ]; then
  timeout 600 python bench.py --config 5 --steps 3 --no-cpu-baseline > gpurun_out/bench_c5.json 2> gpurun_out/bench_c5.err
  echo "bench c5 exit $?"; cat gpurun_out/bench_c5.json; tail -5 gpurun_out/bench_c5.err
fi
if [ "$3" = "ref" ]; then
  timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
  echo "bench ref exit $?"; cat gpurun_out/bench_ref.json | cut -c1-600; tail -3 gpurun_out/bench_ref.err
fi
