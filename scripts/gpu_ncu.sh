#!/bin/bash
# one GPU call: ncu --set full of the five level launches of one step (after warm-up), report in gpurun_out/
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:k_level_persistent -s ${SKIP:-10} -c 5 -f -o gpurun_out/${NAME:-prof} \
  python scripts/dev_bench.py 512 1 > gpurun_out/ncu_run.log 2>&1
tail -3 gpurun_out/ncu_run.log
ls -la gpurun_out/*.ncu-rep
