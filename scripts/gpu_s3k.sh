#!/bin/bash
# final bench lines of this build: default N=1 (with the stamped traffic), the reference arm on the same box
mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench rc=$?"; cut -c1-300 gpurun_out/bench_n1.json; tail -2 gpurun_out/bench_n1.err
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref rc=$?"; cut -c1-500 gpurun_out/bench_ref.json; tail -2 gpurun_out/bench_ref.err
