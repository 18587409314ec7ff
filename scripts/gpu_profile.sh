#!/bin/bash
# One GPU call that regenerates the committed profile evidence of the CURRENT build (run through gpurun):
#   gpurun_out/${NAME}_launches.csv  : ncu launch list (gpu__time_duration) of `python bench.py --steps 2 --warmup 1`, this
#                                      library's kernels only (the harness renders its synthetic frames with torch kernels)
#   gpurun_out/${NAME}_full.ncu-rep  : ncu --set full of the k_level_persistent launches of two steps of the same workload
# Afterwards, here:  python scripts/make_traffic_json.py gpurun_out/${NAME}_full.ncu-rep ${NAME}
NAME=${NAME:-r02}
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"^k_" -c 4000 --csv --log-file gpurun_out/${NAME}_launches.csv \
  python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/${NAME}_launches_run.log 2>&1
echo "launch list rc=$?"; tail -2 gpurun_out/${NAME}_launches_run.log | cut -c1-300
ncu --set full --clock-control none --import-source on -k regex:k_level_persistent -c ${COUNT:-2} -f -o gpurun_out/${NAME}_full \
  python scripts/profile_run.py 512 1 > gpurun_out/${NAME}_full_run.log 2>&1
echo "full rc=$?"; tail -2 gpurun_out/${NAME}_full_run.log
ls -la gpurun_out/${NAME}_*
