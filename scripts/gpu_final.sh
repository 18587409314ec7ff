#!/bin/bash
# One GPU call that refreshes every measured artefact of the CURRENT sources: ncu --set full of one 512-pair step -> stamped
# traffic json (generated here so that the bench line below quotes it), ncu launch list of a short bench run, the default bench
# line, and the -m gpu suite.  Everything lands in gpurun_out/; copy the summaries to profiles/ afterwards.
NAME=${NAME:-r02}
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k regex:k_level_persistent -c 1 -f -o gpurun_out/${NAME}_full \
  python scripts/profile_run.py 512 1 > gpurun_out/${NAME}_full_run.log 2>&1; echo "full rc=$?"
python scripts/make_traffic_json.py gpurun_out/${NAME}_full.ncu-rep ${NAME} | cut -c1-160
cp profiles/${NAME}_traffic.json profiles/${NAME}_level_kernels.txt gpurun_out/
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"^k_" -c 4000 --csv --log-file gpurun_out/${NAME}_launches.csv \
  python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/${NAME}_launches_run.log 2>&1; echo "launch list rc=$?"
timeout 900 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench rc=$?"; cut -c1-260 gpurun_out/bench_n1.json; tail -2 gpurun_out/bench_n1.err
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
