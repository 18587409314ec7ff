#!/bin/bash
# compute-sanitizer over the smoke alignment (320x240, 3 levels) and over one 80-pair 640x480 batch (the fused two-segment
# launch, raw-input pyramids); only this library's kernels are instrumented
mkdir -p gpurun_out
for tool in memcheck initcheck; do
  timeout 900 compute-sanitizer --tool $tool --error-exitcode 9 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/sanitize_$tool.log 2>&1
  echo "$tool smoke exit $?"; grep -E "ERROR SUMMARY" gpurun_out/sanitize_$tool.log | head -3
  timeout 1500 compute-sanitizer --tool $tool --kernel-regex kns=k_ --error-exitcode 9 python scripts/sanitize_batch.py > gpurun_out/sanitize_batch_$tool.log 2>&1
  echo "$tool batch exit $?"; grep -E "ERROR SUMMARY|Invalid|Uninit|done" gpurun_out/sanitize_batch_$tool.log | head -6
done
timeout 1500 compute-sanitizer --tool racecheck --kernel-regex kns=k_level --error-exitcode 9 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/sanitize_racecheck.log 2>&1
echo "racecheck smoke exit $?"; grep -E "ERROR SUMMARY|RACECHECK SUMMARY|hazard" gpurun_out/sanitize_racecheck.log | sort | uniq -c | head -8
