#!/bin/bash
# developer sweep: squad size of the fine level group x coarse-group threshold
mkdir -p gpurun_out
: > gpurun_out/fineg.log
for CT in ${CTS:-40 110}; do
for G in ${GS:-1 2 3}; do
  export DVO_B200_COARSE_TILES=$CT
  if [ $G = 0 ]; then unset DVO_B200_FINE_G; else export DVO_B200_FINE_G=$G; fi
  echo "=== coarse tiles <= $CT, fine g=$G" >> gpurun_out/fineg.log
  timeout 300 python scripts/dev_bench.py ${B:-512} 3 0 2>&1 | grep -E "batch=|kernels" >> gpurun_out/fineg.log
done; done
cat gpurun_out/fineg.log
