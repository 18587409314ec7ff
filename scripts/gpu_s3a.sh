#!/bin/bash
# session-3 GPU call A: timing breakdown of the product build, squad-size knobs, pipeline-depth variants
mkdir -p gpurun_out; L=gpurun_out/s3a.log; : > $L
echo "=== timing" >> $L
DVO_B200_LIB=$PWD/dvo_slam_b200/variants/timing.so DVO_B200_TIMING=1 timeout 300 python scripts/dev_bench.py 512 2 0 2>&1 | grep -vE "^\[dvo_b200 timing\]   consumer" >> $L
for g in 1 2 4; do echo "=== FINE_G=$g" >> $L; DVO_B200_FINE_G=$g timeout 300 python scripts/dev_bench.py 512 3 0 2>&1 | tail -1 >> $L; done
for v in s3r16 s3r16c144 s2r18; do echo "=== variant $v" >> $L; DVO_B200_LIB=$PWD/dvo_slam_b200/variants/$v.so timeout 300 python scripts/dev_bench.py 512 3 0 2>&1 | tail -1 >> $L; done
cat $L
