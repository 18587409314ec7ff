#!/bin/bash
# session-3 GPU call E: prologue hoist (default) vs base, try_wait hints, tail-slice / squad-size knobs
mkdir -p gpurun_out; L=gpurun_out/s3e.log; : > $L
timeout 600 python -m pytest tests/test_gpu_parity.py -x -q -k "records or linearisation or batch_equals or pose_within or golden or odd or degenerate or config5" > gpurun_out/s3e_pytest.log 2>&1; echo "pytest rc=$?" >> $L; tail -3 gpurun_out/s3e_pytest.log >> $L
run() { echo "=== $1" >> $L; shift; env "$@" timeout 300 python scripts/dev_bench.py 512 3 0 2>&1 | tail -3 | grep -v iterations >> $L; }
run default A=1
run base DVO_B200_LIB=$PWD/dvo_slam_b200/variants/base.so
run hint20 DVO_B200_LIB=$PWD/dvo_slam_b200/variants/hint20.so
run hint2 DVO_B200_LIB=$PWD/dvo_slam_b200/variants/hint2.so
run tail60,30 DVO_B200_TAIL=60,30
run tail40,20 DVO_B200_TAIL=40,20
run tail75,30 DVO_B200_TAIL=75,30
run g2_tail140,72 DVO_B200_FINE_G=2 DVO_B200_TAIL=140,72
run g2_tail100,50 DVO_B200_FINE_G=2 DVO_B200_TAIL=100,50
run g4_tail60,0 DVO_B200_FINE_G=4 DVO_B200_TAIL=60,0
cat $L
