#!/bin/bash
# developer sweep: library build variants x squad size target
for v in "$@"; do
  for r in 64 128; do
    echo "LIB=$v RPW=$r"; DVO_B200_LIB=$PWD/dvo_slam_b200/variants/$v DVO_B200_RPW=$r timeout 300 python bench.py --steps 3 --warmup 2 --no-cpu-baseline 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['value']), round(d['ms_per_step'],2), round(d['roofline']['frac'],4), round(d['e2e']['value']))"
  done
done
