#!/bin/bash
# developer sweep: squad size target (rounds per warp) vs throughput
for r in "$@"; do
  echo "RPW=$r"; DVO_B200_RPW=$r timeout 300 python bench.py --steps 3 --warmup 2 --no-cpu-baseline 2>&1 | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(round(d['value']), round(d['ms_per_step'],2), round(d['roofline']['frac'],4), round(d['e2e']['value']))"
done
