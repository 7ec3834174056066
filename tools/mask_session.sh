#!/bin/bash
for m in 0 2 4 8 6 12 10 14; do
  BEATRICE_B200_UPS_MASK_D2=$m timeout 200 python bench.py --steps 300 --warmup 30 --no-cpu-baseline 2>gpurun_out/m.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('mask $m depth2', d['ms_per_step'], 'depth1', d['latency_mode']['ms_per_step'])"
done
