#!/bin/bash
for m in 0 15 14 2 4 8 1 6 12; do
  BEATRICE_B200_MRF_LATE=$m timeout 200 python bench.py --steps 300 --warmup 30 --no-cpu-baseline 2>gpurun_out/m.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('late mask $m depth2', d['ms_per_step'], 'depth1', d['latency_mode']['ms_per_step'])"
done
