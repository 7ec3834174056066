#!/bin/bash
# Developer session: per-op times + short bench lines of several MRF launch plans
# (BEATRICE_B200_MRF_PLAN="<C=64>;<C=32>;<C=16>").   gpurun --timeout 900 -- 'bash tools/plan_session.sh'
O=gpurun_out
mkdir -p $O
run() {
  tag=$1; plan=$2
  echo "=== $tag plan=$plan"
  BEATRICE_B200_MRF_PLAN="$plan" timeout 120 python tools/op_profile.py 2 256 6 2>&1 | grep -E "mrf|serial"
  BEATRICE_B200_MRF_PLAN="$plan" timeout 200 python bench.py --steps 300 --warmup 30 --no-cpu-baseline 2>$O/plan_$tag.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('bench', d['ms_per_step'], d['value'], d.get('latency_mode',{}).get('ms_per_step'))"
}
run base ""
run a ";11@128/7,3;"
run b ";11@114/7,3;"
run c ";11@114/7|3;"
run d ";7|11|3;"
run e ";11|3|7;"
