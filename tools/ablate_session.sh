#!/bin/bash
# Developer session: marginal in-graph cost of every vocoder op (hop time with the op's launches dropped).
O=gpurun_out
mkdir -p $O
one() {
  BEATRICE_B200_SKIP_OPS="$1" timeout 200 python bench.py --steps 300 --warmup 30 --no-cpu-baseline 2>$O/abl.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('skip=%-28s depth2 %.4f  depth1 %.4f' % ('$1', d['ms_per_step'], d.get('latency_mode',{}).get('ms_per_step')))"
}
one ""
for op in wave.cond wave.ups0 wave.mrf0 wave.ups1 wave.mrf1 wave.ups2 wave.mrf2 wave.ups3 wave.mrf3 wave.post "wave.ups" "wave.mrf" "phone.,pitch." "wave."; do one "$op"; done
