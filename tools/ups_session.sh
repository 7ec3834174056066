#!/bin/bash
# Developer session: fused upsampler on/off -- parity probe, per-op times, bench lines.
O=gpurun_out
mkdir -p $O
export BEATRICE_B200_MRF_PLAN="11|7|3;11|7|3;11|7|3"
run() {
  tag=$1; fuse=$2
  echo "=== $tag fuse=$fuse"
  BEATRICE_B200_FUSE_UPS=$fuse timeout 120 python tools/mrf_probe.py 2 9 4 2>&1 | grep -E "probe|rror|latched" | tail -4
  BEATRICE_B200_FUSE_UPS=$fuse timeout 120 python tools/op_profile.py 2 256 6 2>&1 | grep -E "wave|serial"
  BEATRICE_B200_FUSE_UPS=$fuse timeout 200 python bench.py --steps 300 --warmup 30 --no-cpu-baseline 2>$O/ups_$tag.err | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print('bench', d['ms_per_step'], d['value'], d.get('latency_mode',{}).get('ms_per_step'), d.get('parity'))"
}
run off 0
run on 1
