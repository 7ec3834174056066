#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/h.log
: > $L
timeout 120 python tools/mrf_probe.py 2 40 4 >> $L 2>&1
BEATRICE_B200_MRF_TRACE=1 timeout 120 python tools/op_profile.py 2 256 2 > gpurun_out/h_tmp.log 2>&1
grep "mrf trace" gpurun_out/h_tmp.log | tail -21 | cut -c1-200 >> $L
grep "mrfc trace" gpurun_out/h_tmp.log | tail -7 | cut -c1-250 >> $L
timeout 120 python tools/op_profile.py 2 256 >> $L 2>&1
timeout 300 python bench.py --steps 200 --warmup 20 --no-cpu-baseline 2>&1 | cut -c1-200 >> $L
cat $L
