#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/k.log
: > $L
run() { echo "== $*" >> $L; ( timeout 300 env "$@" ) 2>&1 | cut -c1-400 >> $L; echo "rc=$?" >> $L; }
run python tools/mrf_probe.py 2 40 6
run python tools/mrf_probe.py 1 40 6
run python bench.py --steps 300 --warmup 20 --no-cpu-baseline
( BEATRICE_B200_MRF_TRACE=1 timeout 300 python tools/op_profile.py 2 256 2 ) > gpurun_out/trace.log 2>&1
grep "mrfc trace" gpurun_out/trace.log | tail -7 >> $L
cat $L
