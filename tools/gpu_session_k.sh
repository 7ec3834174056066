#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/k.log
: > $L
run() { echo "== $*" >> $L; ( timeout 300 env "$@" ) 2>&1 | cut -c1-400 >> $L; echo "rc=$?" >> $L; }
run python tools/mrf_probe.py 2 40 6
run python bench.py --steps 300 --warmup 20 --no-cpu-baseline
for op in phone.fe3 phone.fe1 wave.ups1 wave.ups3; do
( BEATRICE_B200_TC_TRACE=$op timeout 300 python tools/op_profile.py 2 256 2 ) > gpurun_out/trace.log 2>&1
echo "== trace $op" >> $L
grep "tc trace" gpurun_out/trace.log | tail -12 | cut -c1-600 >> $L
done
grep "wave.post" gpurun_out/trace.log >> $L
cat $L
