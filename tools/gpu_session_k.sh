#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/k.log
: > $L
for op in phone.fe3 phone.fe1; do
( BEATRICE_B200_REPEAT_OP=$op BEATRICE_B200_TC_TRACE=$op timeout 300 python tools/op_profile.py 2 256 2 ) > gpurun_out/trace.log 2>&1
echo "== trace $op" >> $L
grep "tc trace\] KS\|tc trace\] wall" gpurun_out/trace.log | tail -4 | cut -c1-600 >> $L
done
cat $L
