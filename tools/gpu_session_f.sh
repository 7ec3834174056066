#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/f.log
: > $L
for op in phone.res0.conv wave.pre wave.ups0 phone.fe3 pitch.head wave.ups2; do
  echo "== $op" >> $L
  BEATRICE_B200_TC_TRACE=$op timeout 120 python tools/op_profile.py 2 256 2 > gpurun_out/f_tmp.log 2>&1
  grep "tc trace" gpurun_out/f_tmp.log | awk '/grid/{buf=""} {buf=buf"\n"$0} END{print buf}' >> $L
done
cat $L
