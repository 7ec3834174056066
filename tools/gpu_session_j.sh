#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/j.log
: > $L
for op in wave.ups2 phone.fe1; do
  echo "== $op (PDL off)" >> $L
  BEATRICE_B200_NO_PDL=1 BEATRICE_B200_TC_TRACE=$op timeout 120 python tools/op_profile.py 2 256 2 > gpurun_out/j_tmp.log 2>&1
  grep "tc trace" gpurun_out/j_tmp.log | awk '/grid/{buf=""} {buf=buf"\n"$0} END{print buf}' >> $L
done
BEATRICE_B200_NO_PDL=1 timeout 300 python bench.py --steps 200 --warmup 20 --no-cpu-baseline 2>&1 | cut -c1-200 >> $L
cut -c1-400 $L
