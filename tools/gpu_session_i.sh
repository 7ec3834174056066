#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/i.log
: > $L
run() { echo "== $*" >> $L; ( timeout 300 env "$@" ) >> $L 2>&1; echo "rc=$?" >> $L; }
run python tools/mrf_probe.py 2 40 4
(timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15) >> $L
run python bench.py --steps 200 --warmup 20 --no-cpu-baseline
run python tools/op_profile.py 2 256
for op in phone.fe1 phone.fe3 wave.ups2 wave.ups1; do
  echo "== $op" >> $L
  BEATRICE_B200_TC_TRACE=$op timeout 120 python tools/op_profile.py 2 256 2 > gpurun_out/i_tmp.log 2>&1
  grep "tc trace" gpurun_out/i_tmp.log | awk '/grid/{buf=""} {buf=buf"\n"$0} END{print buf}' >> $L
done
cut -c1-330 $L
