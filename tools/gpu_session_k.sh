#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/k.log
: > $L
run() { echo "== $*" >> $L; ( timeout 300 env "$@" ) 2>&1 | cut -c1-1500 >> $L; echo "rc=$?" >> $L; }
run BEATRICE_B200_UPS0_BN=64 python bench.py --steps 400 --warmup 30 --no-cpu-baseline
run python bench.py --steps 400 --warmup 30 --no-cpu-baseline
run BEATRICE_B200_UPS0_BN=32 python bench.py --steps 400 --warmup 30 --no-cpu-baseline
cat $L
