#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/k.log
: > $L
run() { echo "== $*" >> $L; ( timeout 200 env "$@" ) 2>&1 | grep '^{' | cut -c1-900 >> $L; echo "rc=$?" >> $L; }
run BEATRICE_B200_PRECISION=bf16x3 python tools/config_bench.py latency 3000
run python tools/config_bench.py sweep 1000 128
cat $L
