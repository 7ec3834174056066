#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/k.log
: > $L
run() { echo "== $*" >> $L; ( timeout 300 env "$@" ) 2>&1 | cut -c1-1500 >> $L; echo "rc=$?" >> $L; }
run python tools/mrf_probe.py 2 40 6
run python bench.py --steps 500 --warmup 30 --no-cpu-baseline
run python tools/op_profile.py 2 256
cat $L
