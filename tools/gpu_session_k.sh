#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/k.log
: > $L
run() { echo "== $*" >> $L; ( timeout 300 env "$@" ) 2>&1 | cut -c1-1500 >> $L; echo "rc=$?" >> $L; }
(timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -6) >> $L
run python bench.py --steps 400 --warmup 30 --no-cpu-baseline
cat $L
