#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/k.log
: > $L
run() { echo "== $*" >> $L; ( timeout 300 env "$@" ) 2>&1 | cut -c1-1800 >> $L; echo "rc=$?" >> $L; }
(timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5) >> $L
run python bench.py --steps 500 --warmup 30 --no-cpu-baseline
cat $L
