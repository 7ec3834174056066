#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/k.log
: > $L
run() { echo "== $*" >> $L; ( timeout 300 env "$@" ) 2>&1 | cut -c1-600 >> $L; echo "rc=$?" >> $L; }
run python tools/mrf_probe.py 2 40 6
run BEATRICE_B200_NO_COND_CHAIN=1 python tools/mrf_probe.py 2 40 6
(timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8) >> $L
run python bench.py --steps 300 --warmup 20 --no-cpu-baseline
run BEATRICE_B200_NO_COND_CHAIN=1 python bench.py --steps 300 --warmup 20 --no-cpu-baseline
run python tools/op_profile.py 2 256
cat $L
