#!/bin/bash
# one GPU box session: cluster MRF kernel bring-up (stage 0)
mkdir -p gpurun_out
L=gpurun_out/c_probe.log
: > $L
run() { echo "== $*" >> $L; ( timeout 120 env "$@" ) >> $L 2>&1; echo "rc=$?" >> $L; }
run BEATRICE_B200_MRF_STAGES=14 python tools/mrf_probe.py 2 5 4
run python tools/mrf_probe.py 2 5 4
run python tools/mrf_probe.py 2 40 6
run python tools/mrf_probe.py 1 40 6
run python tools/mrf_probe.py 2 256 3
timeout 300 python bench.py --steps 200 --warmup 20 --no-cpu-baseline > gpurun_out/c_bench_x3.json 2> gpurun_out/c_bench_x3.err
cat $L; head -c 700 gpurun_out/c_bench_x3.json; tail -3 gpurun_out/c_bench_x3.err
