#!/bin/bash
# one GPU box session: fused-MRF bring-up (probe with bisect configurations, tests, bench)
mkdir -p gpurun_out
L=gpurun_out/a_probe.log
: > $L
run() { echo "== $*" >> $L; ( timeout 120 env "$@" ) >> $L 2>&1; echo "rc=$?" >> $L; }
run python tools/mrf_probe.py 2 5 6
run BEATRICE_B200_MRF_STAGES=8 python tools/mrf_probe.py 2 5 6
run BEATRICE_B200_MRF_STAGES=4 python tools/mrf_probe.py 2 5 6
run BEATRICE_B200_MRF_STAGES=2 python tools/mrf_probe.py 2 5 6
run BEATRICE_B200_MRF_STAGES=0 python tools/mrf_probe.py 2 5 6
run BEATRICE_B200_MRF_S=8,8,8 python tools/mrf_probe.py 2 16 6
run python tools/mrf_probe.py 1 5 6
run python tools/mrf_probe.py 2 20 6
(timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -15) > gpurun_out/a_tests.log
timeout 300 python bench.py --steps 200 --warmup 20 > gpurun_out/a_bench_x3.json 2> gpurun_out/a_bench_x3.err
timeout 300 python bench.py --steps 200 --warmup 20 --precision bf16 > gpurun_out/a_bench_bf16.json 2> gpurun_out/a_bench_bf16.err
BEATRICE_B200_NO_FUSED_MRF=1 timeout 300 python bench.py --steps 100 --warmup 20 > gpurun_out/a_bench_x3_nofused.json 2> /dev/null
cat $L; cat gpurun_out/a_tests.log; head -c 600 gpurun_out/a_bench_x3.json; tail -3 gpurun_out/a_bench_x3.err
