#!/bin/bash
# one GPU box session: state check after restore (tests, bench, launch list)
mkdir -p gpurun_out
(timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -15) > gpurun_out/b_tests.log
timeout 300 python bench.py --steps 200 --warmup 20 > gpurun_out/b_bench_x3.json 2> gpurun_out/b_bench_x3.err
timeout 300 python bench.py --steps 200 --warmup 20 --precision bf16 > gpurun_out/b_bench_bf16.json 2> gpurun_out/b_bench_bf16.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/b_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/b_ncu.log 2>&1
cat gpurun_out/b_tests.log; head -c 1500 gpurun_out/b_bench_x3.json; echo; head -c 800 gpurun_out/b_bench_bf16.json; tail -3 gpurun_out/b_bench_x3.err
