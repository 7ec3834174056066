#!/bin/bash
mkdir -p gpurun_out
timeout 300 python bench.py --no-cpu-baseline > gpurun_out/bench_x3n.json 2> gpurun_out/bench_x3.err; echo "x3 rc=$?"
timeout 300 python bench.py --precision bf16 --no-cpu-baseline > gpurun_out/bench_bf16.json 2> gpurun_out/bench_bf16.err; echo "bf16 rc=$?"
timeout 120 python tools/op_profile.py 2 256 > gpurun_out/ops_x3.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r1g_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_a.log 2>&1
echo "launch list rc=$?"
head -c 600 gpurun_out/bench_x3n.json
