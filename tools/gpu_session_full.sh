#!/bin/bash
# full check: GPU tests, smoke, bench (all precisions), latency mode, op profile
mkdir -p gpurun_out
L=gpurun_out/full.log
: > $L
(timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5) >> $L
(timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3) >> $L
timeout 600 python bench.py > gpurun_out/bench_x3.json 2> gpurun_out/bench_x3.err; echo "bench x3 rc=$?" >> $L
timeout 600 python bench.py --precision bf16 --no-cpu-baseline > gpurun_out/bench_bf16.json 2> gpurun_out/bench_bf16.err; echo "bench bf16 rc=$?" >> $L
timeout 600 python bench.py --precision f32 --no-cpu-baseline > gpurun_out/bench_f32.json 2> gpurun_out/bench_f32.err; echo "bench f32 rc=$?" >> $L
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "bench ref rc=$?" >> $L
timeout 120 python tools/op_profile.py 2 256 > gpurun_out/ops_x3.log 2>&1
cat $L; head -c 3000 gpurun_out/bench_x3.json; echo; head -c 400 gpurun_out/bench_bf16.json; echo; head -c 400 gpurun_out/bench_f32.json; echo; head -c 600 gpurun_out/bench_ref.json
