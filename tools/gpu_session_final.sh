#!/bin/bash
# round-end artefacts: tests, smoke, bench lines (3 precisions + CPU arm), op profile, ncu launch list + full captures
mkdir -p gpurun_out
L=gpurun_out/final.log
: > $L
(timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5) >> $L
(timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3) >> $L
timeout 600 python bench.py > gpurun_out/bench_x3.json 2> gpurun_out/bench_x3.err; echo "bench x3 rc=$?" >> $L
timeout 600 python bench.py --precision bf16 --no-cpu-baseline > gpurun_out/bench_bf16.json 2> gpurun_out/bench_bf16.err; echo "bench bf16 rc=$?" >> $L
timeout 600 python bench.py --precision f32 --no-cpu-baseline > gpurun_out/bench_f32.json 2> gpurun_out/bench_f32.err; echo "bench f32 rc=$?" >> $L
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "bench ref rc=$?" >> $L
timeout 120 python tools/op_profile.py 2 256 > gpurun_out/ops_x3.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r1g_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_a.log 2>&1
echo "launch list rc=$?" >> $L
timeout 900 ncu --set full --clock-control none --import-source on -k regex:mrf_cluster_kernel -s 6 -c 1 -f -o gpurun_out/r1g_mrf_cluster python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_b.log 2>&1
echo "mrf_cluster rc=$?" >> $L
timeout 900 ncu --set full --clock-control none --import-source on -k regex:mrf_branch_kernel -s 24 -c 4 -f -o gpurun_out/r1g_mrf_branch python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_c.log 2>&1
echo "mrf_branch rc=$?" >> $L
timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv_gemm_tc_kernel -s 200 -c 4 -f -o gpurun_out/r1g_conv_tc python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_d.log 2>&1
echo "conv_tc rc=$?" >> $L
timeout 900 ncu --set full --clock-control none --import-source on -k regex:enc_res_stack_kernel -s 6 -c 3 -f -o gpurun_out/r1g_enc_res_stack python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_e.log 2>&1
echo "enc_res_stack rc=$?" >> $L
cat $L; head -c 2500 gpurun_out/bench_x3.json
