#!/bin/bash
# ncu artefacts for profiles/: launch list of a short bench run + full captures of the fused MRF kernels
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r1c_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_a.log 2>&1
echo "launch list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:mrf_cluster_kernel -s 6 -c 2 -f -o gpurun_out/r1c_mrf_cluster python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_b.log 2>&1
echo "mrf_cluster rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:mrf_branch_kernel -s 18 -c 3 -f -o gpurun_out/r1c_mrf_branch python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_c.log 2>&1
echo "mrf_branch rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv_gemm_tc_kernel -s 200 -c 4 -f -o gpurun_out/r1c_conv_tc python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_d.log 2>&1
echo "conv_tc rc=$?"
ls -la gpurun_out/*.ncu-rep gpurun_out/r1c_launches.csv
tail -3 gpurun_out/ncu_a.log
