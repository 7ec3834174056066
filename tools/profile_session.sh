#!/bin/bash
# The GPU half of the round's evidence.  Run on the box:   gpurun --timeout 1500 -- 'bash tools/profile_session.sh r2'
# then, here:                                              python tools/profile_collect.py r2
# Everything lands in gpurun_out/ (scratch); profile_collect.py turns it into the tracked summaries under profiles/.
T=${1:-r2}
O=gpurun_out
mkdir -p $O
# 1. bench lines: driver-sized (20 steps) and long (1000 steps), depth 2 headline + depth 1 latency mode inside each
python bench.py --steps 20 --warmup 5 > $O/${T}_bench_20.json 2> $O/${T}_bench_20.err
python bench.py --steps 1000 --warmup 50 > $O/${T}_bench_1000.json 2> $O/${T}_bench_1000.err
python bench.py --impl reference --steps 20 --warmup 5 > $O/${T}_bench_reference.json 2> $O/${T}_bench_reference.err
python bench.py --steps 300 --warmup 20 --precision bf16 --no-cpu-baseline > $O/${T}_bench_bf16.json 2> $O/${T}_bench_bf16.err
python bench.py --steps 100 --warmup 10 --precision f32 --no-cpu-baseline > $O/${T}_bench_f32.json 2> $O/${T}_bench_f32.err
# 2. every launch of the SAME bench command with its device time (cold-cache, serialised: shares, not absolutes)
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/${T}_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $O/${T}_launches.log 2>&1
# 3. full captures of the hot kernels (one hop of a 256-stream engine, un-graphed so that -k / -s / -c address launches)
#    form 0 = the upsamplers as launches of their own (the depth-2 headline); mrf_branch_ups = the depth-1 form of the same kernel
for spec in "mrf_cluster:mrf_cluster_kernel:2:2:0" "mrf_branch:mrf_branch_kernel:4:4:0" "mrf_branch_ups:mrf_branch_kernel:4:4:1" "conv_tc:conv_gemm_tc_kernel:17:17:0" "enc_res_stack:enc_res_stack_kernel:3:3:0"; do
  IFS=: read name pat skip count form <<< "$spec"
  ncu --set full --clock-control none --import-source on -k regex:$pat -s $skip -c $count -o $O/${T}_$name -f \
      python tools/tc_probe.py 2 256 3 $form > $O/${T}_ncu_$name.log 2>&1
done
# 4. per-op device times of one hop (CUDA events around every launch) and the in-kernel timelines of the MRF kernels
python tools/op_profile.py 2 256 8 2 > $O/${T}_ops_bf16x3.txt 2>&1
python tools/op_profile.py 2 256 8 1 > $O/${T}_ops_bf16x3_depth1.txt 2>&1
BEATRICE_B200_MRF_TRACE=1 python tools/tc_probe.py 2 256 3 0 2>&1 | grep "mrfc\? trace" | tail -35 > $O/${T}_mrf_timeline.txt
BEATRICE_B200_MRF_TRACE=1 python tools/tc_probe.py 2 256 3 1 2>&1 | grep "mrf trace" | tail -24 > $O/${T}_mrf_timeline_ups.txt
# 4b. marginal cost of every vocoder op inside the hop graph (hop time with the op's launches dropped), both depths
bash tools/ablate_session.sh > $O/${T}_ablation.txt 2>&1
# 5. configs 4 and 5 (one GPU here; the 8-GPU sweep is a separate --gpus 8 call)
python tools/config_bench.py latency 10000 --out $O/${T}_config4_latency.json > /dev/null 2>&1
python tools/config_bench.py sweep 1000 128 --depth 2 --out $O/${T}_config5_1gpu.json > /dev/null 2>&1
ls -la $O | tail -40
