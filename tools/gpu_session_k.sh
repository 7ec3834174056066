#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/k.log
: > $L
run() { echo "== $*" >> $L; ( timeout 300 env "$@" ) 2>&1 | cut -c1-1500 >> $L; echo "rc=$?" >> $L; }
run python tools/op_profile.py 2 256 12
run python bench.py --steps 500 --warmup 30 --no-cpu-baseline
cp beatrice_vst_b200/csrc/libbeatrice_b200.so /tmp/new.so
cp tools/scratch/old_lib.so beatrice_vst_b200/csrc/libbeatrice_b200.so
run python tools/op_profile.py 2 256 12
run python bench.py --steps 500 --warmup 30 --no-cpu-baseline
cp /tmp/new.so beatrice_vst_b200/csrc/libbeatrice_b200.so
run python bench.py --steps 500 --warmup 30 --no-cpu-baseline
cat $L
