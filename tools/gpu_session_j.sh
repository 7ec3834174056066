#!/bin/bash
mkdir -p gpurun_out
BEATRICE_B200_TC_TRACE=phone. timeout 120 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/j_tmp.log 2>&1
grep "tc trace\] wall\|tc trace\] KS" gpurun_out/j_tmp.log | tail -24 | cut -c1-200
