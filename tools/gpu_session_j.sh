#!/bin/bash
mkdir -p gpurun_out
for op in phone.res0.conv phone.fe4; do
BEATRICE_B200_TC_TRACE=$op timeout 120 python tools/op_profile.py 2 256 2 > gpurun_out/j_tmp.log 2>&1
grep "tc trace" gpurun_out/j_tmp.log | awk '/wall/{buf=""} {buf=buf"\n"$0} END{print buf}' | cut -c1-330
done
grep "ops\]" gpurun_out/j_tmp.log | head -24
