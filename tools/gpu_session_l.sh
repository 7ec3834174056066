#!/bin/bash
mkdir -p gpurun_out
BEATRICE_B200_MRF_TRACE=-1 timeout 120 python tools/op_profile.py 2 256 2 > gpurun_out/l_tmp.log 2>&1
python tools/cta_windows.py gpurun_out/l_tmp.log
BEATRICE_B200_MRF_TRACE=-1 timeout 120 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/l_tmp2.log 2>&1
echo "--- in graph (bench)"; python tools/cta_windows.py gpurun_out/l_tmp2.log
