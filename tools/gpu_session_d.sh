#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/d_ops.log
: > $L
run() { echo "== $*" >> $L; ( timeout 120 env "$@" ) >> $L 2>&1; echo "rc=$?" >> $L; }
run BEATRICE_B200_MRF_TRACE=1 python tools/op_profile.py 2 256 2
grep -E "^==|mrfc trace|wave\.|serial|rc=" $L | tail -60
