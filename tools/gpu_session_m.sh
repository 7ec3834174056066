#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/m.log 2>&1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 100 --warmup 10 >> gpurun_out/m.log 2>&1
echo "rc=$?" >> gpurun_out/m.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 >> gpurun_out/m.log 2>&1
echo "rc=$?" >> gpurun_out/m.log
tail -c 6000 gpurun_out/m.log
