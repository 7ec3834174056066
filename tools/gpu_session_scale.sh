#!/bin/bash
# weak-scaling check: bench.py on every GPU of the box gpurun gave us (run with --gpus 2, 4, 8 in turn; the driver
# runs its own 1 -> 8 sweep at round end)
mkdir -p gpurun_out
N=$(nvidia-smi -L | wc -l)
echo "gpus visible: $N" > gpurun_out/scale.log
for n in ${SCALE_NS:-$N}; do
  if [ "$n" -le "$N" ]; then
    timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29600+n)) bench.py --gpus $n --steps 500 --warmup 50 --no-cpu-baseline 2>/dev/null | grep '^{' > gpurun_out/scale_$n.json
    echo "n=$n rc=$? $(head -c 330 gpurun_out/scale_$n.json)" >> gpurun_out/scale.log
  fi
done
cat gpurun_out/scale.log
