#!/bin/bash
# multi-GPU bench exactly as the driver launches it.  usage: tools/gpu_scale.sh <tag> <N> [extra bench args]
TAG=$1; N=$2; shift; shift
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $N --steps 5 --warmup 3 "$@" > gpurun_out/${TAG}_scale_n$N.json 2> gpurun_out/${TAG}_scale_n$N.err
tail -c 1500 gpurun_out/${TAG}_scale_n$N.json; tail -3 gpurun_out/${TAG}_scale_n$N.err
