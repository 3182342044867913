#!/bin/bash
# ncu full-set capture of selected kernels on a workload.  usage: tools/gpu_prof.sh <tag> <kernel-regex> [workload] [psms]
TAG=$1; RE=$2; W=${3:-lowres_phospho}; N=${4:-262144}
mkdir -p gpurun_out
PROF="python bench.py --workload $W --psms $N --steps 1 --warmup 3 --no-cpu-baseline --no-e2e"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"$RE" -s 6 -c 3 \
    -f -o gpurun_out/${TAG}_prof $PROF > gpurun_out/${TAG}_ncu_full.log 2>&1
tail -3 gpurun_out/${TAG}_ncu_full.log | cut -c1-300
