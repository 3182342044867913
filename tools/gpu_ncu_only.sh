#!/bin/bash
# ncu launch list + full-set capture of one timed step (no tests, no bench).  usage: tools/gpu_ncu_only.sh <tag>
TAG=${1:-run}
mkdir -p gpurun_out
PROF="python bench.py --psms 262144 --steps 1 --warmup 3 --no-cpu-baseline --no-e2e"
timeout 600 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none -c 200 --csv \
    --log-file gpurun_out/${TAG}_launches.csv $PROF > gpurun_out/${TAG}_ncu_launch.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_bin_|k_count_score|k_select|k_ascore$' -s 24 -c 8 \
    -f -o gpurun_out/${TAG}_prof $PROF > gpurun_out/${TAG}_ncu_full.log 2>&1
ls -la gpurun_out | grep ${TAG}
