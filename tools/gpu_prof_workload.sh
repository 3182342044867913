#!/bin/bash
# ncu launch list + full-set capture of ONE timed step of a workload (all the path's kernels).
# usage: tools/gpu_prof_workload.sh <tag> <workload> [psms]
TAG=$1; W=$2; N=${3:-262144}
mkdir -p gpurun_out
PROF="python bench.py --workload $W --psms $N --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-configs"
timeout 600 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none -c 400 --csv \
    --log-file gpurun_out/${TAG}_launches_${W}.csv $PROF > gpurun_out/${TAG}_ncu_launch_${W}.log 2>&1
# kernels of the path per step: find how many matching launches one step has from the launch list, then
# capture the 4th step (3 warm-ups come first)
PER=$(python - <<PY
import csv,re
rows=[r for r in csv.reader(open("gpurun_out/${TAG}_launches_${W}.csv")) if len(r)>5 and re.search(r"k_bin_|k_count_score|k_select|k_ascore", r[4] if len(r)>4 else "")]
names=[r[4].split("(")[0] for r in rows if r[-3].startswith("gpu__time_duration")]
# one step = from one k_bin_rows to the next
idx=[i for i,n in enumerate(names) if "k_bin_rows" in n]
print(idx[1]-idx[0] if len(idx)>1 else len(names))
PY
)
echo "kernels per step: $PER"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_bin_|k_count_score|k_select|k_ascore' -s $((PER*3)) -c $PER \
    -f -o gpurun_out/${TAG}_prof_${W} $PROF > gpurun_out/${TAG}_ncu_full_${W}.log 2>&1
tail -2 gpurun_out/${TAG}_ncu_full_${W}.log | cut -c1-200
ls -la gpurun_out | grep ${TAG}
