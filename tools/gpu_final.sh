#!/bin/bash
# the round's evidence in one go: GPU tests, the default bench line (all configs), the reference arm, the ncu launch list of
# the bench command and a full-set capture of one timed step.  usage: tools/gpu_final.sh <tag>
TAG=$1
mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/${TAG}_pytest_gpu.log 2>&1; tail -3 gpurun_out/${TAG}_pytest_gpu.log
python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; tail -c 600 gpurun_out/${TAG}_bench.json; echo
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${TAG}_bench_reference.json 2> gpurun_out/${TAG}_bench_reference.err; tail -c 700 gpurun_out/${TAG}_bench_reference.json; echo
PROF="python bench.py --psms 262144 --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-configs"
timeout 600 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none -c 400 --csv \
    --log-file gpurun_out/${TAG}_launches_lowres_262144.csv $PROF > gpurun_out/${TAG}_ncu_launch.log 2>&1
PER=$(python - <<PY
import csv,re
rows=[r for r in csv.reader(open("gpurun_out/${TAG}_launches_lowres_262144.csv")) if len(r)>12 and r[0].isdigit() and r[12].startswith("gpu__time_duration")]
names=[r[4].split("(")[0] for r in rows if re.search(r"k_bin_|k_count_score|k_select|k_ascore", r[4])]
idx=[i for i,n in enumerate(names) if "k_bin_rows" in n]
print(idx[1]-idx[0] if len(idx)>1 else len(names))
PY
)
echo "kernels per step: $PER"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_bin_|k_count_score|k_select|k_ascore' -s $((PER*3)) -c $PER \
    -f -o gpurun_out/${TAG}_prof_lowres $PROF > gpurun_out/${TAG}_ncu_full.log 2>&1
tail -2 gpurun_out/${TAG}_ncu_full.log | cut -c1-200
