#!/bin/bash
# per-class K3b kernel times (ncu launch list of one timed step) for the old and the new form.  usage: tools/gpu_asc_classes.sh <tag>
TAG=$1
mkdir -p gpurun_out
for V1 in 0 1; do
for W in lowres_phospho hires_phospho_nl acetyl_k; do
  PA_ASC_V1=$V1 timeout 600 ncu --metrics gpu__time_duration.sum,smsp__thread_inst_executed_per_inst_executed.ratio,sm__warps_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:'k_ascore' -c 40 --csv \
    --log-file gpurun_out/${TAG}_cls_v${V1}_$W.csv python bench.py --workload $W --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-configs > /dev/null 2>&1
  python - <<PY
import csv
rows=[r for r in csv.reader(open("gpurun_out/${TAG}_cls_v${V1}_$W.csv")) if len(r)>12 and r[0].isdigit()]
by={}
for r in rows:
    by.setdefault(int(r[0]),{})["name"]=r[4].split("(")[0].replace("void ","")
    by[int(r[0])][r[12]]=float(r[14].replace(",",""))
ids=sorted(by)
# last step = last group of launches: take the final occurrence of each kernel name
last={}
for i in ids: last[by[i]["name"]]=by[i]
print("V1=$V1 $W: "+"; ".join("%s %.2f ms thr %.1f warps %.0f%%" % (n.replace("k_ascore",""), v.get("gpu__time_duration.sum",0)/1e6, v.get("smsp__thread_inst_executed_per_inst_executed.ratio",0), v.get("sm__warps_active.avg.pct_of_peak_sustained_active",0)) for n,v in last.items()))
PY
done; done
