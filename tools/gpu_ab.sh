#!/bin/bash
# A/B of the in-tree library against every build/variants/libpa_*.so (kernel times only, no parity tests).
# usage: tools/gpu_ab.sh <tag> [workloads...]
TAG=$1; shift
mkdir -p gpurun_out
for W in ${@:-lowres_phospho}; do
  python bench.py --workload $W --steps 3 --no-cpu-baseline --no-e2e > gpurun_out/${TAG}_main_$W.json 2> gpurun_out/${TAG}_main_$W.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/${TAG}_main_$W.json").read().strip().splitlines()[-1])
    print("%-14s %-18s value %.4g same %s %s" % ("main", "$W", d["value"], d["host_and_device_paths_bit_identical"], {k: round(v,2) for k,v in d["kernel_ms_per_step"].items()}))
except Exception as e:
    print("main $W failed", e)
PY
done
bash tools/gpu_variants.sh $TAG "$@"
