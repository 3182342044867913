#!/bin/bash
# bench every build/variants/libpa_*.so on the given workloads (kernel times only).  usage: tools/gpu_variants.sh <tag> [workloads...]
TAG=$1; shift
mkdir -p gpurun_out
for SO in build/variants/libpa_*.so; do
  V=$(basename $SO .so | sed s/libpa_//)
  for W in ${@:-lowres_phospho}; do
    PYASCORE_B200_LIB=$PWD/$SO python bench.py --workload $W --steps 3 --no-cpu-baseline --no-e2e --no-configs > gpurun_out/${TAG}_${V}_$W.json 2> gpurun_out/${TAG}_${V}_$W.err
    python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/${TAG}_${V}_$W.json").read().strip().splitlines()[-1])
    print("%-14s %-18s value %.4g same %s %s" % ("$V", "$W", d["value"], d["host_and_device_paths_bit_identical"], {k: round(v,2) for k,v in d["kernel_ms_per_step"].items()}))
except Exception as e:
    print("$V $W failed", e)
PY
  done
done
