#!/bin/bash
# parity tests, then kernel times with K2 in thread form (default) and warp form.  usage: tools/gpu_k2ab.sh <tag> [workloads...]
TAG=$1; shift
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/${TAG}_pytest.log
tail -4 gpurun_out/${TAG}_pytest.log
for W in ${@:-lowres_phospho}; do
  for M in thread warp; do
    PYASCORE_B200_K2=$M python bench.py --workload $W --steps 3 --no-cpu-baseline --no-e2e > gpurun_out/${TAG}_${M}_$W.json 2> gpurun_out/${TAG}_${M}_$W.err
    python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/${TAG}_${M}_$W.json").read().strip().splitlines()[-1])
    print("%-8s %-18s value %.4g same %s %s lookups %d" % ("$M", "$W", d["value"], d["host_and_device_paths_bit_identical"], {k: round(v,2) for k,v in d["kernel_ms_per_step"].items()}, d["fragment_lookups_per_step"]))
except Exception as e:
    print("$M $W failed", e); print(open("gpurun_out/${TAG}_${M}_$W.err").read()[-1500:])
PY
  done
done
