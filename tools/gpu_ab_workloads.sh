#!/bin/bash
# GPU tests + a short kernel-time line per workload.  usage: tools/gpu_ab_workloads.sh <tag> [env assignments ...]
TAG=${1:-run}; shift
mkdir -p gpurun_out
for kv in "$@"; do export "$kv"; done
python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/${TAG}_pytest.log
tail -12 gpurun_out/${TAG}_pytest.log
for W in lowres_phospho hires_phospho_nl acetyl_k stress; do
  python bench.py --workload $W --steps 3 --no-configs --no-cpu-baseline > gpurun_out/${TAG}_bench_$W.json 2> gpurun_out/${TAG}_bench_$W.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/${TAG}_bench_$W.json").read().strip().splitlines()[-1])
    print("$W", "value %.4g e2e %.4g kernels %s same %s bad %s" % (d["value"], d["e2e"]["value"], {k: round(v,2) for k,v in d["kernel_ms_per_step"].items()}, d["host_and_device_paths_bit_identical"], d["psms_not_scored"]))
except Exception as e:
    print("$W failed", e); print(open("gpurun_out/${TAG}_bench_$W.err").read()[-1500:])
PY
done
