#!/bin/bash
# GPU parity tests + one bench line per BASELINE workload (configs 2-5).  usage: tools/gpu_workloads.sh <tag>
TAG=${1:-run}
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/${TAG}_pytest.log
tail -5 gpurun_out/${TAG}_pytest.log
for W in hires_phospho_nl acetyl_k stress; do
  python bench.py --workload $W --steps 3 > gpurun_out/${TAG}_bench_$W.json 2> gpurun_out/${TAG}_bench_$W.err
  python - <<PY
import json
d=json.loads(open("gpurun_out/${TAG}_bench_$W.json").read().strip().splitlines()[-1])
print("$W", "value %.3g e2e %.3g cpu %s kernels %s" % (d["value"], d["e2e"]["value"], d.get("cpu_baseline",{}).get("value"), {k: round(v,2) for k,v in d["kernel_ms_per_step"].items()}))
PY
done
