#!/bin/bash
# GPU parity tests + short bench lines (no CPU baseline).  usage: tools/gpu_quick.sh <tag> [workloads...]
TAG=${1:-run}; shift
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/${TAG}_pytest.log
tail -8 gpurun_out/${TAG}_pytest.log
for W in ${@:-lowres_phospho}; do
  python bench.py --workload $W --steps 3 --no-cpu-baseline > gpurun_out/${TAG}_bench_$W.json 2> gpurun_out/${TAG}_bench_$W.err
  python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/${TAG}_bench_$W.json").read().strip().splitlines()[-1])
    print("$W", "value %.4g e2e %.4g same %s kernels %s" % (d["value"], d["e2e"]["value"], d["host_and_device_paths_bit_identical"], {k: round(v,2) for k,v in d["kernel_ms_per_step"].items()}))
except Exception as e:
    print("$W bench failed", e); print(open("gpurun_out/${TAG}_bench_$W.err").read()[-2000:])
PY
done
