#!/bin/bash
# API tests + single-call latency.  usage: tools/gpu_single.sh <tag>
TAG=$1
mkdir -p gpurun_out
python -m pytest tests/test_gpu_pyascore_api.py tests/test_gpu_cli.py tests/test_gpu_secondary.py tests/test_gpu_parity.py -x -q > gpurun_out/${TAG}_pytest.log 2>&1; tail -6 gpurun_out/${TAG}_pytest.log | cut -c1-300
python - <<PY
import sys, json
sys.path.insert(0, ".")
import bench
print("ours", json.dumps(bench.gpu_single_call(0, n_rep=20)))
print("ref ", json.dumps(bench.cpu_single_call(n_rep=5)))
PY
