#!/bin/bash
# usage: tools/gpu_dbg2.sh <tag>
TAG=$1
mkdir -p gpurun_out
python -m pytest tests/test_gpu_cli.py tests/test_gpu_parity.py -x -q > gpurun_out/${TAG}_pytest.log 2>&1; tail -8 gpurun_out/${TAG}_pytest.log | cut -c1-400
for W in lowres_phospho hires_phospho_nl acetyl_k; do
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
PROF="python bench.py --workload hires_phospho_nl --psms 262144 --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-configs"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_ascore_items' -s 3 -c 1 -f -o gpurun_out/${TAG}_prof_items $PROF > gpurun_out/${TAG}_ncu.log 2>&1
tail -2 gpurun_out/${TAG}_ncu.log
