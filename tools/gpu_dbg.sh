#!/bin/bash
# usage: tools/gpu_dbg.sh <tag>
TAG=$1
mkdir -p gpurun_out
timeout 600 compute-sanitizer --tool memcheck --print-limit 5 python -m pytest tests/test_gpu_cli.py -x -q -k "hit_depth2" > gpurun_out/${TAG}_sanitizer.log 2>&1
grep -E "Invalid|at 0x|pa_|=========     in" gpurun_out/${TAG}_sanitizer.log | head -30
python -m pytest tests/test_gpu_big.py -x -q > gpurun_out/${TAG}_pytest_big.log 2>&1; tail -5 gpurun_out/${TAG}_pytest_big.log
python bench.py --workload stress --steps 3 --no-configs --no-cpu-baseline > gpurun_out/${TAG}_bench_stress.json 2> gpurun_out/${TAG}_bench_stress.err
python -c "
import json
d=json.loads(open('gpurun_out/${TAG}_bench_stress.json').read().strip().splitlines()[-1])
print('stress', d['value'], d['kernel_ms_per_step'])"
PROF="python bench.py --workload hires_phospho_nl --psms 262144 --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-configs"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_ascore_items' -s 3 -c 1 -f -o gpurun_out/${TAG}_prof_items $PROF > gpurun_out/${TAG}_ncu.log 2>&1
tail -2 gpurun_out/${TAG}_ncu.log
