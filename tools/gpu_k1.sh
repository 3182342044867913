#!/bin/bash
# K1 work loop: binning / parity tests, config-2 and config-5 kernel times, ncu full-set capture of K1.  usage: tools/gpu_k1.sh <tag>
TAG=$1
mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py tests/test_gpu_secondary.py tests/test_gpu_shard.py -x -q > gpurun_out/${TAG}_pytest.log 2>&1; tail -4 gpurun_out/${TAG}_pytest.log | cut -c1-300
for W in lowres_phospho acetyl_k; do
  python bench.py --workload $W --steps 3 --no-configs --no-cpu-baseline --no-e2e > gpurun_out/${TAG}_bench_$W.json 2> gpurun_out/${TAG}_bench_$W.err
  python -c "
import json
d=json.loads(open('gpurun_out/${TAG}_bench_$W.json').read().strip().splitlines()[-1])
print('$W', 'value %.4g' % d['value'], {k: round(v,2) for k,v in d['kernel_ms_per_step'].items()}, d['host_and_device_paths_bit_identical'])"
done
PROF="python bench.py --psms 262144 --steps 1 --warmup 3 --no-cpu-baseline --no-e2e --no-configs"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:${K1_KERNEL:-k_bin_rows} -s 3 -c 1 -f -o gpurun_out/${TAG}_prof_k1 $PROF > gpurun_out/${TAG}_ncu.log 2>&1
tail -1 gpurun_out/${TAG}_ncu.log
