#!/bin/bash
# e2e of config 2 with the host narrowing pass on a few thread budgets / off.  usage: tools/gpu_narrow.sh <tag> [ENV=VALUE ...]
TAG=$1; shift
CFGS=("$@"); [ ${#CFGS[@]} -eq 0 ] && CFGS=("PA_NARROW=0" "PA_HOST_THREADS=16" "PA_HOST_THREADS=8" "PA_HOST_THREADS=4" "PA_HOST_THREADS=2")
mkdir -p gpurun_out
nproc
for CFG in "${CFGS[@]}"; do
  env $CFG python bench.py --steps 4 --no-configs --no-cpu-baseline > gpurun_out/${TAG}_n.json 2> gpurun_out/${TAG}_n.err
  python - <<PY
import json
d=json.loads(open("gpurun_out/${TAG}_n.json").read().strip().splitlines()[-1])
e=d["e2e"]
print("$CFG", "e2e %.4g M  ms %.1f  h2d %.1f GB/s  wait %.1f ms  exact %d  f32 %.4g" % (e["value"]/1e6, e["ms_per_step"], e["h2d_achieved_gbs"], e["host_narrowing"]["ms_waited_per_step"], e["host_narrowing"]["spectra_kept_exact_per_step"], d["e2e_f32_intensity"]["value"]/1e6))
PY
done
