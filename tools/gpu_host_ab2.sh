#!/bin/bash
# A/B of two builds on the end-to-end leg of the DEFAULT bench run (configs block on, default narrowing policy).  usage: tools/gpu_host_ab2.sh <tag>
TAG=$1
mkdir -p gpurun_out
for V in nthost oldhost nthost oldhost; do
  PYASCORE_B200_LIB=$PWD/build/variants/libpa_$V.so python bench.py --steps 5 --no-cpu-baseline > gpurun_out/${TAG}_$V.json 2> gpurun_out/${TAG}_$V.err
  python - <<PY
import json
d=json.loads(open("gpurun_out/${TAG}_$V.json").read().strip().splitlines()[-1])
e=d["e2e"]
print("$V", "e2e %.4g M  ms %.1f  h2d %.1f GB/s  wait %.1f ms  f32 %.4g" % (e["value"]/1e6, e["ms_per_step"], e["h2d_achieved_gbs"], e["host_narrowing"]["ms_waited_per_step"], d["e2e_f32_intensity"]["value"]/1e6))
PY
done
