#!/bin/bash
# A/B of builds on the end-to-end leg of config 2 (narrowing forced on, so that the policy's own timing does not interfere), then
# the narrowing / sharding parity tests with the last build.  usage: tools/gpu_host_ab2.sh <tag> <variant> [<variant> ...]
TAG=$1; shift
mkdir -p gpurun_out
for V in "$@"; do
  PA_NARROW=1 PYASCORE_B200_LIB=$PWD/build/variants/libpa_$V.so python bench.py --steps 5 --no-configs --no-cpu-baseline > gpurun_out/${TAG}_$V.json 2> gpurun_out/${TAG}_$V.err
  python - <<PY
import json
d=json.loads(open("gpurun_out/${TAG}_$V.json").read().strip().splitlines()[-1])
e=d["e2e"]
print("$V", "e2e %.4g M  ms %.1f  h2d %.1f GB/s  wait %.1f ms  f32 %.4g same %s" % (e["value"]/1e6, e["ms_per_step"], e["h2d_achieved_gbs"], e["host_narrowing"]["ms_waited_per_step"], d["e2e_f32_intensity"]["value"]/1e6, d["host_and_device_paths_bit_identical"]))
PY
done
python -m pytest tests/test_gpu_shard.py -q -x 2>&1 | tail -2
