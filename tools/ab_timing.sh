#!/bin/bash
# stage clocks of the leader loop (-DML_TIMING=1 build): tools/ab_timing.sh <workload> <tag> [env assignments...]
wl=$1; tag=$2; shift 2
YABPE_NVCC_EXTRA="-DML_TIMING=1" python yet-another-bpe_b200/build.py --force > /dev/null 2>&1 || echo build failed
env "$@" YABPE_DUMP_STATE=1 timeout 400 python bench.py --workload $wl --skip-cpu --skip-e2e --steps 1 --warmup 1 --encode-mb 0 2> gpurun_out/abt_${tag}.err | tee gpurun_out/abt_${tag}.raw | grep -a "^{" > gpurun_out/abt_${tag}.json; grep -a "state\[" gpurun_out/abt_${tag}.raw | tail -n 1; true 2> gpurun_out/abt_${tag}.err || tail -c 800 gpurun_out/abt_${tag}.err
python yet-another-bpe_b200/build.py --force > /dev/null 2>&1
python - <<PY
import json
d = json.loads(open("gpurun_out/abt_${tag}.json").read().strip().splitlines()[-1])
k = [x for x in d if x.startswith("leader_cycles")][0]
print("${tag}", d["ms_per_step"], "ms/step", d["us_per_merge"], "us/merge")
print("   ", k, d[k])
print("   ", d["merge_phase_ms"])
print("   ", d["merge_loop"])
PY
