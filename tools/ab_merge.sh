#!/bin/bash
# A/B of the merge loop on one workload: tools/ab_merge.sh <workload> <tag> [env assignments...]
# prints ms_per_step, us_per_merge, the phase clocks and the batch statistics of bench.py's JSON line
wl=$1; tag=$2; shift 2
env "$@" timeout 400 python bench.py --workload $wl --skip-cpu --skip-e2e --steps 2 --warmup 1 > gpurun_out/ab_${tag}.json 2> gpurun_out/ab_${tag}.err || tail -c 800 gpurun_out/ab_${tag}.err
python - <<PY
import json
d = json.loads(open("gpurun_out/ab_${tag}.json").read().strip().splitlines()[-1])
print("${tag}", d["ms_per_step"], "ms/step", d["us_per_merge"], "us/merge", d.get("digest"))
print("   ", d["merge_phase_ms"])
print("   ", d["merge_loop"])
PY
