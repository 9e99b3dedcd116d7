#!/bin/bash
# select_batch timed by thread 0 of CTA 0 (-DML_RW_TRACE=2): cycles in state[48..53] = list + barrier | refresh | rank + fields |
# tie order | members | final barrier.   usage: tools/ab_seltrace.sh <workload> [env assignments...]
wl=$1; shift
YABPE_NVCC_EXTRA="-DML_RW_TRACE=2" python yet-another-bpe_b200/build.py --force > /dev/null 2>&1 || echo build failed
env "$@" timeout 400 python bench.py --workload $wl --skip-cpu --skip-e2e --steps 1 --warmup 1 --encode-mb 0 > gpurun_out/seltrace.json 2> gpurun_out/seltrace.err || tail -c 800 gpurun_out/seltrace.err
python yet-another-bpe_b200/build.py --force > /dev/null 2>&1
python - <<PY
import json
d = json.loads(open("gpurun_out/seltrace.json").read().strip().splitlines()[-1])
ph = d["merge_phase_ms"]
print("select_batch cycles [list+barrier, refresh, rank+fields] / ms [tie order, members, final barrier]:", ph.get("grid_merges_by_size[n<=2368,n<=18944,more]"), ph.get("grid_cycles_by_size"))
print(d["merge_loop"])
PY
