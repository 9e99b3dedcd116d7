#!/bin/bash
# chain of one rewrite_batch call in grid mode, timed by thread 0 of CTA 0 (-DML_RW_TRACE=1): cycles in state[48..51] =
# item fetch | claim + header | symbols + member mask | rewrites.   usage: tools/ab_rwtrace.sh <workload> [env assignments...]
wl=$1; shift
YABPE_NVCC_EXTRA="-DML_RW_TRACE=1" python yet-another-bpe_b200/build.py --force > /dev/null 2>&1 || echo build failed
env "$@" timeout 400 python bench.py --workload $wl --skip-cpu --skip-e2e --steps 1 --warmup 1 --encode-mb 0 > gpurun_out/rwtrace.json 2> gpurun_out/rwtrace.err || tail -c 800 gpurun_out/rwtrace.err
python yet-another-bpe_b200/build.py --force > /dev/null 2>&1
python - <<PY
import json
d = json.loads(open("gpurun_out/rwtrace.json").read().strip().splitlines()[-1])
ph = d["merge_phase_ms"]
print("grid merges by size (trace build: = fetch, claim+header, mask, rewrite in ms):", ph.get("grid_merges_by_size[n<=2368,n<=18944,more]"), ph.get("grid_cycles_by_size"))
print(ph); print(d["merge_loop"])
PY
