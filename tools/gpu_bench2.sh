cd $GRAFT_REPO_ROOT
summ() { python - "$1" <<'PY'
import json,sys
f=sys.argv[1]
try:
    d=json.loads(open(f).read().strip().splitlines()[-1])
    print(f, "value", d["value"], "ms", d["ms_per_step"], "us/merge", d["us_per_merge"], d["stage_ms"], d["merge_loop"], [d[k] for k in d if k.startswith("leader")], "e2e", d.get("e2e"))
except Exception as e: print(f, "ERR", e, open(f).read()[-2000:])
PY
}
timeout 600 python bench.py --steps 2 --warmup 1 --skip-cpu --skip-e2e > gpurun_out/b_ts.log 2>&1; summ gpurun_out/b_ts.log
if [ "$1" != "ts" ]; then timeout 600 python bench.py --workload owt-1g-v32k --steps 1 --warmup 1 --skip-cpu --skip-e2e > gpurun_out/b_owt.log 2>&1; summ gpurun_out/b_owt.log; fi
