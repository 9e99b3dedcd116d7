cd $GRAFT_REPO_ROOT
timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "train or merge or adversarial or rebuild" 2>&1 | tail -2
timeout 600 python bench.py --workload owt-1g-v32k --steps 2 --warmup 1 --skip-cpu --skip-e2e > gpurun_out/b_owt.log 2>&1
python - <<'PY'
import json
d=json.loads(open("gpurun_out/b_owt.log").read().strip().splitlines()[-1])
print("owt ms", d["ms_per_step"], "us/merge", d["us_per_merge"], d["stage_ms"], d["merge_loop"])
PY
