# round-1 final, call A: the pipelined host-buffer encode (test + bench, pipelined and serial end-to-end)
cd $GRAFT_REPO_ROOT
timeout 150 python -m pytest tests/test_gpu_parity.py -x -q -k "encode_pinned or encode_iterable" > gpurun_out/r1f_test_pinned.log 2>&1; echo "test rc=$?"; tail -15 gpurun_out/r1f_test_pinned.log
timeout 120 python bench.py --workload gpt2-encode-1g --steps 3 --warmup 2 --cpu-sample-mb 16 > gpurun_out/r1f_encode.json 2> gpurun_out/r1f_encode.err; echo "encode rc=$?"; tail -c 600 gpurun_out/r1f_encode.err
timeout 90 python bench.py --workload gpt2-encode-1g --steps 3 --warmup 2 --skip-cpu --encode-e2e serial > gpurun_out/r1f_encode_serial.json 2> gpurun_out/r1f_encode_serial.err; echo "serial rc=$?"
timeout 90 python bench.py --workload gpt2-encode-1g --steps 3 --warmup 2 --skip-cpu --piece-mb 64 > gpurun_out/r1f_encode_p64.json 2> gpurun_out/r1f_encode_p64.err; echo "p64 rc=$?"
python - <<'PY'
import json
for f in ("r1f_encode", "r1f_encode_serial", "r1f_encode_p64"):
    try:
        d = json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1])
        print(f, "value", d["value"], "e2e", d["e2e"], "cpu", d["cpu_baseline"])
    except Exception as e:
        print(f, "ERR", e)
PY
