# round-1 final: whole GPU suite, the bench lines of the final build (both workloads, both arms), file-path timing
cd $GRAFT_REPO_ROOT
timeout 300 python -m pytest tests -m gpu -x -q > gpurun_out/r1g_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r1g_tests.log
timeout 200 python bench.py > gpurun_out/r1g_default.json 2> gpurun_out/r1g_default.err; echo "default rc=$?"
timeout 150 python bench.py --workload gpt2-encode-1g --steps 3 --warmup 3 > gpurun_out/r1g_encode.json 2> gpurun_out/r1g_encode.err; echo "encode rc=$?"
timeout 120 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r1g_reference.json 2> gpurun_out/r1g_reference.err; echo "ref rc=$?"
timeout 90 python bench.py --impl reference --workload gpt2-encode-1g --steps 2 --warmup 1 > gpurun_out/r1g_reference_encode.json 2> gpurun_out/r1g_reference_encode.err; echo "ref-encode rc=$?"
timeout 150 python tools/prof_file_train.py > gpurun_out/r1g_file_train.log 2>&1; echo "file rc=$?"; cat gpurun_out/r1g_file_train.log | tail -6
tail -c 300 gpurun_out/r1g_*.err
