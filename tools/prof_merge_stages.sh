cd $GRAFT_REPO_ROOT
YABPE_NVCC_EXTRA="-DML_TIMING=1" python yet-another-bpe_b200/build.py --force > /dev/null 2>&1 || echo build failed
python bench.py --workload owt-1g-v32k --steps 1 --warmup 1 --skip-e2e --skip-cpu --encode-mb 0 > gpurun_out/r2_t6_timing_owt1g.json 2> gpurun_out/r2_t6_timing.err
python bench.py --workload tinystories-2g-v10k --steps 1 --warmup 1 --skip-e2e --skip-cpu --encode-mb 0 > gpurun_out/r2_t6_timing_ts.json 2>> gpurun_out/r2_t6_timing.err
for m in 3000 12000 25000; do
YABPE_NVCC_EXTRA="-DML_TRACE=$m" python yet-another-bpe_b200/build.py --force > /dev/null 2>&1 || echo build failed
YABPE_TRACE=1 python bench.py --workload owt-1g-v32k --steps 1 --warmup 0 --skip-cpu --skip-e2e --encode-mb 0 2>&1 | grep trace > gpurun_out/r2_t6_trace_owt1g_$m.txt
done
python yet-another-bpe_b200/build.py --force > /dev/null 2>&1
grep -o '"leader_cycles[^]]*]' gpurun_out/r2_t6_timing_*.json
