cd $GRAFT_REPO_ROOT
python bench.py --workload tinystories-256m-v10k --steps 1 --warmup 1 --skip-cpu --skip-e2e > gpurun_out/prof_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k_merge_loop -s 1 -c 1 -o gpurun_out/prof_merge -f python bench.py --workload tinystories-256m-v10k --steps 1 --warmup 1 --skip-cpu --skip-e2e > gpurun_out/ncu_run.log 2>&1
tail -c 300 gpurun_out/prof_plain.log; tail -3 gpurun_out/ncu_run.log
