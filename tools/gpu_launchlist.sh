cd $GRAFT_REPO_ROOT
python bench.py --steps 2 --warmup 3 --skip-cpu --skip-e2e --encode-mb 0 > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'^k_|FillFunctor' -c 600 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --skip-cpu --skip-e2e --encode-mb 0 > gpurun_out/ncu_ll.log 2>&1
echo "rc=$?"; tail -c 200 gpurun_out/ncu_ll.log; wc -l gpurun_out/launches.csv
