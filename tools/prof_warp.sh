cd $GRAFT_REPO_ROOT
K=${1:-tinystories}
python tools/prof_pretok.py $K 256000000 > gpurun_out/prof_plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:k_pretok_warp -s 1 -c 1 -o gpurun_out/prof_warp_$K -f python tools/prof_pretok.py $K 256000000 > gpurun_out/ncu_run.log 2>&1
cat gpurun_out/prof_plain.log; tail -3 gpurun_out/ncu_run.log
