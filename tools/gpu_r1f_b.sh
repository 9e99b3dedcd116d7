# round-1 final, call B: ncu --set full of the encode tile passes, then the whole GPU test suite with durations
cd $GRAFT_REPO_ROOT
python tools/prof_encode.py 256000000 > gpurun_out/prof_encode_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_encode_tiles -s 2 -c 2 -o gpurun_out/r1_ncu_encode_tiles -f python tools/prof_encode.py 256000000 > gpurun_out/ncu_encode_run.log 2>&1
echo "ncu rc=$?"; cat gpurun_out/prof_encode_plain.log; tail -3 gpurun_out/ncu_encode_run.log
timeout 420 python -m pytest tests -m gpu -x -q --durations=12 > gpurun_out/r1f_tests.log 2>&1; echo "tests rc=$?"; tail -22 gpurun_out/r1f_tests.log
