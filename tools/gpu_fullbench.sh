cd $GRAFT_REPO_ROOT
python bench.py > gpurun_out/r1c_default.json 2> gpurun_out/r1c_default.err; echo "default rc=$?"
python bench.py --workload gpt2-encode-1g --steps 3 --warmup 2 > gpurun_out/r1c_encode.json 2> gpurun_out/r1c_encode.err; echo "encode rc=$?"
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r1c_reference.json 2> gpurun_out/r1c_reference.err; echo "ref rc=$?"
python bench.py --workload owt-11g-v32k --steps 1 --warmup 1 --skip-cpu > gpurun_out/r1c_owt11g.json 2> gpurun_out/r1c_owt11g.err; echo "owt rc=$?"
python bench.py --workload owt-1g-v32k --steps 2 --warmup 1 --skip-cpu > gpurun_out/r1c_owt1g.json 2> gpurun_out/r1c_owt1g.err; echo "owt1g rc=$?"
tail -c 300 gpurun_out/r1c_*.err
