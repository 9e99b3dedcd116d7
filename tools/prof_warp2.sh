# usage: prof_warp2.sh <kind> <bytes> <tag>   -- ncu --set full of k_pretok_warp on a synthetic corpus; text summaries only
cd $GRAFT_REPO_ROOT
K=${1:-owt}; N=${2:-2000000000}; TAG=${3:-$K}
python tools/prof_pretok.py $K $N > gpurun_out/prof_${TAG}_plain.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:k_pretok_warp -s 1 -c 1 -o /tmp/prof_$TAG -f python tools/prof_pretok.py $K $N > gpurun_out/prof_${TAG}_ncu_run.log 2>&1
ncu -i /tmp/prof_$TAG.ncu-rep --page details > gpurun_out/prof_${TAG}_details.txt 2>/dev/null
ncu -i /tmp/prof_$TAG.ncu-rep --page raw --csv > gpurun_out/prof_${TAG}_raw.csv 2>/dev/null
ncu -i /tmp/prof_$TAG.ncu-rep --page source --print-source cuda,sass --csv > /tmp/prof_${TAG}_src.csv 2>/dev/null
python tools/ncu_lines.py /tmp/prof_${TAG}_src.csv 60 > gpurun_out/prof_${TAG}_lines.txt
tail -n 2 gpurun_out/prof_${TAG}_plain.log
