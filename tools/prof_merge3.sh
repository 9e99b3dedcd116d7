#!/bin/bash
# source-level stall samples of k_merge_loop on a bench workload (two ncu sections only: few replay passes).
# usage: tools/prof_merge3.sh <workload> <tag>
cd $GRAFT_REPO_ROOT
WL=${1:-owt-1g-v32k}; TAG=${2:-merge}
ncu --section SourceCounters --section WarpStateStats --section SchedulerStats --clock-control none --import-source on -k regex:k_merge_loop -s 1 -c 1 -o /tmp/prof_$TAG -f python bench.py --workload $WL --steps 1 --warmup 1 --skip-cpu --skip-e2e --encode-mb 0 > gpurun_out/prof_${TAG}_ncu_run.log 2>&1
ncu -i /tmp/prof_$TAG.ncu-rep --page details > gpurun_out/prof_${TAG}_details.txt 2>/dev/null
ncu -i /tmp/prof_$TAG.ncu-rep --page source --print-source cuda,sass --csv > /tmp/prof_${TAG}_src.csv 2>/dev/null
python tools/ncu_lines.py /tmp/prof_${TAG}_src.csv 120 > gpurun_out/prof_${TAG}_lines.txt
cp /tmp/prof_${TAG}_src.csv gpurun_out/prof_${TAG}_src.csv; gzip -f gpurun_out/prof_${TAG}_src.csv
head -n 60 gpurun_out/prof_${TAG}_lines.txt
