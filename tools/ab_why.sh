#!/bin/bash
# why leader batches end (-DML_BATCH_WHY=1 build): tools/ab_why.sh <workload> <tag> [env assignments...]
wl=$1; tag=$2; shift 2
YABPE_NVCC_EXTRA="-DML_BATCH_WHY=1" python yet-another-bpe_b200/build.py --force > /dev/null 2>&1 || echo build failed
env "$@" YABPE_DUMP_STATE=1 timeout 400 python bench.py --workload $wl --skip-cpu --skip-e2e --steps 1 --warmup 0 --encode-mb 0 2>&1 | grep -a "state\[5\|ms_per_step" | cut -c 1-400
python yet-another-bpe_b200/build.py --force > /dev/null 2>&1
