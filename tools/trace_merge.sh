cd $GRAFT_REPO_ROOT
YABPE_NVCC_EXTRA="-DML_TRACE=${1:-4000}" python yet-another-bpe_b200/build.py --force > /dev/null 2>&1 || echo build failed
YABPE_TRACE=1 python bench.py --workload ${2:-tinystories-256m-v10k} --steps 1 --warmup 0 --skip-cpu --skip-e2e --encode-mb 0 2>&1 | grep trace
python yet-another-bpe_b200/build.py --force > /dev/null 2>&1
