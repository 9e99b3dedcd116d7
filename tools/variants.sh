set -e
cd $GRAFT_REPO_ROOT
for v in "-DPT_TILE=8192 -DPT_CACHE_N=2048 -DPT_MIN_BLOCKS=2" "-DPT_TILE=8192 -DPT_CACHE_N=1024 -DPT_MIN_BLOCKS=3" "-DPT_TILE=4096 -DPT_CACHE_N=2048 -DPT_MIN_BLOCKS=3" "-DPT_TILE=4096 -DPT_CACHE_N=1024 -DPT_MIN_BLOCKS=4" "-DPT_TILE=4096 -DPT_CACHE_N=1024 -DPT_MIN_BLOCKS=5" "-DPT_TILE=4096 -DPT_CACHE_N=512 -DPT_MIN_BLOCKS=5"; do
  YABPE_NVCC_EXTRA="$v" python yet-another-bpe_b200/build.py --force > /dev/null
  echo "== $v"
  python tools/prof_pretok.py tinystories 256000000 2>&1 | tail -1
  python tools/prof_pretok.py owt 256000000 2>&1 | tail -1
done
