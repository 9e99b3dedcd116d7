# usage: variants.sh "<flags1>" "<flags2>" ...   (rebuilds libyabpe.so per variant ON THE GPU BOX and times the pre-tokeniser)
cd $GRAFT_REPO_ROOT
for v in "$@"; do
  echo "=== variant: $v"
  YABPE_NVCC_EXTRA="$v" python yet-another-bpe_b200/build.py --force > /dev/null 2>&1 || { echo build failed; continue; }
  timeout 300 python tools/prof_pretok.py tinystories 1000000000 2>&1 | tail -1
  timeout 300 python tools/prof_pretok.py owt 1000000000 2>&1 | tail -1
done
python yet-another-bpe_b200/build.py --force > /dev/null 2>&1
