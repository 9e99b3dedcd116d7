# A/B of compile-time variants of the pre-tokeniser on the GPU box: usage variants_pretok.sh "<flags A>" "<flags B>" ...
cd $GRAFT_REPO_ROOT
for flags in "$@"; do
  YABPE_NVCC_EXTRA="$flags" python yet-another-bpe_b200/build.py --force > /dev/null 2>&1 || { echo "build failed: $flags"; continue; }
  echo "== [$flags]"
  python tools/prof_pretok.py owt 2000000000 | tail -n 1
  python tools/prof_pretok.py tinystories 1000000000 | tail -n 1
done
python yet-another-bpe_b200/build.py --force > /dev/null 2>&1
