# usage: variants_merge.sh "<flags1>" ... : rebuild per variant on the GPU box, run the TinyStories bench (merge loop timing)
cd $GRAFT_REPO_ROOT
for v in "$@"; do
  echo "=== variant: $v"
  YABPE_NVCC_EXTRA="$v" python yet-another-bpe_b200/build.py --force > /dev/null 2>&1 || { echo build failed; continue; }
  bash tools/gpu_bench2.sh ${ONLY:-ts}
done
python yet-another-bpe_b200/build.py --force > /dev/null 2>&1
