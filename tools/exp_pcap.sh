cd $GRAFT_REPO_ROOT
for f in 4 1 0.5 0.25; do echo "== pcap factor $f"; YABPE_PCAP_FACTOR=$f bash tools/gpu_bench2.sh ts; done
