"""First difference between the GPU merge list and the oracle's on a synthetic corpus (debugging aid).
usage: python tools/debug_merge_diff.py [kind] [bytes] [vocab]"""
import sys, os, tempfile
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path[:0] = [str(ROOT), str(ROOT / "yet-another-bpe_b200"), str(ROOT / "tests")]
import common
from oracle import oracle
import yabpe

kind = sys.argv[1] if len(sys.argv) > 1 else "tinystories"
size = int(sys.argv[2]) if len(sys.argv) > 2 else 3_000_000
vocab = int(sys.argv[3]) if len(sys.argv) > 3 else 2000
gen = {"tinystories": common.synth_tinystories, "owt": common.synth_owt, "adversarial": common.synth_adversarial}[kind]
data = gen(size)
p = Path(tempfile.mkdtemp()) / "c.txt"
p.write_bytes(data)
wv, wm = oracle.train_bpe(p, vocab, ["<|endoftext|>"], fast=True)
tr = yabpe.BBPETrainer(yabpe.BBPETrainerConfig(vocab_size=vocab, min_frequency=1, max_workers=1, chunk_size_bytes=1 << 30, special_tokens=["<|endoftext|>"]))
model = tr.train([p])
gm = model.merges
st = tr.last_stats
print("stats", st)
n = min(len(gm), len(wm))
first = next((i for i in range(n) if gm[i] != wm[i]), None)
print("merges", len(gm), len(wm), "first difference at", first)
if first is not None:
    for i in range(max(0, first - 3), min(n, first + 6)):
        print(i, "gpu", gm[i], "oracle", wm[i], "" if gm[i] == wm[i] else "<<<")
    pos = {mm: i for i, mm in enumerate(wm)}
    print("where the GPU's merges sit in the oracle's list:", [pos.get(gm[i], -1) for i in range(first, min(n, first + 12))])
