"""GPT-2 encode of an OWT-shaped text on the device, three times (the target of the ncu captures of k_encode_tiles).
usage: python tools/prof_encode.py [bytes]"""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path[:0] = [str(ROOT), str(ROOT / "yet-another-bpe_b200"), str(ROOT / "tools"), str(ROOT / "tests")]
import torch
import common, yabpe
from synth_gpu import synth_corpus_device

torch.cuda.set_device(0)
vocab, merges = common.gpt2_vocab_and_merges()
tok = yabpe.Tokenizer(vocab, merges, ["<|endoftext|>"]).inner
dev, n = synth_corpus_device(torch, int(sys.argv[1]) if len(sys.argv) > 1 else 256_000_000, "owt", 20260103)
tok.profile = True
for it in range(3):
    ids, _ = tok.encode_device(dev, n)
    print("iter", it, "bytes", n, "ids", int(ids.numel()), {k: round(v, 3) if isinstance(v, float) else v for k, v in tok.timing.items()})
