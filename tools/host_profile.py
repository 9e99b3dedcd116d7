"""Where does the host spend time in one train_device step?  (cProfile, device-resident corpus)"""
import cProfile, pstats, sys, io
sys.path.insert(0, '/root/repo/yet-another-bpe_b200'); sys.path.insert(0, '/root/repo/tools')
import torch
import yabpe
from synth_gpu import synth_corpus_device
kind = sys.argv[1] if len(sys.argv) > 1 else "tinystories"
nbytes = int(sys.argv[2]) if len(sys.argv) > 2 else 2_000_000_000
vocab = int(sys.argv[3]) if len(sys.argv) > 3 else 10_000
text, n = synth_corpus_device(torch, nbytes, kind, 20260101)
cfg = yabpe.BBPETrainerConfig(vocab_size=vocab, min_frequency=1, max_workers=1, chunk_size_bytes=1 << 30, special_tokens=["<|endoftext|>"])
for _ in range(2):
    yabpe.BBPETrainer(cfg).train_device(text, n)
torch.cuda.synchronize()
pr = cProfile.Profile()
import time
t0 = time.perf_counter()
pr.enable()
m = yabpe.BBPETrainer(cfg).train_device(text, n)
torch.cuda.synchronize()
pr.disable()
print("wall ms", 1e3 * (time.perf_counter() - t0))
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(28)
print(s.getvalue()[:6000])
