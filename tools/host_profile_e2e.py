"""Host-side profile of the end-to-end path (pinned host bytes -> train_from_buffers)."""
import cProfile, pstats, sys, io, time
sys.path.insert(0, '/root/repo/yet-another-bpe_b200'); sys.path.insert(0, '/root/repo/tools')
import torch
import yabpe
from synth_gpu import synth_corpus_device
text, n = synth_corpus_device(torch, 2_000_000_000, "tinystories", 20260101)
host = torch.empty(n, dtype=torch.uint8).pin_memory(); host.copy_(text[:n]); host_np = host.numpy()
del text; torch.cuda.empty_cache()
cfg = yabpe.BBPETrainerConfig(vocab_size=10000, min_frequency=1, max_workers=1, chunk_size_bytes=1 << 30, special_tokens=["<|endoftext|>"])
for _ in range(2):
    yabpe.BBPETrainer(cfg).train_from_buffers([host_np])
torch.cuda.synchronize()
pr = cProfile.Profile()
t0 = time.perf_counter()
pr.enable()
m = yabpe.BBPETrainer(cfg).train_from_buffers([host_np])
torch.cuda.synchronize()
pr.disable()
print("wall ms", 1e3 * (time.perf_counter() - t0))
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(16)
print(s.getvalue()[:4000])
