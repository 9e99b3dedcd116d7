"""Where encode_pinned's time goes: per-piece device time and host time of encode_device, the bare copies, the whole call.
usage: python tools/prof_pinned.py [piece_mb ...]"""
import sys, time
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path[:0] = [str(ROOT), str(ROOT / "yet-another-bpe_b200"), str(ROOT / "tools"), str(ROOT / "tests")]
import torch
import common, yabpe
from synth_gpu import synth_corpus_device

torch.cuda.set_device(0)
vocab, merges = common.gpt2_vocab_and_merges()
tok = yabpe.Tokenizer(vocab, merges, ["<|endoftext|>"]).inner
dev, n = synth_corpus_device(torch, 1_000_000_000, "owt", 20260103)
host = torch.empty(n, dtype=torch.uint8).pin_memory(); host.copy_(dev[:n])
ids, _ = tok.encode_device(dev, n)
n_ids = ids.numel()
out = torch.empty(n_ids, dtype=torch.int32).pin_memory()
torch.cuda.synchronize()

def wall(f, reps=3):
    f(); torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps): f()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps * 1e3

print(f"n {n} ids {n_ids}")
print("H2D only ms", round(wall(lambda: dev[:n].copy_(host, non_blocking=True)), 2))
print("D2H only ms", round(wall(lambda: out.copy_(ids, non_blocking=True)), 2))
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def both():
    with torch.cuda.stream(s1): dev[:n].copy_(host, non_blocking=True)
    with torch.cuda.stream(s2): out.copy_(ids, non_blocking=True)
print("H2D + D2H on two streams ms", round(wall(both), 2))
print("encode_device whole ms", round(wall(lambda: tok.encode_device(dev, n)), 2))
del ids
real = tok.encode_device
for piece_mb in [int(x) for x in sys.argv[1:]] or [128, 256]:
    rows = []
    def traced(text_dev, ln, *a, **k):
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter(); e0.record()
        r = real(text_dev, ln, *a, **k)
        e1.record(); rows.append((ln, e0, e1, (time.perf_counter() - t0) * 1e3))
        return r
    tok.encode_device = real
    w = wall(lambda: tok.encode_pinned(host, out=out, piece_bytes=piece_mb << 20))
    tok.encode_device = traced
    t0 = time.perf_counter(); tok.encode_pinned(host, out=out, piece_bytes=piece_mb << 20); torch.cuda.synchronize()
    w1 = (time.perf_counter() - t0) * 1e3
    print(f"piece {piece_mb} MiB: encode_pinned {w:.2f} ms (traced run {w1:.2f}); per piece (MB, device ms, host ms):",
          [(ln >> 20, round(a.elapsed_time(b), 2), round(h, 2)) for ln, a, b, h in rows])
