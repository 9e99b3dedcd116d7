"""Python-list API of the tokenizer ("Python-list API timed separately"): `encode(str) -> list[int]` latency
from 10 bytes to 64 MB, `encode_batch` of short lines, `decode(list[int])`, with the CPU port of the reference next to each figure.
Short strings run as ONE launch (yabpe_encode_small, up to 32 KiB); the sweep also has sizes between the decades around 1 KB.
usage: python tools/bench_latency.py [max_bytes]"""
import sys, time
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path[:0] = [str(ROOT), str(ROOT / "yet-another-bpe_b200"), str(ROOT / "tests")]
import torch
import common, yabpe
from oracle import oracle
from yabpe import _ffi

torch.cuda.set_device(0)
vocab, merges = common.gpt2_vocab_and_merges()
tok = yabpe.Tokenizer(vocab, merges, ["<|endoftext|>"])
otok = oracle.Tokenizer(vocab, merges, ["<|endoftext|>"])
max_bytes = int(sys.argv[1]) if len(sys.argv) > 1 else 64 << 20
text = common.synth_owt(max_bytes, seed=20260103).decode("utf-8")


def timed(f, reps):
    f()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        r = f()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps, r


print(f"{'chars':>10} {'gpu encode ms':>14} {'launches':>9} {'cpu port ms':>12} {'equal':>6}")
sizes = sorted({11} | {10 ** k for k in range(1, 9)} | {30, 300, 500, 2000, 4000, 8000, 16000, 30000, 40000})
for n in sizes:
    if n > len(text):
        break
    s = text[:n] if n != 11 else "hello world"
    reps = 200 if n <= 10_000 else 20 if n <= 1_000_000 else 2
    l0 = _ffi.launch_count()
    dt, ids = timed(lambda: tok.encode(s), reps)
    launches = (_ffi.launch_count() - l0) // (reps + 1)
    t0 = time.perf_counter()
    want = otok.encode(s)
    dc = time.perf_counter() - t0
    print(f"{n:>10} {dt * 1e3:>14.3f} {launches:>9} {dc * 1e3:>12.3f} {str(ids == want):>6}", flush=True)
lines = text[:4_000_000].split("\n")
dt, out = timed(lambda: tok.encode_batch(lines), 3)
dt1, _ = timed(lambda: [tok.encode(x) for x in lines[:300]], 1)
print(f"encode_batch of {len(lines)} lines ({sum(map(len, lines))} chars): {dt * 1e3:.1f} ms; one encode() per line: "
      f"{dt1 / 300 * 1e3:.3f} ms per line -> {dt1 / 300 * len(lines) * 1e3:.0f} ms for all of them")
ids = tok.encode(text[:8_000_000])
for k in sorted({min(k, len(ids)) for k in (10, 1000, 100_000, len(ids))}):
    dt, s = timed(lambda: tok.decode(ids[:k]), 20 if k <= 100_000 else 2)
    print(f"decode of {k} ids: {dt * 1e3:.3f} ms")
