"""Line-level wall-clock of BBPETrainer._train_on_device (sys.setprofile-free: a tracing wrapper on line events)."""
import sys, time, collections
sys.path.insert(0, '/root/repo/yet-another-bpe_b200'); sys.path.insert(0, '/root/repo/tools')
import torch
import yabpe
from yabpe import trainer as T
from synth_gpu import synth_corpus_device
kind = sys.argv[1] if len(sys.argv) > 1 else "owt"
nbytes = int(sys.argv[2]) if len(sys.argv) > 2 else 11_000_000_000
vocab = int(sys.argv[3]) if len(sys.argv) > 3 else 32_000
text, n = synth_corpus_device(torch, nbytes, kind, 20260102)
cfg = yabpe.BBPETrainerConfig(vocab_size=vocab, min_frequency=1, max_workers=1, chunk_size_bytes=1 << 30, special_tokens=["<|endoftext|>"])
yabpe.BBPETrainer(cfg).train_device(text, n)
torch.cuda.synchronize()
acc = collections.defaultdict(float)
last = [None, 0.0]
code = T.BBPETrainer._train_on_device.__code__
def tracer(frame, event, arg):
    if frame.f_code is not code:
        return None
    now = time.perf_counter()
    if last[0] is not None:
        acc[last[0]] += now - last[1]
    last[0], last[1] = frame.f_lineno, now
    return tracer
sys.settrace(tracer)
t0 = time.perf_counter()
m = yabpe.BBPETrainer(cfg).train_device(text, n)
sys.settrace(None)
print("wall ms", 1e3 * (time.perf_counter() - t0))
src = open(T.__file__).read().splitlines()
for ln, t in sorted(acc.items(), key=lambda x: -x[1])[:12]:
    print(f"{1e3*t:8.2f} ms  line {ln}: {src[ln-1].strip()[:110]}")
