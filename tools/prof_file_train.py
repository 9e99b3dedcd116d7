"""train(files) from a file on disk: whole-file read + pageable upload vs the streamed pinned-staging upload.
usage: python tools/prof_file_train.py [bytes] [dir]"""
import os, sys, time
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path[:0] = [str(ROOT), str(ROOT / "yet-another-bpe_b200"), str(ROOT / "tools")]
import torch
import yabpe
from synth_gpu import synth_corpus_device

nbytes = int(sys.argv[1]) if len(sys.argv) > 1 else 2_000_000_000
d = Path(sys.argv[2] if len(sys.argv) > 2 else ("/dev/shm" if os.path.isdir("/dev/shm") else "/tmp"))
dev, n = synth_corpus_device(torch, nbytes, "tinystories", 20260101)
path = d / "yabpe_prof_corpus.txt"
dev[:n].cpu().numpy().tofile(path)
del dev
torch.cuda.empty_cache()
cfg = yabpe.BBPETrainerConfig(vocab_size=10_000, min_frequency=1, max_workers=1, chunk_size_bytes=1 << 30, special_tokens=["<|endoftext|>"])
try:
    res = {}
    for name, min_bytes in (("whole-file", 1 << 62), ("streamed", 0), ("whole-file", 1 << 62), ("streamed", 0)):
        tr = yabpe.BBPETrainer(cfg)
        tr.stream_min_bytes = min_bytes
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        m = tr.train([path])
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        res.setdefault(name, []).append((dt, m.merges))
        print(f"{name:11s} train([{path}]) {n} bytes: {dt * 1e3:.1f} ms ({n / dt / 1e9:.2f} GB/s), {len(m.merges)} merges", flush=True)
    assert res["whole-file"][0][1] == res["streamed"][0][1] == res["streamed"][1][1], "merges differ between the two upload paths"
    print("merges identical")
finally:
    path.unlink(missing_ok=True)
