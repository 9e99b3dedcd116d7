import sys
sys.path.insert(0, '/root/repo/yet-another-bpe_b200'); sys.path.insert(0, '/root/repo/tools')
import torch, numpy as np
import yabpe
from yabpe import _ffi, engine
from synth_gpu import synth_corpus_device
torch = _ffi.require_cuda()
kind = sys.argv[1] if len(sys.argv) > 1 else "tinystories"
nbytes = int(sys.argv[2]) if len(sys.argv) > 2 else 256_000_000
text, n = synth_corpus_device(torch, nbytes, kind, 20260101)
for it in range(3):
    ev = []
    res = engine.pretok_count(torch, text, n, None, [b"<|endoftext|>"], 0, stage_events=ev)
    torch.cuda.synchronize()
    st = res.stats_host()
    print("iter", it, "specials %.3f ms tiles %.3f ms long %.3f ms" % (ev[0].elapsed_time(ev[1]), ev[1].elapsed_time(ev[2]), ev[2].elapsed_time(ev[3])),
          "GB/s %.1f" % (n / ev[1].elapsed_time(ev[2]) / 1e6), "ntok", st[0], "uniq", st[1], st[2], "cache-hit %.3f slow-items %d" % (st[10] / max(st[0], 1), st[9]))
