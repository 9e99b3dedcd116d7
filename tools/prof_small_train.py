"""Merge-loop phase breakdown on small samples of the OWT-shaped corpus (the like-for-like sizes of bench.py).
usage: python tools/prof_small_train.py [bytes ...]"""
import sys, time
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path[:0] = [str(ROOT), str(ROOT / "yet-another-bpe_b200"), str(ROOT / "tools")]
import torch
import yabpe
from synth_gpu import synth_corpus_device

torch.cuda.set_device(0)
sizes = [int(x) for x in sys.argv[1:]] or [65536, 2 << 20]
text, n_all = synth_corpus_device(torch, max(sizes) + 64, "owt", 20260102, piece_bytes=max(max(sizes), 1 << 20))
for nb in sizes:
    n = nb
    while n > 0 and (int(text[n].item()) & 0xC0) == 0x80:
        n -= 1
    sl = text[:((n + 15) // 16) * 16 + 64].clone(); sl[n:].zero_()
    cfg = yabpe.BBPETrainerConfig(vocab_size=32000, min_frequency=1, max_workers=1, chunk_size_bytes=1 << 30, special_tokens=["<|endoftext|>"])
    yabpe.BBPETrainer(cfg).train_device(sl, n)
    tr = yabpe.BBPETrainer(cfg); tr.profile = True
    torch.cuda.synchronize(); t0 = time.perf_counter()
    m = tr.train_device(sl, n)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    ph = tr.timing["merge_phase_cycles"]; tot = ph["total"] or 1; ms = tr.timing["merge_loop_ms"]
    print(nb, "bytes:", f"{dt * 1e3:.1f} ms wall, merge loop {ms:.1f} ms,", len(m.merges), "merges,",
          {k: (round(v / tot * ms, 2) if k != "n_top_rebuilds" else v) for k, v in ph.items() if not isinstance(v, list)},
          "leader", tr.last_stats.leader_merges, "grid", tr.last_stats.grid_merges, flush=True)
