"""Multi-GPU training: one process per GPU (torch.distributed, NCCL over NVLink / NVSwitch).

What shards: pre-tokenise + count.  Every rank counts the pre-tokens of its own
shard (an independent text, exactly like one file of `BBPETrainer.train([f0, f1, ...])`,
trainer.py:200-214), the unique (word, count) lists are hash-partitioned and exchanged with ONE
all-to-all, every rank merges the duplicates of its partition on the device, and the partitions are
gathered on rank 0, where the inherently sequential merge loop runs (replicas would compute the
same thing).  Payload is O(unique words): tens of MB, latency- not bandwidth-bound.

The exchange logic is backend agnostic (torch tensors on any device) so that the gloo / CPU tests
cover it; only `count_local` and `reduce_packed` touch CUDA.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np

from . import _ffi, engine


@dataclass
class Packed:
    """A list of byte strings with counts: lens[i] bytes of `data` each, in order."""
    lens: "object"      # int32 [W]
    cnts: "object"      # int64 [W]
    data: "object"      # uint8 [sum(lens)]


def word_hash(torch, p: Packed):
    """Position-mixed additive hash of every word (int64 wrap-around arithmetic, device agnostic)."""
    W = p.lens.numel()
    dev = p.lens.device
    if W == 0:
        return torch.zeros(0, dtype=torch.int64, device=dev)
    lens64 = p.lens.to(torch.int64)
    start = torch.cumsum(lens64, 0) - lens64
    wid = torch.repeat_interleave(torch.arange(W, device=dev), lens64)
    pos = torch.arange(p.data.numel(), device=dev) - start[wid]
    term = (p.data.to(torch.int64) + 1) * (pos * 0x9E3779B1 + 0x7F4A7C15)
    h = torch.zeros(W, dtype=torch.int64, device=dev).index_add_(0, wid, term)
    h = h ^ (h >> 29)
    h = h * 0x2545F4914F6CDD1D
    return (h >> 20) & 0x7FFFFFFF


def reorder(torch, p: Packed, order) -> Packed:
    """Words of `p` in the given order (gathers the ragged byte ranges)."""
    lens64 = p.lens.to(torch.int64)
    start = torch.cumsum(lens64, 0) - lens64
    nl = lens64[order]
    nstart = torch.cumsum(nl, 0) - nl
    wid = torch.repeat_interleave(torch.arange(order.numel(), device=order.device), nl)
    idx = torch.arange(int(nl.sum().item()), device=order.device) - nstart[wid] + start[order][wid]
    return Packed(p.lens[order], p.cnts[order], p.data[idx])


def exchange(torch, dist, p: Packed, dest) -> Packed:
    """Send word i to rank dest[i]; returns everything this rank received (all-to-all, variable sizes)."""
    G = dist.get_world_size()
    dev = p.lens.device
    order = torch.argsort(dest, stable=True)
    q = reorder(torch, p, order)
    d_sorted = dest[order]
    nw_to = torch.bincount(d_sorted, minlength=G).to(torch.int64)
    nb_to = torch.zeros(G, dtype=torch.int64, device=dev).index_add_(0, d_sorted, q.lens.to(torch.int64))
    send_meta = torch.stack([nw_to, nb_to], 1).contiguous()
    recv_meta = torch.empty_like(send_meta)
    dist.all_to_all_single(recv_meta, send_meta)
    sm, rm = send_meta.cpu().numpy(), recv_meta.cpu().numpy()

    def a2a(x, col, dtype):
        out = torch.empty(int(rm[:, col].sum()), dtype=dtype, device=dev)
        dist.all_to_all_single(out, x.contiguous(), rm[:, col].tolist(), sm[:, col].tolist())
        return out

    return Packed(a2a(q.lens, 0, torch.int32), a2a(q.cnts, 0, torch.int64), a2a(q.data, 1, torch.uint8))


def shard_exchange(torch, dist, local: Packed, reduce_fn) -> Packed | None:
    """hash-partition -> all-to-all -> per-rank duplicate merge -> gather on rank 0 (None elsewhere)."""
    G = dist.get_world_size()
    dest = word_hash(torch, local) % G
    mine = reduce_fn(exchange(torch, dist, local, dest))
    root = exchange(torch, dist, mine, torch.zeros(mine.lens.numel(), dtype=torch.int64, device=mine.lens.device))
    return root if dist.get_rank() == 0 else None


# ------------------------------------------------------------------------------------ CUDA side
def packed_from_words(torch, words: engine.WordArrays) -> Packed:
    """Device word table -> packed list (bytes in slot order, i.e. words sorted by their first slot)."""
    W = words.n_words
    order = torch.argsort(words.woff[:W])
    return Packed(words.wlen[:W][order].contiguous(), words.wcnt[:W][order].contiguous(),
                  words.wsym[:words.n_syms].to(torch.uint8))


def reduce_packed_cuda(p: Packed) -> Packed:
    """Merge duplicate words of a packed list on the device (yabpe_insert_words + compaction)."""
    torch = _ffi.require_cuda()
    L = _ffi.load()
    W = int(p.lens.numel())
    if W == 0:
        return p
    nbytes = int(p.data.numel())
    blob = torch.zeros(((nbytes + 15) // 16) * 16 + 64, dtype=torch.uint8, device="cuda")
    blob[:nbytes] = p.data
    lens64 = p.lens.to(torch.int64)
    offs = (torch.cumsum(lens64, 0) - lens64).contiguous()
    has_long = int(bool((p.lens > 256).any().item()))
    short_cap = engine._pow2_at_least(max(4 * W, 1 << 12))
    long_cap = engine._pow2_at_least(max(4 * W, 1 << 8))
    for _ in range(4):
        res = engine.pretok_count(torch, blob, 0, None, [], 0, short_cap=short_cap, long_cap=long_cap)   # n=0: allocates only
        res.args.n = max(nbytes, 1)
        res.args.own_hi = max(nbytes, 1)
        _ffi.check(L.yabpe_insert_words(C.byref(res.args), offs.data_ptr(), p.lens.data_ptr(), p.cnts.data_ptr(), W,
                                        has_long, _ffi.stream_ptr(torch)))
        st = res.stats_host()
        if st[_ffi.ST_TABLE_FULL] == 0:
            break
        short_cap, long_cap = short_cap * 4, long_cap * 4
    else:
        raise _ffi.YabpeError("exchange tables kept overflowing")
    res.text = blob
    words = engine.compact_words(torch, res, st, with_maps=False)
    return packed_from_words(torch, words)


def words_from_packed(torch, p: Packed) -> engine.WordArrays:
    """Packed list (already unique) -> the flat word arrays the merge loop consumes."""
    W = int(p.lens.numel())
    n_syms = int(p.data.numel())
    dev = p.lens.device
    lens64 = p.lens.to(torch.int64)
    wsym = torch.zeros(n_syms + 8, dtype=torch.int32, device=dev)
    wsym[:n_syms] = p.data.to(torch.int32)
    sym_word = torch.zeros(n_syms + 8, dtype=torch.int32, device=dev)
    sym_word[:n_syms] = torch.repeat_interleave(torch.arange(W, dtype=torch.int32, device=dev), lens64)
    woff = torch.zeros(W + 1, dtype=torch.int64, device=dev)
    woff[:W] = torch.cumsum(lens64, 0) - lens64
    wlen = torch.zeros(W + 1, dtype=torch.int32, device=dev)
    wlen[:W] = p.lens
    wcnt = torch.zeros(W + 1, dtype=torch.int64, device=dev)
    wcnt[:W] = p.cnts
    counters = torch.tensor([W, n_syms], dtype=torch.int64, device=dev)
    t = _ffi.WordTable()
    t.wsym, t.sym_word, t.woff, t.wlen, t.wcnt = (x.data_ptr() for x in (wsym, sym_word, woff, wlen, wcnt))
    t.sword = None; t.lword = None; t.counters = counters.data_ptr()
    return engine.WordArrays(table=t, n_words=W, n_syms=n_syms, keep=[wsym, sym_word, woff, wlen, wcnt, counters],
                             wsym=wsym, woff=woff, wlen=wlen, wcnt=wcnt)


def train_device_sharded(trainer, text_dev, n: int, name: str = "<shard>"):
    """Train on the union of every rank's shard (rank r's `text_dev[:n]` is file r of the corpus).
    Returns the BBPEModel on rank 0 and None on the other ranks."""
    import torch.distributed as dist
    torch = _ffi.require_cuda()
    cfg = trainer.config
    rank = dist.get_rank()
    specials = [s.encode("utf-8") for s in cfg.special_tokens]
    # P1 chunk cuts inside this rank's shard (each shard is its own file)
    from .trainer import device_chunk_cuts
    cuts = device_chunk_cuts(text_dev, n, int(cfg.chunk_size_bytes))
    ev = [] if trainer.profile else None
    if n > 0:
        res = engine.pretok_count(torch, text_dev, n, np.asarray(cuts, dtype=np.int64) if cuts else None, specials, 0,
                                  stage_events=ev)
        st = res.stats_host()
        if st[_ffi.ST_TABLE_FULL] != 0:
            del res
            res, st = engine.pretok_count_checked(torch, text_dev, n, np.asarray(cuts, dtype=np.int64) if cuts else None, specials, 0)
        err = int(st[_ffi.ST_ERR_POS])
    else:
        err = _ffi.INT64_MAX
    errs = [None] * dist.get_world_size()
    dist.all_gather_object(errs, err)
    for r, e in enumerate(errs):
        if e != _ffi.INT64_MAX:
            raise ValueError(f"File {name}[rank {r}] contains invalid UTF-8 at position {e}.")
    if n > 0:
        local = packed_from_words(torch, engine.compact_words(torch, res, st, with_maps=False))
        n_pretok = int(st[_ffi.ST_NTOK])
    else:
        local = Packed(torch.zeros(0, dtype=torch.int32, device="cuda"), torch.zeros(0, dtype=torch.int64, device="cuda"),
                       torch.zeros(0, dtype=torch.uint8, device="cuda"))
        n_pretok = 0
    if trainer.profile:
        e0 = torch.cuda.Event(enable_timing=True); e0.record()
    root = shard_exchange(torch, dist, local, reduce_packed_cuda)
    if trainer.profile:
        e1 = torch.cuda.Event(enable_timing=True); e1.record()
        torch.cuda.synchronize()
        trainer.timing["exchange_ms"] = e0.elapsed_time(e1)
        if ev and len(ev) == 4:
            trainer.timing.update(specials_ms=ev[0].elapsed_time(ev[1]), pretok_tiles_ms=ev[1].elapsed_time(ev[2]),
                                  long_tokens_ms=ev[2].elapsed_time(ev[3]))
    if rank != 0:
        return None
    base_vocab = trainer._init_base_vocab()
    num_merges = max(0, cfg.vocab_size - len(base_vocab))
    if root.lens.numel() == 0 or num_merges == 0:
        return trainer._finish(base_vocab, [])
    words = words_from_packed(torch, root)
    mr = engine.merge_loop(torch, words, list(base_vocab.keys()), num_merges, int(cfg.min_frequency),
                           restore=lambda: words_restore(torch, words, root), timing=trainer.timing if trainer.profile else None)
    from .trainer import TrainStats
    trainer.last_stats = TrainStats(n_bytes=n, n_pretokens=n_pretok, n_words=words.n_words, n_syms=words.n_syms,
                                    n_merges=len(mr.merge_new), index_rebuilds=int(mr.state[_ffi.MS_REBUILDS]),
                                    threshold_rebuilds=int(mr.state[_ffi.MS_TREBUILDS]), n_pairs=int(mr.state[_ffi.MS_NPAIRS]),
                                    leader_merges=int(mr.state[_ffi.MS_LEADER_MERGES]), grid_merges=int(mr.state[_ffi.MS_GRID_MERGES]))
    vocab = {b: i for i, b in enumerate(mr.tokens)}
    toks = mr.tokens
    merges = [(toks[a], toks[b]) for a, b in mr.merges.tolist()]
    return trainer._finish(vocab, merges)


def words_restore(torch, words: engine.WordArrays, p: Packed) -> None:
    """Undo the in-place rewrites of a failed merge-loop attempt (capacity retry)."""
    words.wsym[:words.n_syms] = p.data.to(torch.int32)
    words.wlen[:words.n_words] = p.lens
