"""Multi-GPU training: one process per GPU (torch.distributed, NCCL over NVLink / NVSwitch).

What shards: pre-tokenise + count, by BYTE RANGE of one corpus.  `train_files_sharded`
/ `train_range_sharded` give every rank the bytes [edge[r], edge[r+1] + halo) of the corpus, where the edges are
safe edges (yabpe/sharding.py: cutting there changes no pre-token) or reference chunk cuts (hard boundaries
anyway, trainer.py:172-198); the rank counts the pre-tokens that START in its range.  The result equals
single-GPU training on the same file(s) bit for bit.  (`train_device_sharded` is the older form where rank r's
text is FILE r of `BBPETrainer.train([f0, f1, ...])`, trainer.py:200-214.)

The unique (word, count) lists are hash-partitioned and packed by a kernel (yabpe_partition_words), exchanged
with ONE variable-size NCCL all-to-all, every rank merges the duplicates of its partition on the device
(yabpe_insert_words), and the partitions are gathered on rank 0, where the inherently sequential merge loop
runs (replicas would compute the same thing).  Payload is O(unique words): tens of MB.

The exchange logic is backend agnostic (torch tensors on any device) so that the gloo / CPU tests
cover it; `partition_words_cuda`, `count_local` and `reduce_packed_cuda` touch CUDA.
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
    order = torch.argsort(dest, stable=True)
    q = reorder(torch, p, order)
    d_sorted = dest[order]
    nw_to = torch.bincount(d_sorted, minlength=G).to(torch.int64)
    nb_to = torch.zeros(G, dtype=torch.int64, device=p.lens.device).index_add_(0, d_sorted, q.lens.to(torch.int64))
    return exchange_sorted(torch, dist, q, nw_to, nb_to)


def exchange_sorted(torch, dist, q: Packed, nw_to, nb_to) -> Packed:
    """all-to-all of a list already grouped by destination: nw_to[d] words / nb_to[d] bytes go to rank d."""
    dev = q.lens.device
    send_meta = torch.stack([nw_to, nb_to], 1).contiguous()
    recv_meta = torch.empty_like(send_meta)
    dist.all_to_all_single(recv_meta, send_meta)
    sm, rm = send_meta.cpu().numpy(), recv_meta.cpu().numpy()

    def a2a(x, col, dtype):
        out = torch.empty(int(rm[:, col].sum()), dtype=dtype, device=dev)
        dist.all_to_all_single(out, x.contiguous(), rm[:, col].tolist(), sm[:, col].tolist())
        return out

    return Packed(a2a(q.lens, 0, torch.int32), a2a(q.cnts, 0, torch.int64), a2a(q.data, 1, torch.uint8))


def gather_root(torch, dist, p: Packed) -> Packed | None:
    """Concatenation of every rank's list on rank 0 (None elsewhere): an all-to-all whose only destination is rank 0."""
    G = dist.get_world_size()
    dev = p.lens.device
    nw = torch.zeros(G, dtype=torch.int64, device=dev)
    nb = torch.zeros(G, dtype=torch.int64, device=dev)
    nw[0], nb[0] = p.lens.numel(), p.data.numel()
    root = exchange_sorted(torch, dist, p, nw, nb)
    return root if dist.get_rank() == 0 else None


def partition_torch(torch, p: Packed, G: int):
    """(list grouped by destination, words per destination, bytes per destination) with torch ops: the device-agnostic
    form the gloo tests run; the GPU path packs with yabpe_partition_words instead (partition_words_cuda)."""
    dest = word_hash(torch, p) % G
    order = torch.argsort(dest, stable=True)
    q = reorder(torch, p, order)
    d_sorted = dest[order]
    nw_to = torch.bincount(d_sorted, minlength=G).to(torch.int64)
    nb_to = torch.zeros(G, dtype=torch.int64, device=p.lens.device).index_add_(0, d_sorted, q.lens.to(torch.int64))
    return q, nw_to, nb_to


def shard_exchange(torch, dist, local, reduce_fn, partition_fn=None) -> Packed | None:
    """hash-partition -> all-to-all -> per-rank duplicate merge -> gather on rank 0 (None elsewhere).
    `local` is whatever `partition_fn(local, G)` understands (default: a Packed list, partitioned with torch ops)."""
    G = dist.get_world_size()
    q, nw_to, nb_to = (partition_fn or (lambda p, g: partition_torch(torch, p, g)))(local, G)
    mine = reduce_fn(exchange_sorted(torch, dist, q, nw_to, nb_to))
    return gather_root(torch, dist, mine)


# ------------------------------------------------------------------------------------ CUDA side
def packed_from_words(torch, words: engine.WordArrays) -> Packed:
    """Device word table -> packed list (bytes in slot order, i.e. words sorted by their first slot)."""
    W = words.n_words
    order = torch.argsort(words.woff[:W])
    return Packed(words.wlen[:W][order].contiguous(), words.wcnt[:W][order].contiguous(),
                  words.wsym[:words.n_syms].to(torch.uint8))


def partition_words_cuda(words: "engine.WordArrays | None", G: int):
    """yabpe_partition_words: the unique words of a fresh word table grouped by hash(bytes) mod G and packed for the
    all-to-all (one host sync for the per-destination totals, which the all-to-all needs anyway)."""
    torch = _ffi.require_cuda()
    L = _ffi.load()
    z = lambda n, dt: torch.zeros(n, dtype=dt, device="cuda")  # noqa: E731
    if words is None or words.n_words == 0:
        return Packed(z(0, torch.int32), z(0, torch.int64), z(0, torch.uint8)), z(G, torch.int64), z(G, torch.int64)
    W = words.n_words
    a = _ffi.PartitionArgs()
    a.words = words.table; a.n_words = W; a.n_ranks = G
    dest = torch.empty(W, dtype=torch.uint8, device="cuda")
    totals = z(G, torch.int64)
    a.dest, a.totals = dest.data_ptr(), totals.data_ptr()
    stream = _ffi.stream_ptr(torch)
    _ffi.check(L.yabpe_partition_words(C.byref(a), 0, stream))
    nw_to = totals & ((1 << 26) - 1)
    nb_to = totals >> 26
    base_w = (torch.cumsum(nw_to, 0) - nw_to).contiguous()
    base_b = (torch.cumsum(nb_to, 0) - nb_to).contiguous()
    cursor = z(G, torch.int64)
    out_lens = torch.empty(W, dtype=torch.int32, device="cuda")
    out_cnts = torch.empty(W, dtype=torch.int64, device="cuda")
    out_data = torch.empty(max(words.n_syms, 1), dtype=torch.uint8, device="cuda")
    a.base_w, a.base_b, a.cursor = base_w.data_ptr(), base_b.data_ptr(), cursor.data_ptr()
    a.out_lens, a.out_cnts, a.out_data = out_lens.data_ptr(), out_cnts.data_ptr(), out_data.data_ptr()
    _ffi.check(L.yabpe_partition_words(C.byref(a), 1, stream))
    p = Packed(out_lens, out_cnts, out_data[:words.n_syms])
    p._keep = (words, dest, totals, base_w, base_b, cursor)
    return p, nw_to, nb_to


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


def _count_exchange_merge(trainer, text_dev, n: int, cuts: list[int], own: tuple[int, int] | None, err_base: int,
                          describe_error):
    """Shared tail of the sharded trainers: local count of `text_dev[:n]` (pre-tokens starting in `own`), error
    agreement, exchange, merge loop on rank 0.  `err_base` turns a local error offset into a corpus offset;
    `describe_error(corpus_offset)` builds the ValueError message."""
    import torch.distributed as dist
    torch = _ffi.require_cuda()
    cfg = trainer.config
    rank = dist.get_rank()
    specials = [s.encode("utf-8") for s in cfg.special_tokens]
    ev = [] if trainer.profile else None
    own_hi = own[1] if own is not None else n
    res = st = None
    err = _ffi.INT64_MAX
    if n > 0 and own_hi > 0:
        cuts_np = np.asarray(cuts, dtype=np.int64) if cuts else None
        res = engine.pretok_count(torch, text_dev, n, cuts_np, specials, 0, own=own, stage_events=ev)
        st = res.stats_host()
        if st[_ffi.ST_TABLE_FULL] != 0:
            del res
            res, st = engine.pretok_count_checked(torch, text_dev, n, cuts_np, specials, 0, own=own)
        e = int(st[_ffi.ST_ERR_POS])
        if e != _ffi.INT64_MAX and e < own_hi:              # an ill-formed byte beyond the owned range is its owner's to report
            err = err_base + e                              # (the halo may end inside a code point)
    errs = [None] * dist.get_world_size()
    dist.all_gather_object(errs, err)
    if min(errs) != _ffi.INT64_MAX:
        raise ValueError(describe_error(min(errs)))
    words_local = engine.compact_words(torch, res, st, with_maps=False) if res is not None else None
    n_pretok = int(st[_ffi.ST_NTOK]) if st is not None else 0
    if trainer.profile:
        e0 = torch.cuda.Event(enable_timing=True); e0.record()
    root = shard_exchange(torch, dist, words_local, reduce_packed_cuda, partition_fn=partition_words_cuda)
    if trainer.profile:
        e1 = torch.cuda.Event(enable_timing=True); e1.record()
        torch.cuda.synchronize()
        trainer.timing["exchange_ms"] = e0.elapsed_time(e1)
        if ev and len(ev) == 4:
            trainer.timing.update(specials_ms=ev[0].elapsed_time(ev[1]), pretok_tiles_ms=ev[1].elapsed_time(ev[2]),
                                  long_tokens_ms=ev[2].elapsed_time(ev[3]))
    tot = torch.tensor([n_pretok], dtype=torch.int64, device="cuda")
    dist.all_reduce(tot)
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
    trainer.last_stats = TrainStats(n_bytes=n, n_pretokens=int(tot.item()), n_words=words.n_words, n_syms=words.n_syms,
                                    n_merges=len(mr.merge_new), index_rebuilds=int(mr.state[_ffi.MS_REBUILDS]),
                                    threshold_rebuilds=int(mr.state[_ffi.MS_TREBUILDS]), n_pairs=int(mr.state[_ffi.MS_NPAIRS]),
                                    leader_merges=int(mr.state[_ffi.MS_LEADER_MERGES]), grid_merges=int(mr.state[_ffi.MS_GRID_MERGES]))
    trainer.timing['leader_cycles'] = [int(x) for x in mr.state[20:29]] + [int(mr.state[12]), int(mr.state[13]), int(mr.state[17])]
    from .trainer import _phase_cycles
    trainer.timing['merge_phase_cycles'] = _phase_cycles(mr.state)
    _toks, vocab, merges = mr.materialise()
    return trainer._finish(vocab, merges)


def train_device_sharded(trainer, text_dev, n: int, name: str = "<shard>"):
    """Train on the union of every rank's shard, rank r's `text_dev[:n]` being FILE r of the corpus (each shard an
    independent text, trainer.py:200-214).  Returns the BBPEModel on rank 0 and None on the other ranks."""
    import torch.distributed as dist
    from .trainer import device_chunk_cuts
    cuts = device_chunk_cuts(text_dev, n, int(trainer.config.chunk_size_bytes)) if n > 0 else []
    rank = dist.get_rank()
    return _count_exchange_merge(trainer, text_dev, n, cuts, None, rank << 48,
                                 lambda e: f"File {name}[rank {e >> 48}] contains invalid UTF-8 at position {e & ((1 << 48) - 1)}.")


def train_range_sharded(trainer, local_dev, start: int, own_len: int, n_local: int, hard_cuts, file_starts=(0,), names=("<corpus>",)):
    """Byte-range form: this rank holds corpus bytes [start, start + n_local) in `local_dev` (padded like
    engine.to_device_text) and owns the pre-tokens that start in its first `own_len` bytes; `start` and
    `start + own_len` come from sharding.plan_shards.  `hard_cuts`: reference chunk cuts / file ends as CORPUS offsets
    (trainer.py:172-198).  Equals single-GPU training on the whole corpus.  BBPEModel on rank 0, None elsewhere."""
    cuts = sorted({int(c) - start for c in hard_cuts if start < int(c) < start + n_local})
    fs = list(file_starts)

    def describe(e: int) -> str:
        fi = int(np.searchsorted(np.asarray(fs), e, side="right")) - 1
        return f"File {names[fi]} contains invalid UTF-8 at position {e - fs[fi]}."      # trainer.py:157-160

    return _count_exchange_merge(trainer, local_dev, n_local, cuts, (0, own_len), start, describe)


def train_files_sharded(trainer, files):
    """`BBPETrainer.train(files)` on all ranks of the process group: every rank reads its byte range of the corpus
    (the concatenation of the files, file ends being hard cuts) straight into pinned memory and uploads it.  Same
    exceptions as the reference (trainer.py:72-73,204-205).  BBPEModel on rank 0, None elsewhere."""
    import torch.distributed as dist
    from pathlib import Path
    from . import sharding
    torch = _ffi.require_cuda()
    if not files:
        raise ValueError("At least one file must be provided")
    paths = [Path(f) if isinstance(f, str) else f for f in files]
    for p in paths:
        if not p.exists():
            raise FileNotFoundError(f"File not found: {p}")
    cat = sharding.FileConcat(paths, [p.stat().st_size for p in paths])
    if cat.total == 0:
        return trainer._finish(trainer._init_base_vocab(), []) if dist.get_rank() == 0 else None
    specials = [s.encode("utf-8") for s in trainer.config.special_tokens]
    hard = cat.hard_cuts(int(trainer.config.chunk_size_bytes))
    edges = sharding.plan_shards(cat.read, cat.total, dist.get_world_size(), specials, hard)
    start, own_len, n_local = sharding.shard_window(edges, dist.get_rank(), cat.total)
    host = torch.empty(max(n_local, 1), dtype=torch.uint8).pin_memory()
    if n_local:
        cat.readinto(start, start + n_local, host.numpy()[:n_local])
    local_dev, _ = engine.to_device_text(torch, host[:n_local])
    return train_range_sharded(trainer, local_dev, start, own_len, n_local, hard, cat.starts[:-1], [str(p) for p in paths])


def words_restore(torch, words: engine.WordArrays, p: Packed) -> None:
    """Undo the in-place rewrites of a failed merge-loop attempt (capacity retry)."""
    words.wsym[:words.n_syms] = p.data.to(torch.int32)
    words.wlen[:words.n_words] = p.lens


# ------------------------------------------------------------------------------------ document-sharded encode
def document_shard(tok, read, n: int, rank: int, world: int) -> tuple[int, int]:
    """Byte range [lo, hi) of rank `rank` for a document-sharded encode of an n-byte text: interior edges lie
    right AFTER an occurrence of the special token at or after r * n / world.  tokenizer.py:171-189 encodes the parts
    between specials independently, so concat(encode(shard_r)) == encode(text).  Exact when occurrences cannot
    overlap each other -- ONE special token no proper prefix of which is also a suffix (the same condition as
    BBPETokenizer._piece_ends); anything else, or a text without specials, stays on rank 0.
    `read(lo, hi) -> bytes` serves the few windows searched; every rank computes the same edges."""
    sps = tok._sp_bytes
    edges = [0]
    ok = len(sps) == 1 and not any(sps[0][:k] == sps[0][-k:] for k in range(1, len(sps[0])))
    for r in range(1, world):
        edge = n
        if ok:
            sp = sps[0]
            ideal = max((n * r) // world, edges[-1])
            lo, win = max(0, ideal - len(sp)), 1 << 16
            while lo < n:
                hi = min(n, lo + win)
                at = read(lo, hi).find(sp)
                if at >= 0:
                    edge = lo + at + len(sp)
                    break
                if hi >= n:
                    break
                lo, win = hi - len(sp) + 1, min(win * 4, 1 << 26)
        edges.append(max(edge, edges[-1]))
    edges.append(n)
    return edges[rank], edges[rank + 1]


def encode_sharded(tok, source, gather: bool = False, piece_bytes: int = 128 << 20):
    """`BBPETokenizer.encode` of one large text on all ranks of the process group, sharded by document.

    source: a file path, or a 1-D uint8 host tensor / numpy array every rank can see (e.g. a memory map).
    Every rank reads only its byte range (document_shard), encodes it with `encode_pinned` (H2D, encode and D2H of
    consecutive pieces overlapped) and returns (ids, offset, total): its ids as an int32 host tensor, the global
    index of its first id and the id count of the whole text.  There is no exchange on the data path; with
    gather=True rank 0 additionally receives every shard (NCCL send / recv through device staging) and returns
    all ids -- the same list `encode` of the whole text gives."""
    import torch.distributed as dist
    from pathlib import Path
    torch = _ffi.require_cuda()
    rank, world = dist.get_rank(), dist.get_world_size()
    if isinstance(source, (str, Path)):
        path = Path(source)
        n = path.stat().st_size

        def read(lo: int, hi: int) -> bytes:
            with open(path, "rb") as f:
                f.seek(lo)
                return f.read(hi - lo)

        lo, hi = document_shard(tok, read, n, rank, world)
        host = torch.empty(max(hi - lo, 1), dtype=torch.uint8).pin_memory()
        if hi > lo:
            with open(path, "rb", buffering=0) as f:
                f.seek(lo)
                mv, pos = memoryview(host.numpy()), 0
                while pos < hi - lo:
                    got = f.readinto(mv[pos:hi - lo])
                    if not got:
                        raise OSError(f"{path} shrank while it was read")
                    pos += got
        host = host[:hi - lo]
    else:
        arr = source if isinstance(source, np.ndarray) else source.numpy()
        n = int(arr.size)
        lo, hi = document_shard(tok, lambda a, b: arr[a:b].tobytes(), n, rank, world)
        host = torch.from_numpy(np.ascontiguousarray(arr[lo:hi]))
    ids = tok.encode_pinned(host, piece_bytes=piece_bytes)
    counts = [None] * world
    dist.all_gather_object(counts, int(ids.numel()))
    offset, total = sum(counts[:rank]), sum(counts)
    if not gather:
        return ids, offset, total
    if rank == 0:
        out = torch.empty(max(total, 1), dtype=torch.int32).pin_memory()
        out[:counts[0]].copy_(ids)
        pos = counts[0]
        for r in range(1, world):
            if counts[r]:
                buf = torch.empty(counts[r], dtype=torch.int32, device="cuda")
                dist.recv(buf, src=r)
                out[pos:pos + counts[r]].copy_(buf)
                pos += counts[r]
        return out[:total], 0, total
    if ids.numel():
        dist.send(ids.cuda(), dst=0)
    return ids, offset, total
