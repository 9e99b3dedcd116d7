"""Host-side orchestration of the CUDA stages (allocation, sizing, retries).

PyTorch is used for device memory, streams and host<->device copies only; every byte of
compute runs in libyabpe.so.  Shared by the trainer (trainer.py) and the tokenizer
(tokenizer.py).
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass, field

import numpy as np

from . import _ffi

TOK_HASH_B = 0x100000001B3
MASK64 = (1 << 64) - 1
PT_TILE = 2048          # lower bound of the kernel's tile size (sizes the over-long token list)


def _pow2_at_least(x: int) -> int:
    return 1 << max(0, int(x - 1).bit_length())


def mix64(x: int) -> int:
    x &= MASK64
    x ^= x >> 33
    x = (x * 0xFF51AFD7ED558CCD) & MASK64
    x ^= x >> 33
    x = (x * 0xC4CEB9FE1A85EC53) & MASK64
    x ^= x >> 33
    return x


def pack_specials(specials: list[bytes]) -> tuple[np.ndarray, np.ndarray]:
    for s in specials:
        if len(s) == 0:
            raise ValueError("empty special tokens are not supported")
    blob = np.frombuffer(b"".join(specials) + b"\0", dtype=np.uint8).copy()
    offs = np.zeros(len(specials) + 1, dtype=np.int32)
    if specials:
        np.cumsum([len(s) for s in specials], out=offs[1:])
    return blob, offs


def to_device_text(torch, host: np.ndarray | "torch.Tensor", non_blocking: bool = False):
    """Copy `host` bytes into a padded device buffer (capacity round_up(n,16)+64, tail zeroed)."""
    if isinstance(host, np.ndarray):
        import warnings
        with warnings.catch_warnings():           # a read-only view (np.frombuffer of bytes) is only ever read here
            warnings.simplefilter("ignore", UserWarning)
            host = torch.from_numpy(host)
    n = int(host.numel())
    cap = ((n + 15) // 16) * 16 + 64
    dev = torch.empty(cap, dtype=torch.uint8, device="cuda")
    dev[n:].zero_()
    if n:
        dev[:n].copy_(host, non_blocking=non_blocking)
    return dev, n


class Mailbox:
    """Pinned host words that a kernel writes directly (yabpe_publish, mapped memory under unified addressing).
    Reading a few device counters this way is stream-ordered like `.cpu()` but does not use a copy engine, so it is
    not held up by bulk transfers running on other streams (BBPETokenizer.encode_pinned)."""

    def __init__(self, torch, n_words: int = 64) -> None:
        self.torch = torch
        self.buf = torch.zeros(n_words, dtype=torch.int64).pin_memory()
        self.event = torch.cuda.Event()

    def read(self, src, n_words: int) -> np.ndarray:
        """The first `n_words` int64 words of the device tensor `src`, once the current stream has reached this point."""
        assert src.dtype == self.torch.int64 and src.is_cuda and n_words <= self.buf.numel()
        _ffi.check(_ffi.load().yabpe_publish(self.buf.data_ptr(), src.data_ptr(), n_words, _ffi.stream_ptr(self.torch)))
        self.event.record()
        self.event.synchronize()
        return self.buf[:n_words].numpy().copy()


@dataclass
class PretokResult:
    args: _ffi.PretokArgs
    keep: list = field(default_factory=list)     # tensors / arrays that must outlive the async calls
    stats: "object" = None
    short_cap: int = 0
    long_cap: int = 0
    text: "object" = None
    n: int = 0
    mailbox: "Mailbox | None" = None
    cold: bool = False                           # short-table layout chosen (interleaved slots)
    hot: "object" = None                         # hot set handed to the warp kernel's cache

    def stats_host(self) -> np.ndarray:
        if self.mailbox is not None:
            return self.mailbox.read(self.stats, 16)
        return self.stats.cpu().numpy()

    def sizing(self) -> tuple:
        """(short_cap, long_cap, cold, hot): lets the next, similar text skip the sizing sample."""
        return (self.short_cap, self.long_cap, self.cold, self.hot)


def pretok_count(torch, text_dev, n: int, cuts: np.ndarray | None, specials: list[bytes], mode: int,
                 own: tuple[int, int] | None = None, short_cap: int | None = None,
                 long_cap: int | None = None, stage_events: list | None = None,
                 generic_only: bool = False, mailbox: Mailbox | None = None, sizing: tuple | None = None,
                 launch: bool = True, max_cuts: int = 0) -> PretokResult:
    """Launch special resolution + the tile kernel + the long-token kernel (all async).
    `sizing` = PretokResult.sizing() of an earlier, similar text: its layout and hot set are reused and its capacities
    are the default, so no sizing sample is counted.  `mailbox`: statistics reach the host through yabpe_publish."""
    L = _ffi.load()
    dev = text_dev.device
    cold_hint, hot = False, None
    if sizing is not None:
        short_cap, long_cap = short_cap or sizing[0], long_cap or sizing[1]
        cold_hint, hot = sizing[2], sizing[3]
    elif short_cap is None or long_cap is None:
        est_s, est_l, cold_hint, hot = estimate_table_sizes(torch, text_dev, n, cuts, specials, mode, mailbox)
        short_cap = short_cap or est_s
        long_cap = long_cap or est_l
    if os.environ.get("YABPE_SHORT_CAP_LOG2") and n > 4 * _SAMPLE_BYTES:          # experiments only: force the big table's size
        short_cap = 1 << int(os.environ["YABPE_SHORT_CAP_LOG2"])
    n_cuts = 0 if cuts is None else int(len(cuts))
    cuts_t = torch.from_numpy(np.ascontiguousarray(cuts, dtype=np.int64)).to(dev) if n_cuts and launch else None
    blob, offs = pack_specials(specials)
    words32 = (n + 63) // 32 + 1
    cand = torch.zeros(words32, dtype=torch.int32, device=dev) if specials else None
    rec = torch.zeros(words32, dtype=torch.int32, device=dev) if specials else None
    # small tables (hot, L2-resident): keys and counts apart, probes never queue behind count atomics;
    # large tables (DRAM-resident): 32-byte slots {key, count, -}, one sector per probe + update
    layout = os.environ.get("YABPE_SHORT_LAYOUT")              # tests force either layout on small inputs
    interleaved = layout == "interleaved" if layout else cold_hint
    skeys = torch.zeros(short_cap * (4 if interleaved else 2), dtype=torch.int64, device=dev)
    scounts = None if interleaved else torch.zeros(short_cap, dtype=torch.int64, device=dev)
    lent = torch.zeros(long_cap * 4, dtype=torch.int64, device=dev)
    ovf_cap = n // 992 + 16         # at most one over-long pre-token per 992-byte chunk of the warp kernel
    ovf = torch.empty(ovf_cap, dtype=torch.int64, device=dev)
    work_cap = 4 * max(n_cuts, max_cuts) + 64
    work = torch.empty(3 * work_cap, dtype=torch.int64, device=dev)
    if mailbox is not None:                      # no host->device copy either: it would queue behind a bulk upload
        stats = torch.zeros(16, dtype=torch.int64, device=dev)
        stats[_ffi.ST_ERR_POS:_ffi.ST_ERR_POS + 1].fill_(_ffi.INT64_MAX)
    else:
        stats_np = np.zeros(16, dtype=np.int64)
        stats_np[_ffi.ST_ERR_POS] = _ffi.INT64_MAX
        stats = torch.from_numpy(stats_np).to(dev)
    a = _ffi.PretokArgs()
    a.text = text_dev.data_ptr(); a.n = n
    a.cuts = cuts_t.data_ptr() if cuts_t is not None else None; a.n_cuts = n_cuts if cuts_t is not None else 0; a.mode = mode
    a.sp_blob = blob.ctypes.data; a.sp_offs = offs.ctypes.data; a.n_sp = len(specials)
    a.own_lo, a.own_hi = own if own is not None else (0, n)
    a.cand_bits = cand.data_ptr() if specials else None
    a.rec_bits = rec.data_ptr() if specials else None
    a.short_keys = skeys.data_ptr(); a.short_counts = None if interleaved else scounts.data_ptr(); a.short_cap = short_cap
    a.long_entries = lent.data_ptr(); a.long_cap = long_cap
    a.ovf_pos = ovf.data_ptr(); a.ovf_cap = ovf_cap
    a.stats = stats.data_ptr()
    a.work = work.data_ptr(); a.work_cap = work_cap
    a.hot_keys = hot.data_ptr() if hot is not None else None
    # YABPE_HOT_TABLE=1: a DRAM-sized table (interleaved layout) gets an L2-resident table in front of it (2^20 slots of 32
    # bytes).  Exact (tests force it), but measured SLOWER on the OWT-shaped corpus (264 -> 248 GB/s): off
    hot_tab = None
    if interleaved and short_cap > _HOT_TABLE_SLOTS * _HOT_TABLE_MIN_FACTOR and os.environ.get("YABPE_HOT_TABLE", "0") == "1":
        hot_tab = torch.zeros(_HOT_TABLE_SLOTS * 4, dtype=torch.int64, device=dev)
        a.hot_table = hot_tab.data_ptr(); a.hot_cap = _HOT_TABLE_SLOTS
    res = PretokResult(args=a, keep=[text_dev, cuts_t, blob, offs, cand, rec, skeys, scounts, lent, ovf, work, hot, hot_tab], stats=stats,
                       short_cap=short_cap, long_cap=long_cap, text=text_dev, n=n, mailbox=mailbox, cold=interleaved, hot=hot)
    if n > 0 and launch:
        extra = 8 if generic_only else 0        # stages bit 3: generic tile kernel only (A/B parity tests)
        a.stages = extra
        if stage_events is None:
            _ffi.check(L.yabpe_pretok_count(C.byref(a), _ffi.stream_ptr(torch)))
        else:
            # one call per stage with a CUDA event after each (bench: per-kernel durations)
            ev = torch.cuda.Event(enable_timing=True); ev.record(); stage_events.append(ev)
            for bit in (1, 2, 4):
                a.stages = bit | extra
                _ffi.check(L.yabpe_pretok_count(C.byref(a), _ffi.stream_ptr(torch)))
                ev = torch.cuda.Event(enable_timing=True); ev.record(); stage_events.append(ev)
            a.stages = 0
    return res


def pretok_count_pieces(torch, text_dev, n: int, pieces: list[tuple[int, int, int, list[int]]], specials: list[bytes],
                        ready: list, sample_cuts: np.ndarray | None = None) -> tuple[PretokResult, np.ndarray] | None:
    """Count a text WHILE it is being uploaded: piece k = (own_lo, own_hi, n_k, cuts_k) is counted as soon as `ready[k]`
    (a CUDA event: bytes [0, n_k) are on the device) has fired, all pieces into ONE set of tables.  own_lo is 0 or one of
    cuts_k (a text start), n_k the end of the bytes the piece may look at -- sharding.plan_shards edges with their halo,
    exactly the per-rank windows of the multi-GPU path, laid out in one buffer.  Nothing here uses a copy engine (the
    upload owns it): counters travel through the Mailbox, the cut lists are uploaded once, up front.
    Returns None when the tables overflowed (the caller recounts the resident text in one go)."""
    L = _ffi.load()
    mailbox = Mailbox(torch)
    all_cuts = np.asarray([c for p in pieces for c in p[3]] + [0], dtype=np.int64)
    cuts_dev = torch.from_numpy(all_cuts).to(text_dev.device)               # before the bulk upload is enqueued
    torch.cuda.current_stream().wait_event(ready[0])
    sizing = None
    if n > 4 * _SAMPLE_BYTES and pieces[0][2] >= _SAMPLE_BYTES + 64:
        sizing = estimate_table_sizes(torch, text_dev, n, sample_cuts, specials, 0, mailbox)
    res = pretok_count(torch, text_dev, n, None, specials, 0, mailbox=mailbox, sizing=sizing, launch=False,
                       max_cuts=max(len(p[3]) for p in pieces))
    res.keep.append(cuts_dev)
    a = res.args
    off = 0
    stream = _ffi.stream_ptr(torch)
    for k, (lo, hi, nk, cuts_k) in enumerate(pieces):
        torch.cuda.current_stream().wait_event(ready[k])
        a.n, a.own_lo, a.own_hi = nk, lo, hi
        a.cuts = cuts_dev.data_ptr() + 8 * off if cuts_k else None
        a.n_cuts = len(cuts_k)
        off += len(cuts_k)
        a.stages = 16 | 7
        if k:
            res.stats[_ffi.ST_OVF_N:_ffi.ST_OVF_N + 1].zero_()               # per-call lists: over-long pre-tokens, boundary work items
            res.stats[_ffi.ST_SLOW_N:_ffi.ST_SLOW_N + 1].zero_()
        _ffi.check(L.yabpe_pretok_count(C.byref(a), stream))
    a.n, a.own_lo, a.own_hi, a.cuts, a.n_cuts, a.stages = n, 0, n, None, 0, 0
    st = res.stats_host()
    if st[_ffi.ST_TABLE_FULL] != 0:
        return None
    return res, st


_SAMPLE_BYTES = 16 << 20
_HOT_TABLE_SLOTS = int(os.environ.get("YABPE_HOT_TABLE_SLOTS", str(1 << 20)))
_HOT_TABLE_MIN_FACTOR = 4          # the big table must be at least this many times the hot one (else it is L2-sized itself)


def estimate_table_sizes(torch, text_dev, n: int, cuts, specials, mode, mailbox: Mailbox | None = None):
    """Hash-table capacities.  Small inputs: proportional to n.  Large inputs: count the unique
    pre-tokens of a 16 MB prefix and extrapolate (Heaps' law, exponent 0.75), so that a 2 GB corpus
    with 10^5 word types does not zero and scan GB-sized tables.  Overflow is detected and retried.
    The third value picks the short-table layout: interleaved {key, count} slots when the sample says the table
    will be large and cold (more than 5 % of the sample's pre-tokens are first occurrences), split arrays for a
    small table that every SM hammers (see ShortTab in pretok.cuh).
    Fourth value: the hot set of the sample for the warp kernel's shared-memory cache (device tensor or None)."""
    if n <= 4 * _SAMPLE_BYTES:
        return (_pow2_at_least(min(max(n // 4, 1 << 12), 1 << 26)), _pow2_at_least(min(max(n // 32, 1 << 8), 1 << 24)), False, None)
    m = _SAMPLE_BYTES
    if mailbox is not None:                      # the four bytes at the sample's end in one mailbox read
        tail = mailbox.read(text_dev[m - 8:m + 8].view(torch.int64), 2).view(np.uint8)
        while m > _SAMPLE_BYTES - 4 and (int(tail[8 + m - _SAMPLE_BYTES]) & 0xC0) == 0x80:
            m -= 1
    else:                                        # one 8-byte read; a code point has at most 3 continuation bytes, so the
        tail = text_dev[m - 4:m + 4].cpu().numpy()       # walk is bounded even on input that is not UTF-8 (the kernel reports it)
        while m > _SAMPLE_BYTES - 4 and (int(tail[4 + m - _SAMPLE_BYTES]) & 0xC0) == 0x80:
            m -= 1
    sub_cuts = None if cuts is None else np.asarray([c for c in cuts if 0 < c < m], dtype=np.int64)
    res = pretok_count(torch, text_dev, m, sub_cuts if sub_cuts is not None and len(sub_cuts) else None, specials, mode,
                       short_cap=1 << 22, long_cap=1 << 20, mailbox=mailbox)     # the prefix as a text of its own: an estimate only
    st = res.stats_host()
    scale = (n / m) ** 0.75
    us = max(int(st[_ffi.ST_UNIQ_SHORT]), 1 << 10) * scale
    ul = max(int(st[_ffi.ST_UNIQ_LONG]), 1 << 6) * scale
    if st[_ffi.ST_TABLE_FULL] != 0:
        us, ul = 1 << 24, 1 << 22
    cold = int(st[_ffi.ST_UNIQ_SHORT]) > 0.05 * max(int(st[_ffi.ST_NTOK]), 1)
    L = _ffi.load()
    nc = int(L.yabpe_hot_cache_entries())
    hot = torch.empty(2 * nc, dtype=torch.int64, device=text_dev.device)
    scratch = torch.empty(nc, dtype=torch.int64, device=text_dev.device)
    _ffi.check(L.yabpe_select_hot(C.byref(res.args), hot.data_ptr(), scratch.data_ptr(), _ffi.stream_ptr(torch)))
    hot._yabpe_keep = (res, scratch)              # the sample tables must outlive the (stream-ordered) selection
    # Short table: twice the estimate (load factor 25 - 50 %), not four times: probes into a 2 GB table cost more than the
    # longer probe chains of a 1 GB one (OWT 11 GB: 2^26 slots 41.1 ms, 2^25 38.1 ms, 2^24 36.4 ms); a table that fills up
    # is detected and the count repeated with a larger one.
    return (_pow2_at_least(int(min(max(2 * us, 1 << 16), 1 << 26))), _pow2_at_least(int(min(max(4 * ul, 1 << 12), 1 << 24))), cold, hot)


def pretok_count_checked(torch, text_dev, n, cuts, specials, mode, own=None,
                         generic_only: bool = False, mailbox: Mailbox | None = None,
                         sizing: tuple | None = None) -> tuple[PretokResult, np.ndarray]:
    """pretok_count + one host sync; grows the tables and retries when they overflow."""
    short_cap = long_cap = None
    for _ in range(8):
        res = pretok_count(torch, text_dev, n, cuts, specials, mode, own, short_cap, long_cap, generic_only=generic_only,
                           mailbox=mailbox, sizing=sizing)
        st = res.stats_host()
        if st[_ffi.ST_TABLE_FULL] == 0:
            return res, st
        short_cap, long_cap = res.short_cap * 4, res.long_cap * 4
        del res
    raise _ffi.YabpeError("pre-token hash tables kept overflowing")


def token_starts(torch, res: PretokResult) -> np.ndarray:
    """Byte offsets of every pre-token start of the text `res` was counted on, ascending (yabpe_token_starts)."""
    n = int(res.args.n)
    bits = torch.empty((n + 31) // 32, dtype=torch.int32, device=res.text.device)
    _ffi.check(_ffi.load().yabpe_token_starts(C.byref(res.args), bits.data_ptr(), _ffi.stream_ptr(torch)))
    words = bits.cpu().numpy().view(np.uint8)
    return np.flatnonzero(np.unpackbits(words, bitorder="little")[:n]).astype(np.int64)


@dataclass
class WordArrays:
    table: _ffi.WordTable
    n_words: int
    n_syms: int
    keep: list = field(default_factory=list)
    wsym: "object" = None
    woff: "object" = None
    wlen: "object" = None
    wcnt: "object" = None


def compact_words(torch, res: PretokResult, st: np.ndarray, with_maps: bool) -> WordArrays:
    L = _ffi.load()
    dev = res.text.device
    n_words = int(st[_ffi.ST_UNIQ_SHORT] + st[_ffi.ST_UNIQ_LONG])
    n_syms = int(st[_ffi.ST_UNIQ_BYTES])
    wsym = torch.empty(n_syms + 8, dtype=torch.int32, device=dev)
    sym_word = torch.empty(n_syms + 8, dtype=torch.int32, device=dev)
    woff = torch.empty(n_words + 1, dtype=torch.int64, device=dev)
    wlen = torch.empty(n_words + 1, dtype=torch.int32, device=dev)
    wcnt = torch.empty(n_words + 1, dtype=torch.int64, device=dev)
    sword = torch.empty(res.short_cap, dtype=torch.int32, device=dev) if with_maps else None
    lword = torch.empty(res.long_cap, dtype=torch.int32, device=dev) if with_maps else None
    counters = torch.zeros(2, dtype=torch.int64, device=dev)
    w = _ffi.WordTable()
    w.wsym = wsym.data_ptr(); w.sym_word = sym_word.data_ptr(); w.woff = woff.data_ptr()
    w.wlen = wlen.data_ptr(); w.wcnt = wcnt.data_ptr()
    w.sword = sword.data_ptr() if with_maps else None
    w.lword = lword.data_ptr() if with_maps else None
    w.counters = counters.data_ptr()
    if n_words > 0:
        _ffi.check(L.yabpe_compact_words(C.byref(res.args), C.byref(w), _ffi.stream_ptr(torch)))
    return WordArrays(table=w, n_words=n_words, n_syms=n_syms,
                      keep=[wsym, sym_word, woff, wlen, wcnt, sword, lword, counters],
                      wsym=wsym, woff=woff, wlen=wlen, wcnt=wcnt)


def tok_hash(b: bytes) -> tuple[int, int]:
    """Polynomial hash of a token's bytes and base^len (must match merge.cuh phase 2)."""
    h, p = 0, 1
    for x in b:
        h = (h * TOK_HASH_B + x + 1) & MASK64
        p = (p * TOK_HASH_B) & MASK64
    return h, p


def _load_hostlist():
    """yabpe/_hostlist.so (csrc/hostlist.c, built by build.py with gcc): result objects built in C; None when it is not there."""
    try:
        from . import _hostlist
        return _hostlist
    except ImportError:
        return None


_HOSTLIST = _load_hostlist()


@dataclass
class MergeResult:
    merges: np.ndarray          # (n_merges, 2) int32 token ids
    merge_new: np.ndarray       # (n_merges,) int32 resulting id
    state: np.ndarray
    pool: bytes                 # the token bytes, back to back
    offs: np.ndarray            # int64 [n_tokens + 1]: token i = pool[offs[i]:offs[i + 1]]
    _tokens: list | None = None

    @property
    def tokens(self) -> list[bytes]:
        """id -> bytes, all tokens."""
        if self._tokens is None:
            self._tokens = self.materialise()[0]
        return self._tokens

    def materialise(self) -> tuple[list[bytes], dict[bytes, int], list[tuple[bytes, bytes]]]:
        """(tokens, vocab, merges) as the Python objects the reference API returns (trainer.py:94-134, 296-300): ~65 000
        small objects for a 32 000-merge model -- built in C when yabpe/_hostlist.so is there (a tenth of the training
        step otherwise), else with C-level loops over plain lists."""
        offs = np.ascontiguousarray(self.offs, dtype=np.int64)
        mg = np.ascontiguousarray(self.merges, dtype=np.int32).reshape(-1, 2)
        if _HOSTLIST is not None:
            toks, vocab, merges = _HOSTLIST.materialise(self.pool, offs, mg)
        else:
            o = offs.tolist()
            toks = list(map(self.pool.__getitem__, map(slice, o[:-1], o[1:])))
            vocab = dict(zip(toks, range(len(toks))))
            tok_at = toks.__getitem__
            merges = list(zip(map(tok_at, mg[:, 0].tolist()), map(tok_at, mg[:, 1].tolist())))
        self._tokens = toks
        return toks, vocab, merges


def rebuild_period(n_syms: int) -> int:
    """Merges between rebuilds of the pair -> words index.  A rebuild costs O(symbols + table) (0.08 ms on the
    2 GB TinyStories-shaped corpus, 1.8 ms on the 11 GB OWT-shaped one), staleness costs candidates that turn out not to
    contain the pair; measured optimum ~1 500 merges for the former.  With batched merges the big corpora want fewer rebuilds
    than before (OWT 11 GB: 6 000 -> 165 ms of merge loop, 9 000 -> 162, 12 000 -> 159; the tie-heavy 1 GB corpus 293 -> 289 ms
    a step at 12 000), but no period at all -- rebuild on demand only -- costs that 1 GB corpus 10 %."""
    env = os.environ.get("YABPE_REBUILD_EVERY")
    if env is not None:
        return int(env)
    return int(min(max(40.0 * float(max(n_syms, 1)) ** 0.3, 500.0), 12000.0))


def merge_loop(torch, words: WordArrays, base_tokens: list[bytes], num_merges: int, min_frequency: int,
               pcap: int | None = None, pool_cap: int | None = None,
               restore=None, timing: dict | None = None) -> MergeResult:
    """Run the persistent merge kernel.  `restore()` must rebuild `words` in place for a retry."""
    L = _ffi.load()
    dev = words.wsym.device
    n_base = len(base_tokens)
    max_tokens = n_base + num_merges + 2
    for attempt in range(6):
        if pcap is None:
            # distinct pairs stay well below the symbol count (0.07-0.2 x on the bench corpora): half the symbol count keeps
            # the load factor below 40 % and halves what every rebuild scans; the kernel reports a table that fills beyond
            # 3/4 and the loop below retries with a four times larger one
            pcap = _pow2_at_least(min(max(words.n_syms >> int(os.environ.get("YABPE_PCAP_SHIFT", "1")), 1 << 16), 1 << 26))
        if pool_cap is None:
            pool_cap = (4 << 20) + 32 * max_tokens + min(words.n_syms, 1 << 30)
        alog_cap = max(2 * words.n_words, 1 << 16) + 4096
        tset_cap = _pow2_at_least(4 * max_tokens)
        # base tokens on the host
        tok_bytes = np.zeros(sum(len(b) for b in base_tokens) + 16, dtype=np.uint8)      # only the used prefix is uploaded
        tok_off = np.zeros(max_tokens + 1, dtype=np.int64)
        th = np.zeros(max_tokens, dtype=np.uint64)
        tp = np.zeros(max_tokens, dtype=np.uint64)
        tset = np.zeros(tset_cap, dtype=np.uint64)
        pos = 0
        for i, b in enumerate(base_tokens):
            tok_bytes[pos:pos + len(b)] = np.frombuffer(b, dtype=np.uint8)
            pos += len(b)
            tok_off[i + 1] = pos
            h, p = tok_hash(b)
            th[i], tp[i] = h, p
            slot = mix64(h) & (tset_cap - 1)
            while tset[slot] != 0:
                slot = (slot + 1) & (tset_cap - 1)
            tset[slot] = (h & 0xFFFFFFFF00000000) | (i + 1)
        t = lambda a: torch.from_numpy(a).to(dev)  # noqa: E731
        d_tok_bytes = torch.zeros(pool_cap, dtype=torch.uint8, device=dev)
        d_tok_bytes[:len(tok_bytes)].copy_(torch.from_numpy(tok_bytes))
        d_tok_off, d_th, d_tp, d_tset = t(tok_off), t(th.view(np.int64)), t(tp.view(np.int64)), t(tset.view(np.int64))
        z = lambda n, dt: torch.zeros(n, dtype=dt, device=dev)  # noqa: E731
        wstamp = z(words.n_words + 1, torch.int32)
        wslot = z(words.n_syms + 8, torch.int32)
        newp = torch.empty(words.n_syms + 8, dtype=torch.int32, device=dev)
        pkey, pcnt = z(pcap, torch.int64), z(pcap, torch.int64)
        ioff, icnt = z(pcap + 1, torch.int32), z(pcap, torch.int32)
        ipost = z(words.n_syms + 8, torch.int64)
        inact, act = z((pcap + 31) // 32 + 1, torch.int32), z(pcap, torch.int32)
        intop = z((pcap + 31) // 32 + 1, torch.int32)
        top_slot, top_key, hist = z(1024, torch.int32), z(1024, torch.int64), z(1024, torch.int32)
        alog_word = z(alog_cap, torch.int64)
        nm1 = max(num_merges, 1)
        seg_start, seg_end, merge_next = z(nm1, torch.int32), z(nm1, torch.int32), z(nm1, torch.int32)
        tok_first = z(max_tokens, torch.int32)
        tok_head = z(4 * max_tokens + 4, torch.int32)
        partial, bsum = z(1024 * 3, torch.int64), z(1024, torch.int64)
        merges, merge_new = z(2 * max(num_merges, 1), torch.int32), z(max(num_merges, 1), torch.int32)
        state_np = np.zeros(64, dtype=np.int64)
        state_np[_ffi.MS_NTOK] = n_base
        state = t(state_np)
        m = _ffi.MergeArgs()
        m.words = words.table; m.n_words = words.n_words; m.n_syms = words.n_syms
        m.wstamp = wstamp.data_ptr(); m.wslot = wslot.data_ptr(); m.newp = newp.data_ptr()
        m.tok_bytes = d_tok_bytes.data_ptr(); m.tok_bytes_cap = pool_cap
        m.tok_off = d_tok_off.data_ptr(); m.tok_hash = d_th.data_ptr(); m.tok_pow = d_tp.data_ptr()
        tok_pre = z(max_tokens, torch.int64); m.tok_pre = tok_pre.data_ptr()
        m.tset = d_tset.data_ptr(); m.tset_cap = tset_cap; m.max_tokens = max_tokens
        m.pkey = pkey.data_ptr(); m.pcnt = pcnt.data_ptr(); m.pcap = pcap
        m.ioff = ioff.data_ptr(); m.icnt = icnt.data_ptr(); m.ipost = ipost.data_ptr()
        m.inact = inact.data_ptr(); m.intop = intop.data_ptr(); m.act = act.data_ptr()
        m.top_slot = top_slot.data_ptr(); m.top_key = top_key.data_ptr(); m.hist = hist.data_ptr()
        m.alog_word = alog_word.data_ptr(); m.alog_cap = alog_cap
        m.seg_start = seg_start.data_ptr(); m.seg_end = seg_end.data_ptr()
        m.merge_next = merge_next.data_ptr(); m.tok_first = tok_first.data_ptr(); m.tok_head = tok_head.data_ptr()
        m.partial = partial.data_ptr(); m.bsum = bsum.data_ptr()
        m.merges = merges.data_ptr(); m.merge_new = merge_new.data_ptr(); m.state = state.data_ptr()
        m.num_merges = num_merges; m.min_frequency = min_frequency
        m.rebuild_every = rebuild_period(words.n_syms)
        m.helper_mode = int(os.environ.get("YABPE_HELPER_MODE", "0"))
        m.batch_max = int(os.environ.get("YABPE_BATCH_MAX", "0"))                  # 1: one merge per iteration (A/B runs, tests)
        m.helper_min_syms = int(os.environ.get("YABPE_HELPER_MIN_SYMS", "0"))     # tests force the prefetch helpers on small inputs (-1)
        if timing is not None:
            t0 = torch.cuda.Event(enable_timing=True); t0.record()
        _ffi.check(L.yabpe_merge_loop(C.byref(m), _ffi.stream_ptr(torch)))
        if timing is not None:
            t1 = torch.cuda.Event(enable_timing=True); t1.record()
        st = state.cpu().numpy()
        if os.environ.get("YABPE_DUMP_STATE"):        # tuning builds (-DML_BATCH_WHY): raw counters
            print("state[37:40]", st[37:40].tolist(), "state[54:64]", st[54:64].tolist(), flush=True)
        import os as _os
        if _os.environ.get('YABPE_TRACE'):           # only meaningful with a -DML_TRACE=<merge> build (tools/trace_merge.sh)
            tr = bsum[512:512 + 8 * 24].cpu().numpy().reshape(8, 24)
            for row in tr:
                base = row[0]
                print('[trace]', [int(x - base) for x in row[:10]], 'g8', [int(x - base) for x in row[13:16]], 'items', int(row[10]), 'rewritten', int(row[11]), 'new pairs', int(row[12]), flush=True)
        if timing is not None:
            timing["merge_loop_ms"] = t0.elapsed_time(t1)
        err = int(st[_ffi.MS_ERROR])
        if err == 0:
            nm = int(st[_ffi.MS_NMERGES])
            ntok = int(st[_ffi.MS_NTOK])
            mg = merges[:2 * nm].cpu().numpy().reshape(-1, 2)
            mn = merge_new[:nm].cpu().numpy()
            used = int(st[_ffi.MS_POOL_USED]) if ntok > n_base else int(tok_off[n_base])
            pool = d_tok_bytes[:max(used, 1)].cpu().numpy().tobytes()
            offs = d_tok_off[:ntok + 1].cpu().numpy()
            return MergeResult(merges=mg, merge_new=mn, state=st, pool=pool, offs=offs)
        if err & _ffi.ME_INTERNAL:
            raise _ffi.YabpeError(f"merge loop internal error (state={st.tolist()})")
        if restore is None:
            raise _ffi.YabpeError(f"merge loop capacity error {err} and no restore callback")
        if err & _ffi.ME_PAIR_TABLE_FULL:
            pcap *= 4
        if err & _ffi.ME_TOK_POOL_FULL:
            pool_cap *= 4
        restore()
    raise _ffi.YabpeError("merge loop kept running out of capacity")
