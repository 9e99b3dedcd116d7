"""B200 tokenizer: the reference's `BBPETokenizer` interface over the CUDA encode pipeline.

Mirrors /root/reference/src/yet_another_bpe/tokenizer.py (encode / decode / encode_batch /
decode_batch / from_file / vocab_size / special_tokens / get_vocab) and the adapter's
`encode_iterable` (tests/adapters.py:30-34).  Pre-tokenisation, per-word BPE by rank and the
id scatter all run in libyabpe.so; `decode` is a host-side byte gather.
"""
from __future__ import annotations

import ctypes as C
import json
import os
from collections.abc import Iterable, Iterator, Sequence
from pathlib import Path

import numpy as np

from . import _ffi, engine

_ITER_BATCH_BYTES = 4 << 20
_ITER_BATCH_ITEMS = 1 << 16
_FUSED_IDS = os.environ.get("YABPE_ENCODE_FUSED", "1") != "0"      # ids in one pass over the text (chained scan) instead of count + write
_SMALL_MAX_BYTES = 1 << 15        # encode(str) up to this many bytes runs as one launch (yabpe_encode_small)
_DECODE_DEVICE_MIN = 1 << 16       # ids; shorter lists are gathered on the host
_DECODE_DEVICE_MAX_ID = 1 << 24    # offsets are a dense array over the id range


class BBPETokenizer:
    """Byte-level BPE tokenizer on one B200."""

    def __init__(self, vocab: dict[bytes, int] | None = None, merges: list[tuple[bytes, bytes]] | None = None,
                 special_tokens: list[str] | None = None) -> None:
        self._vocab: dict[bytes, int] = vocab or {}
        self._vocab_inv: dict[int, bytes] = {v: k for k, v in self._vocab.items()}    # tokenizer.py:63
        self._merges: list[tuple[bytes, bytes]] = merges or []
        self._special_tokens: list[str] = special_tokens or []
        # tokenizer.py:74-76: rank = index, the LAST duplicate wins
        ranks: dict[tuple[bytes, bytes], int] = {p: i for i, p in enumerate(self._merges)}
        # symbols = distinct byte strings among single bytes, merge operands and results
        sym: dict[bytes, int] = {bytes([b]): b for b in range(256)}

        sid = sym.setdefault                              # sid(b, len(sym)): the id of b, a new one on first sight
        rows = []                                         # a, b, rank, result (plain ints: numpy row writes cost 10x)
        for (a, b), r in ranks.items():
            sa = sid(a, len(sym))
            sb = sid(b, len(sym))
            rows.append((sa, sb, r, sid(a + b, len(sym))))
        m = np.asarray(rows, dtype=np.int64).reshape(len(rows), 4)
        unk = self._vocab.get(b"[UNK]", 0)                                            # tokenizer.py:299
        self._sym_out = np.fromiter((self._vocab.get(b, unk) for b in sym), dtype=np.int32, count=len(sym))
        # batch rule is exact iff every merge that uses a produced symbol ranks after all merges producing it
        maxprod = np.full(len(sym), -1, dtype=np.int64)
        if len(m):
            np.maximum.at(maxprod, m[:, 3], m[:, 2])
            self._consistent = bool(np.all(m[:, 2] > maxprod[m[:, 0]]) and np.all(m[:, 2] > maxprod[m[:, 1]]))
        else:
            self._consistent = True
        # device merge table: open addressing, key = 1<<63 | a<<32 | b, value = rank<<32 | result
        mcap = engine._pow2_at_least(max(16, 2 * len(m) + 2))
        keys, vals, mask, mix64 = [0] * mcap, [0] * mcap, mcap - 1, engine.mix64
        for a, b, r, c in rows:                           # Python ints: numpy scalar indexing costs 10x per probe
            key = (1 << 63) | (a << 32) | b
            slot = mix64(key) & mask
            while keys[slot]:
                slot = (slot + 1) & mask
            keys[slot] = key
            vals[slot] = (r << 32) | c
        self._mkey, self._mval, self._mcap = np.asarray(keys, dtype=np.uint64), np.asarray(vals, dtype=np.uint64), mcap
        # tokenizer.py:99: longest first (len of the str), stable
        sp_sorted = sorted(self._special_tokens, key=len, reverse=True)
        self._sp_bytes = [s.encode("utf-8") for s in sp_sorted]
        self._sp_ids = np.asarray([self._vocab.get(s, -1) for s in self._sp_bytes] + [-1], dtype=np.int32)
        self._dev = None
        self.last_launches = 0

    # -- persistence (tokenizer.py:106-150, literal format) -------------------------------------
    @classmethod
    def from_file(cls, model_dir: str | Path) -> "BBPETokenizer":
        model_path = Path(model_dir)
        with open(model_path / "vocab.json", encoding="utf-8") as f:
            vocab = {k.encode("latin-1"): v for k, v in json.load(f).items()}
        merges: list[tuple[bytes, bytes]] = []
        with open(model_path / "merges.txt", encoding="utf-8") as f:
            for line in f:
                line = line.rstrip("\n")
                if not line:
                    continue
                parts = line.split(" ", 1)
                if len(parts) == 2:
                    merges.append((parts[0].encode("latin-1"), parts[1].encode("latin-1")))
        special_tokens: list[str] = []
        sp_file = model_path / "special_tokens.json"
        if sp_file.exists():
            with open(sp_file, encoding="utf-8") as f:
                special_tokens = json.load(f)
        return cls(vocab=vocab, merges=merges, special_tokens=special_tokens)

    # -- device model ---------------------------------------------------------------------------
    def _device_model(self, torch):
        dev = torch.cuda.current_device()
        if self._dev is None or self._dev[0] != dev:
            t = lambda a: torch.from_numpy(a).cuda()  # noqa: E731
            tensors = [t(self._mkey.view(np.int64)), t(self._mval.view(np.int64)), t(np.arange(256, dtype=np.int32)),
                       t(self._sym_out), t(self._sp_ids)]
            e = _ffi.EncodeModel()
            e.mkey, e.mval, e.mcap = tensors[0].data_ptr(), tensors[1].data_ptr(), self._mcap
            e.byte_sym, e.sym_out, e.sp_ids = tensors[2].data_ptr(), tensors[3].data_ptr(), tensors[4].data_ptr()
            e.consistent = 1 if self._consistent else 0
            self._dev = (dev, e, tensors)
        return self._dev[1]

    def encode_device(self, text_dev, n: int, cuts: np.ndarray | None = None, own: tuple[int, int] | None = None,
                      reuse_output: bool = False, mailbox: "engine.Mailbox | None" = None, sizing: tuple | None = None):
        """Encode `n` bytes resident on the device; `cuts` = interior document boundaries.
        Returns (ids tensor int32 on device, doc_off tensor int64 or None).  One host sync
        (table sizes) + one (id count).  reuse_output=True writes the ids into a buffer kept by the
        tokenizer (sized from the previous call, grown on demand): no allocation and no host sync between
        the two tile passes; the returned view is valid until the next call.
        `mailbox` / `sizing` (encode_pinned): counters reach the host through yabpe_publish instead of copies, and
        the table sizing of an earlier piece (`self.last_sizing`) replaces the sizing sample."""
        torch = _ffi.require_cuda()
        L = _ffi.load()
        if n == 0:
            return torch.empty(0, dtype=torch.int32, device="cuda"), None
        e = self._device_model(torch)
        launches0 = _ffi.launch_count()
        prof = getattr(self, "profile", False)
        evs = []

        def mark():
            if prof:
                ev = torch.cuda.Event(enable_timing=True); ev.record(); evs.append(ev)

        mark()
        res, st = engine.pretok_count_checked(torch, text_dev, n, cuts, self._sp_bytes, mode=1, own=own,
                                              mailbox=mailbox, sizing=sizing)
        self.last_sizing = res.sizing()
        mark()
        words = engine.compact_words(torch, res, st, with_maps=True)
        stream = _ffi.stream_ptr(torch)
        _ffi.check(L.yabpe_encode_words(C.byref(e), C.byref(words.table), words.n_words, stream))
        _ffi.check(L.yabpe_encode_finalize(C.byref(res.args), C.byref(e), C.byref(words.table), words.n_words, stream))
        mark()
        lo, hi = own if own is not None else (0, n)
        n_tiles = int(L.yabpe_num_tiles(lo, hi))
        tile_count = torch.zeros(n_tiles + 2, dtype=torch.int64, device="cuda")
        n_cuts = 0 if cuts is None else len(cuts)
        doc_off = torch.full((n_cuts + 2,), -1, dtype=torch.int64, device="cuda") if n_cuts else None
        o = _ffi.EncodeOut()
        o.tile_count = tile_count.data_ptr(); o.out_ids = None; o.out_cap = 0
        o.doc_off = doc_off.data_ptr() if n_cuts else None

        def read_total() -> int:
            return int(mailbox.read(tile_count[n_tiles:], 1)[0]) if mailbox is not None else int(tile_count[n_tiles].item())

        if _FUSED_IDS:
            # ONE pass over the text (counts, chained scan, ids): the id buffer is sized by a guess -- ids per byte of the last
            # call, else 0.6 -- and the pass is repeated with the exact size in the rare case the guess was short
            mark()                                                # (no separate count pass: its stage time reads 0)
            buf = getattr(self, "_ids_buf", None) if reuse_output else None
            want = int(n * getattr(self, "_ids_per_byte", 0.6) * 1.05) + 4096
            if buf is None or buf.numel() < want:
                buf = torch.empty(want, dtype=torch.int32, device="cuda")
            while True:
                o.out_ids = buf.data_ptr(); o.out_cap = buf.numel()
                _ffi.check(L.yabpe_encode_ids(C.byref(res.args), C.byref(e), C.byref(words.table), C.byref(o), 2, stream))
                total = read_total()
                if total <= buf.numel():
                    break
                buf = torch.empty(total + total // 16 + 4096, dtype=torch.int32, device="cuda")
                tile_count.zero_()
                if doc_off is not None:
                    doc_off.fill_(-1)
            self._ids_per_byte = max(total / max(n, 1), 0.05)
            if reuse_output:
                self._ids_buf = buf
            ids = buf
        else:
            _ffi.check(L.yabpe_encode_ids(C.byref(res.args), C.byref(e), C.byref(words.table), C.byref(o), 0, stream))
            mark()
            if reuse_output:
                buf = getattr(self, "_ids_buf", None)
                want = n // 2 + 4096
                if buf is None or buf.numel() < want:
                    buf = torch.empty(want, dtype=torch.int32, device="cuda")
                while True:                                   # writes beyond out_cap are dropped by the kernel
                    o.out_ids = buf.data_ptr(); o.out_cap = buf.numel()
                    _ffi.check(L.yabpe_encode_ids(C.byref(res.args), C.byref(e), C.byref(words.table), C.byref(o), 1, stream))
                    total = int(tile_count[n_tiles].item())
                    if total <= buf.numel():
                        break
                    buf = torch.empty(total + total // 8, dtype=torch.int32, device="cuda")
                self._ids_buf = buf
                ids = buf
            else:
                total = read_total()
                ids = torch.empty(max(total, 1), dtype=torch.int32, device="cuda")
                o.out_ids = ids.data_ptr(); o.out_cap = total
                _ffi.check(L.yabpe_encode_ids(C.byref(res.args), C.byref(e), C.byref(words.table), C.byref(o), 1, stream))
        mark()
        self.last_launches = _ffi.launch_count() - launches0
        if prof:
            torch.cuda.synchronize()
            names = ["pretok_count_ms", "words_ms", "count_pass_ms", "write_pass_ms"]
            self.timing = {k: evs[i].elapsed_time(evs[i + 1]) for i, k in enumerate(names)}
            self.timing["unique_words"] = words.n_words
        self._keep = (res, words, tile_count)
        return ids[:total], doc_off

    # -- host buffers in, host buffers out ------------------------------------------------------
    def _piece_ends(self, host_np: np.ndarray, n: int, piece_bytes: int) -> list[int]:
        """Ends of the pieces `encode_pinned` streams: every interior end lies right after an occurrence of the
        special token, so that encode(text) == concat(encode(piece)) (tokenizer.py:171-189: the parts between
        specials are independent texts).  Exact only when occurrences cannot overlap each other -- one special
        token without a border (no proper prefix that is also a suffix); otherwise the text stays one piece."""
        if n <= piece_bytes or len(self._sp_bytes) != 1:
            return [n]
        sp = self._sp_bytes[0]
        if any(sp[:k] == sp[-k:] for k in range(1, len(sp))):
            return [n]
        ends: list[int] = []
        lo = 0
        while n - lo > piece_bytes + piece_bytes // 8:      # a short last piece is fine: its ids are the un-overlapped tail
            target, window, at = lo + piece_bytes, 1 << 16, -1
            while at < 0 and window <= 2 * piece_bytes:
                a = max(lo, target - window)
                at = host_np[a:target].tobytes().rfind(sp)
                if at >= 0:
                    at += a
                elif a == lo:
                    break
                window *= 4
            if at < 0:                            # no special inside this piece: look forward instead
                at = host_np[target:n].tobytes().find(sp)
                if at < 0:
                    break
                at += target
            lo = at + len(sp)
            if lo >= n:
                break
            ends.append(lo)
        ends.append(n)
        return ends

    def encode_pinned(self, host, out=None, piece_bytes: int = 128 << 20, id_dtype=None):
        """ids of the UTF-8 bytes in `host` (1-D uint8 torch tensor, ideally pinned) as an int32 host tensor.

        The text is streamed through the device in pieces cut after special tokens (`_piece_ends`): the
        host->device copy of piece i+1 and the device->host copy of the ids of piece i-1 run on their own
        streams while piece i is encoded, so a large buffer costs about max(copy in, encode, copy out)
        instead of their sum.  `out` (pinned int32, optional) receives the ids when it is large enough; the
        returned tensor is a view of it.  Same ids as `encode` of the whole text.
        id_dtype=torch.uint16: the ids are narrowed on the device (yabpe_narrow_ids) and come back as uint16 -- half the
        download, which is what bounds this path; only for vocabularies whose ids are all below 65 536 (ValueError otherwise)."""
        torch = _ffi.require_cuda()
        n = int(host.numel())
        id_dtype = id_dtype or torch.int32
        if id_dtype not in (torch.int32, torch.uint16):
            raise ValueError("id_dtype must be torch.int32 or torch.uint16")
        narrow = id_dtype == torch.uint16
        if narrow and max(list(self._vocab.values()) + [0]) >= 1 << 16:
            raise ValueError("uint16 ids need a vocabulary whose ids are all below 65536")
        if n == 0:
            return torch.empty(0, dtype=id_dtype)
        assert host.dtype == torch.uint8 and host.dim() == 1 and not host.is_cuda
        ends = self._piece_ends(host.numpy(), n, int(piece_bytes))
        starts = [0] + ends[:-1]
        cap = max(((e - s + 15) // 16) * 16 + 64 for s, e in zip(starts, ends))
        cur = torch.cuda.current_stream()
        s_in, s_out = torch.cuda.Stream(), torch.cuda.Stream()
        s_in.wait_stream(cur)
        bufs = [torch.empty(cap, dtype=torch.uint8, device="cuda") for _ in range(min(2, len(ends)))]
        free_ev = [None] * len(bufs)               # the encode of the piece that used the buffer last
        in_ev: list = [None] * len(ends)

        def copy_in(i: int) -> None:
            b, ln = bufs[i % len(bufs)], ends[i] - starts[i]
            with torch.cuda.stream(s_in):
                if free_ev[i % len(bufs)] is not None:
                    s_in.wait_event(free_ev[i % len(bufs)])
                b[:ln].copy_(host[starts[i]:ends[i]], non_blocking=True)
                b[ln:((ln + 15) // 16) * 16 + 64].zero_()
                in_ev[i] = torch.cuda.Event()
                in_ev[i].record(s_in)

        def ensure_out(need: int, done: int, used: int):
            """A pinned landing buffer for `need` ids; the `used` ids already copied are carried over."""
            nonlocal out
            if out is not None and out.numel() >= need:
                return
            est = int(need * (n / done) * 1.05) + 4096        # ids per byte so far, scaled to the whole text
            new = torch.empty(max(est, need), dtype=id_dtype).pin_memory()
            if used:
                s_out.synchronize()
                new[:used].copy_(out[:used])
            out = new

        if out is not None:
            assert out.dtype == id_dtype and out.dim() == 1 and not out.is_cuda
        pos = 0
        mailbox = engine.Mailbox(torch)            # counters bypass the copy engines the bulk transfers occupy
        sizing = None                              # table sizes, layout and hot set of the first piece serve the others
        copy_in(0)
        for i in range(len(ends)):
            if i + 1 < len(ends):
                copy_in(i + 1)
            cur.wait_event(in_ev[i])
            ln = ends[i] - starts[i]
            ids, _ = self.encode_device(bufs[i % len(bufs)][:((ln + 15) // 16) * 16 + 64], ln, mailbox=mailbox, sizing=sizing)
            if sizing is None and len(ends) > 1 and ends[1] - starts[1] <= 2 * ln:
                sizing = self.last_sizing
            done = torch.cuda.Event()
            done.record(cur)
            free_ev[i % len(bufs)] = done
            k = int(ids.numel())
            if narrow and k:
                ids16 = torch.empty(k, dtype=torch.uint16, device="cuda")
                _ffi.check(_ffi.load().yabpe_narrow_ids(ids.data_ptr(), ids16.data_ptr(), k, _ffi.stream_ptr(torch)))
                ids = ids16
                done = torch.cuda.Event()
                done.record(cur)
                free_ev[i % len(bufs)] = done
            ensure_out(pos + k, ends[i], pos)
            with torch.cuda.stream(s_out):
                s_out.wait_event(done)
                out[pos:pos + k].copy_(ids, non_blocking=True)
            ids.record_stream(s_out)
            pos += k
        s_out.synchronize()
        cur.wait_stream(s_in)
        return out[:pos]

    # -- reference API --------------------------------------------------------------------------
    def _encode_small(self, torch, raw: bytes) -> list[int] | None:
        """Short texts: ONE launch (yabpe_encode_small), text and ids through mapped pinned memory -- no copies, one event
        wait.  None when the kernel declines (a pre-token longer than 64 bytes): the caller takes the batched path."""
        L = _ffi.load()
        dev = torch.cuda.current_device()
        st = getattr(self, "_small", None)
        if st is None or st[0] != dev:
            cap = int(L.yabpe_encode_small_max_bytes())
            blob, offs = engine.pack_specials(self._sp_bytes)
            st = (dev, cap, torch.zeros(cap + 64, dtype=torch.uint8).pin_memory(), torch.zeros(cap + 8, dtype=torch.int32).pin_memory(),
                  torch.empty(cap + 64, dtype=torch.int32, device="cuda"), blob, offs, torch.cuda.Event())
            self._small = st
        _, cap, tin, tout, scratch, blob, offs, ev = st
        n = len(raw)
        tin.numpy()[:n] = np.frombuffer(raw, dtype=np.uint8)
        e = self._device_model(torch)
        out = tout.numpy()
        out[0] = -2                                  # the kernel stores the id count here LAST (after a system-wide fence)
        _ffi.check(L.yabpe_encode_small(C.byref(e), tin.data_ptr(), n, blob.ctypes.data, offs.ctypes.data, len(self._sp_bytes),
                                        scratch.data_ptr(), tout.data_ptr(), n + 1, _ffi.stream_ptr(torch)))
        # the result lands in mapped host memory: polling that word is cheaper than an event round trip; the event is the
        # fallback when the kernel takes long (or failed: the synchronize then raises)
        cnt = -2
        for _ in range(2000):
            cnt = int(out[0])
            if cnt != -2:
                break
        if cnt == -2:
            ev.record()
            ev.synchronize()
            cnt = int(out[0])
        if cnt < 0:
            return None
        self.last_launches = 1
        return out[1:1 + cnt].tolist()

    def encode(self, text: str) -> list[int]:
        if not text:
            return []
        torch = _ffi.require_cuda()
        data = text.encode("utf-8")
        if len(data) <= _SMALL_MAX_BYTES:
            ids = self._encode_small(torch, data)
            if ids is not None:
                return ids
        raw = np.frombuffer(data, dtype=np.uint8)
        text_dev, n = engine.to_device_text(torch, raw)
        ids, _ = self.encode_device(text_dev, n)
        return ids.cpu().tolist()

    def encode_batch(self, texts: Sequence[str]) -> list[list[int]]:
        """== [encode(t) for t in texts] (tokenizer.py:351-360), one device pass for the batch."""
        torch = _ffi.require_cuda()
        raws = [t.encode("utf-8") for t in texts]
        lens = np.asarray([len(r) for r in raws], dtype=np.int64)
        total = int(lens.sum())
        if total == 0:
            return [[] for _ in texts]
        ends = np.cumsum(lens)
        cuts = np.unique(ends[:-1][(ends[:-1] > 0) & (ends[:-1] < total)])
        text_dev, n = engine.to_device_text(torch, np.frombuffer(b"".join(raws), dtype=np.uint8))
        ids, doc_off = self.encode_device(text_dev, n, cuts if len(cuts) else None)
        ids_h = ids.cpu().numpy()
        # id offset of every byte boundary that is a document start
        starts = np.concatenate([[0], ends[:-1]])
        if len(cuts):
            off = doc_off.cpu().numpy()
            off[0] = 0
            off[len(cuts) + 1] = len(ids_h)
            for k in range(len(cuts), 0, -1):          # a cut whose first token produced no write keeps -1
                if off[k] < 0:
                    off[k] = off[k + 1]
            cut_off = dict(zip(cuts.tolist(), off[1:len(cuts) + 1].tolist()))
        else:
            cut_off = {}
        cut_off[0] = 0
        cut_off[total] = len(ids_h)
        out = []
        for s, ln in zip(starts.tolist(), lens.tolist()):
            out.append(ids_h[cut_off[s]:cut_off[s + ln]].tolist() if ln else [])
        return out

    def encode_iterable(self, iterable: Iterable[str]) -> Iterator[int]:
        """Lazy, bounded-memory flattening of encode(item) for every item (adapters.py:30-34)."""
        batch: list[str] = []
        size = 0
        for item in iterable:
            batch.append(item)
            size += len(item)
            if size >= _ITER_BATCH_BYTES or len(batch) >= _ITER_BATCH_ITEMS:
                for ids in self.encode_batch(batch):
                    yield from ids
                batch, size = [], 0
        if batch:
            for ids in self.encode_batch(batch):
                yield from ids

    def _device_vocab(self, torch):
        """Vocabulary bytes as (offsets int64 [cap + 1], pool uint8) on the device; ids without bytes get an empty range."""
        dev = torch.cuda.current_device()
        dv = getattr(self, "_dev_vocab", None)
        if dv is None or dv[0] != dev:
            inv = self._vocab_inv
            cap = max((i for i in inv if isinstance(i, int) and 0 <= i < (1 << 31) - 1), default=-1) + 1
            lens = np.zeros(cap + 1, dtype=np.int64)
            for i, b in inv.items():
                if 0 <= i < cap:
                    lens[i + 1] = len(b)
            off = np.cumsum(lens)
            pool = np.frombuffer(b"".join(inv[i] for i in range(cap) if i in inv) + b"\0" * 16, dtype=np.uint8).copy()
            dv = (dev, cap, torch.from_numpy(off).cuda(), torch.from_numpy(pool).cuda())
            self._dev_vocab = dv
        return dv[1], dv[2], dv[3]

    def decode_device(self, ids_dev, n_ids: int | None = None):
        """Bytes of the ids in an int32 device tensor, as a uint8 device tensor (yabpe_decode_ids: ids that are not in
        the vocabulary are skipped, tokenizer.py:335-339).  One host sync (the byte count)."""
        torch = _ffi.require_cuda()
        L = _ffi.load()
        n_ids = int(ids_dev.numel()) if n_ids is None else int(n_ids)
        if n_ids == 0:
            return torch.empty(0, dtype=torch.uint8, device="cuda")
        assert ids_dev.dtype == torch.int32 and ids_dev.is_cuda and ids_dev.is_contiguous()
        cap, off, pool = self._device_vocab(torch)
        n_blocks = int(L.yabpe_decode_blocks(n_ids))
        counts = torch.empty(n_blocks + 1, dtype=torch.int64, device="cuda")
        d = _ffi.DecodeArgs()
        d.ids, d.n_ids, d.tok_off, d.tok_bytes, d.vocab_cap = ids_dev.data_ptr(), n_ids, off.data_ptr(), pool.data_ptr(), cap
        d.block_count, d.out, d.out_cap = counts.data_ptr(), None, 0
        stream = _ffi.stream_ptr(torch)
        _ffi.check(L.yabpe_decode_ids(C.byref(d), 0, stream))
        total = int(counts[n_blocks].item())
        out = torch.empty(max(total, 1), dtype=torch.uint8, device="cuda")
        d.out, d.out_cap = out.data_ptr(), total
        _ffi.check(L.yabpe_decode_ids(C.byref(d), 1, stream))
        return out[:total]

    def decode(self, ids: Sequence[int]) -> str:
        """tokenizer.py:323-349: unknown ids skipped; strict UTF-8, else errors='replace'.  The byte gather runs on
        the device from _DECODE_DEVICE_MIN ids up (below that a kernel launch costs more than the join)."""
        if not len(ids):
            return ""
        buf = None
        if len(ids) >= _DECODE_DEVICE_MIN and max(self._vocab_inv, default=0) < _DECODE_DEVICE_MAX_ID:
            try:
                arr = np.asarray(ids, dtype=np.int64)
            except (OverflowError, TypeError, ValueError):
                arr = None
            if arr is not None and arr.ndim == 1:
                torch = _ffi.require_cuda()
                arr = np.where((arr >= 0) & (arr < (1 << 31) - 1), arr, -1).astype(np.int32)
                buf = self.decode_device(torch.from_numpy(arr).cuda()).cpu().numpy().tobytes()
        if buf is None:
            inv = self._vocab_inv
            buf = b"".join(inv[i] for i in ids if i in inv)
        try:
            return buf.decode("utf-8")
        except UnicodeDecodeError:
            return buf.decode("utf-8", errors="replace")

    def decode_batch(self, ids_batch: Sequence[Sequence[int]]) -> list[str]:
        return [self.decode(ids) for ids in ids_batch]

    @property
    def vocab_size(self) -> int:
        return len(self._vocab)

    @property
    def special_tokens(self) -> list[str]:
        return self._special_tokens.copy()

    def get_vocab(self) -> dict[str, int]:
        return {k.decode("latin-1"): v for k, v in self._vocab.items()}

    def clear_cache(self) -> None:
        """The reference clears its LRU word cache here; the GPU path keeps no cache between calls."""

    def cache_info(self) -> str:
        return "hits=0, misses=0, size=0/0"


class Tokenizer:
    """`Tokenizer(vocab, merges, special_tokens)` with the adapter's id->bytes vocab (adapters.py:37-63)."""

    def __init__(self, vocab: dict[int, bytes], merges: list[tuple[bytes, bytes]],
                 special_tokens: list[str] | None = None) -> None:
        self._tokenizer = BBPETokenizer(vocab={v: k for k, v in vocab.items()}, merges=merges,
                                        special_tokens=special_tokens or [])

    def encode(self, text: str) -> list[int]:
        return self._tokenizer.encode(text)

    def decode(self, ids: list[int]) -> str:
        return self._tokenizer.decode(ids)

    def encode_iterable(self, iterable: Iterable[str]) -> Iterator[int]:
        return self._tokenizer.encode_iterable(iterable)

    def encode_batch(self, texts: Sequence[str]) -> list[list[int]]:
        return self._tokenizer.encode_batch(texts)

    @property
    def inner(self) -> BBPETokenizer:
        return self._tokenizer
