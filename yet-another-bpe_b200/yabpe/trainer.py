"""B200 trainer: the reference's `BBPETrainer` interface over the CUDA pipeline.

Mirrors /root/reference/src/yet_another_bpe/trainer.py (same class / field / method names,
same exceptions) and the adapter entry point tests/adapters.py:66-99 (`train_bpe`).
All compute happens in libyabpe.so; this module only reads files, sizes buffers and turns
device arrays back into `dict[bytes, int]` / `list[tuple[bytes, bytes]]`.
"""
from __future__ import annotations

import ctypes as C
import json
import os
from collections.abc import Mapping, Sequence
from dataclasses import dataclass, field
from pathlib import Path

import numpy as np

from . import _ffi, engine


@dataclass
class BBPETrainerConfig:
    """Same fields and defaults as trainer.py:17-38.  `max_workers` and `seed` are accepted and
    ignored (the GPU pipeline has no thread pool); `chunk_size_bytes` keeps its SEMANTIC meaning:
    files larger than it are cut into independent texts."""

    vocab_size: int = 32000
    min_frequency: int = 2
    max_workers: int = 8
    chunk_size_bytes: int = 8 * 1024 * 1024
    seed: int = 42
    special_tokens: Sequence[str] = field(default_factory=lambda: ["[PAD]", "[UNK]", "[BOS]", "[EOS]"])


class BBPEModel:
    """Container for a trained model (trainer.py:41-52)."""

    def __init__(self, vocab: Mapping[bytes, int], merges: Sequence[tuple[bytes, bytes]],
                 special_tokens: Sequence[str]) -> None:
        self.vocab: dict[bytes, int] = dict(vocab)
        self.merges: list[tuple[bytes, bytes]] = list(merges)
        self.special_tokens: list[str] = list(special_tokens)


def chunk_cuts(data: np.ndarray, chunk_size: int) -> list[int]:
    """Reference chunk END offsets for one file (trainer.py:139-144,172-198): cuts every
    `chunk_size` bytes, moved back (<= 4 bytes) to a byte that is not a UTF-8 continuation."""
    n = int(data.size)
    if n == 0:
        return []
    if n <= chunk_size:
        return [n]
    cuts: list[int] = []
    start = 0
    while start < n:
        tentative = min(start + chunk_size, n)
        if tentative < n:
            bstart = max(0, tentative - 4)
            pos = tentative - bstart
            while pos > 0 and (int(data[bstart + pos]) & 0xC0) == 0x80:
                pos -= 1
            actual = bstart + pos
        else:
            actual = n
        if actual > start:
            cuts.append(actual)
            start = actual
        else:
            start += 1
    return cuts


def device_chunk_cuts(text_dev, n: int, chunk_size: int) -> list[int]:
    """`chunk_cuts` for bytes resident on the device: reads only the <= 5 bytes around each cut."""
    cuts: list[int] = []
    if n <= chunk_size:
        return cuts
    start = 0
    while start < n:
        tentative = min(start + chunk_size, n)
        if tentative < n:
            bstart = max(0, tentative - 4)
            window = text_dev[bstart:tentative + 1].cpu().numpy()
            pos = tentative - bstart
            while pos > 0 and (int(window[pos]) & 0xC0) == 0x80:
                pos -= 1
            actual = bstart + pos
        else:
            actual = n
        if actual > start:
            cuts.append(actual)
            start = actual
        else:
            start += 1
    return [c for c in cuts if 0 < c < n]


def _phase_cycles(state) -> dict:
    """Phase clocks of the merge kernel (cycles of CTA 0's SM; include/yabpe.h MS_CLK_*)."""
    names = ["pair_histogram", "first_index_and_active_set", "top_list_rebuilds", "index_rebuilds", "leader_sessions", "grid_merges", "total", "n_top_rebuilds"]
    out = {k: int(state[40 + i]) for i, k in enumerate(names)}
    out["grid_merges_by_size[n<=2368,n<=18944,more]"] = [int(x) for x in state[48:51]]
    out["grid_cycles_by_size"] = [int(x) for x in state[51:54]]
    out["grid_batch_cycles[select,barrier1,ranges,rewrite,barrier2]"] = [int(x) for x in state[58:63]]
    return out


@dataclass
class TrainStats:
    n_bytes: int = 0
    n_pretokens: int = 0
    n_words: int = 0
    n_syms: int = 0
    n_merges: int = 0
    index_rebuilds: int = 0
    threshold_rebuilds: int = 0
    n_pairs: int = 0
    n_specials: int = 0
    launches: int = 0
    leader_merges: int = 0
    leader_iterations: int = 0      # an iteration of the leader merges a batch of 1 .. 8 pairs (csrc/merge.cuh)
    batched_merges: int = 0         # merges done as members of a batch of two or more
    grid_batches: int = 0           # grid-mode iterations that merged two or more pairs
    grid_batched_merges: int = 0
    grid_merges: int = 0


class BBPETrainer:
    """Byte-level BPE trainer on one B200 (trainer.py:55-302)."""

    def __init__(self, config: BBPETrainerConfig | None = None) -> None:
        self.config: BBPETrainerConfig = config or BBPETrainerConfig()
        self._vocab: dict[bytes, int] = {}
        self._merges: list[tuple[bytes, bytes]] = []
        self.last_stats = TrainStats()
        self.profile = False            # True: record CUDA-event stage durations into self.timing
        self.stream_min_bytes = 64 << 20          # train(files): inputs this large are streamed through pinned staging
        self.stream_piece_bytes = 32 << 20
        self.pipeline_min_bytes = 256 << 20       # train_from_buffers: inputs this large are counted while they are uploaded
        self.pipeline_piece_bytes = 256 << 20
        self.timing: dict[str, float] = {}

    # -- reference API ---------------------------------------------------------------------
    def train(self, files: Sequence[str | Path]) -> BBPEModel:
        if not files:
            raise ValueError("At least one file must be provided")       # trainer.py:72-73
        paths = [Path(f) if isinstance(f, str) else f for f in files]
        for p in paths:
            if not p.exists():
                raise FileNotFoundError(f"File not found: {p}")           # trainer.py:204-205
        sizes = [p.stat().st_size for p in paths]
        if sum(sizes) >= self.stream_min_bytes:
            return self._train_streamed_files(paths, sizes)
        blobs = [np.fromfile(p, dtype=np.uint8) for p in paths]
        return self.train_from_buffers(blobs, [str(p) for p in paths])

    def _train_streamed_files(self, paths: list[Path], sizes: list[int]) -> BBPEModel:
        """Large inputs: the files are read piece by piece straight into two pinned staging buffers and copied to their
        place in the device text on a copy stream, so reading piece k+1 overlaps the upload of piece k and the corpus
        never exists as a pageable host array (np.fromfile + a pageable copy costs a second pass over the bytes).
        Chunk cuts (P1) are then taken from the bytes around each cut on the device, per file."""
        torch = _ffi.require_cuda()
        total = sum(sizes)
        if total == 0:
            return self._finish(self._init_base_vocab(), [])              # trainer.py:81-85
        text_dev = torch.empty(((total + 15) // 16) * 16 + 64, dtype=torch.uint8, device="cuda")
        text_dev[total:].zero_()
        piece = max(int(self.stream_piece_bytes), 1)
        stage = [torch.empty(piece, dtype=torch.uint8).pin_memory() for _ in range(2)]
        in_flight: list = [None, None]
        cur, copy = torch.cuda.current_stream(), torch.cuda.Stream()
        copy.wait_stream(cur)
        off, k = 0, 0
        for p, size in zip(paths, sizes):
            done = 0
            with open(p, "rb", buffering=0) as f:
                while done < size:
                    buf = stage[k % 2]
                    if in_flight[k % 2] is not None:
                        in_flight[k % 2].synchronize()                    # its previous upload has left the buffer
                    got = f.readinto(memoryview(buf.numpy())[:min(piece, size - done)])
                    if not got:
                        raise OSError(f"{p} shrank while it was read ({done} of {size} bytes)")
                    with torch.cuda.stream(copy):
                        text_dev[off + done:off + done + got].copy_(buf[:got], non_blocking=True)
                        in_flight[k % 2] = torch.cuda.Event()
                        in_flight[k % 2].record(copy)
                    done += got
                    k += 1
            off += size
        cur.wait_stream(copy)
        file_starts, cuts, off = [], [], 0
        for size in sizes:
            file_starts.append(off)
            if size:
                cuts += [off + c for c in device_chunk_cuts(text_dev[off:off + size], size, int(self.config.chunk_size_bytes))]
                cuts.append(off + size)                                   # a file end is a hard boundary too
            off += size
        return self._train_on_device(torch, text_dev, total, cuts, file_starts, [str(p) for p in paths])

    def train_from_buffers(self, blobs: Sequence[np.ndarray], names: Sequence[str] | None = None) -> BBPEModel:
        """Train from host byte buffers, one per file (pinned memory makes the H2D copy fastest)."""
        torch = _ffi.require_cuda()
        cfg = self.config
        names = list(names) if names is not None else [f"<buffer {i}>" for i in range(len(blobs))]
        # P1: chunk cuts per file; file ends are hard boundaries too
        file_starts, cuts, total = [], [], 0
        for b in blobs:
            file_starts.append(total)
            cuts += [total + c for c in chunk_cuts(b, cfg.chunk_size_bytes)]
            total += int(b.size)
        if total == 0:
            return self._finish(self._init_base_vocab(), [])              # trainer.py:81-85
        if total >= self.pipeline_min_bytes and not self.profile:
            return self._train_pipelined(torch, [b for b in blobs], cuts, file_starts, names, total)
        host = blobs[0] if len(blobs) == 1 else np.concatenate([b for b in blobs if b.size])
        text_dev, n = engine.to_device_text(torch, host)
        return self._train_on_device(torch, text_dev, n, cuts, file_starts, names)

    def _train_pipelined(self, torch, blobs: list[np.ndarray], cuts: list[int], file_starts: list[int], names: list[str],
                         total: int) -> BBPEModel:
        """Large host buffers: the upload (PCIe, ~55 GB/s) and the pre-tokenise + count kernels (hundreds of GB/s) overlap.
        The text is cut into pieces at safe edges (sharding.plan_shards: the same edges the multi-GPU path shards at);
        piece k is copied on a copy stream and counted on the compute stream as soon as piece k + 1 -- its look-ahead
        halo -- has landed, every piece into the same tables.  Same result as uploading first: tests compare both."""
        from . import sharding
        import warnings
        specials = [s.encode("utf-8") for s in self.config.special_tokens]
        hard = sorted({c for c in cuts if 0 < c < total})
        starts = np.asarray(file_starts + [total], dtype=np.int64)

        def read(lo: int, hi: int) -> bytes:
            out = []
            for b, s0 in zip(blobs, file_starts):
                a, e = max(lo, s0), min(hi, s0 + int(b.size))
                if a < e:
                    out.append(b[a - s0:e - s0].tobytes())
            return b"".join(out)

        k_pieces = max(2, -(-total // max(int(self.pipeline_piece_bytes), 1 << 20)))
        edges = sharding.plan_shards(read, total, k_pieces, specials, hard)
        edges = sorted(set(edges))
        text_dev = torch.empty(((total + 15) // 16) * 16 + 64, dtype=torch.uint8, device="cuda")
        cur, copy = torch.cuda.current_stream(), torch.cuda.Stream()
        pieces, ready = [], []
        for lo, hi in zip(edges, edges[1:]):
            nk = min(total, hi + sharding.HALO)
            pieces.append((lo, hi, nk, sorted({c for c in hard if c < nk} | ({lo} if lo > 0 else set()))))
        sample_cuts = np.asarray([c for c in hard if c < (16 << 20)], dtype=np.int64)
        # cut lists first (a small synchronous copy), then the bulk upload on its own stream
        text_dev[total:].zero_()
        copy.wait_stream(cur)
        uploaded = []
        with torch.cuda.stream(copy), warnings.catch_warnings():
            warnings.simplefilter("ignore", UserWarning)                  # read-only numpy views are only read
            for lo, hi in zip(edges, edges[1:]):
                for b, s0 in zip(blobs, file_starts):
                    a, e = max(lo, s0), min(hi, s0 + int(b.size))
                    if a < e:
                        text_dev[a:e].copy_(torch.from_numpy(b[a - s0:e - s0]), non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(copy)
                uploaded.append(ev)
        # piece k may look HALO bytes into piece k + 1
        ready = [uploaded[min(k + 1, len(uploaded) - 1)] for k in range(len(pieces))]
        counted = engine.pretok_count_pieces(torch, text_dev, total, pieces, specials, ready,
                                             sample_cuts if len(sample_cuts) else None)
        cur.wait_stream(copy)
        return self._train_on_device(torch, text_dev, total, cuts, file_starts, names, counted=counted)

    def train_device(self, text_dev, n: int, name: str = "<device buffer>") -> BBPEModel:
        """Train on `n` bytes already resident in HBM (capacity >= round_up(n,16)+16, 16-byte aligned),
        treated as ONE file: reference chunk cuts are derived from the bytes around each cut."""
        torch = _ffi.require_cuda()
        if n == 0:
            return self._finish(self._init_base_vocab(), [])
        cuts = device_chunk_cuts(text_dev, n, int(self.config.chunk_size_bytes))
        return self._train_on_device(torch, text_dev, n, cuts, [0], [name])

    def _train_on_device(self, torch, text_dev, n: int, cuts: list[int], file_starts: list[int],
                         names: list[str], counted=None) -> BBPEModel:
        cfg = self.config
        specials = [s.encode("utf-8") for s in cfg.special_tokens]
        base_vocab = self._init_base_vocab()
        cuts = sorted({c for c in cuts if 0 < c < n})
        cuts_np = np.asarray(cuts, dtype=np.int64) if cuts else None
        ev: list = [] if self.profile else None                           # type: ignore[assignment]
        launches0 = _ffi.launch_count()
        if counted is not None:                                           # counted while it was uploaded (_train_pipelined)
            res, st = counted
        else:
            res = engine.pretok_count(torch, text_dev, n, cuts_np, specials, 0, stage_events=ev)
            st = res.stats_host()
        if st[_ffi.ST_TABLE_FULL] != 0:
            del res
            res, st = engine.pretok_count_checked(torch, text_dev, n, cuts_np, specials, mode=0)
        err = int(st[_ffi.ST_ERR_POS])
        if err != _ffi.INT64_MAX:                                         # trainer.py:157-160
            fi = int(np.searchsorted(np.asarray(file_starts), err, side="right")) - 1
            raise ValueError(f"File {names[fi]} contains invalid UTF-8 at position {err - file_starts[fi]}.")
        if self.profile:
            e0 = torch.cuda.Event(enable_timing=True); e0.record()
        words = engine.compact_words(torch, res, st, with_maps=False)
        if self.profile:
            e1 = torch.cuda.Event(enable_timing=True); e1.record()
        stats = TrainStats(n_bytes=n, n_pretokens=int(st[_ffi.ST_NTOK]), n_words=words.n_words, n_syms=words.n_syms,
                           n_specials=int(st[_ffi.ST_NSPECIAL]))
        base_tokens = list(base_vocab.keys())
        num_merges = max(0, cfg.vocab_size - len(base_vocab))            # trainer.py:238
        if words.n_words == 0 or num_merges == 0:
            self.last_stats = stats
            return self._finish(base_vocab, [])

        def restore() -> None:
            words.keep[-1].zero_()
            _ffi.check(_ffi.load().yabpe_compact_words(C.byref(res.args), C.byref(words.table), _ffi.stream_ptr(torch)))

        mr = engine.merge_loop(torch, words, base_tokens, num_merges, int(cfg.min_frequency), restore=restore,
                               timing=self.timing if self.profile else None)
        stats.n_merges = len(mr.merge_new)
        stats.index_rebuilds = int(mr.state[_ffi.MS_REBUILDS])
        stats.threshold_rebuilds = int(mr.state[_ffi.MS_TREBUILDS])
        stats.n_pairs = int(mr.state[_ffi.MS_NPAIRS])
        stats.leader_merges = int(mr.state[_ffi.MS_LEADER_MERGES])
        stats.grid_merges = int(mr.state[_ffi.MS_GRID_MERGES])
        stats.leader_iterations = int(mr.state[54])
        stats.batched_merges = int(mr.state[55])
        stats.grid_batches = int(mr.state[56])
        stats.grid_batched_merges = int(mr.state[57])
        self.timing['leader_cycles'] = [int(x) for x in mr.state[20:29]] + [int(mr.state[12]), int(mr.state[13]), int(mr.state[17])]
        self.timing['merge_phase_cycles'] = _phase_cycles(mr.state)
        stats.launches = _ffi.launch_count() - launches0
        self.last_stats = stats
        if self.profile and ev and len(ev) == 4:
            torch.cuda.synchronize()
            self.timing.update(specials_ms=ev[0].elapsed_time(ev[1]), pretok_tiles_ms=ev[1].elapsed_time(ev[2]),
                               long_tokens_ms=ev[2].elapsed_time(ev[3]), compact_ms=e0.elapsed_time(e1))
        _toks, vocab, merges = mr.materialise()           # the result's Python objects (C API when yabpe/_hostlist.so is built)
        return self._finish(vocab, merges)

    def _preprocess_corpus(self, files: Sequence[str | Path]) -> list[list[int]]:
        """Every pre-token occurrence as a list of byte values, files and text in order (trainer.py:200-214;
        tests/test_trainer.py:27-42).  The training path never materialises this list (it only needs the multiset);
        here the device marks the pre-token starts (yabpe_token_starts) and the host slices the bytes."""
        torch = _ffi.require_cuda()
        cfg = self.config
        specials = [s.encode("utf-8") for s in cfg.special_tokens]
        out: list[list[int]] = []
        for f in files:
            p = Path(f) if isinstance(f, str) else f
            if not p.exists():
                raise FileNotFoundError(f"File not found: {p}")           # trainer.py:204-205
            blob = np.fromfile(p, dtype=np.uint8)
            n = int(blob.size)
            if n == 0:
                continue
            cuts = [c for c in chunk_cuts(blob, cfg.chunk_size_bytes) if 0 < c < n]
            text_dev, _ = engine.to_device_text(torch, blob)
            res, st = engine.pretok_count_checked(torch, text_dev, n, np.asarray(cuts, dtype=np.int64) if cuts else None,
                                                  specials, mode=0)
            err = int(st[_ffi.ST_ERR_POS])
            if err != _ffi.INT64_MAX:                                     # trainer.py:157-160
                raise ValueError(f"File {p} contains invalid UTF-8 at position {err}.")
            starts = engine.token_starts(torch, res).tolist()
            raw = blob.tobytes()
            out.extend(list(raw[s:e]) for s, e in zip(starts, starts[1:] + [n]))
        return out

    def _merge_loop(self, sequences: Sequence[Sequence[int]]) -> tuple[dict[bytes, int], list[tuple[bytes, bytes]]]:
        """The merge loop on a caller-supplied list of byte-value sequences, one per pre-token occurrence
        (trainer.py:216-302; tests/test_trainer.py:214,245).  The occurrences are de-duplicated on the device
        (yabpe_insert_words: the word_freq dict of trainer.py:221-225), then k_merge_loop runs as in train()."""
        torch = _ffi.require_cuda()
        from . import distributed as D
        cfg = self.config
        base_vocab = self._init_base_vocab()
        num_merges = max(0, cfg.vocab_size - len(base_vocab))            # trainer.py:238
        seqs = [bytes(bytearray(s)) for s in sequences if len(s)]         # ValueError for values outside 0..255
        if not seqs or num_merges == 0:
            self._vocab, self._merges = base_vocab, []
            return base_vocab, []
        lens = torch.tensor([len(s) for s in seqs], dtype=torch.int32, device="cuda")
        data = torch.from_numpy(np.frombuffer(b"".join(seqs), dtype=np.uint8).copy()).cuda()
        packed = D.reduce_packed_cuda(D.Packed(lens, torch.ones(len(seqs), dtype=torch.int64, device="cuda"), data))
        words = D.words_from_packed(torch, packed)
        mr = engine.merge_loop(torch, words, list(base_vocab.keys()), num_merges, int(cfg.min_frequency),
                               restore=lambda: D.words_restore(torch, words, packed))
        _toks, vocab, merges = mr.materialise()
        self._vocab, self._merges = vocab, merges
        return vocab, merges

    def save(self, output_dir: str | Path) -> None:
        """Same on-disk format as trainer.py:94-117 (latin-1 keys, "a b" merge lines)."""
        if not self._vocab:
            raise ValueError("Model has not been trained yet. Call train() first.")
        out = Path(output_dir)
        out.mkdir(parents=True, exist_ok=True)
        with open(out / "vocab.json", "w", encoding="utf-8") as f:
            json.dump({k.decode("latin-1"): v for k, v in self._vocab.items()}, f, ensure_ascii=False, indent=2)
        with open(out / "merges.txt", "w", encoding="utf-8") as f:
            for a, b in self._merges:
                f.write(f"{a.decode('latin-1')} {b.decode('latin-1')}\n")
        with open(out / "special_tokens.json", "w", encoding="utf-8") as f:
            json.dump(list(self.config.special_tokens), f, ensure_ascii=False, indent=2)

    # -- helpers ---------------------------------------------------------------------------
    def _init_base_vocab(self) -> dict[bytes, int]:
        """256 bytes, then each special unless its bytes are already a key (trainer.py:119-134)."""
        vocab: dict[bytes, int] = {bytes([i]): i for i in range(256)}
        for s in self.config.special_tokens:
            b = s.encode("utf-8")
            if b not in vocab:
                vocab[b] = len(vocab)
        return vocab

    def _finish(self, vocab: dict[bytes, int], merges: list[tuple[bytes, bytes]]) -> BBPEModel:
        self._vocab, self._merges = vocab, merges
        return BBPEModel(vocab=vocab, merges=merges, special_tokens=list(self.config.special_tokens))


def train_bpe(input_path: str | os.PathLike, vocab_size: int, special_tokens: list[str]
              ) -> tuple[dict[int, bytes], list[tuple[bytes, bytes]]]:
    """Drop-in for tests/adapters.py:66-99 `run_train_bpe` (min_frequency=1, 1 GiB chunks)."""
    config = BBPETrainerConfig(vocab_size=vocab_size, min_frequency=1, max_workers=1,
                               chunk_size_bytes=1024 * 1024 * 1024, seed=42, special_tokens=special_tokens)
    trainer = BBPETrainer(config)
    path = Path(input_path) if not isinstance(input_path, Path) else input_path
    model = trainer.train([path])
    return {v: k for k, v in model.vocab.items()}, model.merges


def pretoken_counts(data: bytes | np.ndarray, special_tokens: Sequence[str] = (), *, mode: str = "train",
                    chunk_size_bytes: int = 1 << 30, cuts: Sequence[int] | None = None,
                    generic_only: bool = False, stats_out: dict | None = None) -> dict[bytes, int]:
    """Device word-count table of `data` as a dict (the multiset trainer.py:221-225 builds).
    mode="encode" applies the tokenizer's special handling (specials split first and not counted)."""
    torch = _ffi.require_cuda()
    raw = np.frombuffer(data, dtype=np.uint8) if not isinstance(data, np.ndarray) else data
    if raw.size == 0:
        return {}
    if mode == "train":
        sp = [s.encode("utf-8") for s in special_tokens]
        cut_list = chunk_cuts(raw, chunk_size_bytes) if cuts is None else list(cuts)
    else:
        sp = [s.encode("utf-8") for s in sorted(special_tokens, key=len, reverse=True)]
        cut_list = [] if cuts is None else list(cuts)
    cut_list = [c for c in cut_list if 0 < c < raw.size]
    text_dev, n = engine.to_device_text(torch, raw)
    res, st = engine.pretok_count_checked(torch, text_dev, n, np.asarray(cut_list, dtype=np.int64) if cut_list else None,
                                          sp, 0 if mode == "train" else 1, generic_only=generic_only)
    if stats_out is not None:
        stats_out.update(slow_items=int(st[_ffi.ST_SLOW_N]), cache_hits=int(st[_ffi.ST_CACHE_HIT]), n_tok=int(st[_ffi.ST_NTOK]))
    if int(st[_ffi.ST_ERR_POS]) != _ffi.INT64_MAX:
        raise ValueError(f"invalid UTF-8 at position {int(st[_ffi.ST_ERR_POS])}")
    words = engine.compact_words(torch, res, st, with_maps=False)
    out: dict[bytes, int] = {}
    if words.n_words:
        wsym = words.wsym.cpu().numpy()
        woff = words.woff.cpu().numpy()
        wlen = words.wlen.cpu().numpy()
        wcnt = words.wcnt.cpu().numpy()
        for w in range(words.n_words):
            b = wsym[woff[w]:woff[w] + wlen[w]].astype(np.uint8).tobytes()
            assert b not in out, "duplicate word in device table"
            out[b] = int(wcnt[w])
    out["__n_pretokens__"] = int(st[_ffi.ST_NTOK])  # type: ignore[index]
    return out
