"""ctypes wrapper around the CPU oracle (oracle/bpe_oracle.c).

TEST INFRASTRUCTURE ONLY -- see the header of bpe_oracle.c.  Nothing under
yet-another-bpe_b200/ imports this module; tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs do.

The public functions mirror the reference's adapter boundary
(/root/reference/tests/adapters.py:37-99): `train_bpe` == run_train_bpe,
`Tokenizer` == get_tokenizer(...)'s TokenizerAdapter.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from collections.abc import Iterable, Iterator
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
_LIB: C.CDLL | None = None

c_u8p = C.POINTER(C.c_uint8)
c_i32p = C.POINTER(C.c_int32)
c_i64p = C.POINTER(C.c_int64)


def build(force: bool = False) -> Path:
    """Compile liboracle.so next to the sources (gcc, a second or two)."""
    so = _HERE / "liboracle.so"
    src = _HERE / "bpe_oracle.c"
    inc = _HERE / "unicode_tables.inc"
    if force or not so.exists() or so.stat().st_mtime < max(src.stat().st_mtime, inc.stat().st_mtime):
        subprocess.check_call(["make", "-s", "-B", "-C", str(_HERE), "liboracle.so"])
    return so


def lib() -> C.CDLL:
    global _LIB
    if _LIB is None:
        L = C.CDLL(str(build()))
        L.orc_class_of.restype = C.c_int
        L.orc_class_of.argtypes = [C.c_uint32]
        L.orc_utf8_first_error.restype = C.c_int64
        L.orc_utf8_first_error.argtypes = [C.c_void_p, C.c_int64]
        L.orc_chunk_cuts.restype = C.c_int64
        L.orc_chunk_cuts.argtypes = [C.c_void_p, C.c_int64, C.c_int64, C.c_void_p, C.c_int64]
        L.orc_pretokenize.restype = C.c_int64
        L.orc_pretokenize.argtypes = [C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_void_p, C.c_void_p,
                                      C.c_int32, C.c_int, C.c_void_p, C.c_void_p, C.c_int64]
        L.orc_trainer_new.restype = C.c_void_p
        L.orc_trainer_new.argtypes = [C.c_void_p, C.c_void_p, C.c_int32]
        L.orc_trainer_free.argtypes = [C.c_void_p]
        L.orc_trainer_feed.restype = C.c_int
        L.orc_trainer_feed.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_int64, c_i64p]
        L.orc_trainer_feed_word.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_int64]
        for name in ("num_words", "num_pretokens", "words_bytes", "vocab_size", "vocab_bytes"):
            fn = getattr(L, f"orc_trainer_{name}")
            fn.restype = C.c_int64
            fn.argtypes = [C.c_void_p]
        L.orc_trainer_get_words.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_trainer_run.restype = C.c_int64
        L.orc_trainer_run.argtypes = [C.c_void_p, C.c_int64, C.c_int64, C.c_int]
        L.orc_trainer_get_vocab.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_trainer_get_merges.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_tok_new.restype = C.c_void_p
        L.orc_tok_new.argtypes = [C.c_int32] + [C.c_void_p] * 6 + [C.c_int64] + [C.c_void_p] * 3 + [C.c_int32]
        L.orc_tok_free.argtypes = [C.c_void_p]
        L.orc_tok_encode.restype = C.c_int64
        L.orc_tok_encode.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p, C.c_int64]
        _LIB = L
    return _LIB


def _ptr(a: np.ndarray) -> int:
    return a.ctypes.data


def _pack_specials(specials: list[bytes]) -> tuple[np.ndarray, np.ndarray]:
    blob = np.frombuffer(b"".join(specials) + b"\0", dtype=np.uint8).copy()
    offs = np.zeros(len(specials) + 1, dtype=np.int32)
    np.cumsum([len(s) for s in specials], out=offs[1:])
    return blob, offs


def _as_u8(data: bytes | bytearray | memoryview | np.ndarray) -> np.ndarray:
    if isinstance(data, np.ndarray):
        return np.ascontiguousarray(data, dtype=np.uint8)
    return np.frombuffer(data, dtype=np.uint8)


def class_of(cp: int) -> int:
    return lib().orc_class_of(cp)


def utf8_first_error(data: bytes) -> int:
    a = _as_u8(data)
    return lib().orc_utf8_first_error(_ptr(a) if a.size else None, a.size)


def chunk_cuts(data, chunk_size: int) -> list[int]:
    a = _as_u8(data)
    if a.size == 0:
        return []
    n = lib().orc_chunk_cuts(_ptr(a), a.size, chunk_size, None, 0)
    cuts = np.zeros(n, dtype=np.int64)
    lib().orc_chunk_cuts(_ptr(a), a.size, chunk_size, _ptr(cuts), n)
    return cuts.tolist()


def pretokenize(data, special_tokens: list[str] | None = None, mode: str = "train",
                chunk_size: int = 1 << 30) -> list[bytes]:
    """Pre-tokens of `data` as byte strings.

    mode="train": trainer.py:163-170 semantics (specials are leading alternatives, list
    order, P1 chunk cuts apply).  mode="encode": tokenizer.py:169-186 semantics (split at
    specials first, longest-first); specials that match are returned too.
    """
    starts, _ = pretokenize_spans(data, special_tokens, mode, chunk_size)
    a = _as_u8(data)
    raw = a.tobytes()
    ends = starts[1:] + [len(raw)]
    return [raw[s:e] for s, e in zip(starts, ends)]


def pretokenize_spans(data, special_tokens=None, mode="train", chunk_size=1 << 30):
    a = _as_u8(data)
    sp = [s.encode("utf-8") for s in (special_tokens or [])]
    if mode == "encode":
        # tokenizer.py:99: sorted by len(str) descending, stable
        sp = [s.encode("utf-8") for s in sorted(special_tokens or [], key=len, reverse=True)]
    blob, offs = _pack_specials(sp)
    if a.size == 0:
        return [], []
    cuts = np.asarray(chunk_cuts(a, chunk_size) if mode == "train" else [a.size], dtype=np.int64)
    cuts = cuts[:-1]  # interior cuts only; the last chunk ends at n
    L = lib()
    m = 0 if mode == "train" else 1
    n = L.orc_pretokenize(_ptr(a), a.size, _ptr(cuts) if cuts.size else None, cuts.size,
                          _ptr(blob), _ptr(offs), len(sp), m, None, None, 0)
    starts = np.zeros(n, dtype=np.int64)
    kinds = np.zeros(n, dtype=np.int32)
    L.orc_pretokenize(_ptr(a), a.size, _ptr(cuts) if cuts.size else None, cuts.size,
                      _ptr(blob), _ptr(offs), len(sp), m, _ptr(starts), _ptr(kinds), n)
    return starts.tolist(), kinds.tolist()


class Trainer:
    """Restatement of BBPETrainer (trainer.py:55-302) over the C oracle."""

    def __init__(self, special_tokens: list[str]):
        self._sp = [s.encode("utf-8") for s in special_tokens]
        blob, offs = _pack_specials(self._sp)
        self._h = lib().orc_trainer_new(_ptr(blob), _ptr(offs), len(self._sp))

    def __del__(self):
        if getattr(self, "_h", None):
            lib().orc_trainer_free(self._h)
            self._h = None

    def feed_bytes(self, data, chunk_size: int = 1 << 30, name: str = "<bytes>") -> None:
        a = _as_u8(data)
        if a.size == 0:
            return
        err = C.c_int64(-1)
        rc = lib().orc_trainer_feed(self._h, _ptr(a), a.size, chunk_size, C.byref(err))
        if rc != 0:
            raise ValueError(f"File {name} contains invalid UTF-8 at position {err.value}.")

    def feed_word(self, word: bytes, freq: int = 1) -> None:
        a = _as_u8(word)
        lib().orc_trainer_feed_word(self._h, _ptr(a), a.size, freq)

    def word_counts(self) -> dict[bytes, int]:
        L = lib()
        n = L.orc_trainer_num_words(self._h)
        nb = L.orc_trainer_words_bytes(self._h)
        blob = np.zeros(max(nb, 1), dtype=np.uint8)
        offs = np.zeros(n + 1, dtype=np.int64)
        freqs = np.zeros(max(n, 1), dtype=np.int64)
        L.orc_trainer_get_words(self._h, _ptr(blob), _ptr(offs), _ptr(freqs))
        raw = blob.tobytes()
        return {raw[offs[i]:offs[i + 1]]: int(freqs[i]) for i in range(n)}

    @property
    def num_pretokens(self) -> int:
        return lib().orc_trainer_num_pretokens(self._h)

    def run(self, vocab_size: int, min_frequency: int = 1, fast: bool = False):
        L = lib()
        nm = L.orc_trainer_run(self._h, vocab_size, min_frequency, 1 if fast else 0)
        nv = L.orc_trainer_vocab_size(self._h)
        nb = L.orc_trainer_vocab_bytes(self._h)
        blob = np.zeros(max(nb, 1), dtype=np.uint8)
        offs = np.zeros(nv + 1, dtype=np.int64)
        L.orc_trainer_get_vocab(self._h, _ptr(blob), _ptr(offs))
        raw = blob.tobytes()
        vocab = {i: raw[offs[i]:offs[i + 1]] for i in range(nv)}
        pairs = np.zeros((max(nm, 1), 2), dtype=np.int32)
        newids = np.zeros(max(nm, 1), dtype=np.int32)
        if nm:
            L.orc_trainer_get_merges(self._h, _ptr(pairs), _ptr(newids))
        merges = [(vocab[int(pairs[i, 0])], vocab[int(pairs[i, 1])]) for i in range(nm)]
        return vocab, merges


def train_bpe(input_path: str | os.PathLike, vocab_size: int, special_tokens: list[str], *,
              min_frequency: int = 1, chunk_size_bytes: int = 1 << 30, fast: bool = False):
    """== tests/adapters.py:66-99 run_train_bpe (min_frequency=1, 1 GiB chunks)."""
    path = Path(input_path)
    if not path.exists():
        raise FileNotFoundError(f"File not found: {path}")
    tr = Trainer(special_tokens)
    data = np.fromfile(path, dtype=np.uint8)
    tr.feed_bytes(data, chunk_size_bytes, str(path))
    return tr.run(vocab_size, min_frequency, fast)


def train_bpe_bytes(data, vocab_size, special_tokens, *, min_frequency=1, chunk_size_bytes=1 << 30, fast=False):
    tr = Trainer(special_tokens)
    tr.feed_bytes(data, chunk_size_bytes)
    return tr.run(vocab_size, min_frequency, fast)


class Tokenizer:
    """Restatement of BBPETokenizer + TokenizerAdapter (tokenizer.py:36-349, adapters.py:16-34)."""

    def __init__(self, vocab: dict[int, bytes], merges: list[tuple[bytes, bytes]],
                 special_tokens: list[str] | None = None):
        self._vocab = {v: k for k, v in vocab.items()}           # adapters.py:56
        self._vocab_inv = {v: k for k, v in self._vocab.items()}  # tokenizer.py:63
        specials = list(special_tokens or [])
        ranks: dict[tuple[bytes, bytes], int] = {p: i for i, p in enumerate(merges)}  # tokenizer.py:74-76
        sym: dict[bytes, int] = {bytes([b]): b for b in range(256)}

        def sid(b: bytes) -> int:
            s = sym.get(b)
            if s is None:
                s = sym[b] = len(sym)
            return s

        m_a, m_b, m_rank, m_res = [], [], [], []
        for (a, b), r in ranks.items():
            m_a.append(sid(a)); m_b.append(sid(b)); m_rank.append(r); m_res.append(sid(a + b))
        unk = self._vocab.get(b"[UNK]", 0)                        # tokenizer.py:299
        sym_out = np.zeros(len(sym), dtype=np.int32)
        for b, s in sym.items():
            sym_out[s] = self._vocab.get(b, unk)
        byte_sym = np.arange(256, dtype=np.int32)
        sp_sorted = sorted(specials, key=len, reverse=True)       # tokenizer.py:99
        sp_bytes = [s.encode("utf-8") for s in sp_sorted]
        blob, offs = _pack_specials(sp_bytes)
        sp_ids = np.asarray([self._vocab.get(s, -1) for s in sp_bytes] + [0], dtype=np.int32)
        arr = [np.asarray(x + [0], dtype=np.int32) for x in (m_a, m_b, m_rank, m_res)]
        self._h = lib().orc_tok_new(len(sym), _ptr(byte_sym), _ptr(sym_out), _ptr(arr[0]), _ptr(arr[1]),
                                    _ptr(arr[2]), _ptr(arr[3]), len(m_a), _ptr(blob), _ptr(offs), _ptr(sp_ids),
                                    len(sp_bytes))

    def __del__(self):
        if getattr(self, "_h", None):
            lib().orc_tok_free(self._h)
            self._h = None

    def encode_bytes(self, data) -> np.ndarray:
        """ids of UTF-8 bytes as an int32 array (large inputs: no Python list of 10^8 ints)."""
        a = _as_u8(data)
        if a.size == 0:
            return np.zeros(0, dtype=np.int32)
        out = np.zeros(a.size, dtype=np.int32)
        n = lib().orc_tok_encode(self._h, _ptr(a), a.size, _ptr(out), out.size)
        assert n <= out.size
        return out[:n]

    def encode(self, text: str) -> list[int]:
        if not text:
            return []
        return self.encode_bytes(text.encode("utf-8")).tolist()

    def encode_iterable(self, iterable: Iterable[str]) -> Iterator[int]:
        for line in iterable:                                     # adapters.py:30-34
            yield from self.encode(line)

    def decode(self, ids) -> str:
        if not ids:
            return ""
        buf = b"".join(self._vocab_inv[i] for i in ids if i in self._vocab_inv)  # tokenizer.py:337-341
        try:
            return buf.decode("utf-8")
        except UnicodeDecodeError:
            return buf.decode("utf-8", errors="replace")
