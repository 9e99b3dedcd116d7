"""Shared test helpers: golden-vector loading, GPT-2 fixture reconstruction, synthetic corpora."""
from __future__ import annotations

import base64
import json
import pickle
from functools import lru_cache
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parents[1]
FIXTURES = ROOT / "tests" / "fixtures_gpt2"
DATA = ROOT / "tests" / "data"
GOLDEN = ROOT / "tests" / "golden"
SNAPSHOTS = ROOT / "tests" / "_snapshots"


@lru_cache
def gpt2_bytes_to_unicode() -> dict[int, str]:
    """The standard GPT-2 printable-byte map (same table as the reference's tests/common.py:9-54)."""
    keep = list(range(ord("!"), ord("~") + 1)) + list(range(0xA1, 0xAD)) + list(range(0xAE, 0x100))
    out = {b: chr(b) for b in keep}
    extra = 0
    for b in range(256):
        if b not in out:
            out[b] = chr(256 + extra)
            extra += 1
    # insertion order matters: ids 0..255 of the GPT-2 vocab follow it (printable bytes first)
    ordered = {b: out[b] for b in keep}
    for b in range(256):
        if b not in ordered:
            ordered[b] = out[b]
    return ordered


@lru_cache
def gpt2_vocab_and_merges() -> tuple[dict[int, bytes], list[tuple[bytes, bytes]]]:
    """gpt2_vocab.json is git-ignored upstream; rebuild it from gpt2_merges.txt."""
    b2u = gpt2_bytes_to_unicode()
    u2b = {v: k for k, v in b2u.items()}
    vocab = {i: bytes([b]) for i, b in enumerate(b2u.keys())}
    merges: list[tuple[bytes, bytes]] = []
    for line in (FIXTURES / "gpt2_merges.txt").read_text(encoding="utf-8").split("\n"):
        parts = line.rstrip().split(" ")
        if len(parts) != 2:
            continue
        a = bytes(u2b[c] for c in parts[0])
        b = bytes(u2b[c] for c in parts[1])
        merges.append((a, b))
        vocab[len(vocab)] = a + b
    vocab[len(vocab)] = b"<|endoftext|>"
    return vocab, merges


def reference_merges_corpus_en() -> list[tuple[bytes, bytes]]:
    """tests/fixtures_gpt2/train-bpe-reference-merges.txt decoded to bytes (test_train_bpe_gpt2.py:44-53)."""
    u2b = {v: k for k, v in gpt2_bytes_to_unicode().items()}
    out = []
    for line in (FIXTURES / "train-bpe-reference-merges.txt").read_text(encoding="utf-8").split("\n"):
        parts = line.rstrip().split(" ")
        if len(parts) == 2:
            out.append((bytes(u2b[c] for c in parts[0]), bytes(u2b[c] for c in parts[1])))
    return out


def load_snapshot() -> dict:
    with open(SNAPSHOTS / "test_train_bpe_special_tokens.pkl", "rb") as f:
        return pickle.load(f)


def load_train_cases() -> list[dict]:
    cases = json.loads((GOLDEN / "train_cases.json").read_text())
    for c in cases:
        if "input_file" in c:
            c["inputs"] = [(ROOT / c["input_file"]).read_bytes()]
        elif "input_b64" in c:
            c["inputs"] = [base64.b64decode(c["input_b64"])]
        else:
            c["inputs"] = [base64.b64decode(x) for x in c["inputs_b64"]]
        c["merges_b"] = [(bytes.fromhex(a), bytes.fromhex(b)) for a, b in c["merges"]]
        c["vocab_b"] = {i: bytes.fromhex(v) for i, v in enumerate(c["vocab"])}
    return cases


def load_pretok_cases() -> list[dict]:
    return json.loads((GOLDEN / "pretokenize_cases.json").read_text())


def load_class_api_cases() -> dict:
    """tests/golden/class_api_cases.json (tools/make_golden.py make_class_api): reference outputs of
    _preprocess_corpus, _merge_loop(sequences), save() and from_file()."""
    d = json.loads((GOLDEN / "class_api_cases.json").read_text())
    for c in d["preprocess"]:
        c["files"] = [base64.b64decode(x) for x in c["files_b64"]]
        c["want"] = [bytes.fromhex(x) for x in c["sequences"]]
    for c in d["merge_loop"]:
        c["seqs"] = [bytes.fromhex(x) for x in c["sequences"]]
        c["want_vocab"] = [bytes.fromhex(x) for x in c["vocab"]]
        c["want_merges"] = [(bytes.fromhex(a), bytes.fromhex(b)) for a, b in c["merges"]]
    for c in d["persist"]:
        c["input"] = (ROOT / c["input_file"]).read_bytes() if "input_file" in c else base64.b64decode(c["input_b64"])
        c["files"] = {k: base64.b64decode(v) for k, v in c["files_b64"].items()}
    return d


def load_prefix_special_cases() -> dict:
    """tests/golden/prefix_specials_cases.json (tools/make_golden.py make_prefix_specials): prefix-related special tokens
    right before chunk cuts, file ends and document ends -- reference outputs of findall / split, train() and encode()."""
    d = json.loads((GOLDEN / "prefix_specials_cases.json").read_text())
    for c in d["train"]:
        c["inputs"] = [base64.b64decode(x) for x in c["inputs_b64"]]
        c["merges_b"] = [(bytes.fromhex(a), bytes.fromhex(b)) for a, b in c["merges"]]
        c["vocab_b"] = {i: bytes.fromhex(v) for i, v in enumerate(c["vocab"])}
    m = d["encode_model"]
    d["encode_model"] = ({i: bytes.fromhex(v) for i, v in enumerate(m["vocab"])},
                         [(bytes.fromhex(a), bytes.fromhex(b)) for a, b in m["merges"]])
    return d


def load_encode_cases() -> tuple[dict, list[dict]]:
    d = json.loads((GOLDEN / "encode_cases.json").read_text())
    models = {}
    for name, m in d["models"].items():
        models[name] = ({i: bytes.fromhex(v) for i, v in enumerate(m["vocab"])},
                        [(bytes.fromhex(a), bytes.fromhex(b)) for a, b in m["merges"]])
    v, m = gpt2_vocab_and_merges()
    models["gpt2"] = (v, m)
    v2 = dict(v)
    v2[50257] = b"<|endoftext|><|endoftext|>"
    models["gpt2+double"] = (v2, m)
    return models, d["cases"]


# --------------------------------------------------------------------------------------
# synthetic corpora (small CPU versions of the BASELINE.json configs)
# --------------------------------------------------------------------------------------

def _lexicon(rng: np.random.Generator, n_types: int) -> list[bytes]:
    letters = np.frombuffer(b"etaoinshrdlcumwfgypbvkjxqz", dtype=np.uint8)
    w = 1.0 / np.arange(1, 27)
    w /= w.sum()
    lens = rng.integers(1, 13, size=n_types)
    out = []
    seen = set()
    for L in lens:
        b = bytes(rng.choice(letters, size=int(L), p=w))
        if b not in seen:
            seen.add(b)
            out.append(b)
    return out


def synth_tinystories(n_bytes: int, seed: int = 20260101, n_types: int = 4000) -> bytes:
    """TinyStories-shaped text: Zipf words, sentence capitals, punctuation, <|endoftext|> separators."""
    rng = np.random.default_rng(seed)
    lex = _lexicon(rng, n_types)
    p = 1.0 / np.arange(1, len(lex) + 1) ** 1.05
    p /= p.sum()
    out = bytearray()
    while len(out) < n_bytes:
        n_words = int(rng.integers(150, 251))
        ids = rng.choice(len(lex), size=n_words, p=p)
        r = rng.random(n_words)
        cap = True
        for k, wid in enumerate(ids):
            w = lex[wid]
            if cap:
                w = w[:1].upper() + w[1:]
                cap = False
            out += w
            if r[k] < 0.03:
                out += b"'s"
            if r[k] > 0.9:
                out += rng.choice([b".", b",", b"!", b"?"])
                cap = out[-1:] != b","
            out += b"\n" if r[k] > 0.99 else b" "
        out += b"\n<|endoftext|>\n"
    return bytes(out[:n_bytes])


def synth_owt(n_bytes: int, seed: int = 20260102, n_types: int = 30000) -> bytes:
    """OWT-shaped text: bigger lexicon, digits, URLs/punctuation runs, ~2 % non-ASCII, \\n\\n paragraphs."""
    rng = np.random.default_rng(seed)
    lex = _lexicon(rng, n_types)
    extra = ["é", "naïve", "über", "中文", "日本語", "привет", "мир", "\U0001f643", "café", "—", "…"]
    p = 1.0 / (np.arange(1, len(lex) + 1) + 2.7)
    p /= p.sum()
    out = bytearray()
    while len(out) < n_bytes:
        n_words = int(rng.integers(200, 1500))
        ids = rng.choice(len(lex), size=n_words, p=p)
        r = rng.random(n_words)
        for k, wid in enumerate(ids):
            if r[k] < 0.03:
                out += str(int(rng.integers(0, 100000))).encode()
            elif r[k] < 0.05:
                out += extra[int(rng.integers(0, len(extra)))].encode("utf-8")
            elif r[k] < 0.055:
                out += b"http://www." + lex[wid] + b".com/" + lex[ids[(k * 7) % n_words]] + b"?x=1&y=2"
            else:
                out += lex[wid]
            if r[k] > 0.88:
                out += rng.choice([b".", b",", b";", b":", b")", b"!!", b"...", b"\""])
            out += b"\n\n" if r[k] > 0.985 else b" "
        out += b"<|endoftext|>"
    raw = bytes(out[:n_bytes])
    # never cut inside a UTF-8 sequence
    while raw and (raw[-1] & 0xC0) == 0x80:
        raw = raw[:-1]
    if raw and raw[-1] >= 0xC0:
        raw = raw[:-1]
    return raw


def synth_adversarial(n_bytes: int, seed: int = 20260104) -> bytes:
    """BASELINE.json config 5 in miniature: long runs, dense specials, odd whitespace, tie farms."""
    rng = np.random.default_rng(seed)
    pieces = [
        lambda: "a" * int(rng.integers(1, 3000)),
        lambda: "9" * int(rng.integers(1, 700)),
        lambda: "!" * int(rng.integers(1, 500)),
        lambda: "中文字符串" * int(rng.integers(1, 200)),
        lambda: "<|endoftext|>" * int(rng.integers(1, 6)),
        lambda: rng.choice(["", " ", "!", "\n", "x", "  ", "\t", "'s", "."]) + "<|endoftext|>",
        lambda: rng.choice(["\U0001f643", "\U00016ea0", "\u00a0", "\u2003", "\u2028", "\u3000", "\u0085", "\u001c", "\u0301", "e"]),
        lambda: rng.choice(["don't", " we've", "I'll", "'S", "'re", "''s", " 's"]),
        lambda: rng.choice(["ab", "abc", "bc", "aaa", "aaaa", "ab ab", "éa", "aé", " xy", " xz", " yx", " zx"]),
        lambda: rng.choice([" ", "  ", "\n", "\r\n", " \n ", "\t"]),
    ]
    weights = np.array([1, 1, 1, 1, 3, 6, 8, 8, 40, 30], dtype=float)
    weights /= weights.sum()
    out = bytearray()
    while len(out) < n_bytes:
        out += pieces[int(rng.choice(len(pieces), p=weights))]().encode("utf-8")
    return bytes(out)
