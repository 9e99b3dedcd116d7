"""Byte-range sharding of ONE corpus (yabpe/sharding.py) against the oracle's scanner, on CPU.

Claim under test: with edges from `plan_shards`, the pre-tokens of shard r's window text[e_r, e_{r+1} + HALO) taken as
a text of its own, restricted to those that START before e_{r+1}, are exactly the pre-tokens of the whole text
(trainer.py:146-170 semantics, reference chunk cuts included) that start in [e_r, e_{r+1}).
"""
from __future__ import annotations

import random
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parents[1]
for p in (ROOT, ROOT / "yet-another-bpe_b200", ROOT / "tests"):
    if str(p) not in sys.path:
        sys.path.insert(0, str(p))

import common  # noqa: E402
from oracle import oracle  # noqa: E402
from yabpe import sharding  # noqa: E402


def whole_spans(text: bytes, specials: list[str], chunk: int) -> list[tuple[int, int]]:
    starts, _ = oracle.pretokenize_spans(text, specials, "train", chunk)
    return list(zip(starts, starts[1:] + [len(text)]))


def shard_spans(text: bytes, specials: list[str], chunk: int, world: int) -> tuple[list[tuple[int, int]], list[int]]:
    sp_b = [s.encode() for s in specials]
    hard = [c for c in oracle.chunk_cuts(text, chunk) if 0 < c < len(text)]
    edges = sharding.plan_shards(lambda a, b: text[a:b], len(text), world, sp_b, hard)
    assert edges[0] == 0 and edges[-1] == len(text) and all(a <= b for a, b in zip(edges, edges[1:]))
    out: list[tuple[int, int]] = []
    for r in range(world):
        start, own_len, n_local = sharding.shard_window(edges, r, len(text))
        if own_len == 0:
            continue
        local = text[start:start + n_local]
        cuts = [c - start for c in hard if start < c < start + n_local]
        # the shard as a text of its own, with the hard cuts that fall inside it (what the device kernel is given)
        pieces, prev = [], 0
        for c in cuts + [n_local]:
            s, _ = oracle.pretokenize_spans(local[prev:c], specials, "train", 1 << 40)
            pieces += [prev + x for x in s]
            prev = c
        ends = pieces[1:] + [n_local]
        own = [(start + s, start + e) for s, e in zip(pieces, ends) if s < own_len]
        assert own and own[-1][1] == start + own_len, "the last owned pre-token must end at the edge"
        out += own
    return out, edges


TEXTS = {
    "owt": lambda: common.synth_owt(300_000, seed=7),
    "tinystories": lambda: common.synth_tinystories(200_000, seed=8),
    "adversarial": lambda: common.synth_adversarial(200_000, seed=9),
    "crlf": lambda: (ROOT / "tests" / "fixtures_gpt2" / "corpus.en").read_bytes(),
}


@pytest.mark.parametrize("name", sorted(TEXTS))
@pytest.mark.parametrize("world", [2, 3, 8])
def test_shards_reproduce_whole_text(name, world):
    text = TEXTS[name]()
    for specials in (["<|endoftext|>"], [], ["<|endoftext|>", "\n\n", " the"]):
        for chunk in (1 << 40, 50_021):
            got, edges = shard_spans(text, specials, chunk, world)
            assert got == whole_spans(text, specials, chunk), (name, world, specials, chunk, edges)


def test_fuzz_dense_specials_and_contractions():
    rng = random.Random(1234)
    alphabet = ["a", "b", "'s", "'ll", "'", " ", "  ", "\n", "\r\n", "\t", "<|endoftext|>", "<|end", "|>", "1", "é", "中",
                " ", " ", "!", "x y", "\n\n", "<|a|>", "<|a|>x"]
    for it in range(60):
        text = "".join(rng.choice(alphabet) for _ in range(rng.randrange(200, 3000))).encode()
        specials = rng.choice([["<|endoftext|>"], ["<|a|>", "<|a|>x"], ["<|endoftext|>", "\n"], ["x y", "<|end"]])
        world = rng.choice([2, 3, 5])
        got, edges = shard_spans(text, specials, 1 << 40, world)
        assert got == whole_spans(text, specials, 1 << 40), (it, specials, edges)


def test_no_safe_edge_collapses_instead_of_guessing():
    text = b"a" * 100_000                                   # one pre-token: nowhere to cut
    edges = sharding.plan_shards(lambda a, b: text[a:b], len(text), 4, [b"<|endoftext|>"])
    assert edges == [0, len(text), len(text), len(text), len(text)]
    got, _ = shard_spans(text, ["<|endoftext|>"], 1 << 40, 4)
    assert got == [(0, len(text))]


def test_edges_avoid_special_tokens():
    sp = b"<|endoftext|>"
    text = (b"word\n" + sp) * 3000                          # every newline is followed by a special: not a safe edge...
    edges = sharding.plan_shards(lambda a, b: text[a:b], len(text), 3, [sp])
    for e in edges[1:-1]:
        assert sp not in text[max(0, e - len(sp)):e + len(sp) + 1] or e in (0, len(text))
    got, _ = shard_spans(text, [sp.decode()], 1 << 40, 3)
    assert got == whole_spans(text, [sp.decode()], 1 << 40)


def test_file_concat_reader(tmp_path):
    blobs = [b"first file\nwith lines\n", b"", "zweite Datei äöü\n".encode(), b"x" * 5000 + b"\nend"]
    paths = []
    for i, b in enumerate(blobs):
        p = tmp_path / f"f{i}.txt"
        p.write_bytes(b)
        paths.append(p)
    cat = sharding.FileConcat(paths, [len(b) for b in blobs])
    whole = b"".join(blobs)
    assert cat.total == len(whole) and cat.read(0, cat.total) == whole
    for lo, hi in [(0, 5), (15, 40), (20, 25), (len(whole) - 7, len(whole)), (3, 3)]:
        assert cat.read(lo, hi) == whole[lo:hi]
        buf = np.zeros(hi - lo, dtype=np.uint8)
        cat.readinto(lo, hi, buf)
        assert buf.tobytes() == whole[lo:hi]
    # hard cuts = reference chunk cuts per file + file ends
    want = []
    off = 0
    for b in blobs:
        if b:
            want += [off + c for c in oracle.chunk_cuts(b, 1000)]
        off += len(b)
    assert cat.hard_cuts(1000) == sorted({c for c in want if 0 < c < len(whole)})


def test_document_shards_encode_like_the_whole_text():
    """distributed.document_shard: concat(encode(shard)) == encode(text) (tokenizer.py:171-189), via the oracle's tokenizer."""
    from types import SimpleNamespace

    from yabpe import distributed as D
    vocab, merges = oracle.train_bpe_bytes(common.synth_owt(60_000, seed=3), 600, ["<|endoftext|>"], fast=True)
    otok = oracle.Tokenizer(vocab, merges, ["<|endoftext|>"])
    docs = [common.synth_owt(3000 + 500 * i, seed=40 + i).replace(b"<|endoftext|>", b" ") for i in range(12)]
    text = b"  \n<|endoftext|>".join(docs) + b" trailing   "
    want = otok.encode(text.decode())
    tok = SimpleNamespace(_sp_bytes=[b"<|endoftext|>"])
    for world in (1, 2, 3, 8, 32):
        got, prev_hi = [], 0
        for r in range(world):
            lo, hi = D.document_shard(tok, lambda a, b: text[a:b], len(text), r, world)
            assert lo == prev_hi and lo <= hi
            assert lo in (0, len(text)) or text[lo - 13:lo] == b"<|endoftext|>"
            prev_hi = hi
            got += otok.encode(text[lo:hi].decode())
        assert prev_hi == len(text) and got == want, world
    # several specials, or one with a border (occurrences may overlap): everything stays on rank 0
    for sps in ([b"<|a|>", b"<|b|>"], [b"abab"]):
        t2 = SimpleNamespace(_sp_bytes=sps)
        assert D.document_shard(t2, lambda a, b: text[a:b], len(text), 0, 4) == (0, len(text))
        assert D.document_shard(t2, lambda a, b: text[a:b], len(text), 3, 4) == (len(text), len(text))
