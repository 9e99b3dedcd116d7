"""Pin the CPU oracle against the reference's own golden vectors and the committed
reference-generated vectors (tests/golden/, made by tools/make_golden.py).  CPU only."""
from __future__ import annotations

import pytest

import common
from oracle import oracle


def test_corpus_en_500_matches_reference_fixture():
    """== /root/reference/tests/test_train_bpe_gpt2.py:27-62 (merges file; vocab json rebuilt)."""
    vocab, merges = oracle.train_bpe(common.FIXTURES / "corpus.en", 500, ["<|endoftext|>"])
    ref = common.reference_merges_corpus_en()
    assert len(ref) == 243
    assert merges == ref
    want_vocab = {bytes([i]) for i in range(256)} | {b"<|endoftext|>"} | {a + b for a, b in ref}
    assert set(vocab.values()) == want_vocab
    assert set(vocab.keys()) == set(range(500))


def test_snapshot_structure_documents_special_token_quirk():
    """tests/_snapshots/test_train_bpe_special_tokens.pkl: 999 ids / 743 merges.
    Its input (tinystories_sample_5M.txt) is missing upstream, so only the structure and the
    quirk it proves can be checked: the special's bytes are merged like any word and the final
    merge re-creates bytes already in the vocab (no new id)."""
    snap = common.load_snapshot()
    assert len(snap["vocab_keys"]) == 999 and len(snap["merges"]) == 743
    assert (b"<", b"|endoftext|>") in snap["merges"]
    # the same quirk reproduced by the oracle on the small TinyStories fixture
    vocab, merges = oracle.train_bpe(common.FIXTURES / "tinystories_sample.txt", 1000, ["<|endoftext|>"], fast=True)
    assert any(a + b == b"<|endoftext|>" for a, b in merges)
    assert len(vocab) < 256 + 1 + len(merges) + 1 and len(set(vocab.values())) == len(vocab)


@pytest.mark.parametrize("fast", [False, True])
def test_train_golden(fast):
    for c in common.load_train_cases():
        tr = oracle.Trainer(c["specials"])
        for blob in c["inputs"]:
            tr.feed_bytes(blob, c["chunk_size"])
        vocab, merges = tr.run(c["vocab_size"], c["min_frequency"], fast)
        assert merges == c["merges_b"], c["name"]
        assert vocab == c["vocab_b"], c["name"]


def test_prefix_related_specials_at_hard_boundaries():
    """A special is never matched across a chunk cut / file end / document end, even when a longer special that
    starts with it would fit across the boundary (reference-generated, tools/make_golden.py make_prefix_specials)."""
    d = common.load_prefix_special_cases()
    for c in d["pretok"]:
        for doc, want in zip(c["docs"], c["tokens"]):
            assert oracle.pretokenize(doc.encode("utf-8"), c["specials"], c["mode"]) == [t.encode("utf-8") for t in want], (c, doc)
    for c in d["train"]:
        tr = oracle.Trainer(c["specials"])
        for blob in c["inputs"]:
            tr.feed_bytes(blob, c["chunk_size"])
        vocab, merges = tr.run(c["vocab_size"], c["min_frequency"], False)
        assert merges == c["merges_b"] and vocab == c["vocab_b"], (c["specials"], c["chunk_size"])
    v, m = d["encode_model"]
    for c in d["encode_docs"]:
        t = oracle.Tokenizer(v, m, c["specials"])
        assert [t.encode(x) for x in c["docs"]] == c["ids"]
        assert list(t.encode_iterable(c["docs"])) == [i for ids in c["ids"] for i in ids]


def test_pretokenize_golden():
    for c in common.load_pretok_cases():
        got = oracle.pretokenize(c["text"].encode("utf-8"), c["specials"], c["mode"])
        assert got == [t.encode("utf-8") for t in c["tokens"]], c


def test_encode_golden():
    models, cases = common.load_encode_cases()
    toks = {}
    for c in cases:
        key = (c["model"], tuple(c["specials"]))
        if key not in toks:
            v, m = models[c["model"]]
            toks[key] = oracle.Tokenizer(v, m, c["specials"])
        t = toks[key]
        assert t.encode(c["text"]) == c["ids"], c["text"][:60]
        assert t.decode(c["ids"]) == c["decoded"]


def test_gpt2_known_ids():
    """ids known from the published GPT-2 vocabulary."""
    v, m = common.gpt2_vocab_and_merges()
    t = oracle.Tokenizer(v, m, ["<|endoftext|>"])
    assert t.encode("Hello world") == [15496, 995]
    assert t.encode("Hello, how are you?") == [15496, 11, 703, 389, 345, 30]
    assert t.encode("\n\n") == [628]
    assert t.encode("<|endoftext|>") == [50256]


def test_encode_matches_offline_tiktoken():
    """Mirror of tests/test_tokenizer_gpt2.py *_matches_tiktoken with an offline Encoding."""
    tiktoken = pytest.importorskip("tiktoken")
    v, m = common.gpt2_vocab_and_merges()
    enc = tiktoken.Encoding("gpt2-local", pat_str=r"""'(?:[sdmt]|ll|ve|re)| ?\p{L}+| ?\p{N}+| ?[^\s\p{L}\p{N}]+|\s+(?!\S)|\s+""",
                            mergeable_ranks={b: i for i, b in v.items() if i < 50256},
                            special_tokens={"<|endoftext|>": 50256})
    t = oracle.Tokenizer(v, m, ["<|endoftext|>"])
    for name in ["address.txt", "german.txt", "tinystories_sample.txt", "corpus.en"]:
        with open(common.FIXTURES / name) as f:
            text = f.read()
        ids = t.encode(text)
        assert ids == enc.encode(text, allowed_special={"<|endoftext|>"}), name
        assert t.decode(ids) == text


def test_utf8_validation_matches_python():
    import random
    rng = random.Random(9)
    pool = [b"a", b"\xc3\xa9", b"\xe4\xb8\xad", b"\xf0\x9f\x99\x83", b"\x80", b"\xc0\x80", b"\xed\xa0\x80", b"\xf4\x90\x80\x80",
            b"\xe0\x80\x80", b"\xf0\x80\x80\x80", b"\xc3", b"\xe4\xb8", b"\xff", b"\xf5\x80\x80\x80", b" "]
    for _ in range(3000):
        b = b"".join(rng.choice(pool) for _ in range(rng.randint(0, 8)))
        try:
            b.decode("utf-8")
            want = -1
        except UnicodeDecodeError as e:
            want = e.start
        assert oracle.utf8_first_error(b) == want, b


def test_invalid_utf8_raises_like_reference(tmp_path):
    p = tmp_path / "bad.txt"
    p.write_bytes(b"hello \xff world")
    with pytest.raises(ValueError, match="invalid UTF-8 at position 6"):
        oracle.train_bpe(p, 300, [])
    with pytest.raises(FileNotFoundError):
        oracle.train_bpe(tmp_path / "nope.txt", 300, [])


def test_empty_and_tiny_inputs(tmp_path):
    p = tmp_path / "e.txt"
    p.write_bytes(b"")
    vocab, merges = oracle.train_bpe(p, 300, ["<|endoftext|>"])
    assert merges == [] and len(vocab) == 257
    p.write_bytes(b"a")
    vocab, merges = oracle.train_bpe(p, 300, ["<|endoftext|>"])
    assert merges == [] and len(vocab) == 257
    vocab, merges = oracle.train_bpe(common.FIXTURES / "corpus.en", 100, ["<|endoftext|>"])
    assert merges == [] and len(vocab) == 257
