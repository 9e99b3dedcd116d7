"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle, the reference's
golden vectors and the reference-generated vectors in tests/golden/.  Bit-exact everywhere."""
from __future__ import annotations

import random
from collections import Counter

import numpy as np
import pytest

import common
from oracle import oracle

pytestmark = pytest.mark.gpu

ALPHABET = list("ab Z'sdmtlvre19!<|>\n\t \r") + [
    "é", "中", "１", " ", " ", "​", "\U0001f643", "", "　", "",
    "́", "\U00016ea0", "'ll", "'ve", " '", "<|endoftext|>", "<|e|>", "don't", " we've",
]


@pytest.fixture(scope="module")
def yabpe():
    import yabpe as y
    from yabpe import _ffi
    _ffi.require_cuda()          # fails loudly when the library or the device is missing
    return y


def _oracle_counts(data: bytes, specials, mode="train", chunk_size=1 << 30):
    toks = oracle.pretokenize(data, specials, mode, chunk_size)
    if mode == "encode":
        _, kinds = oracle.pretokenize_spans(data, specials, mode, chunk_size)
        toks = [t for t, k in zip(toks, kinds) if k < 0]
    return dict(Counter(toks))


def _device_counts(data, specials, mode="train", chunk_size=1 << 30):
    from yabpe.trainer import pretoken_counts
    d = pretoken_counts(data, specials, mode=mode, chunk_size_bytes=chunk_size)
    n = d.pop("__n_pretokens__", 0)
    assert n == sum(d.values())
    return d


# ------------------------------------------------------------------------------- pre-tokeniser
def test_pretok_counts_fixtures(yabpe):
    for name in ["corpus.en", "tinystories_sample.txt", "address.txt", "german.txt",
                 "special_token_trailing_newlines.txt", "special_token_double_newlines_non_whitespace.txt"]:
        data = (common.FIXTURES / name).read_bytes()
        for sp in ([], ["<|endoftext|>"]):
            assert _device_counts(data, sp) == _oracle_counts(data, sp), (name, sp)
    data = (common.DATA / "unicode.txt").read_bytes()
    assert _device_counts(data, []) == _oracle_counts(data, [])


def test_pretok_counts_golden_cases_concatenated(yabpe):
    """All reference-generated pre-tokenisation vectors, batched as independent texts via cuts."""
    from yabpe.trainer import pretoken_counts
    by_key = {}
    for c in common.load_pretok_cases():
        by_key.setdefault((c["mode"], tuple(c["specials"])), []).append(c)
    for (mode, sp), cases in by_key.items():
        blobs = [c["text"].encode("utf-8") for c in cases if c["text"]]
        want = Counter()
        for c in cases:
            toks = [t.encode("utf-8") for t in c["tokens"]]
            if mode == "encode":
                toks = [t for t in toks if t.decode() not in sp]
            want.update(toks)
        data = b"".join(blobs)
        cuts = np.cumsum([len(b) for b in blobs])[:-1].tolist()
        got = pretoken_counts(data, list(sp), mode=mode, cuts=cuts)
        got.pop("__n_pretokens__")
        assert got == dict(want), (mode, sp)


def test_pretok_counts_fuzz_and_adversarial(yabpe):
    rng = random.Random(123)
    for sp in ([], ["<|endoftext|>"], ["<|e|>", "<|endoftext|>"], [" <", "<|e|>"], ["\nb", "ab", "a"]):
        text = "".join(rng.choice(ALPHABET) for _ in range(60000)).encode("utf-8")
        assert _device_counts(text, sp) == _oracle_counts(text, sp), sp
        for cs in (97, 4096):
            assert _device_counts(text[:30000], sp, chunk_size=cs) == _oracle_counts(text[:30000], sp, chunk_size=cs), (sp, cs)
    adv = common.synth_adversarial(300_000)
    for sp in ([], ["<|endoftext|>"]):
        assert _device_counts(adv, sp) == _oracle_counts(adv, sp)
    for sp in (["<|endoftext|>"], ["<|e|>", "<|endoftext|>", "<|endoftext|><|endoftext|>"], ["a", "ab", " "]):
        text = "".join(rng.choice(ALPHABET) for _ in range(40000)).encode("utf-8")
        assert _device_counts(text, sp, mode="encode") == _oracle_counts(text, sp, mode="encode"), sp


def test_prefix_related_specials_at_hard_boundaries(yabpe, tmp_path):
    """every re-match of a recognised special must stop at the same hard boundary as its recognition did
    ('<|eot|>' + CUT + 'x' is not '<|eot|>x').  Reference-generated vectors; documents batched through cuts."""
    from yabpe.trainer import pretoken_counts
    d = common.load_prefix_special_cases()
    for c in d["pretok"]:
        blobs = [x.encode("utf-8") for x in c["docs"]]
        want = Counter(t.encode("utf-8") for toks in c["tokens"] for t in toks
                       if not (c["mode"] == "encode" and t in c["specials"]))
        cuts = np.cumsum([len(b) for b in blobs])[:-1].tolist()
        for generic in (False, True):
            got = pretoken_counts(b"".join(blobs), c["specials"], mode=c["mode"], cuts=cuts, generic_only=generic)
            got.pop("__n_pretokens__")
            assert got == dict(want), (c["mode"], c["specials"], c["docs"], generic)
    for c in d["train"]:
        paths = []
        for i, blob in enumerate(c["inputs"]):
            p = tmp_path / f"pf_{i}.txt"
            p.write_bytes(blob)
            paths.append(p)
        cfg = yabpe.BBPETrainerConfig(vocab_size=c["vocab_size"], min_frequency=c["min_frequency"], max_workers=1,
                                      chunk_size_bytes=c["chunk_size"], special_tokens=c["specials"])
        model = yabpe.BBPETrainer(cfg).train(paths)
        assert model.merges == c["merges_b"], (c["specials"], c["chunk_size"])
        assert {v: k for k, v in model.vocab.items()} == c["vocab_b"]
    v, m = d["encode_model"]
    for c in d["encode_docs"]:
        t = yabpe.Tokenizer(v, m, c["specials"])
        assert t.encode_batch(c["docs"]) == c["ids"], c["docs"]
        assert list(t.encode_iterable(c["docs"])) == [i for ids in c["ids"] for i in ids]
        assert [t.encode(x) for x in c["docs"]] == c["ids"]


def test_pretok_fast_path_large_fuzz(yabpe):
    """Interior tiles take the register-resident fast path: hammer it with every event kind at every alignment."""
    rng = random.Random(321)
    heavy = ALPHABET + ["'s", "'t", "'re", " 'll", "\u00a0", "\u3000", "\u2003", "x'", "''", "<|endoftext|>'s", "é's"]
    for sp, mode in (([], "train"), (["<|endoftext|>"], "train"), (["<|e|>", "<|endoftext|>"], "train"),
                     (["<|endoftext|>"], "encode"), (["<|e|>", "<|endoftext|>", "<|endoftext|><|endoftext|>"], "encode")):
        text = "".join(rng.choice(heavy) for _ in range(400_000)).encode("utf-8")
        assert _device_counts(text, sp, mode=mode) == _oracle_counts(text, sp, mode=mode), (sp, mode)
    # mostly-ASCII prose with sparse events
    words = ["the", "a", "don't", "we've", "I'll", "it's", "naïve", "café", "中文", "x", "1234", "...", "\n", "\n\n", " ", "  "]
    text = " ".join(rng.choice(words) for _ in range(500_000)).encode("utf-8")
    assert _device_counts(text, ["<|endoftext|>"]) == _oracle_counts(text, ["<|endoftext|>"])


def test_pretok_warp_kernel_vs_generic_kernel(yabpe):
    """Trainer mode runs the warp-autonomous kernel on interior 992-byte chunks and the generic tile kernel on
    the rest; stages bit 3 forces the generic kernel everywhere.  Both must give the oracle's table, and the
    warp kernel must actually have run (cache hits > 0, only a handful of boundary work items)."""
    from yabpe.trainer import pretoken_counts
    rng = random.Random(99)
    heavy = ALPHABET + ["'s", "'t", "'re", " the", " and", " of", "\n\n", ". ", "<|endoftext|>\n", "\u00a0", "tion", "ing"]
    text = "".join(rng.choice(heavy) for _ in range(700_000)).encode("utf-8")
    for sp, cs in (([], 1 << 30), (["<|endoftext|>"], 1 << 30), (["<|endoftext|>"], 200_003), (["<|e|>", "<|endoftext|>"], 1 << 30)):
        st = {}
        fast = pretoken_counts(text, sp, chunk_size_bytes=cs, stats_out=st)
        slow = pretoken_counts(text, sp, chunk_size_bytes=cs, generic_only=True)
        assert fast == slow, (sp, cs)
        assert st["cache_hits"] > 0 and 0 < st["slow_items"] <= 4 * (len(text) // cs) + 16, st
        fast.pop("__n_pretokens__")
        assert fast == _oracle_counts(text, sp, chunk_size=cs), (sp, cs)
    # pre-tokens of 15..100 bytes everywhere (long-table path of the warp kernel), and chunk-straddling ones
    words = ["x" * n for n in (15, 16, 17, 31, 32, 33, 40, 64, 100)] + ["ab", "the", "1234567890123456"]
    text = " ".join(rng.choice(words) for _ in range(120_000)).encode()
    assert _device_counts(text, ["<|endoftext|>"]) == _oracle_counts(text, ["<|endoftext|>"])


def test_pretok_long_tokens_and_tile_edges(yabpe):
    """Tokens straddling tile boundaries, longer than the tile window, and MB-long runs."""
    parts = [b"x" * 8191, b" ", b"y" * 9000, b"\n", "中".encode() * 5000, b" 1234567890" * 3, b"!" * 20000, b" ",
             b"z" * 300, b" ", b"z" * 300, b" ", b"q" * 1_200_000, b" tail", b" ", b"q" * 1_200_000]
    data = b"".join(parts)
    assert _device_counts(data, ["<|endoftext|>"]) == _oracle_counts(data, ["<|endoftext|>"])


def test_invalid_utf8_position(yabpe, tmp_path):
    p = tmp_path / "bad.txt"
    for blob, pos in [(b"hello \xff world", 6), (b"a" * 9000 + b"\xe4\xb8" + b"b" * 100, 9000), (b"ok \xc3", 3),
                      (b"\x80abc", 0), (b"x" * 8190 + b"\xf0\x9f\x99" + b"y", 8190)]:
        p.write_bytes(blob)
        with pytest.raises(ValueError, match=f"invalid UTF-8 at position {pos}\\."):
            yabpe.train_bpe(p, 300, [])
    with pytest.raises(FileNotFoundError):
        yabpe.train_bpe(tmp_path / "missing.txt", 300, [])
    with pytest.raises(ValueError):
        yabpe.BBPETrainer().train([])


# ------------------------------------------------------------------------------- trainer
def test_train_corpus_en_matches_reference_fixture(yabpe):
    """The reference's own golden test (tests/test_train_bpe_gpt2.py:27-62)."""
    vocab, merges = yabpe.train_bpe(common.FIXTURES / "corpus.en", 500, ["<|endoftext|>"])
    ref = common.reference_merges_corpus_en()
    assert merges == ref
    assert set(vocab.keys()) == set(range(500))
    assert set(vocab.values()) == {bytes([i]) for i in range(256)} | {b"<|endoftext|>"} | {a + b for a, b in ref}


@pytest.mark.parametrize("streamed", [False, True])
def test_train_golden_cases(yabpe, tmp_path, streamed):
    """streamed=True forces train(files) through the pinned-staging upload (pieces of 1 000 bytes, so every file spans
    several pieces and chunk cuts fall anywhere inside them); the default path reads small files whole."""
    for c in common.load_train_cases():
        paths = []
        for i, blob in enumerate(c["inputs"]):
            p = tmp_path / f"in_{i}.txt"
            p.write_bytes(blob)
            paths.append(p)
        cfg = yabpe.BBPETrainerConfig(vocab_size=c["vocab_size"], min_frequency=c["min_frequency"], max_workers=1,
                                      chunk_size_bytes=c["chunk_size"], special_tokens=c["specials"])
        tr = yabpe.BBPETrainer(cfg)
        if streamed:
            tr.stream_min_bytes, tr.stream_piece_bytes = 0, 1000
        model = tr.train(paths)
        assert model.merges == c["merges_b"], c["name"]
        assert {v: k for k, v in model.vocab.items()} == c["vocab_b"], c["name"]


@pytest.mark.parametrize("kind,size,vocab", [("tinystories", 3_000_000, 2000), ("owt", 3_000_000, 3000),
                                             ("adversarial", 400_000, 1500)])
def test_train_synthetic_vs_oracle(yabpe, tmp_path, kind, size, vocab):
    gen = {"tinystories": common.synth_tinystories, "owt": common.synth_owt, "adversarial": common.synth_adversarial}[kind]
    data = gen(size)
    p = tmp_path / "c.txt"
    p.write_bytes(data)
    want_vocab, want_merges = oracle.train_bpe(p, vocab, ["<|endoftext|>"], fast=True)
    got_vocab, got_merges = yabpe.train_bpe(p, vocab, ["<|endoftext|>"])
    assert got_merges == want_merges
    assert got_vocab == want_vocab


def test_train_medium_corpus_vs_oracle_and_deterministic(yabpe, tmp_path):
    """Big enough for many grid-mode merges, index rebuilds and threshold steps; run twice (races show up
    as run-to-run differences) and check against the oracle."""
    data = common.synth_owt(24_000_000, seed=77)
    p = tmp_path / "m.txt"
    p.write_bytes(data)
    want = oracle.train_bpe(p, 2500, ["<|endoftext|>"], fast=True)
    got1 = yabpe.train_bpe(p, 2500, ["<|endoftext|>"])
    tr = yabpe.BBPETrainer(yabpe.BBPETrainerConfig(vocab_size=2500, min_frequency=1, max_workers=1, chunk_size_bytes=1 << 30,
                                                   special_tokens=["<|endoftext|>"]))
    tr.stream_min_bytes, tr.stream_piece_bytes = 0, 5_000_001          # second run: the streamed upload (5 pieces)
    m2 = tr.train([p])
    got2 = ({v: k for k, v in m2.vocab.items()}, m2.merges)
    assert got1[1] == got2[1]
    assert got1[1] == want[1]
    assert got1[0] == want[0]


@pytest.mark.parametrize("layout", ["interleaved", "split"])
def test_short_table_layouts(yabpe, tmp_path, monkeypatch, layout):
    """The pre-token table has two layouts (32-byte {key, count} slots for DRAM-sized tables, separate key / count
    arrays for small hot ones, chosen by capacity): every consumer -- counting, compaction, training, the multi-GPU
    re-insert kernel's table, encode lookups -- must give the same results with either."""
    monkeypatch.setenv("YABPE_SHORT_LAYOUT", layout)
    rng = random.Random(7)
    text = "".join(rng.choice(ALPHABET + [" the", " of", "ing", "\n"]) for _ in range(150_000)).encode("utf-8")
    for sp, mode in (([], "train"), (["<|endoftext|>"], "train"), (["<|endoftext|>"], "encode")):
        assert _device_counts(text, sp, mode=mode) == _oracle_counts(text, sp, mode=mode), (sp, mode)
    data = common.synth_owt(1_000_000, seed=3)
    p = tmp_path / "l.txt"
    p.write_bytes(data)
    assert yabpe.train_bpe(p, 1200, ["<|endoftext|>"]) == oracle.train_bpe(p, 1200, ["<|endoftext|>"], fast=True)
    v, m = common.gpt2_vocab_and_merges()
    t = yabpe.Tokenizer(v, m, ["<|endoftext|>"])
    o = oracle.Tokenizer(v, m, ["<|endoftext|>"])
    s = data[:300_000].decode("utf-8", errors="ignore")
    assert t.encode(s) == o.encode(s)
    assert t.encode_batch([s[:1000], "", s[1000:5000]]) == [o.encode(s[:1000]), [], o.encode(s[1000:5000])]


@pytest.mark.parametrize("slots", [64, 4096])
def test_hot_table_in_front_of_the_big_table(yabpe, tmp_path, monkeypatch, slots):
    """DRAM-sized count tables get an L2-resident direct-mapped table in front of them (k_pretok_warp<true>, k_hot_flush).
    Forced here on small inputs with a tiny hot table (every slot fought over) and a moderate one; counts, training and
    encode must not move.  Also the piece-wise counting of train_from_buffers, which reuses one hot table across calls."""
    from yabpe import engine
    monkeypatch.setenv("YABPE_SHORT_LAYOUT", "interleaved")
    monkeypatch.setenv("YABPE_HOT_TABLE", "1")
    monkeypatch.setattr(engine, "_HOT_TABLE_SLOTS", slots)
    monkeypatch.setattr(engine, "_HOT_TABLE_MIN_FACTOR", 0)
    rng = random.Random(11)
    text = "".join(rng.choice(ALPHABET + [" the", " of", "ing", "\n"]) for _ in range(200_000)).encode("utf-8")
    for sp, mode in (([], "train"), (["<|endoftext|>"], "train"), (["<|endoftext|>"], "encode")):
        assert _device_counts(text, sp, mode=mode) == _oracle_counts(text, sp, mode=mode), (sp, mode)
    data = common.synth_owt(6_000_000, seed=31)
    p = tmp_path / "hot.txt"
    p.write_bytes(data)
    want = oracle.train_bpe(p, 1500, ["<|endoftext|>"], fast=True)
    assert yabpe.train_bpe(p, 1500, ["<|endoftext|>"]) == want
    tr = yabpe.BBPETrainer(yabpe.BBPETrainerConfig(vocab_size=1500, min_frequency=1, max_workers=1, chunk_size_bytes=1 << 30,
                                                   special_tokens=["<|endoftext|>"]))
    tr.pipeline_min_bytes, tr.pipeline_piece_bytes = 0, 1 << 20
    m = tr.train_from_buffers([np.frombuffer(data, dtype=np.uint8)])
    assert ({v: k for k, v in m.vocab.items()}, m.merges) == want
    v, mg = common.gpt2_vocab_and_merges()
    s = data[:400_000].decode("utf-8", errors="ignore")
    assert yabpe.Tokenizer(v, mg, ["<|endoftext|>"]).inner.encode_device(*engine.to_device_text(__import__("torch"), np.frombuffer(s.encode(), dtype=np.uint8)))[0].cpu().tolist() \
        == oracle.Tokenizer(v, mg, ["<|endoftext|>"]).encode(s)


def test_train_with_frequent_index_rebuilds(yabpe, tmp_path, monkeypatch):
    """The pair -> words index is rebuilt every `rebuild_every` merges (engine.rebuild_period); force a tiny period so
    that dozens of rebuilds (and leader exits for them) happen, and a zero period (rebuild only when the log is full)."""
    data = common.synth_tinystories(2_000_000, seed=5)
    p = tmp_path / "r.txt"
    p.write_bytes(data)
    want = oracle.train_bpe(p, 1800, ["<|endoftext|>"], fast=True)
    for period in ("37", "0"):
        monkeypatch.setenv("YABPE_REBUILD_EVERY", period)
        tr = yabpe.BBPETrainer(yabpe.BBPETrainerConfig(vocab_size=1800, min_frequency=1, max_workers=1, chunk_size_bytes=1 << 30,
                                                       special_tokens=["<|endoftext|>"]))
        model = tr.train([p])
        assert model.merges == want[1], period
        assert {v: k for k, v in model.vocab.items()} == want[0], period
        if period == "37":
            assert tr.last_stats.index_rebuilds >= 30


def test_tables_grow_when_the_estimate_is_too_small(yabpe, tmp_path, monkeypatch):
    """The pre-token tables are sized from an extrapolated estimate (engine.estimate_table_sizes: twice the expected number of
    unique pre-tokens); an estimate that is far too small must be detected (ST_TABLE_FULL) and the count repeated with larger
    tables -- same result."""
    from yabpe import engine
    data = common.synth_owt(3_000_000, seed=21)
    p = tmp_path / "g.txt"
    p.write_bytes(data)
    want = oracle.train_bpe(p, 1500, ["<|endoftext|>"], fast=True)
    calls = []
    monkeypatch.setattr(engine, "estimate_table_sizes", lambda *a, **k: (calls.append(1), (1 << 10, 1 << 6, False, None))[1])
    assert yabpe.train_bpe(p, 1500, ["<|endoftext|>"]) == want
    assert calls


@pytest.mark.parametrize("mode,what", [("1", "one merge per iteration everywhere (trainer.py:241-300 as written)"),
                                       ("264", "batches in leader mode only"),
                                       ("2049", "batches in grid mode only"),
                                       ("65552", "no leader mode: every merge in the grid, batched"),
                                       ("131088", "big batches stay with the leader"),
                                       ("2", "pairs"), ("0", "default")])
def test_train_batched_merges(yabpe, tmp_path, monkeypatch, mode, what):
    """The merge loop takes up to 31 pairs per iteration when it can prove that the sequential loop would pick exactly these,
    in this order (csrc/merge.cuh, "batched leader merges").  Every combination of where batching is allowed must give the
    oracle's merges; the default must actually batch on a corpus of this kind."""
    monkeypatch.setenv("YABPE_BATCH_MAX", mode)
    if not hasattr(test_train_batched_merges, "_want"):
        data = common.synth_owt(24_000_000, seed=78, n_types=200_000)
        test_train_batched_merges._data = data
        p0 = tmp_path / "b0.txt"
        p0.write_bytes(data)
        test_train_batched_merges._want = oracle.train_bpe(p0, 6000, ["<|endoftext|>"], fast=True)
    want = test_train_batched_merges._want
    p = tmp_path / "b.txt"
    p.write_bytes(test_train_batched_merges._data)
    tr = yabpe.BBPETrainer(yabpe.BBPETrainerConfig(vocab_size=6000, min_frequency=1, max_workers=1, chunk_size_bytes=1 << 30,
                                                   special_tokens=["<|endoftext|>"]))
    model = tr.train([p])
    assert model.merges == want[1], what
    assert {v: k for k, v in model.vocab.items()} == want[0], what
    st = tr.last_stats
    if mode == "1":
        assert st.batched_merges == 0 and st.grid_batched_merges == 0
    if mode == "0":
        assert st.batched_merges + st.grid_batched_merges > st.n_merges // 4, st
    if mode == "65552":
        assert st.leader_merges == 0


def test_train_with_prefetch_helpers_and_tie_regime(yabpe, tmp_path, monkeypatch):
    """(1) The leader's prefetch helpers (idle CTAs pulling the next merges' words into the L2) only run on word arrays
    beyond the L2 size; YABPE_HELPER_MIN_SYMS=-1 forces them on a small corpus -- results must not move.
    (2) A small corpus trained to exhaustion spends most merges in the massive-tie regime (thousands of pairs share the
    maximum count, the top list stays disabled between retries): every merge there is decided by the byte-wise tie-break."""
    monkeypatch.setenv("YABPE_HELPER_MIN_SYMS", "-1")
    data = common.synth_owt(3_000_000, seed=11)
    p = tmp_path / "h.txt"
    p.write_bytes(data)
    want = oracle.train_bpe(p, 3000, ["<|endoftext|>"], fast=True)
    for mode in ("1", "2"):
        monkeypatch.setenv("YABPE_HELPER_MODE", mode)
        assert yabpe.train_bpe(p, 3000, ["<|endoftext|>"]) == want, mode
    monkeypatch.delenv("YABPE_HELPER_MIN_SYMS")
    monkeypatch.delenv("YABPE_HELPER_MODE")
    small = common.synth_owt(200_000, seed=12)
    p.write_bytes(small)
    got = yabpe.train_bpe(p, 32000, ["<|endoftext|>"])
    want = oracle.train_bpe(p, 32000, ["<|endoftext|>"], fast=True)
    assert len(want[1]) > 10000 and got == want


def test_train_counted_while_uploaded(yabpe, tmp_path):
    """train_from_buffers on large host buffers counts piece k while piece k + 1 is still being uploaded (pieces cut at the
    safe edges of yabpe/sharding.py, all into one table set): forced here with 1 MiB pieces on two buffers, reference chunk
    cuts inside pieces, a 30 000-byte pre-token across a piece edge and back-to-back specials."""
    blobs = [common.synth_owt(9_000_000, seed=21) + b" " + b"z" * 30000 + b"\n" + common.synth_tinystories(2_000_000, seed=22),
             (b"a<|endoftext|><|endoftext|>b !<|endoftext|>\n" * 3000) + common.synth_adversarial(1_500_000, seed=23)]
    for chunk in (1 << 30, 2_000_003):
        cfg = yabpe.BBPETrainerConfig(vocab_size=1800, min_frequency=1, max_workers=1, chunk_size_bytes=chunk,
                                      special_tokens=["<|endoftext|>"])
        tr = yabpe.BBPETrainer(cfg)
        tr.pipeline_min_bytes, tr.pipeline_piece_bytes = 0, 1 << 20
        model = tr.train_from_buffers([np.frombuffer(b, dtype=np.uint8) for b in blobs])
        o = oracle.Trainer(["<|endoftext|>"])
        for b in blobs:
            o.feed_bytes(b, chunk)
        vocab, merges = o.run(1800, 1, True)
        assert model.merges == merges, chunk
        assert {v: k for k, v in model.vocab.items()} == vocab, chunk
        assert tr.last_stats.n_pretokens == o.num_pretokens
    # invalid UTF-8 is still reported with its file and offset
    bad = bytearray(common.synth_owt(3_000_000, seed=24)); bad[2_500_000] = 0xFF
    tr = yabpe.BBPETrainer(cfg)
    tr.pipeline_min_bytes, tr.pipeline_piece_bytes = 0, 1 << 20
    with pytest.raises(ValueError, match="invalid UTF-8 at position 2500000"):
        tr.train_from_buffers([np.frombuffer(bytes(bad), dtype=np.uint8)], ["bad.txt"])


def test_train_edge_cases(yabpe, tmp_path):
    p = tmp_path / "e.txt"
    p.write_bytes(b"")
    vocab, merges = yabpe.train_bpe(p, 300, ["<|endoftext|>"])
    assert merges == [] and len(vocab) == 257
    p.write_bytes(b"a")
    assert yabpe.train_bpe(p, 300, ["<|endoftext|>"]) == oracle.train_bpe(p, 300, ["<|endoftext|>"])
    p.write_bytes(b"a" * 63)
    assert yabpe.train_bpe(p, 300, []) == oracle.train_bpe(p, 300, [])
    assert yabpe.train_bpe(common.FIXTURES / "corpus.en", 100, ["<|endoftext|>"])[1] == []
    # the special's own bytes are merged like a word and re-created without a new id
    got = yabpe.train_bpe(common.FIXTURES / "tinystories_sample.txt", 1000, ["<|endoftext|>"])
    want = oracle.train_bpe(common.FIXTURES / "tinystories_sample.txt", 1000, ["<|endoftext|>"], fast=True)
    assert got == want
    assert any(a + b == b"<|endoftext|>" for a, b in got[1])


# ------------------------------------------------------------------------------- tokenizer
def test_encode_golden_cases(yabpe):
    models, cases = common.load_encode_cases()
    toks = {}
    for c in cases:
        key = (c["model"], tuple(c["specials"]))
        if key not in toks:
            v, m = models[c["model"]]
            toks[key] = yabpe.Tokenizer(v, m, c["specials"])
        t = toks[key]
        assert t.encode(c["text"]) == c["ids"], c["text"][:60]
        assert t.decode(c["ids"]) == c["decoded"]


def test_encode_fixtures_vs_oracle_and_tiktoken(yabpe):
    v, m = common.gpt2_vocab_and_merges()
    t = yabpe.Tokenizer(v, m, ["<|endoftext|>"])
    o = oracle.Tokenizer(v, m, ["<|endoftext|>"])
    texts = {}
    for name in ["address.txt", "german.txt", "tinystories_sample.txt", "corpus.en"]:
        with open(common.FIXTURES / name) as f:
            texts[name] = f.read()
    for name, text in texts.items():
        ids = t.encode(text)
        assert ids == o.encode(text), name
        assert t.decode(ids) == text
    try:
        import tiktoken
    except ImportError:
        return
    enc = tiktoken.Encoding("gpt2-local", pat_str=r"""'(?:[sdmt]|ll|ve|re)| ?\p{L}+| ?\p{N}+| ?[^\s\p{L}\p{N}]+|\s+(?!\S)|\s+""",
                            mergeable_ranks={b: i for i, b in v.items() if i < 50256},
                            special_tokens={"<|endoftext|>": 50256})
    for name, text in texts.items():
        assert t.encode(text) == enc.encode(text, allowed_special={"<|endoftext|>"}), name


def test_encode_iterable_and_batch(yabpe):
    v, m = common.gpt2_vocab_and_merges()
    t = yabpe.Tokenizer(v, m, ["<|endoftext|>"])
    o = oracle.Tokenizer(v, m, ["<|endoftext|>"])
    with open(common.FIXTURES / "tinystories_sample.txt") as f:
        lines = f.readlines()
    lines += ["", "\n", "<|endoftext|>", " ", "a" * 500, "x<|endoftext|>"]
    assert list(t.encode_iterable(lines)) == list(o.encode_iterable(lines))
    assert t.encode_batch(lines) == [o.encode(x) for x in lines]
    assert t.encode("") == [] and t.decode([]) == ""


def test_encode_synthetic_and_long_words(yabpe):
    v, m = common.gpt2_vocab_and_merges()
    t = yabpe.Tokenizer(v, m, ["<|endoftext|>"])
    o = oracle.Tokenizer(v, m, ["<|endoftext|>"])
    text = common.synth_owt(1_500_000).decode("utf-8")
    assert t.encode(text) == o.encode(text)
    adv = common.synth_adversarial(120_000).decode("utf-8")
    assert t.encode(adv) == o.encode(adv)
    long_text = "a" * 3000 + " " + "ab" * 2500 + " " + "the" * 1000 + "中文" * 700
    assert t.encode(long_text) == o.encode(long_text)


def test_encode_pinned_streams_pieces_exactly(yabpe):
    """encode_pinned (host bytes in, host ids out, pieces pipelined over three streams) == encode of the whole text,
    for piece sizes from a few documents to everything, pinned or pageable input, given or grown output buffer."""
    import torch
    v, m = common.gpt2_vocab_and_merges()
    t = yabpe.Tokenizer(v, m, ["<|endoftext|>"]).inner
    o = oracle.Tokenizer(v, m, ["<|endoftext|>"])
    raw = common.synth_owt(3_000_000, seed=11) + b"<|endoftext|><|endoftext|> tail"
    want = o.encode(raw.decode("utf-8"))
    host = torch.frombuffer(bytearray(raw), dtype=torch.uint8)
    pinned = host.pin_memory()
    for piece in (20_000, 300_000, 1_000_000, 1 << 30):
        assert t.encode_pinned(pinned, piece_bytes=piece).tolist() == want, piece
    out = torch.empty(len(want) + 5, dtype=torch.int32).pin_memory()
    got = t.encode_pinned(pinned, out=out, piece_bytes=250_000)
    assert got.data_ptr() == out.data_ptr() and got.tolist() == want
    small = torch.empty(1000, dtype=torch.int32).pin_memory()          # too small: a larger buffer is allocated
    assert t.encode_pinned(host, out=small, piece_bytes=250_000).tolist() == want
    assert t.encode_pinned(torch.empty(0, dtype=torch.uint8)).numel() == 0
    # ids as uint16 (the GPT-2 vocabulary fits): narrowed on the device, half the download
    got16 = t.encode_pinned(pinned, piece_bytes=300_000, id_dtype=torch.uint16)
    assert got16.dtype == torch.uint16 and got16.to(torch.int32).tolist() == want
    big = yabpe.Tokenizer({**v, 70000: b"\xff\xfe"}, m, ["<|endoftext|>"]).inner
    with pytest.raises(ValueError):
        big.encode_pinned(pinned, id_dtype=torch.uint16)
    assert t.encode(raw.decode("utf-8")) == want                        # the plain path is untouched by the streams


def test_encode_roundtrip_property_large(yabpe):
    """Size-independent property at a larger size: decode(encode(x)) == x, ids are valid."""
    v, m = common.gpt2_vocab_and_merges()
    t = yabpe.Tokenizer(v, m, ["<|endoftext|>"])
    text = common.synth_owt(8_000_000, seed=7).decode("utf-8")
    ids = t.encode(text)
    assert t.decode(ids) == text
    assert min(ids) >= 0 and max(ids) < 50257


# ------------------------------------------------------------------------------- full-size properties
def test_train_properties_at_bench_scale(yabpe):
    """Size-independent properties on 256 MB of the bench generator (no reference chunk cut inside; the 1.2 GB case that
    crosses a real 2^30 cut is compared with the oracle in tests/test_gpu_baseline_sizes.py): conservation of pre-token
    occurrences, run-to-run determinism, dense ids, every merge result present, and agreement of the sharded table
    with the whole."""
    import sys
    import torch
    sys.path.insert(0, str(common.ROOT / "tools"))
    from synth_gpu import synth_corpus_device
    from yabpe import _ffi, engine
    text, n = synth_corpus_device(torch, 256_000_000, "tinystories", 4242)
    cfg = yabpe.BBPETrainerConfig(vocab_size=3000, min_frequency=1, max_workers=1, chunk_size_bytes=1 << 30,
                                  special_tokens=["<|endoftext|>"])
    m1 = yabpe.BBPETrainer(cfg).train_device(text, n)
    tr2 = yabpe.BBPETrainer(cfg)
    m2 = tr2.train_device(text, n)
    assert m1.merges == m2.merges and m1.vocab == m2.vocab
    assert sorted(m1.vocab.values()) == list(range(len(m1.vocab)))
    assert len(m1.merges) == 3000 - 257
    for a, b in m1.merges:
        assert a in m1.vocab and b in m1.vocab and a + b in m1.vocab
    # conservation: the table of the whole text == the sum of the tables of two owned halves (the shard edge is NOT a cut)
    sp = [b"<|endoftext|>"]
    whole, st = engine.pretok_count_checked(torch, text, n, None, sp, 0)
    half = (n // 2) | 7
    tot = 0
    for own in ((0, half), (half, n)):
        part, st_p = engine.pretok_count_checked(torch, text, n, None, sp, 0, own=own)
        tot += int(st_p[_ffi.ST_NTOK])
    assert tot == int(st[_ffi.ST_NTOK]) == tr2.last_stats.n_pretokens
    # the generic kernel alone gives the same statistics as the warp kernel + boundary list
    gen, st_g = engine.pretok_count_checked(torch, text, n, None, sp, 0, generic_only=True)
    for k in (_ffi.ST_NTOK, _ffi.ST_UNIQ_SHORT, _ffi.ST_UNIQ_LONG, _ffi.ST_UNIQ_BYTES, _ffi.ST_NSPECIAL):
        assert int(st_g[k]) == int(st[k]), k
    assert int(st[_ffi.ST_CACHE_HIT]) > 0 and int(st_g[_ffi.ST_CACHE_HIT]) == 0


def test_train_adversarial_large_vocab(yabpe, tmp_path):
    """BASELINE.json configs[4] at a size the oracle still finishes: tie farms, dense specials, long runs, odd
    whitespace -- many merges, so the late low-count phase (ties decided by token bytes) is covered."""
    data = common.synth_adversarial(1_500_000, seed=99)
    p = tmp_path / "adv.txt"
    p.write_bytes(data)
    want = oracle.train_bpe(p, 6000, ["<|endoftext|>"], fast=True)
    got = yabpe.train_bpe(p, 6000, ["<|endoftext|>"])
    assert got[1] == want[1]
    assert got[0] == want[0]
