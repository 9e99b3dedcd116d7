"""SURVEY 8(f) rows 1-2 on the GPU: BBPETrainer._preprocess_corpus / _merge_loop(sequences) (the private entry
points tests/test_trainer.py of the reference calls) and a save() -> from_file() -> encode round trip, against the
reference's outputs (tests/golden/class_api_cases.json) and the CPU oracle."""
from __future__ import annotations

import random
from collections import Counter

import pytest

import common
from oracle import oracle

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def yabpe():
    import yabpe as y
    from yabpe import _ffi
    _ffi.require_cuda()
    return y


def _write(tmp_path, blobs, tag):
    paths = []
    for i, b in enumerate(blobs):
        p = tmp_path / f"{tag}_{i}.txt"
        p.write_bytes(b)
        paths.append(p)
    return paths


def _cfg(yabpe, **kw):
    kw.setdefault("max_workers", 1)
    return yabpe.BBPETrainerConfig(**kw)


# ------------------------------------------------------------------------------- _preprocess_corpus
def test_preprocess_corpus_golden(yabpe, tmp_path):
    for k, c in enumerate(common.load_class_api_cases()["preprocess"]):
        tr = yabpe.BBPETrainer(_cfg(yabpe, chunk_size_bytes=c["chunk_size"], special_tokens=c["specials"]))
        seqs = tr._preprocess_corpus(_write(tmp_path, c["files"], f"g{k}"))
        assert all(isinstance(s, list) for s in seqs)
        assert [bytes(s) for s in seqs] == c["want"], (c["specials"], c["chunk_size"])


def test_preprocess_corpus_order_vs_oracle(yabpe, tmp_path):
    """Text order of the start bitmap (generic per-byte rule) against the oracle's sequential scan, and its multiset
    against the counting kernels (warp + tile) on the same text: three implementations, one answer."""
    from yabpe.trainer import pretoken_counts
    rng = random.Random(77)
    alphabet = list("ab Z'sdmtlvre19!<|>\n\t \r") + ["é", "中", " ", "　", "\U0001f643", "'ll", "don't",
                                                     "<|endoftext|>", "<|e|>", " the", "x" * 40]
    for sp, cs in (([], 1 << 30), (["<|endoftext|>"], 1 << 30), (["<|e|>", "<|endoftext|>"], 4099), (["\nb", "ab", "a"], 513)):
        text = "".join(rng.choice(alphabet) for _ in range(120_000)).encode("utf-8")
        tr = yabpe.BBPETrainer(_cfg(yabpe, chunk_size_bytes=cs, special_tokens=sp))
        got = [bytes(s) for s in tr._preprocess_corpus(_write(tmp_path, [text], "o"))]
        assert b"".join(got) == text
        assert got == oracle.pretokenize(text, sp, "train", cs), (sp, cs)
        counts = pretoken_counts(text, sp, chunk_size_bytes=cs)
        counts.pop("__n_pretokens__")
        assert counts == dict(Counter(got)), (sp, cs)
    for name in ("corpus.en", "tinystories_sample.txt", "german.txt"):
        data = (common.FIXTURES / name).read_bytes()
        tr = yabpe.BBPETrainer(_cfg(yabpe, chunk_size_bytes=1024, special_tokens=["<|endoftext|>"]))
        got = [bytes(s) for s in tr._preprocess_corpus([common.FIXTURES / name])]
        assert got == oracle.pretokenize(data, ["<|endoftext|>"], "train", 1024), name
    # MB-long pre-tokens and 4-byte code points at the end of the text
    data = b"q" * 1_300_000 + b" tail " + "中".encode() * 3000 + b"\n\n" + "\U0001f643".encode()
    tr = yabpe.BBPETrainer(_cfg(yabpe, special_tokens=[]))
    assert [bytes(s) for s in tr._preprocess_corpus(_write(tmp_path, [data], "l"))] == oracle.pretokenize(data, [], "train", 8 << 20)


def test_preprocess_corpus_edge_cases(yabpe, tmp_path):
    tr = yabpe.BBPETrainer(_cfg(yabpe))
    empty, one = _write(tmp_path, [b"", b"Hello world!"], "e")
    assert tr._preprocess_corpus([empty]) == []
    assert tr._preprocess_corpus([empty, one, empty]) == [list(b"Hello"), list(b" world"), list(b"!")]
    assert tr._preprocess_corpus([str(one)]) == [list(b"Hello"), list(b" world"), list(b"!")]
    with pytest.raises(FileNotFoundError):
        tr._preprocess_corpus([tmp_path / "missing.txt"])
    bad = _write(tmp_path, [b"ok \xe4\xb8 broken"], "b")[0]
    with pytest.raises(ValueError, match="invalid UTF-8 at position 3"):
        tr._preprocess_corpus([bad])


# ------------------------------------------------------------------------------- _merge_loop(sequences)
def test_merge_loop_sequences_golden(yabpe):
    for c in common.load_class_api_cases()["merge_loop"]:
        tr = yabpe.BBPETrainer(_cfg(yabpe, **c["config"]))
        vocab, merges = tr._merge_loop([list(s) for s in c["seqs"]])
        assert list(vocab.keys()) == c["want_vocab"] and list(vocab.values()) == list(range(len(vocab))), c["config"]
        assert merges == c["want_merges"], c["config"]
        assert all(isinstance(a, bytes) and isinstance(b, bytes) for a, b in merges)


def test_merge_loop_sequences_vs_oracle_and_train(yabpe, tmp_path):
    """_merge_loop(_preprocess_corpus(files)) is train(files) (trainer.py:75-90), also with min_frequency > 1."""
    data = (common.FIXTURES / "tinystories_sample.txt").read_bytes()[:200_000]
    path = _write(tmp_path, [data], "t")[0]
    for vs, mf, sp in ((700, 1, ["<|endoftext|>"]), (2000, 7, ["<|endoftext|>", "the"])):
        cfg = _cfg(yabpe, vocab_size=vs, min_frequency=mf, chunk_size_bytes=50_000, special_tokens=sp)
        tr = yabpe.BBPETrainer(cfg)
        vocab, merges = tr._merge_loop(tr._preprocess_corpus([path]))
        model = yabpe.BBPETrainer(cfg).train([path])
        assert (vocab, merges) == (model.vocab, model.merges)
        want_vocab, want_merges = oracle.train_bpe(path, vs, sp, min_frequency=mf, chunk_size_bytes=50_000)
        assert merges == want_merges and {v: k for k, v in vocab.items()} == want_vocab
        if mf > 1:
            assert len(merges) < vs - len(tr._init_base_vocab())          # stopped by the frequency floor
    # long words (> 256 symbols) and single-symbol words through the sequence entry point
    rng = random.Random(5)
    seqs = [bytes(rng.choice(b"ab") for _ in range(rng.choice((1, 2, 3, 300, 700)))) for _ in range(400)]
    ot = oracle.Trainer([])
    for w, f in Counter(seqs).items():
        ot.feed_word(w, f)
    want_vocab, want_merges = ot.run(300, 2)
    vocab, merges = yabpe.BBPETrainer(_cfg(yabpe, vocab_size=300, min_frequency=2, special_tokens=[]))._merge_loop([list(s) for s in seqs])
    assert merges == want_merges and {v: k for k, v in vocab.items()} == want_vocab


# ------------------------------------------------------------------------------- save -> from_file -> encode
def test_persisted_model_round_trip(yabpe, tmp_path):
    for i, c in enumerate(common.load_class_api_cases()["persist"]):
        path = _write(tmp_path, [c["input"]], f"p{i}")[0]
        tr = yabpe.BBPETrainer(_cfg(yabpe, vocab_size=c["vocab_size"], min_frequency=1, chunk_size_bytes=1 << 30,
                                    special_tokens=c["specials"]))
        tr.train([path])
        tr.save(tmp_path / f"model{i}")
        for name, want in c["files"].items():                              # trained on the GPU, written like the reference
            assert (tmp_path / f"model{i}" / name).read_bytes() == want, name
        tok = yabpe.BBPETokenizer.from_file(tmp_path / f"model{i}")
        for e in c["encodes"]:
            assert tok.encode(e["text"]) == e["ids"], e["text"]


# ------------------------------------------------------------------------------- decode on the device
def test_decode_device_vs_host_gather(yabpe):
    """yabpe_decode_ids against b"".join(vocab_inv[i] for i in ids if i in vocab_inv) (tokenizer.py:335-339):
    unknown / negative / out-of-range ids are skipped, ids without bytes (gaps in the vocabulary) too."""
    import numpy as np
    import torch
    rng = np.random.default_rng(11)
    vocab, merges = common.gpt2_vocab_and_merges()
    tok = yabpe.Tokenizer(vocab, merges, ["<|endoftext|>"]).inner
    gappy = {i: b for i, b in vocab.items() if i % 7 != 3}                 # ids 3, 10, 17, ... are not in the vocabulary
    gappy[60000] = b"x" * 300                                              # a long token beyond a gap
    tok2 = yabpe.Tokenizer(gappy, [], None).inner
    for t, inv in ((tok, vocab), (tok2, gappy)):
        for n in (1, 2047, 2048, 2049, 300_000):
            ids = rng.integers(-3, 60_010, size=n).astype(np.int32)
            ids[rng.integers(0, n, size=max(1, n // 50))] = 60000
            got = t.decode_device(torch.from_numpy(ids).cuda()).cpu().numpy().tobytes()
            assert got == b"".join(inv[i] for i in ids.tolist() if i in inv), n
    assert tok.decode_device(torch.empty(0, dtype=torch.int32, device="cuda")).numel() == 0
    # nothing to write at all
    assert tok2.decode_device(torch.full((5000,), 3, dtype=torch.int32, device="cuda")).numel() == 0


def test_decode_long_lists_take_the_device_path(yabpe, monkeypatch):
    from yabpe import tokenizer as T
    vocab, merges = common.gpt2_vocab_and_merges()
    tok = yabpe.Tokenizer(vocab, merges, ["<|endoftext|>"])
    text = ("The quick brown fox — naïve café 中文 \U0001f643 don't stop.\r\n\r\n<|endoftext|>" * 3000)
    ids = tok.encode(text)
    assert len(ids) >= T._DECODE_DEVICE_MIN
    calls = []
    orig = T.BBPETokenizer.decode_device
    monkeypatch.setattr(T.BBPETokenizer, "decode_device", lambda self, *a, **k: (calls.append(1), orig(self, *a, **k))[1])
    assert tok.decode(ids) == text and calls
    # truncated inside a multi-byte character: strict decode fails, the whole buffer is re-decoded with replacement
    cut = ids[: len(ids) - 7] + [{b: i for i, b in vocab.items()}[b"\xe4"]]
    host = b"".join(vocab[i] for i in cut)
    assert tok.decode(cut) == host.decode("utf-8", errors="replace")
    assert tok.decode(ids[:50]) == b"".join(vocab[i] for i in ids[:50]).decode("utf-8", errors="replace")
    assert tok.decode([]) == ""


def test_encode_decode_round_trip_on_device(yabpe):
    """Size-independent property at 64 MB: decode(encode(text)) is the text, byte for byte, without leaving the device."""
    import sys
    import torch
    sys.path.insert(0, str(common.ROOT / "tools"))
    from synth_gpu import synth_corpus_device
    vocab, merges = common.gpt2_vocab_and_merges()
    tok = yabpe.Tokenizer(vocab, merges, ["<|endoftext|>"]).inner
    text_dev, n = synth_corpus_device(torch, 64 << 20, "owt", 20260103)
    ids, _ = tok.encode_device(text_dev, n)
    out = tok.decode_device(ids.clone())
    assert out.numel() == n and torch.equal(out, text_dev[:n])


def test_cli_trains_and_saves(yabpe, tmp_path, capsys):
    from yabpe.scripts import train_bpe
    out = tmp_path / "model"
    assert train_bpe.main(["--input", str(common.FIXTURES / "corpus.en"), "--output", str(out), "--vocab-size", "400",
                           "--min-frequency", "2"]) == 0
    assert "Number of merges: 143" in capsys.readouterr().out
    tok = yabpe.BBPETokenizer.from_file(out)
    assert tok.vocab_size == 400 and tok.special_tokens == ["<|endoftext|>"]
    want_vocab, want_merges = oracle.train_bpe(common.FIXTURES / "corpus.en", 400, ["<|endoftext|>"], min_frequency=2,
                                               chunk_size_bytes=20 << 20)
    assert len(want_merges) == 143 and tok.decode(tok.encode("the cat sat")) == "the cat sat"
