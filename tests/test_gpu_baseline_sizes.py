"""Parity against the oracle at BASELINE.json sizes: the CUDA path and the CPU oracle see the
SAME bytes (the bench's GPU generator, copied to the host) and must agree bit for bit.

  * TinyStories-shaped 1.2e9 bytes / vocab 10 000 -- crosses a real 2^30 reference chunk cut (trainer.py:172-198)
  * OWT-shaped 256e6 bytes / vocab 32 000 -- millions of unique pre-tokens, interleaved table layout, non-ASCII
  * adversarial 64e6 bytes / vocab 50 000 (BASELINE configs[4]: tie farms, dense specials, long runs)
  * GPT-2 encode of a 256e6-byte slice of OWT-shaped text

The oracle runs in its `fast` mode (heap instead of the reference's linear max(): same answers, tests/test_oracle_golden.py
pins both modes to each other and to the reference).  Slow (about a minute of host time) but run by default with `-m gpu`.
"""
from __future__ import annotations

import sys

import numpy as np
import pytest

import common
from oracle import oracle

pytestmark = pytest.mark.gpu
SP = ["<|endoftext|>"]


@pytest.fixture(scope="module")
def yabpe():
    import yabpe as y
    return y


def _gen(kind: str, nbytes: int, seed: int):
    import torch
    sys.path.insert(0, str(common.ROOT / "tools"))
    from synth_gpu import synth_corpus_device
    text, n = synth_corpus_device(torch, nbytes, kind, seed)
    return torch, text, n


def _check_train(yabpe, text, n, host: bytes, vocab: int, label: str):
    cfg = yabpe.BBPETrainerConfig(vocab_size=vocab, min_frequency=1, max_workers=1, chunk_size_bytes=1 << 30, special_tokens=SP)
    tr = yabpe.BBPETrainer(cfg)
    model = tr.train_device(text, n)
    o = oracle.Trainer(SP)
    o.feed_bytes(host, 1 << 30)
    want_vocab, want_merges = o.run(vocab, 1, True)
    assert tr.last_stats.n_pretokens == o.num_pretokens, label
    assert len(model.merges) == len(want_merges), label
    for i, (g, w) in enumerate(zip(model.merges, want_merges)):
        assert g == w, f"{label}: merge {i} differs: {g!r} != {w!r}"
    assert {v: k for k, v in model.vocab.items()} == want_vocab, label
    return model


def test_tinystories_1p2g_vocab_10k_crosses_a_real_chunk_cut(yabpe):
    torch, text, n = _gen("tinystories", 1_200_000_000, 20260101)
    assert n > (1 << 30)
    from yabpe.trainer import device_chunk_cuts
    cuts = device_chunk_cuts(text, n, 1 << 30)
    assert len(cuts) == 1 and (1 << 30) - 4 <= cuts[0] <= (1 << 30)
    host = text[:n].cpu().numpy().tobytes()
    assert oracle.chunk_cuts(host, 1 << 30)[:-1] == cuts
    _check_train(yabpe, text, n, host, 10_000, "tinystories 1.2 GB")
    # the same bytes through the host-buffer path (counted while uploaded, pieces of 256 MiB) give the same model
    m2 = yabpe.BBPETrainer(yabpe.BBPETrainerConfig(vocab_size=10_000, min_frequency=1, max_workers=1, chunk_size_bytes=1 << 30,
                                                   special_tokens=SP)).train_from_buffers([np.frombuffer(host, dtype=np.uint8)])
    m1 = yabpe.BBPETrainer(yabpe.BBPETrainerConfig(vocab_size=10_000, min_frequency=1, max_workers=1, chunk_size_bytes=1 << 30,
                                                   special_tokens=SP)).train_device(text, n)
    assert m1.merges == m2.merges and m1.vocab == m2.vocab


def test_owt_256m_vocab_32k(yabpe):
    torch, text, n = _gen("owt", 256_000_000, 20260102)
    host = text[:n].cpu().numpy().tobytes()
    _check_train(yabpe, text, n, host, 32_000, "owt 256 MB")


def test_adversarial_64m_vocab_50k(yabpe):
    """BASELINE configs[4].  The numpy generator makes ~1.3 MB / s, so 16 MB of it are tiled with a changing separator
    (the tie farms and special-token runs repeat; the counts grow; the order of ties stays a byte-order question)."""
    import torch
    from yabpe import engine
    base = common.synth_adversarial(16_000_000, seed=20260104)
    host = b"".join(base + (b"\n%d<|endoftext|>" % i) for i in range(4))
    text, n = engine.to_device_text(torch, np.frombuffer(host, dtype=np.uint8))
    _check_train(yabpe, text, n, host, 50_000, "adversarial 64 MB")


def test_gpt2_encode_256m_slice(yabpe):
    torch, text, n = _gen("owt", 256_000_000, 20260103)
    host = text[:n].cpu().numpy().tobytes()
    v, m = common.gpt2_vocab_and_merges()
    tok = yabpe.Tokenizer(v, m, SP).inner
    ids, _ = tok.encode_device(text, n)
    want = oracle.Tokenizer(v, m, SP).encode_bytes(host)
    got = ids.cpu().numpy()
    assert got.shape == want.shape
    bad = np.flatnonzero(got != want)
    assert bad.size == 0, f"first differing id at index {bad[:1]}"
    out = tok.decode_device(ids)
    assert out.numel() == n and torch.equal(out, text[:n])
