"""SURVEY 8(f) rows 1-2 on the CPU: the oracle against the reference's outputs for the private trainer entry points,
and the host-only persistence code (save / from_file write and parse files; no device work involved)."""
from __future__ import annotations

from collections import Counter

import common
from oracle import oracle


def test_oracle_preprocess_order_matches_reference():
    for c in common.load_class_api_cases()["preprocess"]:
        got = []
        for blob in c["files"]:
            got += oracle.pretokenize(blob, c["specials"], "train", c["chunk_size"])
        assert got == c["want"], (c["specials"], c["chunk_size"])


def test_oracle_merge_loop_on_sequences_matches_reference():
    for c in common.load_class_api_cases()["merge_loop"]:
        cfg = c["config"]
        tr = oracle.Trainer(cfg.get("special_tokens", ["[PAD]", "[UNK]", "[BOS]", "[EOS]"]))
        for w, f in Counter(c["seqs"]).items():
            tr.feed_word(w, f)
        vocab, merges = tr.run(cfg["vocab_size"], cfg["min_frequency"])
        assert [vocab[i] for i in range(len(vocab))] == c["want_vocab"], cfg
        assert merges == c["want_merges"], cfg


def test_save_writes_the_reference_files(tmp_path):
    """trainer.py:94-117: latin-1 keys in vocab.json, "a b" lines in merges.txt, special_tokens.json -- byte for byte."""
    import yabpe
    for i, c in enumerate(common.load_class_api_cases()["persist"]):
        tr = yabpe.BBPETrainer(yabpe.BBPETrainerConfig(vocab_size=c["vocab_size"], special_tokens=c["specials"]))
        tr._finish({bytes.fromhex(h): j for j, h in enumerate(c["trained_vocab"])},
                   [(bytes.fromhex(a), bytes.fromhex(b)) for a, b in c["trained_merges"]])
        out = tmp_path / f"m{i}" / "nested"
        tr.save(out)
        for name, want in c["files"].items():
            assert (out / name).read_bytes() == want, name


def test_from_file_parses_like_the_reference(tmp_path):
    """tokenizer.py:106-150, lossy on purpose: a merge whose left token starts with a space or holds a line break
    does not survive the "a b" text format; the loader must lose exactly what the reference loses."""
    import yabpe
    for i, c in enumerate(common.load_class_api_cases()["persist"]):
        d = tmp_path / f"m{i}"
        d.mkdir()
        for name, blob in c["files"].items():
            (d / name).write_bytes(blob)
        tok = yabpe.BBPETokenizer.from_file(d)
        assert sorted(((k.hex(), v) for k, v in tok._vocab.items()), key=lambda kv: kv[1]) == [tuple(x) for x in c["loaded_vocab"]]
        assert [[a.hex(), b.hex()] for a, b in tok._merges] == c["loaded_merges"]
        assert list(tok._special_tokens) == c["loaded_specials"]
        assert c["loaded_merges"] != c["trained_merges"]          # the format really is lossy on these models
    (d / "special_tokens.json").unlink()
    assert yabpe.BBPETokenizer.from_file(d).special_tokens == []


def test_save_before_training_raises(tmp_path):
    import pytest
    import yabpe
    with pytest.raises(ValueError, match="not been trained"):
        yabpe.BBPETrainer().save(tmp_path / "x")


def test_cli_arguments_and_missing_input(tmp_path, capsys):
    """scripts/train_bpe.py: the reference script's defaults, and its FileNotFoundError before any device work."""
    import pytest
    from yabpe.scripts import train_bpe
    with pytest.raises(SystemExit) as e:
        train_bpe.main(["--help"])
    assert e.value.code == 0 and "--vocab-size" in capsys.readouterr().out
    with pytest.raises(FileNotFoundError, match="Data file not found"):
        train_bpe.main(["--input", str(tmp_path / "nope.txt")])


def test_encode_pinned_piece_cuts_are_exact():
    """BBPETokenizer._piece_ends (host logic of encode_pinned): pieces end right after a special token, so the
    concatenation of the oracle's per-piece ids equals its ids for the whole text; specials that could overlap
    (several, or one with a border) keep the text in one piece."""
    import numpy as np
    import yabpe
    vocab, merges = oracle.train_bpe_bytes(common.synth_tinystories(60_000, seed=3), 400, ["<|endoftext|>"])
    text = common.synth_tinystories(300_000, seed=5) + b"<|endoftext|><|endoftext|>tail without a separator " * 3
    host = np.frombuffer(text, dtype=np.uint8)
    inv = {v: k for k, v in vocab.items()}
    tok = yabpe.BBPETokenizer(vocab=inv, merges=merges, special_tokens=["<|endoftext|>"])
    otok = oracle.Tokenizer(vocab, merges, ["<|endoftext|>"])
    want = otok.encode(text.decode("utf-8"))
    for piece in (1_000, 4_096, 50_000, 250_000, 1 << 20):
        ends = tok._piece_ends(host, len(text), piece)
        assert ends[-1] == len(text) and ends == sorted(set(ends))
        if piece < len(text) // 2:
            assert len(ends) > 1
        got, lo = [], 0
        for e in ends:
            if e != len(text):
                assert text[:e].endswith(b"<|endoftext|>")
            got += otok.encode(text[lo:e].decode("utf-8"))
            lo = e
        assert got == want, piece
    # no special in the text at all / two specials / a special with a border: one piece
    plain = np.frombuffer(b"abc def " * 4000, dtype=np.uint8)
    assert tok._piece_ends(plain, plain.size, 1000) == [plain.size]
    two = yabpe.BBPETokenizer(vocab=inv, merges=merges, special_tokens=["<|endoftext|>", "<|pad|>"])
    assert two._piece_ends(host, len(text), 1000) == [len(text)]
    bordered = yabpe.BBPETokenizer(vocab=inv, merges=merges, special_tokens=["|x|"])
    assert bordered._piece_ends(np.frombuffer(b"a|x|x|b " * 999, dtype=np.uint8), 7992, 100) == [7992]
