"""Adapter functions with the reference's names and signatures (/root/reference/tests/adapters.py),
bound to the B200 implementation.  A user of the reference swaps this file in and keeps their tests."""
from __future__ import annotations

import os
import sys
from collections.abc import Iterable, Iterator
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
if str(ROOT / "yet-another-bpe_b200") not in sys.path:
    sys.path.insert(0, str(ROOT / "yet-another-bpe_b200"))

from yabpe.tokenizer import BBPETokenizer  # noqa: E402
from yabpe.trainer import BBPETrainer, BBPETrainerConfig  # noqa: E402


class TokenizerAdapter:
    def __init__(self, tokenizer: BBPETokenizer) -> None:
        self._tokenizer: BBPETokenizer = tokenizer

    def encode(self, text: str) -> list[int]:
        return self._tokenizer.encode(text)

    def decode(self, ids: list[int]) -> str:
        return self._tokenizer.decode(ids)

    def encode_iterable(self, iterable: Iterable[str]) -> Iterator[int]:
        # adapters.py:30-34: independent encode per item, flattened lazily (batched on the GPU)
        return self._tokenizer.encode_iterable(iterable)


def get_tokenizer(vocab: dict[int, bytes], merges: list[tuple[bytes, bytes]],
                  special_tokens: list[str] | None = None) -> TokenizerAdapter:
    vocab_internal: dict[bytes, int] = {v: k for k, v in vocab.items()}
    return TokenizerAdapter(BBPETokenizer(vocab=vocab_internal, merges=merges, special_tokens=special_tokens or []))


def run_train_bpe(input_path: str | os.PathLike, vocab_size: int, special_tokens: list[str]
                  ) -> tuple[dict[int, bytes], list[tuple[bytes, bytes]]]:
    config = BBPETrainerConfig(vocab_size=vocab_size, min_frequency=1, max_workers=1,
                               chunk_size_bytes=1024 * 1024 * 1024, seed=42, special_tokens=special_tokens)
    trainer = BBPETrainer(config)
    input_file = Path(input_path) if not isinstance(input_path, Path) else input_path
    model = trainer.train([input_file])
    return {v: k for k, v in model.vocab.items()}, model.merges
