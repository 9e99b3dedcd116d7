"""`train-tiny-stories` (reference: src/yet_another_bpe/scripts/train_bpe.py, pyproject.toml:33-34): train a BPE model on a
text file and save it.  Same defaults as the reference's script (vocab 5000, min_frequency 2, 20 MiB chunks, the
`<|endoftext|>` special, input tests/data/TinyStoriesV2-GPT4-valid.txt, output models/tinystories_bpe); here they are
command-line options instead of constants.

    python -m yabpe.scripts.train_bpe --input corpus.txt --output models/my_bpe --vocab-size 10000
"""
from __future__ import annotations

import argparse
import time
from pathlib import Path


def main(argv: list[str] | None = None) -> int:
    ap = argparse.ArgumentParser(prog="train-tiny-stories", description=__doc__.split("\n\n")[0])
    ap.add_argument("--input", type=Path, nargs="+", default=[Path("tests/data/TinyStoriesV2-GPT4-valid.txt")])
    ap.add_argument("--output", type=Path, default=Path("models/tinystories_bpe"))
    ap.add_argument("--vocab-size", type=int, default=5000)
    ap.add_argument("--min-frequency", type=int, default=2)
    ap.add_argument("--chunk-size-bytes", type=int, default=20 * 1024 * 1024)
    ap.add_argument("--special-token", action="append", default=None, help="repeatable; default <|endoftext|>")
    args = ap.parse_args(argv)
    for f in args.input:
        if not f.exists():
            raise FileNotFoundError(f"Data file not found: {f}")

    from yabpe.trainer import BBPETrainer, BBPETrainerConfig
    specials = args.special_token if args.special_token is not None else ["<|endoftext|>"]
    trainer = BBPETrainer(BBPETrainerConfig(vocab_size=args.vocab_size, min_frequency=args.min_frequency, max_workers=8,
                                            chunk_size_bytes=args.chunk_size_bytes, seed=42, special_tokens=specials))
    print(f"Training BPE model on: {', '.join(str(f) for f in args.input)}")
    print(f"Output directory: {args.output}")
    t0 = time.perf_counter()
    model = trainer.train(files=list(args.input))
    dt = time.perf_counter() - t0
    trainer.save(output_dir=args.output)
    st = trainer.last_stats
    print(f"Training complete in {dt:.2f} s ({st.n_bytes / 1e6:.1f} MB, {st.n_pretokens} pre-tokens, {st.n_words} unique)")
    print(f"  Vocabulary size:  {len(model.vocab)}")
    print(f"  Number of merges: {len(model.merges)}")
    print(f"  Special tokens:   {model.special_tokens}")
    return 0


if __name__ == "__main__":
    raise SystemExit(main())
