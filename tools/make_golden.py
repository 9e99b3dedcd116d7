#!/usr/bin/env python3
"""Generate tests/golden/*.json by running the UNMODIFIED reference from /root/reference.

Run in the authoring container only (the reference does not travel to the GPU box):

    python tools/make_golden.py

Every vector stores its own input, so nothing depends on RNG reproducibility.
Calls go through the reference's adapter boundary (tests/adapters.py:37-99) or, for
chunking / min_frequency / multi-file cases, through BBPETrainer directly with the
same arguments the adapter would pass.
"""
from __future__ import annotations

import base64
import json
import random
import sys
import tempfile
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
REF = Path("/root/reference")
sys.path.insert(0, str(REF / "src"))
sys.path.insert(0, str(REF))

import regex  # noqa: E402
from yet_another_bpe.tokenizer import BBPETokenizer  # noqa: E402
from yet_another_bpe.trainer import BBPETrainer, BBPETrainerConfig  # noqa: E402

GOLD = ROOT / "tests" / "golden"
FIX = ROOT / "tests" / "fixtures_gpt2"
GPT2 = r"""'(?:[sdmt]|ll|ve|re)| ?\p{L}+| ?\p{N}+| ?[^\s\p{L}\p{N}]+|\s+(?!\S)|\s+"""


def b64(b: bytes) -> str:
    return base64.b64encode(b).decode("ascii")


def hx(b: bytes) -> str:
    return b.hex()


ALPHABET = list("ab Z'sdmtlvre19!<|>\n\t \r") + [
    "é", "中", "１", " ", " ", "​", "\U0001f643", "", "　", "",
    "́", "\U00016ea0", "'ll", "'ve", " '", "<|endoftext|>", "<|e|>", "don't", " we've", "I'll",
]


def rand_text(rng: random.Random, n: int) -> str:
    return "".join(rng.choice(ALPHABET) for _ in range(n))


def adversarial_text(rng: random.Random, scale: int = 1) -> str:
    """Mini version of BASELINE.json config 5."""
    parts = []
    parts.append("a" * (257 * scale))                     # long single-class runs
    parts.append(" " + "7" * (130 * scale))
    parts.append("!" * (99 * scale) + "\n")
    parts.append("中文字符" * (40 * scale))
    for pre in ["", " ", "!", "\n", "x", "  ", "\t", "'s", "."]:   # F3 cases
        parts.append("end" + pre + "<|endoftext|>")
    parts.append("<|endoftext|>" * 5)
    parts.append("!<|endoftext|><|endoftext|>a<|endoftext|><|endoftext|>")
    parts.append("\U0001f643\U00016ea0 \U0001f643x　y z  w\n\nv")
    parts.append("ábć́ i'm you're they'LL <|endoftext|>'s 'S")
    for _ in range(60 * scale):                           # tie farms + prefix-related tokens
        parts.append(rng.choice(["ab", "abc", "bc", "aaa", "aaaa", "ab ab", "éa", "aé", "\u0080",
                                 " xy", " xz", " yx", " zx", "xy", "zx"]))
        parts.append(rng.choice([" ", "", "\n", "  "]))
    parts.append(rand_text(rng, 400 * scale))
    return "".join(parts)


def train_ref(data_files: list[bytes], vocab_size, specials, min_frequency=1, chunk_size=1 << 30):
    with tempfile.TemporaryDirectory() as td:
        paths = []
        for i, d in enumerate(data_files):
            p = Path(td) / f"f{i}.txt"
            p.write_bytes(d)
            paths.append(p)
        tr = BBPETrainer(BBPETrainerConfig(vocab_size=vocab_size, min_frequency=min_frequency, max_workers=1,
                                           chunk_size_bytes=chunk_size, seed=42, special_tokens=specials))
        model = tr.train(paths)
    vocab = {v: k for k, v in model.vocab.items()}
    return vocab, model.merges


def make_train() -> None:
    rng = random.Random(20260104)
    cases = []
    corpus = (FIX / "corpus.en").read_bytes()
    named = [
        ("corpus.en@500", None, 500, ["<|endoftext|>"], 1, 1 << 30),
        ("corpus.en@1500", None, 1500, ["<|endoftext|>"], 1, 1 << 30),
        ("corpus.en@700/chunk4096", None, 700, ["<|endoftext|>"], 1, 4096),
        ("corpus.en@600/minfreq20/two-specials", None, 600, ["<|endoftext|>", "the"], 20, 1 << 30),
    ]
    for name, _, vs, sp, mf, cs in named:
        vocab, merges = train_ref([corpus], vs, sp, mf, cs)
        cases.append(dict(name=name, input_file="tests/fixtures_gpt2/corpus.en", vocab_size=vs, specials=sp,
                          min_frequency=mf, chunk_size=cs, merges=[[hx(a), hx(b)] for a, b in merges],
                          vocab=[hx(vocab[i]) for i in range(len(vocab))]))
    ts = (FIX / "tinystories_sample.txt").read_bytes()
    vocab, merges = train_ref([ts], 400, ["<|endoftext|>"])
    cases.append(dict(name="tinystories_sample@400", input_file="tests/fixtures_gpt2/tinystories_sample.txt",
                      vocab_size=400, specials=["<|endoftext|>"], min_frequency=1, chunk_size=1 << 30,
                      merges=[[hx(a), hx(b)] for a, b in merges], vocab=[hx(vocab[i]) for i in range(len(vocab))]))
    for k in range(6):
        text = adversarial_text(rng, 1 + (k % 2)).encode("utf-8")
        sp = [["<|endoftext|>"], ["<|endoftext|>", "<|e|>"], [], ["ab", "<|endoftext|>"], ["<|endoftext|>"], ["a"]][k]
        vs = [600, 1000, 300, 450, 5000, 280][k]
        cs = [1 << 30, 1 << 30, 1 << 30, 97, 1 << 30, 1 << 30][k]
        vocab, merges = train_ref([text], vs, sp, 1, cs)
        cases.append(dict(name=f"adversarial{k}", input_b64=b64(text), vocab_size=vs, specials=sp, min_frequency=1,
                          chunk_size=cs, merges=[[hx(a), hx(b)] for a, b in merges],
                          vocab=[hx(vocab[i]) for i in range(len(vocab))]))
    # multi-file + empty file
    f0, f1 = b"hello hello world<|endoftext|>", "héllo wörld hello\n".encode()
    vocab, merges = train_ref([f0, b"", f1], 280, ["<|endoftext|>"])
    cases.append(dict(name="multi-file", inputs_b64=[b64(f0), b64(b""), b64(f1)], vocab_size=280,
                      specials=["<|endoftext|>"], min_frequency=1, chunk_size=1 << 30,
                      merges=[[hx(a), hx(b)] for a, b in merges], vocab=[hx(vocab[i]) for i in range(len(vocab))]))
    (GOLD / "train_cases.json").write_text(json.dumps(cases, indent=0))
    print("train cases:", len(cases))


def make_pretok() -> None:
    rng = random.Random(20260105)
    cases = []
    special_sets = [[], ["<|endoftext|>"], ["<|e|>", "<|endoftext|>"], [" <", "<|e|>"], ["\nb", "ab", "a"]]
    for _ in range(400):
        s = rand_text(rng, rng.randint(0, 40))
        sp = rng.choice(special_sets)
        pat = GPT2 if not sp else "|".join(regex.escape(t) for t in sp) + "|" + GPT2
        toks = [t for t in regex.findall(pat, s) if t]
        cases.append(dict(mode="train", text=s, specials=sp, tokens=toks))
    enc_sets = [["<|endoftext|>"], ["<|e|>", "<|endoftext|>", "<|endoftext|><|endoftext|>"], ["a", "ab", " "]]
    g = regex.compile(GPT2)
    for _ in range(300):
        s = rand_text(rng, rng.randint(0, 40))
        sp = rng.choice(enc_sets)
        srt = sorted(sp, key=len, reverse=True)
        spat = regex.compile("(" + "|".join(regex.escape(t) for t in srt) + ")")
        toks = []
        for part in spat.split(s):
            if not part:
                continue
            toks += [part] if part in sp else g.findall(part)
        cases.append(dict(mode="encode", text=s, specials=sp, tokens=toks))
    (GOLD / "pretokenize_cases.json").write_text(json.dumps(cases, ensure_ascii=True, indent=0))
    print("pretok cases:", len(cases))


def gpt2_vocab_and_merges():
    """Rebuild the git-ignored gpt2_vocab.json from gpt2_merges.txt."""
    sys.path.insert(0, str(REF))
    from tests.common import gpt2_bytes_to_unicode
    b2u = gpt2_bytes_to_unicode()
    u2b = {v: k for k, v in b2u.items()}
    vocab = {i: bytes([b]) for i, b in enumerate(b2u.keys())}
    merges = []
    for line in (FIX / "gpt2_merges.txt").read_text(encoding="utf-8").split("\n"):
        parts = line.rstrip().split(" ")
        if len(parts) != 2:
            continue
        a = bytes(u2b[c] for c in parts[0])
        b = bytes(u2b[c] for c in parts[1])
        merges.append((a, b))
        vocab[len(vocab)] = a + b
    vocab[len(vocab)] = b"<|endoftext|>"
    return vocab, merges


def make_encode() -> None:
    rng = random.Random(20260103)
    vocab, merges = gpt2_vocab_and_merges()
    assert len(vocab) == 50257
    sp = ["<|endoftext|>"]
    tok = BBPETokenizer(vocab={v: k for k, v in vocab.items()}, merges=merges, special_tokens=sp)
    cases = []
    texts = ["", "s", "\U0001f643", "Hello, how are you?", "Héllò hôw <|endoftext|><|endoftext|> are ü? \U0001f643<|endoftext|>",
             "Hello, how <|endoftext|><|endoftext|> are you?<|endoftext|>", "\n\n", "a  b   c\n \n  d ", "don't we've I'LL",
             adversarial_text(rng)]
    for name in ["address.txt", "german.txt", "tinystories_sample.txt", "special_token_trailing_newlines.txt",
                 "special_token_double_newlines_non_whitespace.txt"]:
        with open(FIX / name) as f:   # text mode, like tests/test_tokenizer_gpt2.py:270
            texts.append(f.read())
    with open(FIX / "corpus.en") as f:
        corpus = f.read()
    texts.append(corpus[:20000])
    for _ in range(40):
        texts.append(rand_text(rng, rng.randint(1, 60)))
    for t in texts:
        ids = tok.encode(t)
        cases.append(dict(model="gpt2", specials=sp, text=t, ids=ids, decoded=tok.decode(ids)))
    # overlapping specials, longest first (tests/test_tokenizer_gpt2.py:248-262)
    sp2 = ["<|endoftext|>", "<|endoftext|><|endoftext|>"]
    tok2 = BBPETokenizer(vocab={**{v: k for k, v in vocab.items()}, b"<|endoftext|><|endoftext|>": 50257},
                         merges=merges, special_tokens=sp2)
    for t in ["Hello, how <|endoftext|><|endoftext|> are you?<|endoftext|>", "<|endoftext|><|endoftext|><|endoftext|>x"]:
        ids = tok2.encode(t)
        cases.append(dict(model="gpt2+double", specials=sp2, text=t, ids=ids, decoded=tok2.decode(ids)))
    # special missing from vocab is dropped (tokenizer.py:177-181); [UNK] fallback (tokenizer.py:299)
    tok3 = BBPETokenizer(vocab={v: k for k, v in vocab.items()}, merges=merges, special_tokens=["<|missing|>"])
    t = "a<|missing|>b <|endoftext|>"
    ids = tok3.encode(t)
    cases.append(dict(model="gpt2", specials=["<|missing|>"], text=t, ids=ids, decoded=tok3.decode(ids)))
    # a small reference-trained model with an inconsistent, duplicated merge list
    weird_merges = [(b"ab", b"a"), (b"a", b"b"), (b"b", b"a"), (b"a", b"b"), (b"aba", b"b"), (b"x", b"y")]
    weird_vocab = {i: bytes([i]) for i in range(256)}
    weird_vocab.update({256: b"ab", 257: b"aba", 258: b"ba", 259: b"[UNK]"})
    tokw = BBPETokenizer(vocab={v: k for k, v in weird_vocab.items()}, merges=weird_merges, special_tokens=[])
    for _ in range(60):
        t = "".join(rng.choice("abxy ") for _ in range(rng.randint(0, 14)))
        ids = tokw.encode(t)
        cases.append(dict(model="weird", specials=[], text=t, ids=ids, decoded=tokw.decode(ids)))
    models = {"weird": dict(vocab=[hx(weird_vocab[i]) for i in range(len(weird_vocab))],
                            merges=[[hx(a), hx(b)] for a, b in weird_merges])}
    (GOLD / "encode_cases.json").write_text(json.dumps(dict(models=models, cases=cases), ensure_ascii=True, indent=0))
    print("encode cases:", len(cases))


def make_class_api() -> None:
    """the private trainer entry points the reference's own tests call
    (tests/test_trainer.py:27-42,214,245) and the on-disk format (trainer.py:94-117, tokenizer.py:106-150)."""
    rng = random.Random(20260105)
    out: dict = dict(preprocess=[], merge_loop=[], persist=[])

    def cfg(**kw):
        kw.setdefault("max_workers", 1)
        return BBPETrainerConfig(**kw)

    # -- _preprocess_corpus: per-occurrence byte lists, files and text in order
    corpus = (FIX / "corpus.en").read_bytes()
    pp = [
        ([rand_text(rng, 1500).encode()], ["<|endoftext|>"], 1 << 30),
        ([rand_text(rng, 1500).encode()], ["<|e|>", "<|endoftext|>"], 97),
        ([rand_text(rng, 1200).encode(), b"", rand_text(rng, 700).encode()], [" <", "<|e|>"], 256),
        ([adversarial_text(rng, 1).encode()], ["<|endoftext|>"], 1024),
        ([corpus[:20000]], [], 4096),
        (["hello hello world<|endoftext|>".encode(), "h\u00e9llo w\u00f6rld \u4e2d\u6587 don't\r\n\r\n  x ".encode()], ["<|endoftext|>"], 8),
    ]
    for files, sp, cs in pp:
        with tempfile.TemporaryDirectory() as td:
            paths = []
            for i, d in enumerate(files):
                q = Path(td) / f"f{i}.txt"
                q.write_bytes(d)
                paths.append(q)
            seqs = BBPETrainer(cfg(chunk_size_bytes=cs, special_tokens=sp))._preprocess_corpus(paths)
        out["preprocess"].append(dict(files_b64=[b64(d) for d in files], specials=sp, chunk_size=cs,
                                      sequences=[hx(bytes(x)) for x in seqs]))

    # -- _merge_loop(sequences)
    def run_ml(seqs, **kw):
        vocab, merges = BBPETrainer(cfg(**kw))._merge_loop([list(x) for x in seqs])
        inv = {v: k for k, v in vocab.items()}
        out["merge_loop"].append(dict(sequences=[hx(bytes(x)) for x in seqs], config=kw,
                                      vocab=[hx(inv[i]) for i in range(len(inv))], merges=[[hx(a), hx(b)] for a, b in merges]))

    run_ml([], vocab_size=300, min_frequency=1)
    run_ml([b"Hello", b"Hello"], vocab_size=265, min_frequency=1)
    run_ml([b"AB"] * 100 + [b"CD"] * 50 + [b"EF"] * 10, vocab_size=270, min_frequency=1)
    run_ml([b"AB", b"CD", b"EF", b"GH", b"IJ"] * 10, vocab_size=262, min_frequency=1)
    run_ml([b"AB"] * 10 + [b"CD"] * 5 + [b"EF"] * 4 + [b"GH"], vocab_size=300, min_frequency=5)
    run_ml([], vocab_size=300, min_frequency=1, special_tokens=["[PAD]", "[UNK]", "[BOS]", "[EOS]", "[MASK]"])
    words = [bytes(rng.choice(b"abc\xc3\xa9 ") for _ in range(rng.randint(1, 9))) for _ in range(60)]
    run_ml([rng.choice(words) for _ in range(3000)], vocab_size=330, min_frequency=1, special_tokens=["<|endoftext|>"])
    run_ml([rng.choice(words) for _ in range(3000)], vocab_size=400, min_frequency=25, special_tokens=[])
    run_ml([b"a" * 300, b"a" * 7, b"aaab" * 90, b"b"] * 3, vocab_size=290, min_frequency=2, special_tokens=["a"])

    # -- save() / from_file(): literal files and what the (lossy) loader gets back
    texts = ["Hello, how are you?", " the cat\r\nsat  on\n\nthe mat<|endoftext|> caf\u00e9 \u4e2d\u6587 don't", ""]
    for data, vs, sp in [(corpus, 420, ["<|endoftext|>"]),
                         (("a b\r\nc\rd \u00e9\u00e9 \u0085x\u2028y  z\t\tq " * 40).encode(), 300, ["<|endoftext|>", "[PAD]"])]:
        with tempfile.TemporaryDirectory() as td:
            q = Path(td) / "in.txt"
            q.write_bytes(data)
            tr = BBPETrainer(cfg(vocab_size=vs, min_frequency=1, chunk_size_bytes=1 << 30, special_tokens=sp))
            model = tr.train([q])
            tr.save(Path(td) / "model")
            files = {n: b64((Path(td) / "model" / n).read_bytes()) for n in ("vocab.json", "merges.txt", "special_tokens.json")}
            tok = BBPETokenizer.from_file(Path(td) / "model")
            loaded_vocab = sorted(((hx(k), v) for k, v in tok._vocab.items()), key=lambda kv: kv[1])
            loaded_merges = [[hx(a), hx(b)] for a, b in tok._merges]
            inv = {v: k for k, v in model.vocab.items()}
            out["persist"].append(dict(
                **(dict(input_file="tests/fixtures_gpt2/corpus.en") if data is corpus else dict(input_b64=b64(data))),
                vocab_size=vs, specials=sp, files_b64=files,
                trained_vocab=[hx(inv[i]) for i in range(len(inv))], trained_merges=[[hx(a), hx(b)] for a, b in model.merges],
                loaded_vocab=loaded_vocab, loaded_merges=loaded_merges, loaded_specials=list(tok._special_tokens),
                encodes=[dict(text=t, ids=tok.encode(t)) for t in texts]))
    (GOLD / "class_api_cases.json").write_text(json.dumps(out, ensure_ascii=True, indent=0))
    print("class api cases:", {k: len(v) for k, v in out.items()})


def make_prefix_specials() -> None:
    """Prefix-related special tokens next to hard boundaries (chunk cuts, file ends, document ends): the
    reference never matches a special across a boundary (trainer.py:146-170 scans each chunk as its own text;
    tests/adapters.py:30-34 encodes each item on its own), so '<|eot|>' + CUT + 'x...' must stay '<|eot|>'
    even when '<|eot|>x' is a special of higher priority."""
    out: dict = dict(pretok=[], train=[], encode_docs=[])
    g = regex.compile(GPT2)
    # -- pre-tokenisation of independent texts (documents / chunks), both modes
    doc_sets = [
        ["ab\n<|eot|>", "xyz <|eot|>x <|eot|>", "x<|eot|>", "x"],
        ["<|eot|>", "x<|eot|>x", "<|eot|>"],
        ["hello <|eot|>", "x <|eot|>xx<|eot|>", "xx"],
        ["a\n<|e", "ot|>x b\n<|eot|>", "xb"],
    ]
    for sp in (["<|eot|>x", "<|eot|>"], ["<|eot|>", "<|eot|>x"], ["<|eot|>", "<|eot|>x", "<|eot|>xx"]):
        for docs in doc_sets:
            pat = "|".join(regex.escape(t) for t in sp) + "|" + GPT2
            out["pretok"].append(dict(mode="train", specials=sp, docs=docs,
                                      tokens=[[t for t in regex.findall(pat, d) if t] for d in docs]))
            srt = sorted(sp, key=len, reverse=True)
            spat = regex.compile("(" + "|".join(regex.escape(t) for t in srt) + ")")
            per_doc = []
            for d in docs:
                toks = []
                for part in spat.split(d):
                    if part:
                        toks += [part] if part in sp else g.findall(part)
                per_doc.append(toks)
            out["pretok"].append(dict(mode="encode", specials=sp, docs=docs, tokens=per_doc))
    # -- training: the special ends exactly at a chunk cut (chunk_size 16) / at a file end
    blocks = ["aaaaaaaa\n<|eot|>", "xyz hello x worl", "d\nabc de\n<|eot|>", "x<|eot|>x  tail\n"]
    assert all(len(b) == 16 for b in blocks)
    body = "".join(blocks) * 6
    for sp in (["<|eot|>x", "<|eot|>"], ["<|eot|>", "<|eot|>x"]):
        for files, cs in (([body.encode()], 16), ([body.encode()], 1 << 30),
                          ([b"hello hello\n<|eot|>", b"xyz hello x<|eot|>", b"x<|eot|>x hello"], 1 << 30)):
            vocab, merges = train_ref(files, 300, sp, 1, cs)
            out["train"].append(dict(inputs_b64=[b64(f) for f in files], vocab_size=300, specials=sp, min_frequency=1,
                                     chunk_size=cs, merges=[[hx(a), hx(b)] for a, b in merges],
                                     vocab=[hx(vocab[i]) for i in range(len(vocab))]))
    # -- encode of independent documents (encode_iterable / encode_batch): ids per document
    sp = ["<|eot|>", "<|eot|>x"]
    vocab, merges = train_ref([(body * 3).encode()], 290, sp, 1, 1 << 30)
    tok = BBPETokenizer(vocab={v: k for k, v in vocab.items()}, merges=merges, special_tokens=sp)
    for docs in doc_sets + [["<|eot|>x", "x<|eot|>", "x", "", "<|eot|>"]]:
        out["encode_docs"].append(dict(specials=sp, docs=docs, ids=[tok.encode(d) for d in docs]))
    out["encode_model"] = dict(vocab=[hx(vocab[i]) for i in range(len(vocab))], merges=[[hx(a), hx(b)] for a, b in merges])
    (GOLD / "prefix_specials_cases.json").write_text(json.dumps(out, ensure_ascii=True, indent=0))
    print("prefix-special cases:", {k: len(v) for k, v in out.items() if isinstance(v, list)})


if __name__ == "__main__":
    GOLD.mkdir(parents=True, exist_ok=True)
    only = sys.argv[1:]
    for fn in (make_pretok, make_train, make_encode, make_class_api, make_prefix_specials):
        if not only or fn.__name__ in only:
            fn()
