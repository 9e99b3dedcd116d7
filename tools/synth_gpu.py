"""Synthetic corpora of the BASELINE.json shapes, generated ON THE GPU with torch ops.

Bench / test infrastructure (not part of the product path): an 11 GB corpus cannot be produced
by Python loops in minutes, so words are sampled (Zipf / Zipf-Mandelbrot inverse CDF), decorated
and scattered into a byte tensor with vectorised torch ops, piece by piece.

  kind="tinystories": ~40 K lowercase word types, sentence capitals, . , ! ? on ~10 % of words,
                      'new line' paragraphs, documents of ~200 words ended by "\n<|endoftext|>\n"
  kind="owt":         large lexicon (default 5 M types, Zipf-Mandelbrot q=2.7) with ~3 % digit
                      strings, ~2 % non-ASCII words (Latin-1 accents, CJK, Cyrillic, emoji), URL-like
                      types, "\n\n" paragraphs, documents of ~900 words ended by "<|endoftext|>"
"""
from __future__ import annotations

import numpy as np

LEX_W = 24          # bytes per lexicon row
ITEM_W = 48         # max bytes per emitted item (word + punctuation + separator)

_SEPS = [b" ", b"\n", b"\n\n", b"\n<|endoftext|>\n", b"<|endoftext|>", b". ", b", ", b"! ", b"? ", b".\n", b"; ", b"\" ",
         b") ", b"... ", b".\n\n", b".\n<|endoftext|>\n", b".<|endoftext|>"]


def _lexicon(kind: str, n_types: int, seed: int) -> tuple[np.ndarray, np.ndarray]:
    rng = np.random.default_rng(seed)
    letters = np.frombuffer(b"etaoinshrdlcumwfgypbvkjxqz", dtype=np.uint8)
    w = 1.0 / np.arange(1, 27) ** 0.9
    w /= w.sum()
    lens = np.minimum(12, 1 + rng.poisson(4.2 if kind == "tinystories" else 5.5, size=n_types)).astype(np.int32)
    lens = np.maximum(lens, 1)
    body = rng.choice(letters, size=(n_types, LEX_W), p=w).astype(np.uint8)
    lex = np.where(np.arange(LEX_W)[None, :] < lens[:, None], body, 0).astype(np.uint8)
    if kind == "owt":
        n_special = n_types // 16
        idx = rng.choice(n_types, size=n_special, replace=False)
        extras = ["é", "naïve", "über", "café", "señor", "中文", "日本語", "汉字测试", "привет", "мир", "данные",
                  "\U0001f643", "\U0001f600\U0001f600", "—", "…", "€100", "100%", "x=1&y=2", "e.g.", "U.S.", "don't", "it's"]
        for k, i in enumerate(idx):
            r = k % 10
            if r < 5:      # digit strings
                s = str(int(rng.integers(0, 10 ** int(rng.integers(1, 7))))).encode()
            elif r < 8:    # non-ASCII / punctuation-bearing words
                s = extras[int(rng.integers(0, len(extras)))].encode("utf-8")
            else:          # URL-like
                s = b"www." + bytes(rng.choice(letters, size=int(rng.integers(3, 9)), p=w)) + b".com/" + bytes(rng.choice(letters, size=3, p=w))
            s = s[:LEX_W]
            while s and (s[-1] & 0xC0) == 0x80:
                s = s[:-1]
            if s and s[-1] >= 0xC0:
                s = s[:-1]
            if not s:
                s = b"x"
            lex[i] = 0
            lex[i, :len(s)] = np.frombuffer(s, dtype=np.uint8)
            lens[i] = len(s)
    return lex, lens


def synth_corpus_device(torch, n_bytes: int, kind: str = "tinystories", seed: int = 20260101,
                        n_types: int | None = None, piece_bytes: int = 256 << 20, lex_seed: int | None = None):
    """Return (uint8 device tensor with capacity round_up(n,16)+64, n).
    `seed` drives the text; `lex_seed` (default: seed) the lexicon -- shards of ONE corpus share the lexicon
    and differ in their text, so a multi-GPU run passes lex_seed=<corpus seed>, seed=<corpus seed> + rank."""
    dev = torch.device("cuda")
    if n_types is None:
        n_types = 40_000 if kind == "tinystories" else 5_000_000
    lex_np, len_np = _lexicon(kind, n_types, seed if lex_seed is None else lex_seed)
    lex = torch.from_numpy(lex_np).to(dev)
    lex_len = torch.from_numpy(len_np).to(dev)
    ranks = np.arange(1, n_types + 1, dtype=np.float64)
    pm = 1.0 / ranks ** 1.05 if kind == "tinystories" else 1.0 / (ranks + 2.7)
    cdf = torch.from_numpy(np.cumsum(pm / pm.sum())).to(dev)
    sep_tab_np = np.zeros((len(_SEPS), 24), dtype=np.uint8)
    for i, s in enumerate(_SEPS):
        sep_tab_np[i, :len(s)] = np.frombuffer(s, dtype=np.uint8)
    sep_tab = torch.from_numpy(sep_tab_np).to(dev)
    sep_len = torch.tensor([len(s) for s in _SEPS], dtype=torch.int32, device=dev)
    gen = torch.Generator(device=dev)
    cap = ((n_bytes + 15) // 16) * 16 + 64
    out = torch.zeros(cap, dtype=torch.uint8, device=dev)
    filled = 0
    piece = 0
    avg = 6.0 if kind == "tinystories" else 7.5
    while filled < n_bytes:
        gen.manual_seed(seed * 1000 + piece)
        piece += 1
        want = min(piece_bytes, n_bytes - filled)
        W = int(want / avg * 1.15) + 1024
        u = torch.rand(W, generator=gen, device=dev, dtype=torch.float64)
        wid = torch.searchsorted(cdf, u).clamp_(max=n_types - 1)
        r = torch.rand(W, generator=gen, device=dev)
        wl = lex_len[wid]
        # separator kind per item
        if kind == "tinystories":
            sep = torch.zeros(W, dtype=torch.int64, device=dev)                 # " "
            sep[r > 0.90] = 5                                                    # ". "
            sep[r > 0.93] = 6                                                    # ", "
            sep[r > 0.96] = 7                                                    # "! "
            sep[r > 0.975] = 8                                                   # "? "
            sep[r > 0.985] = 9                                                   # ".\n"
            sep[r > 0.995] = 15                                                  # ".\n<|endoftext|>\n"
        else:
            sep = torch.zeros(W, dtype=torch.int64, device=dev)
            sep[r > 0.88] = 5
            sep[r > 0.92] = 6
            sep[r > 0.95] = 10                                                   # "; "
            sep[r > 0.96] = 11                                                   # "\" "
            sep[r > 0.97] = 12                                                   # ") "
            sep[r > 0.975] = 13                                                  # "... "
            sep[r > 0.98] = 14                                                   # ".\n\n"
            sep[r > 0.9989] = 16                                                 # ".<|endoftext|>"
        sl = sep_len[sep]
        item_len = (wl + sl).to(torch.int64)
        offs = torch.cumsum(item_len, 0) - item_len
        total = int((offs[-1] + item_len[-1]).item())
        # sentence-initial capitals
        ends_sentence = (sep == 5) | (sep == 7) | (sep == 8) | (sep == 9) | (sep >= 13)
        cap_flag = torch.zeros(W, dtype=torch.bool, device=dev)
        cap_flag[1:] = ends_sentence[:-1]
        cap_flag[0] = True
        buf = torch.zeros(max(total, 1) + ITEM_W, dtype=torch.uint8, device=dev)
        for j in range(LEX_W + 20):
            in_word = j < wl
            k = (j - wl).clamp(min=0, max=23).to(torch.int64)
            val = torch.where(in_word, lex[wid, min(j, LEX_W - 1)], sep_tab[sep, k])
            if j == 0:
                lower = (val >= 97) & (val <= 122) & cap_flag
                val = torch.where(lower, val - 32, val)
            m = j < item_len
            buf[(offs + j)[m]] = val[m]
        take = min(total, n_bytes - filled)
        out[filled:filled + take] = buf[:take]
        filled += take
        del buf, u, wid, r, wl, sep, sl, item_len, offs
    # never end inside a UTF-8 sequence
    n = n_bytes
    tail = out[max(0, n - 4):n].cpu().numpy()
    k = len(tail)
    p = k - 1
    while p >= 0 and (tail[p] & 0xC0) == 0x80:
        p -= 1
    if p >= 0:
        b = int(tail[p])
        need = 1 if b < 0x80 else 2 if b < 0xE0 else 3 if b < 0xF0 else 4
        if p + need > k:
            n -= k - p
            out[n:].zero_()
    return out, n
