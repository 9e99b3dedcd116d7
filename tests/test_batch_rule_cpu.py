"""The batching rules of the merge loop (csrc/merge.cuh, "batched leader merges"; DESIGN.md "Batched merges") checked on the
CPU, independently of the CUDA code: a plain Python BPE that takes ONE pair per iteration (trainer.py:241-300: best = max
(count, (left bytes, right bytes)), every merge recorded, a new token only when the bytes are new) against a Python BPE that
takes, per iteration, the longest prefix of the exactly ordered pair list that rules (1) - (3) allow and applies its members
word by word, in order.  Both must produce the same merges on corpora built to be hostile: tiny alphabets (every pair touches
some other), small counts (ties everywhere), repeated letters (a == b pairs), words that re-create existing tokens."""
from __future__ import annotations

import random
from collections import Counter

import pytest


def pair_counts(words: dict[tuple, int]) -> Counter:
    c: Counter = Counter()
    for w, f in words.items():
        for x, y in zip(w, w[1:]):
            c[(x, y)] += f
    return c


def apply_merge(w: tuple, a: bytes, b: bytes) -> tuple:
    out, i = [], 0
    while i < len(w):
        if i + 1 < len(w) and w[i] == a and w[i + 1] == b:
            out.append(a + b); i += 2
        else:
            out.append(w[i]); i += 1
    return tuple(out)


def rewrite(words: dict[tuple, int], members: list[tuple[bytes, bytes]]) -> dict[tuple, int]:
    """Every word gets the members applied in order (a word's state depends on nothing but the word)."""
    new: dict[tuple, int] = {}
    for w, f in words.items():
        for a, b in members:
            w = apply_merge(w, a, b)
        new[w] = new.get(w, 0) + f
    return new


def train_sequential(words, n_merges):
    vocab = {bytes([i]) for i in range(256)}
    merges = []
    while len(merges) < n_merges:
        c = pair_counts(words)
        if not c:
            break
        best = max(c.items(), key=lambda kv: (kv[1], kv[0]))[0]
        merges.append(best)
        vocab.add(best[0] + best[1])
        words = rewrite(words, [best])
    return merges


def select_batch(c: Counter, vocab: set, cap: int, t2: int):
    """Rules (1) - (3); `t2` plays the top-list threshold (pairs below it are invisible to the selection)."""
    order = sorted(((cnt, p) for p, cnt in c.items() if cnt >= t2), reverse=True)     # exact order: count, left bytes, right bytes
    if not order:
        return []
    members: list[tuple[bytes, bytes]] = []
    for cnt, (a, b) in order[:cap]:
        if members:
            if any(ma == mb for ma, mb in members):                      # an a == b pair is only taken as the last member
                break
            if any(b == ma or a == mb for ma, mb in members):            # touches an earlier member
                break
        members.append((a, b))
    # a pair left out with the count of the last member must not touch a member
    while len(members) > 1:
        ck = c[members[-1]]
        rest = [p for cnt, p in order[len(members):] if cnt == ck]
        if any(rb == ma or ra == mb for ra, rb in rest for ma, mb in members):
            members.pop()
        else:
            break
    # (3) every member but the last makes a NEW token; no two members make the same bytes
    out: list[tuple[bytes, bytes]] = []
    made: set = set()
    for a, b in members:
        if a + b in made:
            break
        out.append((a, b))
        made.add(a + b)
        if a + b in vocab:
            break
    return out


def train_batched(words, n_merges, cap, t2_of):
    vocab = {bytes([i]) for i in range(256)}
    merges, sizes = [], []
    while len(merges) < n_merges:
        c = pair_counts(words)
        if not c:
            break
        members = select_batch(c, vocab, min(cap, n_merges - len(merges)), t2_of(c))
        if not members:                                                  # nothing above the threshold: the one-merge path
            members = [max(c.items(), key=lambda kv: (kv[1], kv[0]))[0]]
        merges.extend(members)
        sizes.append(len(members))
        vocab.update(a + b for a, b in members)
        words = rewrite(words, members)
    return merges, sizes


def random_words(rng: random.Random, alphabet: bytes, n_words: int, max_len: int, max_freq: int) -> dict[tuple, int]:
    words: dict[tuple, int] = {}
    for _ in range(n_words):
        n = rng.randint(1, max_len)
        w = tuple(bytes([rng.choice(alphabet)]) for _ in range(n))
        words[w] = words.get(w, 0) + rng.randint(1, max_freq)
    return words


@pytest.mark.parametrize("alphabet,n_words,max_len,max_freq", [
    (b"ab", 40, 9, 3), (b"abc", 120, 8, 2), (b"abcd", 200, 7, 5), (b"abcdefgh", 400, 6, 50), (b"etaoinshrdlu ", 600, 9, 1000),
])
def test_batched_selection_equals_the_sequential_loop(alphabet, n_words, max_len, max_freq):
    total_batched = 0
    for seed in range(16):
        rng = random.Random(seed * 7919 + len(alphabet))
        words = random_words(rng, alphabet, n_words, max_len, max_freq)
        want = train_sequential(dict(words), 300)
        for cap, t2_of in ((31, lambda c: 1), (4, lambda c: 1), (31, lambda c: max(c.values()) // 2 + 1)):
            got, sizes = train_batched(dict(words), 300, cap, t2_of)
            assert got == want, (seed, cap)
            total_batched += sum(s for s in sizes if s > 1)
    assert total_batched > 0                     # the rules must actually allow batches on these corpora


def test_rule_details():
    # "aaaa": (a, a) creates (aa, aa) out of itself -- nothing may follow it in a batch
    c = Counter({(b"a", b"a"): 10, (b"x", b"y"): 9})
    assert select_batch(c, set(), 8, 1) == [(b"a", b"a")]
    # a left-out pair with the count of the last member that touches a member ends the batch before that member
    c = Counter({(b"p", b"q"): 10, (b"u", b"v"): 7, (b"t", b"u"): 7})       # order: (u,v) then (t,u): (t,u) touches (u,v)
    assert select_batch(c, set(), 8, 1) == [(b"p", b"q")]
    # ... but not when it is clear of every member
    c = Counter({(b"p", b"q"): 10, (b"u", b"v"): 7, (b"r", b"s"): 7})
    assert select_batch(c, set(), 2, 1) == [(b"p", b"q"), (b"u", b"v")]
    # a member whose bytes exist already is the last one; the same bytes twice never share a batch
    c = Counter({(b"a", b"b"): 10, (b"c", b"d"): 9, (b"e", b"f"): 8})
    assert select_batch(c, {b"cd"}, 8, 1) == [(b"a", b"b"), (b"c", b"d")]
    c = Counter({(b"ab", b"c"): 10, (b"a", b"bc"): 9, (b"e", b"f"): 8})
    assert select_batch(c, set(), 8, 1) == [(b"ab", b"c")]
