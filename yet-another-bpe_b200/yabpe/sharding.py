"""Byte-range shards of ONE corpus: where the text may be cut between GPUs.

The reference pre-tokenises every chunk of a file as one text (trainer.py:146-170); reference chunk
cuts (trainer.py:172-198) are hard boundaries, GPU shard edges are not.  A shard edge therefore has
to sit where cutting changes nothing.  Position p is a SAFE EDGE when

  type A   text[p-1] is an ASCII white-space character other than U+0020 (\\t \\n \\v \\f \\r) and
           text[p] is an ASCII character that is not white space, or
  type B   text[p] is U+0020, text[p-1] and text[p+1] are ASCII characters that are not white space,

and no special token occurs anywhere in text[p - m, p + m + 1) (m = the longest special, in bytes).
By the reference's pattern such a p always starts a pre-token (Appendix A.1: a non-space character
after non-U+0020 white space starts one; a space after a non-space starts one and takes the next
character as its body), no contraction reaches across it ('s 'd 'm 't 'll 've 're contain no white
space and an apostrophe at p is live both after white space and at a text start), and no special token
candidate chain (Appendix A.2) links the two sides.  The text from p on, taken as a text of its own,
thus has exactly the pre-tokens the whole text has from p on.  The shard BEFORE the edge must not see an end of text at
p (a white-space run that ends the text is ONE token, trainer.py:167 `\\s+(?!\\S)`), so every shard is
handed `HALO` bytes beyond its end and counts only the pre-tokens that START before the edge
(`own = (0, edge - start)` in yabpe_pretok_count).

Pure host logic over a `read(lo, hi) -> bytes` callback (a file, a concatenation of files, or a
device buffer): every rank evaluates the same few windows and arrives at the same plan without
communication.  tests/test_sharding_cpu.py checks the claim above against the oracle's scanner.
"""
from __future__ import annotations

from collections.abc import Callable, Sequence

HALO = 512                     # bytes of look-ahead handed to a shard beyond its edge (>= longest special + a code point)
_WINDOW0 = 1 << 16
_WINDOW_MAX = 1 << 26

_WS_NOT_SPACE = frozenset(b"\t\n\v\f\r")


def _ascii_non_ws(b: int) -> bool:
    return b < 0x80 and b != 0x20 and b not in _WS_NOT_SPACE


def _special_near(buf: bytes, i: int, specials: Sequence[bytes], m: int) -> bool:
    """Does any special occur in buf[i - m, i + m + 1)?  (i is an index into buf.)"""
    if not specials:
        return False
    lo, hi = max(0, i - m), min(len(buf), i + m + 1)
    seg = buf[lo:hi]
    return any(s in seg for s in specials)


def find_safe_edge(buf: bytes, base: int, ideal: int, specials: Sequence[bytes], total: int) -> int | None:
    """First safe edge p with ideal <= p, decided from `buf` = text[base, base + len(buf)); None when the window
    holds none.  Needs m bytes of context on both sides of p inside the window (or the true ends of the text)."""
    m = max((len(s) for s in specials), default=0)
    n = len(buf)
    for p in range(max(ideal, base + 1), base + n - 1):
        i = p - base
        c, prev, nxt = buf[i], buf[i - 1], buf[i + 1]
        ok = (prev in _WS_NOT_SPACE and _ascii_non_ws(c)) or (c == 0x20 and _ascii_non_ws(prev) and _ascii_non_ws(nxt))
        if not ok:
            continue
        if (i - m < 0 and base > 0) or (i + m + 1 > n and base + n < total):
            return None                                   # not enough context in this window to rule specials out
        if _special_near(buf, i, specials, m):
            continue
        return p
    return None


def plan_shards(read: Callable[[int, int], bytes], total: int, world: int, specials: Sequence[bytes],
                hard_cuts: Sequence[int] = ()) -> list[int]:
    """Edges e[0] = 0 <= e[1] <= ... <= e[world] = total: rank r owns the pre-tokens that start in [e[r], e[r+1]).
    Edges are safe edges (see the module docstring) at or after r * total / world, or hard cuts of the reference
    (which are exact by definition).  When a window of `_WINDOW_MAX` bytes holds neither, the shard is merged
    into its left neighbour (an empty range): still exact, just not balanced."""
    specials = [s for s in specials if s]
    m = max((len(s) for s in specials), default=0)
    edges = [0]
    hard = sorted(c for c in hard_cuts if 0 < c < total)
    for r in range(1, world):
        ideal = max((total * r) // world, edges[-1])
        edge = None
        if ideal <= 0:
            edge = 0
        elif ideal >= total:
            edge = total
        else:
            win = _WINDOW0
            while edge is None:
                lo = max(0, ideal - m - 1)
                hi = min(total, ideal + win)
                edge = find_safe_edge(read(lo, hi), lo, ideal, specials, total)
                if edge is None:
                    nh = next((c for c in hard if c >= ideal), None)
                    if nh is not None and nh <= hi:
                        edge = nh
                    elif hi >= total or win >= _WINDOW_MAX:
                        edge = nh if nh is not None else total
                    else:
                        win *= 4
            nh = next((c for c in hard if ideal <= c <= edge), None)      # a hard cut on the way is just as good and nearer
            if nh is not None:
                edge = nh
        edges.append(edge)
    edges.append(total)
    return edges


def shard_window(edges: Sequence[int], rank: int, total: int) -> tuple[int, int, int]:
    """(start, own_len, n_local): the bytes rank `rank` needs are text[start, start + n_local); it owns the first
    own_len of them."""
    start, end = edges[rank], edges[rank + 1]
    n_local = min(total, end + HALO) - start if end > start else 0
    return start, end - start, n_local


def reference_chunk_cuts(read: Callable[[int, int], bytes], size: int, chunk_size: int) -> list[int]:
    """trainer.py:139-144,172-198 for one file of `size` bytes, from the <= 5 bytes around each tentative cut."""
    cuts: list[int] = []
    if size <= chunk_size:
        return cuts
    start = 0
    while start < size:
        tentative = min(start + chunk_size, size)
        if tentative < size:
            bstart = max(0, tentative - 4)
            w = read(bstart, tentative + 1)
            pos = tentative - bstart
            while pos > 0 and (w[pos] & 0xC0) == 0x80:
                pos -= 1
            actual = bstart + pos
        else:
            actual = size
        if actual > start:
            cuts.append(actual)
            start = actual
        else:
            start += 1
    return [c for c in cuts if 0 < c < size]


class FileConcat:
    """`read(lo, hi)` over the concatenation of several files (the corpus of BBPETrainer.train(files))."""

    def __init__(self, paths: Sequence, sizes: Sequence[int]) -> None:
        self.paths, self.sizes = list(paths), list(sizes)
        self.starts = [0]
        for s in self.sizes:
            self.starts.append(self.starts[-1] + s)
        self.total = self.starts[-1]

    def read(self, lo: int, hi: int) -> bytes:
        lo, hi = max(0, lo), min(self.total, hi)
        out = []
        for p, s0, size in zip(self.paths, self.starts, self.sizes):
            a, b = max(lo, s0), min(hi, s0 + size)
            if a < b:
                with open(p, "rb") as f:
                    f.seek(a - s0)
                    out.append(f.read(b - a))
        return b"".join(out)

    def readinto(self, lo: int, hi: int, dst) -> None:
        """Fill the writable buffer `dst` (hi - lo bytes) with text[lo, hi)."""
        mv = memoryview(dst).cast("B")
        pos = 0
        for p, s0, size in zip(self.paths, self.starts, self.sizes):
            a, b = max(lo, s0), min(hi, s0 + size)
            if a < b:
                with open(p, "rb", buffering=0) as f:
                    f.seek(a - s0)
                    want = b - a
                    while want:
                        got = f.readinto(mv[pos:pos + want])
                        if not got:
                            raise OSError(f"{p} shrank while it was read")
                        pos += got
                        want -= got

    def hard_cuts(self, chunk_size: int) -> list[int]:
        """Reference chunk cuts of every file plus the file ends, as offsets into the concatenation."""
        cuts: list[int] = []
        for s0, size in zip(self.starts, self.sizes):
            if size:
                cuts += [s0 + c for c in reference_chunk_cuts(lambda a, b, s0=s0: self.read(s0 + a, s0 + b), size, chunk_size)]
                cuts.append(s0 + size)
        return sorted({c for c in cuts if 0 < c < self.total})
