"""yabpe/_hostlist.so (csrc/hostlist.c): the Python objects of a training result built with the C API must be exactly what the
Python construction gives -- the objects tests/adapters.py:66-99 hands back (dict[bytes, int] turned around by the adapter,
list[tuple[bytes, bytes]])."""
from __future__ import annotations

import importlib.util
import random
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
PKG = ROOT / "yet-another-bpe_b200"
if str(PKG) not in sys.path:
    sys.path.insert(0, str(PKG))


def _build():
    spec = importlib.util.spec_from_file_location("yabpe_build", PKG / "build.py")
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.build_hostlist()


def _result(tokens: list[bytes], merges: np.ndarray):
    from yabpe import engine
    pool = b"".join(tokens)
    offs = np.cumsum([0] + [len(t) for t in tokens]).astype(np.int64)
    return engine.MergeResult(merges=merges, merge_new=np.zeros(max(len(merges), 1), dtype=np.int32),
                              state=np.zeros(64, dtype=np.int64), pool=pool, offs=offs)


def test_c_and_python_construction_agree(monkeypatch):
    if _build() is None:
        pytest.skip("no Python.h / gcc here: the package uses its Python construction")
    from yabpe import engine
    import importlib
    hostlist = importlib.import_module("yabpe._hostlist")
    rng = random.Random(3)
    cases = []
    base = [bytes([i]) for i in range(256)] + [b"<|endoftext|>"]
    for n in (0, 1, 7, 2000):
        toks = base + [bytes(rng.choices(range(256), k=rng.randint(1, 40))) for _ in range(n)]
        mg = np.asarray([[rng.randrange(len(toks)), rng.randrange(len(toks))] for _ in range(n)], dtype=np.int32).reshape(-1, 2)
        cases.append((toks, mg))
    cases.append((base + [b"ab", b"ab", b""], np.asarray([[97, 98], [257, 258], [259, 259]], dtype=np.int32)))   # equal bytes twice, an empty token
    for toks, mg in cases:
        monkeypatch.setattr(engine, "_HOSTLIST", hostlist)
        got = _result(toks, mg).materialise()
        monkeypatch.setattr(engine, "_HOSTLIST", None)
        want = _result(toks, mg).materialise()
        assert got == want
        assert got[0] == toks
        assert got[1] == {b: i for i, b in enumerate(toks)}
        assert got[2] == [(toks[a], toks[b]) for a, b in mg.tolist()]
        assert all(type(t) is tuple and type(t[0]) is bytes and type(t[1]) is bytes for t in got[2])
        assert all(type(v) is int for v in got[1].values())


def test_bad_inputs_raise():
    if _build() is None:
        pytest.skip("no Python.h / gcc here")
    import importlib
    hostlist = importlib.import_module("yabpe._hostlist")
    pool = b"abc"
    with pytest.raises(ValueError):
        hostlist.materialise(pool, np.asarray([0, 2, 9], dtype=np.int64), np.zeros((0, 2), dtype=np.int32))        # offset beyond the pool
    with pytest.raises(ValueError):
        hostlist.materialise(pool, np.asarray([0, 1, 3], dtype=np.int64), np.asarray([[0, 5]], dtype=np.int32))     # unknown token id
    with pytest.raises(ValueError):
        hostlist.materialise(pool, np.asarray([2, 1], dtype=np.int64), np.zeros((0, 2), dtype=np.int32))           # decreasing offsets
