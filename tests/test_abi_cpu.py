"""CPU-only checks of the C-ABI boundary and the host logic (no kernel is launched here).

* libyabpe.so builds (nvcc cross-compiles sm_100a without a GPU), loads, and exports every function that
  include/yabpe.h declares; the ctypes struct layouts match the compiled ones
* the product path fails loudly without a CUDA device (there is no CPU fallback)
* host-side pieces: Unicode classes compiled into the library vs the `regex` module, reference chunk cuts
  (trainer.py:172-198), special-token packing, base vocabulary (trainer.py:119-134)
"""
from __future__ import annotations

import ctypes as C
import re
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "yet-another-bpe_b200"))


@pytest.fixture(scope="module")
def lib():
    import importlib.util
    spec = importlib.util.spec_from_file_location("yabpe_build", ROOT / "yet-another-bpe_b200" / "build.py")
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    mod.build()
    from yabpe import _ffi
    return _ffi.load()


def test_every_declared_function_is_exported(lib):
    header = (ROOT / "include" / "yabpe.h").read_text()
    declared = set(re.findall(r"^(?:const char\*|int|int32_t|int64_t)\s+(yabpe_\w+)\s*\(", header, flags=re.M))
    assert len(declared) >= 12
    from yabpe import _ffi
    assert declared == set(_ffi.EXPORTS)
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.yabpe_abi_version() == _ffi.ABI_VERSION
    assert int(re.search(r"#define YABPE_ABI_VERSION (\d+)", header).group(1)) == _ffi.ABI_VERSION


def test_struct_layouts_match(lib):
    from yabpe import _ffi
    for which, st in enumerate((_ffi.PretokArgs, _ffi.WordTable, _ffi.MergeArgs, _ffi.EncodeModel, _ffi.EncodeOut,
                                _ffi.DecodeArgs, _ffi.PartitionArgs)):
        assert lib.yabpe_sizeof(which) == C.sizeof(st), st.__name__
    assert lib.yabpe_sizeof(99) == -1


def test_integration_md_stub_matches_the_library(lib):
    """INTEGRATION.md shows the ctypes stub a maintainer of the reference would add; its `PretokArgs` must be the struct the
    library was compiled with ("field for field"), and the ABI version it asserts must be the current one."""
    text = (ROOT / "INTEGRATION.md").read_text()
    m = re.search(r"(class PretokArgs\(C\.Structure\):.*?\n)assert lib\.yabpe_sizeof", text, flags=re.S)
    assert m, "stub not found"
    ns = {"C": C}
    exec(m.group(1), ns)
    assert C.sizeof(ns["PretokArgs"]) == lib.yabpe_sizeof(0)
    from yabpe import _ffi
    assert [f[0] for f in ns["PretokArgs"]._fields_] == [f[0] for f in _ffi.PretokArgs._fields_]
    assert int(re.search(r"yabpe_abi_version\(\) == (\d+)", text).group(1)) == _ffi.ABI_VERSION


def test_no_cpu_fallback(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("needs a machine without CUDA")
    import yabpe
    from yabpe import _ffi
    with pytest.raises(_ffi.YabpeUnavailable):
        _ffi.require_cuda()
    with pytest.raises(_ffi.YabpeUnavailable):
        yabpe.Tokenizer({i: bytes([i]) for i in range(256)}, [], []).encode("abc")
    assert "oracle" not in " ".join(m for m in sys.modules if m.startswith("yabpe"))


def test_unicode_classes_match_regex_module(lib):
    regex = pytest.importorskip("regex")
    pats = [regex.compile(r"\p{L}"), regex.compile(r"\p{N}"), regex.compile(r"\s")]
    rng = np.random.default_rng(5)
    cps = list(range(0, 0x3000)) + [int(x) for x in rng.integers(0x3000, 0x110000, 20000)]
    for cp in cps:
        if 0xD800 <= cp <= 0xDFFF:
            continue
        ch = chr(cp)
        want = 1 if pats[0].match(ch) else 2 if pats[1].match(ch) else 3 if pats[2].match(ch) else 0
        assert lib.yabpe_class_of(cp) == want, hex(cp)
    assert lib.yabpe_class_of(0x110000) == 0


def test_chunk_cuts_follow_the_reference_rule():
    from yabpe.trainer import chunk_cuts
    data = np.frombuffer(("ab" + "é" * 50 + "中" * 30 + "\U0001f643" * 20 + "xyz").encode("utf-8"), dtype=np.uint8)
    for cs in (1, 2, 3, 5, 7, 16, 97, 1 << 20):
        cuts = chunk_cuts(data, cs)
        # reference: trainer.py:172-198 (tentative end, moved back <= 4 bytes off continuation bytes)
        want, start, n = [], 0, data.size
        while start < n:
            tent = min(start + cs, n)
            if tent < n:
                b0 = max(0, tent - 4)
                pos = tent - b0
                while pos > 0 and (int(data[b0 + pos]) & 0xC0) == 0x80:
                    pos -= 1
                actual = b0 + pos
            else:
                actual = n
            if actual > start:
                want.append(actual); start = actual
            else:
                start += 1
        assert cuts == want, cs
        assert cuts[-1] == n and all(b > a for a, b in zip(cuts, cuts[1:]))
    assert chunk_cuts(np.zeros(0, dtype=np.uint8), 8) == []


def test_special_packing_and_base_vocab():
    import yabpe
    from yabpe import engine
    blob, offs = engine.pack_specials([b"<|endoftext|>", b"ab"])
    assert offs.tolist() == [0, 13, 15] and bytes(blob[:15]) == b"<|endoftext|>ab"
    with pytest.raises(ValueError):
        engine.pack_specials([b""])
    cfg = yabpe.BBPETrainerConfig()
    assert (cfg.vocab_size, cfg.min_frequency, cfg.chunk_size_bytes) == (32000, 2, 8 * 1024 * 1024)
    tr = yabpe.BBPETrainer(yabpe.BBPETrainerConfig(special_tokens=["<|endoftext|>", "a", "<|endoftext|>"]))
    base = tr._init_base_vocab()
    # 256 bytes, then specials unless their bytes are already a key (duplicate special, 1-byte special): trainer.py:119-134
    assert len(base) == 257 and base[b"<|endoftext|>"] == 256 and base[b"a"] == 97


def test_device_chunk_cuts_equal_host_chunk_cuts_per_file():
    """The streamed upload of train(files) takes the reference chunk cuts from the bytes on the device, file by file
    (trainer.py:_train_streamed_files); the rule only indexes a tensor, so it is checked here on CPU tensors against
    the host rule for several files laid end to end."""
    import torch
    from yabpe.trainer import chunk_cuts, device_chunk_cuts
    rng = np.random.default_rng(5)
    alphabet = ["a", " ", "é", "中", "\U0001f643", "\n"]
    files = ["".join(rng.choice(alphabet, size=int(k))).encode("utf-8") for k in (0, 1, 37, 400, 2, 1500)]
    blob = np.frombuffer(b"".join(files), dtype=np.uint8)
    dev = torch.from_numpy(blob.copy())
    for cs in (1, 3, 4, 7, 64, 97, 1 << 20):
        off, got, want = 0, [], []
        for f in files:
            n = len(f)
            if n:
                got += [off + c for c in device_chunk_cuts(dev[off:off + n], n, cs)] + [off + n]
                want += [off + c for c in chunk_cuts(np.frombuffer(f, dtype=np.uint8), cs)]
            off += n
        assert got == want, cs
