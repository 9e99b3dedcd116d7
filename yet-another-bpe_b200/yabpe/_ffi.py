"""ctypes binding of libyabpe.so (include/yabpe.h).  The ONLY module that touches the C ABI.

There is no CPU fallback: importing the trainer / tokenizer on a machine without the built
library or without a CUDA device fails loudly (`YabpeUnavailable`).
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path

_HERE = Path(__file__).resolve().parent
LIB_PATH = _HERE / "libyabpe.so"

ABI_VERSION = 6

# stats / state slots (include/yabpe.h)
ST_NTOK, ST_UNIQ_SHORT, ST_UNIQ_LONG, ST_UNIQ_BYTES, ST_ERR_POS, ST_TABLE_FULL, ST_OVF_N = 0, 1, 2, 3, 4, 5, 6
ST_NSPECIAL, ST_SLOW_N, ST_CACHE_HIT = 8, 9, 10
MS_NMERGES, MS_NTOK, MS_ERROR, MS_NPAIRS, MS_POOL_USED, MS_REBUILDS, MS_TREBUILDS = 0, 1, 2, 6, 8, 9, 10
MS_LEADER_MERGES, MS_GRID_MERGES = 15, 16
ME_PAIR_TABLE_FULL, ME_TOK_POOL_FULL, ME_INTERNAL = 1, 2, 4
INT64_MAX = (1 << 63) - 1

EXPORTS = [
    "yabpe_last_error", "yabpe_abi_version", "yabpe_device_init", "yabpe_class_of", "yabpe_pretok_count",
    "yabpe_compact_words", "yabpe_merge_loop", "yabpe_encode_words", "yabpe_encode_ids", "yabpe_num_tiles",
    "yabpe_launch_count", "yabpe_insert_words", "yabpe_sizeof", "yabpe_encode_finalize", "yabpe_hot_cache_entries",
    "yabpe_select_hot", "yabpe_token_starts", "yabpe_decode_ids", "yabpe_decode_blocks", "yabpe_publish", "yabpe_partition_words", "yabpe_narrow_ids", "yabpe_encode_small", "yabpe_encode_small_max_bytes",
]


class YabpeUnavailable(RuntimeError):
    pass


class YabpeError(RuntimeError):
    pass


class PretokArgs(C.Structure):
    _fields_ = [
        ("text", C.c_void_p), ("n", C.c_int64),
        ("cuts", C.c_void_p), ("n_cuts", C.c_int32), ("mode", C.c_int32),
        ("sp_blob", C.c_void_p), ("sp_offs", C.c_void_p), ("n_sp", C.c_int32), ("stages", C.c_int32),
        ("own_lo", C.c_int64), ("own_hi", C.c_int64),
        ("cand_bits", C.c_void_p), ("rec_bits", C.c_void_p),
        ("short_keys", C.c_void_p), ("short_counts", C.c_void_p), ("short_cap", C.c_int64),
        ("long_entries", C.c_void_p), ("long_cap", C.c_int64),
        ("ovf_pos", C.c_void_p), ("ovf_cap", C.c_int64),
        ("stats", C.c_void_p),
        ("hot_keys", C.c_void_p), ("work", C.c_void_p), ("work_cap", C.c_int64),
        ("hot_table", C.c_void_p), ("hot_cap", C.c_int64),
    ]


class WordTable(C.Structure):
    _fields_ = [
        ("wsym", C.c_void_p), ("sym_word", C.c_void_p), ("woff", C.c_void_p), ("wlen", C.c_void_p),
        ("wcnt", C.c_void_p), ("sword", C.c_void_p), ("lword", C.c_void_p), ("counters", C.c_void_p),
    ]


class MergeArgs(C.Structure):
    _fields_ = [
        ("words", WordTable), ("n_words", C.c_int64), ("n_syms", C.c_int64),
        ("wstamp", C.c_void_p), ("wslot", C.c_void_p), ("newp", C.c_void_p),
        ("tok_bytes", C.c_void_p), ("tok_bytes_cap", C.c_int64),
        ("tok_off", C.c_void_p), ("tok_hash", C.c_void_p), ("tok_pow", C.c_void_p), ("tok_pre", C.c_void_p),
        ("tset", C.c_void_p), ("tset_cap", C.c_int64), ("max_tokens", C.c_int64),
        ("pkey", C.c_void_p), ("pcnt", C.c_void_p), ("pcap", C.c_int64),
        ("ioff", C.c_void_p), ("icnt", C.c_void_p), ("ipost", C.c_void_p), ("inact", C.c_void_p), ("intop", C.c_void_p), ("top_slot", C.c_void_p), ("top_key", C.c_void_p), ("hist", C.c_void_p),
        ("act", C.c_void_p),
        ("alog_word", C.c_void_p), ("alog_cap", C.c_int64), ("seg_start", C.c_void_p), ("seg_end", C.c_void_p),
        ("merge_next", C.c_void_p), ("tok_first", C.c_void_p), ("tok_head", C.c_void_p),
        ("partial", C.c_void_p), ("bsum", C.c_void_p),
        ("merges", C.c_void_p), ("merge_new", C.c_void_p), ("state", C.c_void_p),
        ("num_merges", C.c_int64), ("min_frequency", C.c_int64), ("rebuild_every", C.c_int64), ("helper_min_syms", C.c_int64), ("helper_mode", C.c_int64), ("batch_max", C.c_int64),
    ]


class EncodeModel(C.Structure):
    _fields_ = [
        ("mkey", C.c_void_p), ("mval", C.c_void_p), ("mcap", C.c_int64),
        ("byte_sym", C.c_void_p), ("sym_out", C.c_void_p), ("sp_ids", C.c_void_p),
        ("consistent", C.c_int32), ("_pad", C.c_int32),
    ]


class EncodeOut(C.Structure):
    _fields_ = [("tile_count", C.c_void_p), ("out_ids", C.c_void_p), ("out_cap", C.c_int64), ("doc_off", C.c_void_p)]


class DecodeArgs(C.Structure):
    _fields_ = [("ids", C.c_void_p), ("n_ids", C.c_int64), ("tok_off", C.c_void_p), ("tok_bytes", C.c_void_p),
                ("vocab_cap", C.c_int32), ("_pad", C.c_int32), ("block_count", C.c_void_p),
                ("out", C.c_void_p), ("out_cap", C.c_int64)]


class PartitionArgs(C.Structure):
    _fields_ = [("words", WordTable), ("n_words", C.c_int64), ("n_ranks", C.c_int32), ("_pad", C.c_int32),
                ("dest", C.c_void_p), ("totals", C.c_void_p), ("base_w", C.c_void_p), ("base_b", C.c_void_p),
                ("cursor", C.c_void_p), ("out_lens", C.c_void_p), ("out_cnts", C.c_void_p), ("out_data", C.c_void_p)]


_lib: C.CDLL | None = None
_device_ready: set[int] = set()


def load() -> C.CDLL:
    """dlopen libyabpe.so and declare prototypes.  Does not need a GPU."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise YabpeUnavailable(
            f"{LIB_PATH} is missing: build it with `python yet-another-bpe_b200/build.py` "
            "(nvcc, sm_100a).  There is no CPU fallback.")
    L = C.CDLL(str(LIB_PATH))
    L.yabpe_last_error.restype = C.c_char_p
    L.yabpe_abi_version.restype = C.c_int
    L.yabpe_device_init.restype = C.c_int
    L.yabpe_class_of.restype = C.c_int
    L.yabpe_class_of.argtypes = [C.c_uint32]
    L.yabpe_pretok_count.restype = C.c_int
    L.yabpe_pretok_count.argtypes = [C.POINTER(PretokArgs), C.c_void_p]
    L.yabpe_compact_words.restype = C.c_int
    L.yabpe_compact_words.argtypes = [C.POINTER(PretokArgs), C.POINTER(WordTable), C.c_void_p]
    L.yabpe_insert_words.restype = C.c_int
    L.yabpe_insert_words.argtypes = [C.POINTER(PretokArgs), C.c_void_p, C.c_void_p, C.c_void_p, C.c_int64, C.c_int32, C.c_void_p]
    L.yabpe_merge_loop.restype = C.c_int
    L.yabpe_merge_loop.argtypes = [C.POINTER(MergeArgs), C.c_void_p]
    L.yabpe_encode_words.restype = C.c_int
    L.yabpe_encode_words.argtypes = [C.POINTER(EncodeModel), C.POINTER(WordTable), C.c_int64, C.c_void_p]
    L.yabpe_encode_finalize.restype = C.c_int
    L.yabpe_encode_finalize.argtypes = [C.POINTER(PretokArgs), C.POINTER(EncodeModel), C.POINTER(WordTable), C.c_int64, C.c_void_p]
    L.yabpe_encode_ids.restype = C.c_int
    L.yabpe_encode_ids.argtypes = [C.POINTER(PretokArgs), C.POINTER(EncodeModel), C.POINTER(WordTable),
                                   C.POINTER(EncodeOut), C.c_int32, C.c_void_p]
    L.yabpe_num_tiles.restype = C.c_int64
    L.yabpe_num_tiles.argtypes = [C.c_int64, C.c_int64]
    L.yabpe_launch_count.restype = C.c_int64
    L.yabpe_hot_cache_entries.restype = C.c_int32
    L.yabpe_select_hot.restype = C.c_int
    L.yabpe_select_hot.argtypes = [C.POINTER(PretokArgs), C.c_void_p, C.c_void_p, C.c_void_p]
    L.yabpe_token_starts.restype = C.c_int
    L.yabpe_token_starts.argtypes = [C.POINTER(PretokArgs), C.c_void_p, C.c_void_p]
    L.yabpe_decode_ids.restype = C.c_int
    L.yabpe_decode_ids.argtypes = [C.POINTER(DecodeArgs), C.c_int32, C.c_void_p]
    L.yabpe_decode_blocks.restype = C.c_int64
    L.yabpe_decode_blocks.argtypes = [C.c_int64]
    L.yabpe_publish.restype = C.c_int
    L.yabpe_publish.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p]
    L.yabpe_encode_small.restype = C.c_int
    L.yabpe_encode_small.argtypes = [C.POINTER(EncodeModel), C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_int32,
                                     C.c_void_p, C.c_void_p, C.c_int32, C.c_void_p]
    L.yabpe_encode_small_max_bytes.restype = C.c_int32
    L.yabpe_narrow_ids.restype = C.c_int
    L.yabpe_narrow_ids.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]
    L.yabpe_partition_words.restype = C.c_int
    L.yabpe_partition_words.argtypes = [C.POINTER(PartitionArgs), C.c_int32, C.c_void_p]
    L.yabpe_sizeof.restype = C.c_int64
    L.yabpe_sizeof.argtypes = [C.c_int32]
    for which, st in enumerate((PretokArgs, WordTable, MergeArgs, EncodeModel, EncodeOut, DecodeArgs, PartitionArgs)):
        if L.yabpe_sizeof(which) != C.sizeof(st):
            raise YabpeUnavailable(f"{st.__name__}: ctypes layout ({C.sizeof(st)} B) != libyabpe.so ({L.yabpe_sizeof(which)} B); rebuild")
    if L.yabpe_abi_version() != ABI_VERSION:
        raise YabpeUnavailable(f"libyabpe.so ABI {L.yabpe_abi_version()} != expected {ABI_VERSION}; rebuild")
    _lib = L
    return L


def check(rc: int) -> None:
    if rc != 0:
        raise YabpeError(f"libyabpe error {rc}: {load().yabpe_last_error().decode(errors='replace')}")


def require_cuda():
    """Return torch after making sure a CUDA device and the native library are usable."""
    import torch
    if not torch.cuda.is_available():
        raise YabpeUnavailable("yabpe needs a CUDA device (B200 / sm_100a); there is no CPU fallback")
    L = load()
    dev = torch.cuda.current_device()
    if dev not in _device_ready:
        check(L.yabpe_device_init())
        _device_ready.add(dev)
    return torch


def stream_ptr(torch) -> int:
    return torch.cuda.current_stream().cuda_stream


def launch_count() -> int:
    return int(load().yabpe_launch_count())
