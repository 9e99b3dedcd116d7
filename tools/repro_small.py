import sys, ctypes as C
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path[:0] = [str(ROOT), str(ROOT / "yet-another-bpe_b200"), str(ROOT / "tests")]
import numpy as np
import torch
import common, yabpe
from oracle import oracle
from yabpe import _ffi, engine
vocab, merges = common.gpt2_vocab_and_merges()
tok = yabpe.Tokenizer(vocab, merges, ["<|endoftext|>"]).inner
otok = oracle.Tokenizer(vocab, merges, ["<|endoftext|>"])
torch = _ffi.require_cuda()
L = _ffi.load()
e = tok._device_model(torch)
blob, offs = engine.pack_specials(tok._sp_bytes)
for nsp in (0, 1):
  for mode in ("device", "pinned_out", "pinned_both"):
    for s in ["hello world", "Hello, how are you?<|endoftext|> fine\n\nthanks " * 20]:
        raw = s.encode()
        n = len(raw)
        tin_d = torch.zeros(32768 + 64, dtype=torch.uint8, device="cuda"); tin_d[:n] = torch.from_numpy(np.frombuffer(raw, dtype=np.uint8).copy()).cuda()
        tin_p = torch.zeros(32768 + 64, dtype=torch.uint8).pin_memory(); tin_p.numpy()[:n] = np.frombuffer(raw, dtype=np.uint8)
        out_d = torch.zeros(32768 + 8, dtype=torch.int32, device="cuda")
        out_p = torch.zeros(32768 + 8, dtype=torch.int32).pin_memory()
        scratch = torch.empty(32768 + 64, dtype=torch.int32, device="cuda")
        tin = tin_p if mode == "pinned_both" else tin_d
        out = out_d if mode == "device" else out_p
        print("try", nsp, mode, n, flush=True)
        _ffi.check(L.yabpe_encode_small(C.byref(e), tin.data_ptr(), n, blob.ctypes.data, offs.ctypes.data, nsp, scratch.data_ptr(), out.data_ptr(), n + 1, _ffi.stream_ptr(torch)))
        torch.cuda.synchronize()
        o = out.cpu().numpy()
        print("   ->", o[0], o[1:1 + max(int(o[0]), 0)][:10].tolist(), otok.encode(s)[:10] if nsp else "", flush=True)
