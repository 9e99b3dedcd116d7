"""torchrun worker for tests/test_gpu_distributed.py and bench.py's N > 1 self-check: sharded training == the ORACLE.

1. byte ranges of ONE corpus (train_files_sharded: safe edges + halo) -- two files, a forced small
   reference chunk size so that hard cuts fall inside shards, dense specials, a 20 000-byte pre-token
2. the files-per-rank form (train_device_sharded)
"""
import os
import sys
import tempfile
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
for p in (ROOT, ROOT / "yet-another-bpe_b200", ROOT / "tests"):
    sys.path.insert(0, str(p))

import numpy as np
import torch
import torch.distributed as dist

import common
import yabpe
from oracle import oracle
from yabpe import engine
from yabpe.distributed import train_device_sharded, train_files_sharded

local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
rank, world = dist.get_rank(), dist.get_world_size()
SP = ["<|endoftext|>"]

# ---- 1. byte ranges of one corpus
blobs = [common.synth_owt(3_000_000, seed=700) + b" " + b"q" * 20000 + b"\n" + common.synth_tinystories(1_500_000, seed=701),
         (b"a<|endoftext|><|endoftext|>b !<|endoftext|>\n" * 2000) + common.synth_adversarial(400_000, seed=702)]
tmp = Path(tempfile.gettempdir()) / "yabpe_dist_test"
if rank == 0:
    tmp.mkdir(exist_ok=True)
    for i, b in enumerate(blobs):
        (tmp / f"f{i}.txt").write_bytes(b)
dist.barrier()
files = [tmp / f"f{i}.txt" for i in range(len(blobs))]
for chunk in (1 << 30, 1_000_003):
    cfg = yabpe.BBPETrainerConfig(vocab_size=1500, min_frequency=1, max_workers=1, chunk_size_bytes=chunk, special_tokens=SP)
    model = train_files_sharded(yabpe.BBPETrainer(cfg), files)
    if rank == 0:
        tr = oracle.Trainer(SP)
        for b in blobs:
            tr.feed_bytes(b, chunk)
        vocab, merges = tr.run(1500, 1, True)
        assert model.merges == merges, f"byte-range sharded merges differ from the oracle (chunk {chunk})"
        assert {v: k for k, v in model.vocab.items()} == vocab
        one = yabpe.BBPETrainer(cfg).train(files)
        assert one.merges == model.merges and one.vocab == model.vocab
        print(f"RANGE_OK world={world} chunk={chunk} merges={len(merges)}", flush=True)
    else:
        assert model is None

# ---- 2. rank r's text is file r
shards = [common.synth_owt(2_000_000, seed=500 + r) + (b" " + b"q" * 20000 if r == 1 else b"") for r in range(world)]
cfg = yabpe.BBPETrainerConfig(vocab_size=1500, min_frequency=1, max_workers=1, chunk_size_bytes=1 << 30, special_tokens=SP)
text_dev, n = engine.to_device_text(torch, np.frombuffer(shards[rank], dtype=np.uint8))
model = train_device_sharded(yabpe.BBPETrainer(cfg), text_dev, n)
if rank == 0:
    tr = oracle.Trainer(SP)
    for s in shards:
        tr.feed_bytes(s)
    vocab, merges = tr.run(1500, 1, True)
    assert model.merges == merges, "files-per-rank sharded merges differ from the oracle"
    assert {v: k for k, v in model.vocab.items()} == vocab
    print(f"DIST_OK world={world} merges={len(model.merges)}", flush=True)

# ---- 3. document-sharded encode (encode_sharded) == the oracle's encode of the whole text
from yabpe.distributed import encode_sharded
gv, gm = common.gpt2_vocab_and_merges()
tok = yabpe.Tokenizer(gv, gm, SP).inner
text = common.synth_owt(6_000_000, seed=900)
if rank == 0:
    (tmp / "enc.txt").write_bytes(text)
dist.barrier()
ids, off, total = encode_sharded(tok, tmp / "enc.txt", gather=True, piece_bytes=1 << 20)
if rank == 0:
    want = oracle.Tokenizer(gv, gm, SP).encode(text.decode("utf-8"))
    assert total == len(want) and ids.tolist() == want, "document-sharded encode differs from the oracle"
    print(f"ENCODE_OK world={world} ids={total}", flush=True)
else:
    assert 0 < off < total
dist.barrier()
dist.destroy_process_group()
