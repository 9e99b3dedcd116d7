"""torchrun worker for tests/test_gpu_distributed.py: sharded training == multi-file training."""
import os
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parents[1]
for p in (ROOT, ROOT / "yet-another-bpe_b200", ROOT / "tests"):
    sys.path.insert(0, str(p))

import numpy as np
import torch
import torch.distributed as dist

import common
import yabpe
from yabpe import engine
from yabpe.distributed import train_device_sharded

local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
rank, world = dist.get_rank(), dist.get_world_size()
shards = [common.synth_owt(2_000_000, seed=500 + r) + (b" " + b"q" * 20000 if r == 1 else b"") for r in range(world)]
cfg = yabpe.BBPETrainerConfig(vocab_size=1500, min_frequency=1, max_workers=1, chunk_size_bytes=1 << 30,
                              special_tokens=["<|endoftext|>"])
text_dev, n = engine.to_device_text(torch, np.frombuffer(shards[rank], dtype=np.uint8))
model = train_device_sharded(yabpe.BBPETrainer(cfg), text_dev, n)
if rank == 0:
    ref = yabpe.BBPETrainer(cfg).train_from_buffers([np.frombuffer(s, dtype=np.uint8) for s in shards])
    assert model.merges == ref.merges, "sharded merges differ from multi-file training"
    assert model.vocab == ref.vocab
    print(f"DIST_OK world={world} merges={len(model.merges)}")
dist.barrier()
dist.destroy_process_group()
