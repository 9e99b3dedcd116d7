"""2-GPU parity: pre-tokenise/count sharded over ranks (byte ranges of one corpus, and files per rank) + NCCL
all-to-all == the oracle and == single-GPU training.  bench.py runs the same worker at N = 2 (`--self-check`)."""
from __future__ import annotations

import socket
import subprocess
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]
pytestmark = pytest.mark.gpu


def test_sharded_training_two_gpus():
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (run with gpurun --gpus 2)")
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    res = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                          "--master-addr", "127.0.0.1", "--master-port", str(port), str(ROOT / "tests" / "dist_gpu_worker.py")],
                         capture_output=True, text=True, timeout=900)
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-3000:]
    assert "DIST_OK" in res.stdout and res.stdout.count("RANGE_OK") == 2 and "ENCODE_OK" in res.stdout
