#!/usr/bin/env python3
"""bench.py -- throughput of the BPE training hot path on B200 (contract: see the task statement).

A "step" is one full `train_bpe` (pre-tokenise + count + word table + merge loop) over one batch
of synthetic text.  Default workload (N=1): BASELINE.json configs[1] -- TinyStories-shaped
synthetic corpus, 2e9 bytes, vocab 10 000, special token <|endoftext|>.

  value        corpus MB / s, inputs already resident in HBM when the timed region starts
  e2e          the same through the host-buffer API (pinned host bytes -> H2D -> train -> D2H of
               merges / vocab), copies inside the timed region
  roofline     dominant HBM kernel (k_pretok_warp): corpus bytes / its CUDA-event duration vs the
               measured HBM copy peak in MEASURED_PEAKS.json
  cpu_baseline the CPU oracle port (oracle/bpe_oracle.c, reference algorithm) on a bounded sample
  --impl reference   the same oracle port on the host cores (the reference is pure Python and is
               not present on the GPU box; the port follows trainer.py line by line)
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
for p in (ROOT, ROOT / "yet-another-bpe_b200", ROOT / "tools"):
    if str(p) not in sys.path:
        sys.path.insert(0, str(p))

WORKLOADS = {
    # name: (kind, bytes, vocab, seed)
    "tinystories-2g-v10k": ("tinystories", 2_000_000_000, 10_000, 20260101),
    "owt-11g-v32k": ("owt", 11_000_000_000, 32_000, 20260102),
    "tinystories-256m-v10k": ("tinystories", 256_000_000, 10_000, 20260101),
    "owt-1g-v32k": ("owt", 1_000_000_000, 32_000, 20260102),
    # BASELINE.json configs[3]: GPT-2 50257-vocab encode of 1e9 bytes of OWT-shaped text, sharded by document
    # (strong scaling: every rank encodes 1e9 / world bytes); a different metric (encode MB/s), selected explicitly
    "gpt2-encode-1g": ("owt", 1_000_000_000, 50_257, 20260103),
}
SPECIALS = ["<|endoftext|>"]
# (dram__bytes_read + dram__bytes_write) / corpus bytes of k_pretok_warp from the committed ncu --set full captures
# (profiles/r1_ncu_pretok_warp_details.txt, profiles/r1_ncu_pretok_warp_owt.txt): the OWT-shaped corpus has 50x the unique
# pre-tokens, so its count-table traffic dwarfs the text itself
NCU_TRAFFIC_RATIO = {"tinystories": 1.24, "owt": 7.0}
# the same for the encode tile passes (profiles/r1_ncu_encode_tiles.txt: 256 MB of OWT-shaped text, 3.0 M unique words): DRAM bytes per
# TEXT byte -- the table probes (one 32-byte sector per token for the key + lookup record, another for its ids) dominate
NCU_ENCODE_TRAFFIC_PER_TEXT_BYTE = {"count_pass_ms": 4.42, "write_pass_ms": 11.24}
METRIC = "train_bpe corpus throughput (pretokenize+count+merge loop)"
ENCODE_METRIC = "GPT-2 encode throughput (pretokenize + BPE by rank + ids in text order)"
UNIT = "MB/s"


def measured_peak_gbs() -> tuple[float, str]:
    f = ROOT / "MEASURED_PEAKS.json"
    if f.exists():
        try:
            return float(json.loads(f.read_text())["hbm_gbs"]), "measured"
        except Exception:
            pass
    return 6650.0, "fallback"


class ClockSampler:
    """SM clock + throttle reasons during the timed region: NVML (pynvml, ~20 samples per 100 ms) when it loads, else one
    `nvidia-smi` query after the other (each takes longer than a short timed region, so few samples)."""

    def __init__(self, index: int = 0):
        self.rows: list[list[str]] = []
        self._stop = threading.Event()
        self.index = index
        self.source = "nvidia-smi"
        self._nvml = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self._handle = self._cuda_device_handle(pynvml, index)
            pynvml.nvmlDeviceGetClockInfo(self._handle, pynvml.NVML_CLOCK_SM)
            self._nvml, self.source = pynvml, "nvml"
        except Exception:
            self._nvml = None
        self._thread = threading.Thread(target=self._run_nvml if self._nvml else self._run, daemon=True)

    @staticmethod
    def _cuda_device_handle(pynvml, index: int):
        """NVML handle of CUDA device `index` (by UUID: CUDA_VISIBLE_DEVICES may renumber the devices)."""
        try:
            import torch
            uuid = str(torch.cuda.get_device_properties(index).uuid)
            return pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + uuid if not uuid.startswith("GPU-") else uuid).encode())
        except Exception:
            return pynvml.nvmlDeviceGetHandleByIndex(index)

    def _run_nvml(self):
        nv, h = self._nvml, self._handle
        bits = [(nv.nvmlClocksThrottleReasonHwSlowdown, 2), (nv.nvmlClocksThrottleReasonHwThermalSlowdown, 3),
                (nv.nvmlClocksThrottleReasonSwThermalSlowdown, 4), (nv.nvmlClocksThrottleReasonSwPowerCap, 5)]
        while not self._stop.is_set():
            try:
                row = [str(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)), str(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)),
                       "Not Active", "Not Active", "Not Active", "Not Active"]
                mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                for bit, col in bits:
                    if mask & bit:
                        row[col] = "Active"
                self.rows.append(row)
            except Exception:
                pass
            self._stop.wait(0.005)

    def _run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        while not self._stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--id={self.index}", f"--query-gpu={q}", "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            self._stop.wait(0.2)

    def __enter__(self):
        self._thread.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        self._thread.join(timeout=5)

    def summary(self) -> dict:
        sm = sorted(int(r[0]) for r in self.rows if r and r[0].isdigit())
        mx = [int(r[1]) for r in self.rows if len(r) > 1 and r[1].isdigit()]
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            for k, nme in enumerate(names):
                if len(r) > 2 + k and r[2 + k].lower().startswith("active"):
                    reasons.add(nme)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(self.rows), "source": self.source}


def _trim_utf8(b: bytes) -> bytes:
    while b and (b[-1] & 0xC0) == 0x80:
        b = b[:-1]
    if b and b[-1] >= 0xC0:
        b = b[:-1]
    return b


def cpu_port_run(sample: bytes, vocab: int) -> tuple[float, int]:
    """One timed run of the oracle port on `sample`; returns (seconds, merges)."""
    from oracle import oracle
    t0 = time.perf_counter()
    _, merges = oracle.train_bpe_bytes(sample, vocab, SPECIALS, fast=False)
    return time.perf_counter() - t0, len(merges)


def run_reference(args) -> None:
    """--impl reference: the CPU port of the reference algorithm on the host cores (rank 0 only)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    kind, nbytes, vocab, seed = WORKLOADS[args.workload]
    sys.path.insert(0, str(ROOT / "tests"))
    import common
    sample_bytes = args.cpu_sample_mb << 20
    sample = None
    try:                                      # same generator as the GPU arm when a device is present
        import torch
        if torch.cuda.is_available():
            from synth_gpu import synth_corpus_device
            dev_text, dn = synth_corpus_device(torch, sample_bytes, kind, seed, piece_bytes=min(sample_bytes, 256 << 20))
            sample = _trim_utf8(dev_text[:dn].cpu().numpy().tobytes())
            del dev_text
            torch.cuda.empty_cache()
    except Exception:
        sample = None
    if sample is None:
        gen = common.synth_tinystories if kind == "tinystories" else common.synth_owt
        sample = gen(sample_bytes, seed=seed)
    if args.workload.startswith("gpt2-encode"):
        # BASELINE.json configs[3]: the reference's encode (tokenizer.py:152-308) as the C port, GPT-2 vocabulary
        from oracle import oracle
        gv, gm = common.gpt2_vocab_and_merges()
        otok = oracle.Tokenizer(gv, gm, SPECIALS)
        text = sample.decode("utf-8")
        for _ in range(args.warmup):
            otok.encode(text[: 1 << 20])
        t0 = time.perf_counter()
        for _ in range(args.steps):
            n_ids = len(otok.encode(text))
        t = time.perf_counter() - t0
        mbps = len(sample) * args.steps / t / 1e6
        print(json.dumps({
            "impl": "reference", "metric": ENCODE_METRIC, "value": round(mbps, 3), "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(1000 * t / args.steps, 2), "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "u8/int32", "data": "synthetic",
            "config": {"workload": args.workload, "vocab_size": len(gv), "merges": len(gm), "special_tokens": SPECIALS,
                       "note": "reference is pure Python (not on the GPU box); timed: C port of tokenizer.py, one core -- the Python reference "
                               "itself measured 2.1-2.5 MB/s in the authoring container (BASELINE.md section 2)"},
            "cpu_baseline": {"value": round(mbps, 3), "unit": UNIT, "cores": 1, "kind": "port",
                             "sample": f"{len(sample)} bytes of the {kind}-shaped generator (seed {seed}), {n_ids} ids per step"},
            "e2e": {"value": round(mbps, 3), "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}))
        return
    for _ in range(args.warmup):
        cpu_port_run(sample[: 1 << 20], vocab)
    t = 0.0
    nm = 0
    for _ in range(args.steps):
        dt, nm = cpu_port_run(sample, vocab)
        t += dt
    mbps = len(sample) * args.steps / t / 1e6
    line = {
        "impl": "reference", "metric": METRIC, "value": round(mbps, 3), "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(1000 * t / args.steps, 2),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8/int64", "data": "synthetic",
        "config": {"workload": args.workload, "vocab_size": vocab, "special_tokens": SPECIALS,
                   "note": "reference is pure Python (not on the GPU box); timed: C port of trainer.py, linear max() scan -- the Python "
                           "reference itself measured 2.8-3.0 MB/s on this stage in the authoring container (BASELINE.md section 2)"},
        "cpu_baseline": {"value": round(mbps, 3), "unit": UNIT, "cores": 1, "kind": "port",
                         "sample": f"{len(sample)} bytes of the {kind}-shaped generator (numpy, seed {seed}), "
                                   f"vocab {vocab}, {nm} merges; reference uses max_workers=1 (threads are GIL-bound)"},
        "e2e": {"value": round(mbps, 3), "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def run_encode(args) -> None:
    """--workload gpt2-encode-1g: batched encode with the GPT-2 vocabulary / merges (tests/fixtures_gpt2)."""
    import numpy as np
    import torch
    import yabpe
    from synth_gpu import synth_corpus_device
    from yabpe import _ffi, engine
    sys.path.insert(0, str(ROOT / "tests"))
    import common
    world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    kind, nbytes, vocab_n, seed = WORKLOADS[args.workload]
    vocab, merges = common.gpt2_vocab_and_merges()
    tok = yabpe.Tokenizer(vocab, merges, SPECIALS).inner
    shard = nbytes // world                                  # documents are independent: shard by document, no exchange
    text_dev, n = synth_corpus_device(torch, shard, kind, seed + rank, lex_seed=seed)
    torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        tok.encode_device(text_dev, n, reuse_output=True)
    barrier()
    launches0 = _ffi.launch_count()
    tok.profile = True
    timings = []
    with ClockSampler(local) as clocks:
        ev0 = torch.cuda.Event(enable_timing=True); ev1 = torch.cuda.Event(enable_timing=True)
        ev0.record()
        for _ in range(args.steps):
            ids, _ = tok.encode_device(text_dev, n, reuse_output=True)
            timings.append(dict(tok.timing))
        ev1.record()
        barrier()
        ms_total = ev0.elapsed_time(ev1)
    tok.profile = False
    launches = _ffi.launch_count() - launches0
    n_ids = int(ids.numel())
    tot = torch.tensor([float(n), float(n_ids), ms_total], device="cuda", dtype=torch.float64)
    if world > 1:
        mx = tot.clone(); dist.all_reduce(tot, op=dist.ReduceOp.SUM); dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        ms_total = float(mx[2].item())
    total_bytes, total_ids = float(tot[0].item()), float(tot[1].item())
    ms_step = ms_total / args.steps
    value = total_bytes / (ms_step / 1e3) / 1e6
    peak, peak_kind = measured_peak_gbs()
    stage = {k: round(float(np.mean([t[k] for t in timings])), 3) for k in timings[0] if k.endswith("_ms")}
    # dominant kernel: the write pass reads the text once more and writes 4 bytes per id
    hbm_stages = {k: v for k, v in stage.items() if k != "words_ms"}      # k_encode_words is latency-bound work on the UNIQUE words
    dom = max(hbm_stages, key=hbm_stages.get)
    alg = {"pretok_count_ms": n, "count_pass_ms": n, "write_pass_ms": n + 4 * n_ids}[dom]
    roofline = {"bound": "hbm", "kernel": {"pretok_count_ms": "k_pretok_count (mode 1)", "count_pass_ms": "k_encode_tiles<false>",
                                            "write_pass_ms": "k_encode_tiles<true>", "words_ms": "k_encode_words"}[dom],
                "achieved": round(alg / (stage[dom] / 1e3) / 1e9, 2), "peak": peak, "unit": "GB/s",
                "frac": round(alg / (stage[dom] / 1e3) / 1e9 / peak, 4),
                "traffic": int(n * NCU_ENCODE_TRAFFIC_PER_TEXT_BYTE[dom]) if dom in NCU_ENCODE_TRAFFIC_PER_TEXT_BYTE else None,
                "traffic_source": "dram bytes per text byte from the committed ncu --set full capture (profiles/r1_ncu_encode_tiles.txt), scaled to this launch",
                "peak_source": peak_kind, "algorithmic_bytes_per_launch": int(alg), "ms_per_launch": stage[dom]}
    # the way back (tokenizer.py:323-349 decode): ids -> bytes on the device; every byte must be the input's
    decode = None
    if n_ids:
        ids_keep = ids.clone()
        out = tok.decode_device(ids_keep)
        torch.cuda.synchronize()
        d0 = torch.cuda.Event(enable_timing=True); d1 = torch.cuda.Event(enable_timing=True)
        d0.record()
        for _ in range(args.steps):
            out = tok.decode_device(ids_keep)
        d1.record(); torch.cuda.synchronize()
        dms = d0.elapsed_time(d1) / args.steps
        decode = {"ms": round(dms, 3), "GB/s (4 B per id in + bytes out)": round((4 * n_ids + int(out.numel())) / (dms / 1e3) / 1e9, 1),
                  "round_trip_equal": bool(out.numel() == n and torch.equal(out, text_dev[:n]))}
        del ids_keep, out
    e2e = None
    if not args.skip_e2e:
        host = torch.empty(n, dtype=torch.uint8).pin_memory()
        host.copy_(text_dev[:n])
        ids_h = torch.empty(n_ids, dtype=torch.int32).pin_memory()          # pinned landing buffer for the ids

        def e2e_step():
            if args.encode_e2e == "pipelined":    # the host-buffer API: pieces cut after specials, copies overlap the encode
                return int(tok.encode_pinned(host, out=ids_h, piece_bytes=args.piece_mb << 20).numel())
            dev2, n2 = engine.to_device_text(torch, host, non_blocking=True)
            ids2, _ = tok.encode_device(dev2, n2, reuse_output=True)
            ids_h[: ids2.numel()].copy_(ids2, non_blocking=True)
            return int(ids2.numel())

        ids_ref = ids.clone()                     # device-resident result of the timed region (the reuse buffer is overwritten below)
        e2e_step()                                # warm-up: allocator blocks for the text copy
        torch.cuda.synchronize()
        assert torch.equal(ids_h[:n_ids].cuda(), ids_ref), "end-to-end ids differ from the device-resident run"
        del ids_ref
        reps = max(1, min(args.steps, 3))
        barrier()
        t0 = time.perf_counter()
        for _ in range(reps):
            got = e2e_step()
        barrier()
        dt = (time.perf_counter() - t0) / reps
        assert got == n_ids
        t0 = time.perf_counter()                  # the bare copies on this box (one after the other), for reading the number above
        text_dev[:n].copy_(host, non_blocking=True); torch.cuda.synchronize()
        h2d_only_ms = (time.perf_counter() - t0) * 1e3
        t0 = time.perf_counter()
        ids_h[:n_ids].copy_(ids[:n_ids], non_blocking=True); torch.cuda.synchronize()
        d2h_only_ms = (time.perf_counter() - t0) * 1e3
        if world > 1:
            tm = torch.tensor([dt], device="cuda", dtype=torch.float64); dist.all_reduce(tm, op=dist.ReduceOp.MAX); dt = float(tm.item())
        e2e = {"value": round(total_bytes / dt / 1e6, 2), "unit": UNIT, "h2d_bytes_per_step": int(total_bytes),
               "d2h_bytes_per_step": int(4 * total_ids), "ms_per_step": round(dt * 1e3, 2),
               "h2d_only_ms": round(h2d_only_ms, 2), "d2h_only_ms": round(d2h_only_ms, 2),
               "mode": args.encode_e2e + (f" ({args.piece_mb} MiB pieces, H2D / encode / D2H on three streams)" if args.encode_e2e == "pipelined" else "")}
    cpu = None
    if not args.skip_cpu and rank == 0 and world == 1:
        from oracle import oracle
        sample = _trim_utf8(text_dev[: min(n, args.cpu_sample_mb << 20)].cpu().numpy().tobytes())
        otok = oracle.Tokenizer(vocab, merges, SPECIALS)
        t0 = time.perf_counter()
        want = otok.encode(sample.decode("utf-8"))
        dt = time.perf_counter() - t0
        cpu = {"value": round(len(sample) / dt / 1e6, 3), "unit": UNIT, "cores": 1, "kind": "port",
               "sample": f"first {len(sample)} bytes of the same text, {len(want)} ids, {dt:.1f} s; C port of tokenizer.py (one core, as the reference)"}
    if rank == 0:
        print(json.dumps({
            "metric": ENCODE_METRIC, "value": round(value, 2), "unit": UNIT,
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(ms_step, 2), "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "u8/int32", "data": f"synthetic ({kind}-shaped, torch generator, seed {seed})",
            "config": {"workload": args.workload, "total_bytes": int(total_bytes), "bytes_per_gpu": n, "vocab_size": len(vocab), "merges": len(merges),
                       "special_tokens": SPECIALS, "ids": int(total_ids), "l2": "inputs (>= 125 MB per GPU) larger than the 126 MB L2"},
            "stage_ms": stage, "unique_words": timings[-1].get("unique_words"), "decode": decode,
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks.summary()}))
    if world > 1:
        dist.destroy_process_group()


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="tinystories-2g-v10k", choices=sorted(WORKLOADS))
    ap.add_argument("--cpu-sample-mb", type=int, default=48)
    ap.add_argument("--skip-cpu", action="store_true")
    ap.add_argument("--skip-e2e", action="store_true")
    ap.add_argument("--encode-mb", type=int, default=256)
    ap.add_argument("--encode-e2e", default="pipelined", choices=["pipelined", "serial"])
    ap.add_argument("--piece-mb", type=int, default=128)
    args = ap.parse_args()
    import faulthandler
    faulthandler.dump_traceback_later(int(os.environ.get("YABPE_BENCH_WATCHDOG_S", "600")), exit=True)   # never hang a GPU box
    if args.impl == "reference":
        run_reference(args)
        return
    if args.workload.startswith("gpt2-encode"):
        run_encode(args)
        return

    import numpy as np
    import torch
    import yabpe
    from synth_gpu import synth_corpus_device
    from yabpe import _ffi, engine
    from yabpe.distributed import train_device_sharded

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    kind, nbytes, vocab, seed = WORKLOADS[args.workload]
    text_dev, n = synth_corpus_device(torch, nbytes, kind, seed + rank, lex_seed=seed)     # one corpus: shared lexicon, own text
    torch.cuda.synchronize()

    cfg = yabpe.BBPETrainerConfig(vocab_size=vocab, min_frequency=1, max_workers=1, chunk_size_bytes=1 << 30,
                                  special_tokens=SPECIALS)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step(profile=False):
        tr = yabpe.BBPETrainer(cfg)
        tr.profile = profile
        if world > 1:
            model = train_device_sharded(tr, text_dev, n)
        else:
            model = tr.train_device(text_dev, n)
        return tr, model

    for _ in range(args.warmup):
        step()
    barrier()
    launches0 = _ffi.launch_count()
    timings = []
    with ClockSampler(local) as clocks:
        ev0 = torch.cuda.Event(enable_timing=True); ev1 = torch.cuda.Event(enable_timing=True)
        ev0.record()
        for _ in range(args.steps):
            tr, model = step(profile=True)
            timings.append(dict(tr.timing))
        ev1.record()
        barrier()
        ms_total = ev0.elapsed_time(ev1)
    launches = _ffi.launch_count() - launches0
    if world > 1:
        t = torch.tensor([ms_total], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_total = float(t.item())
    ms_step = ms_total / args.steps
    total_bytes = n * world
    value = total_bytes / (ms_step / 1e3) / 1e6
    stats = tr.last_stats

    # roofline of the dominant kernel: algorithmic bytes = corpus bytes read once per launch
    peak, peak_kind = measured_peak_gbs()
    tile_ms = float(np.mean([t["pretok_tiles_ms"] for t in timings if "pretok_tiles_ms" in t])) if timings and "pretok_tiles_ms" in timings[0] else None
    merge_ms = float(np.mean([t["merge_loop_ms"] for t in timings if "merge_loop_ms" in t])) if timings and "merge_loop_ms" in timings[0] else None
    roofline = None
    if tile_ms:
        achieved = n / (tile_ms / 1e3) / 1e9
        roofline = {"bound": "hbm", "kernel": "k_pretok_warp (+ k_pretok_count on the boundary chunks)", "achieved": round(achieved, 2), "peak": peak, "unit": "GB/s",
                    "frac": round(achieved / peak, 4), "traffic": int(n * NCU_TRAFFIC_RATIO[kind]), "peak_source": peak_kind,
                    "traffic_source": "dram bytes per algorithmic byte from the committed ncu --set full capture (profiles/), scaled to this launch",
                    "algorithmic_bytes_per_launch": n, "ms_per_launch": round(tile_ms, 3)}

    # e2e: pinned host bytes -> H2D -> train -> D2H results (every rank copies its own shard)
    e2e = None
    if not args.skip_e2e:
        host = torch.empty(n, dtype=torch.uint8).pin_memory()
        host.copy_(text_dev[:n])
        host_np = host.numpy()
        reps = max(1, min(args.steps, 3))

        def e2e_step():
            tr2 = yabpe.BBPETrainer(cfg)
            if world > 1:
                dev2, n2 = engine.to_device_text(torch, host, non_blocking=True)
                return train_device_sharded(tr2, dev2, n2)
            return tr2.train_from_buffers([host_np])

        e2e_step()                                # warm-up: the caching allocator gets its 2 GB text block and table blocks
        e2e_step()
        barrier()
        rep_ms = []
        t0 = time.perf_counter()
        for _ in range(reps):
            t1 = time.perf_counter()
            m2 = e2e_step()
            rep_ms.append(round((time.perf_counter() - t1) * 1e3, 2))     # train_from_buffers returns host objects: the step is complete
        barrier()
        dt = (time.perf_counter() - t0) / reps
        t0 = time.perf_counter()                  # the bare upload on this box, for reading the number above (PCIe differs between boxes)
        text_dev[:n].copy_(host)
        torch.cuda.synchronize()
        h2d_only_ms = (time.perf_counter() - t0) * 1e3
        if world > 1:
            tmax = torch.tensor([dt], device="cuda", dtype=torch.float64)
            dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
            dt = float(tmax.item())
        if rank == 0:
            d2h = sum(len(a) + len(b) for a, b in m2.merges) + sum(len(k) for k in m2.vocab)
            e2e = {"value": round(total_bytes / dt / 1e6, 2), "unit": UNIT, "h2d_bytes_per_step": int(total_bytes),
                   "d2h_bytes_per_step": int(d2h), "ms_per_step": round(dt * 1e3, 2), "ms_per_rep": rep_ms, "h2d_only_ms": round(h2d_only_ms, 2)}
            assert m2.merges == model.merges
        del host, host_np

    # secondary: encode MB/s with the trained model on a slice of the corpus (device-resident in, ids out)
    encode = None
    if args.encode_mb > 0 and rank == 0:
        tok = yabpe.BBPETokenizer(vocab=model.vocab, merges=model.merges, special_tokens=SPECIALS)
        en = min(n, args.encode_mb << 20)
        while en > 0 and (int(text_dev[en].item()) & 0xC0) == 0x80:
            en -= 1
        sl = text_dev[:((en + 15) // 16) * 16 + 64].clone()
        sl[en:].zero_()
        tok.encode_device(sl, en)
        torch.cuda.synchronize()
        a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
        a.record()
        ids, _ = tok.encode_device(sl, en)
        b.record(); torch.cuda.synchronize()
        encode = {"MBps": round(en / (a.elapsed_time(b) / 1e3) / 1e6, 2), "bytes": en, "ids": int(ids.numel())}

    cpu = None
    if not args.skip_cpu and rank == 0 and world == 1:
        sample = text_dev[: args.cpu_sample_mb << 20].cpu().numpy().tobytes()
        while sample and (sample[-1] & 0xC0) == 0x80:
            sample = sample[:-1]
        if sample and sample[-1] >= 0xC0:
            sample = sample[:-1]
        dt, nm = cpu_port_run(sample, vocab)
        cpu = {"value": round(len(sample) / dt / 1e6, 3), "unit": UNIT, "cores": 1, "kind": "port",
               "sample": f"first {len(sample)} bytes of the same corpus, vocab {vocab}, {nm} merges, {dt:.1f} s; "
                         f"C port of trainer.py (linear max() scan); host has {os.cpu_count()} cores, the reference uses 1"}

    if rank == 0:
        line = {
            "metric": METRIC, "value": round(value, 2), "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": round(ms_step, 2), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u8/int64", "data": f"synthetic ({kind}-shaped, torch generator, lexicon seed {seed}, text seed {seed}+rank)",
            "config": {"workload": args.workload, "corpus_bytes_per_gpu": n, "vocab_size": vocab, "special_tokens": SPECIALS,
                       "l2": "inputs (>= 256 MB) larger than the 126 MB L2", "n_pretokens": stats.n_pretokens,
                       "unique_words": stats.n_words, "merges": stats.n_merges},
            "train_wall_s": round(ms_step / 1e3, 4),
            "merges_per_s": round(stats.n_merges / (merge_ms / 1e3), 1) if merge_ms else None,
            "us_per_merge": round(1e3 * merge_ms / max(stats.n_merges, 1), 2) if merge_ms else None,
            "pretokenize_GBps": round(n / (tile_ms / 1e3) / 1e9, 2) if tile_ms else None,
            "stage_ms": {k: round(float(np.mean([t[k] for t in timings])), 3) for k in (timings[0] if timings else {}) if k.endswith("_ms")},
            "leader_cycles[argmax,ranges,claim+commit,rewrite,close,sum_act,sum_items,sum_words]": timings[-1].get("leader_cycles") if timings else None,
            "merge_loop": {"index_rebuilds": stats.index_rebuilds, "threshold_rebuilds": stats.threshold_rebuilds,
                           "pairs_created": stats.n_pairs, "leader_mode_merges": stats.leader_merges,
                           "grid_mode_merges": stats.grid_merges},
            "encode": encode,
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e,
            "gpu_launches": int(launches), "clocks": clocks.summary(),
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
