"""BASELINE.json configs[4] (C5): the reference's multi-speaker Glow-TTS training step
(train.py:91-162, ModelConfig(n_speakers>1, gin_channels=256), DistributedDataParallel as
__main__.py:268-271 wraps it) with its own `monotonic_align` and with this repository's module
swapped in -- step time for both, and the share of the step spent inside maximum_path.

    python profiles/c5_train_step.py [--steps 6] [--batch 32] [--t-text 200] [--t-mel 1000]
    python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 profiles/c5_train_step.py ...

Prints one JSON object (rank 0).  The reference package comes from oracle/_ref/refpkg.zip (packed by
oracle/build_ref.py in the build container); without it the script reports that and exits 0.
Synthetic LJSpeech-shaped batches (ragged lengths, one batch per rank: what DistributedSampler
gives each rank, __main__.py:235); random-init weights."""
from __future__ import annotations

import argparse
import importlib
import json
import os
import sys
import time
from pathlib import Path

import torch

REPO = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(REPO))
import __graft_entry__ as entry  # noqa: E402


class TimedModule:
    """A monotonic_align module whose maximum_path is bracketed by CUDA events."""

    def __init__(self, module):
        self.module, self.events = module, []

    def maximum_path(self, value, mask):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = self.module.maximum_path(value, mask)
        e1.record()
        self.events.append((e0, e1))
        return out

    def total_ms(self):
        torch.cuda.synchronize()
        return sum(a.elapsed_time(b) for a, b in self.events)


def run(steps=6, warmup=2, batch=32, t_text=200, t_mel=1000, n_speakers=4, gin_channels=256, mean_only=True):
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist  # noqa: WPS433

        if not dist.is_initialized():
            dist.init_process_group("nccl", device_id=dev)
    oracle = entry.load_oracle()
    rm = importlib.import_module(oracle.__name__ + ".ref_model")
    ref = rm.import_reference()
    if ref is None:
        return {"unavailable": "reference package not staged (oracle/_ref/refpkg.zip)"}
    pkg = entry.load_package()
    train = importlib.import_module(ref.__name__ + ".train")
    theirs = importlib.import_module(ref.__name__ + ".monotonic_align")
    batches = [tuple(t if t is None else t.cpu() for t in
                     rm.synthetic_batch(batch, t_text, t_mel, n_speakers=n_speakers, seed=1234 + 17 * rank + i, device="cpu"))
               for i in range(warmup + steps)]
    out = {}
    # "ours_no_sync": additionally the reference's train_step replaced by this package's restatement without
    # its host synchronisations (.item() per step, and per parameter tensor in clip_grad_value_: SURVEY 8f-3)
    for name, module, step_fn in (("reference", theirs, train.train_step), ("ours", pkg.monotonic_align, train.train_step),
                                  ("ours_no_sync", pkg.monotonic_align, pkg.train_step)):
        config, model, optimizer = rm.make_model(ref, mean_only=mean_only, n_speakers=n_speakers, gin_channels=gin_channels,
                                                 device=dev, seed=1234)
        if world > 1:
            model = torch.nn.parallel.DistributedDataParallel(model, device_ids=[local_rank], output_device=local_rank)   # __main__.py:268-271
        timed = TimedModule(module)
        prev = rm.swap_monotonic_align(ref, timed)
        try:
            step_fn(0, 0, model, optimizer, config, batches[:warmup], fp16_run=False)
            timed.events.clear()
            if dist is not None:
                dist.barrier()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            step_fn(warmup, 0, model, optimizer, config, batches[warmup:], fp16_run=False)
            torch.cuda.synchronize()
            if dist is not None:
                dist.barrier()
            wall = (time.perf_counter() - t0) / steps
            mas_ms = timed.total_ms() / steps
        finally:
            rm.swap_monotonic_align(ref, prev)
        t = torch.tensor([wall * 1e3, mas_ms], dtype=torch.float64, device=dev)
        if dist is not None:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        out[name] = {"step_ms": float(t[0]), "maximum_path_ms": float(t[1]), "path_share": float(t[1] / t[0])}
        del model, optimizer
        torch.cuda.empty_cache()
    cells = batch * (t_text) * (t_mel // 2 * 2)
    out["speedup_step"] = out["reference"]["step_ms"] / out["ours"]["step_ms"]
    out["speedup_step_no_sync"] = out["reference"]["step_ms"] / out["ours_no_sync"]["step_ms"]
    out["config"] = {"what": "reference train_step (train.py:91-162), fp32, multi-speaker Glow-TTS base", "n_gpus": world,
                     "per_gpu_batch": batch, "T_text": t_text, "T_mel": t_mel, "n_speakers": n_speakers, "gin_channels": gin_channels,
                     "mean_only": mean_only, "steps": steps, "ddp": world > 1, "lengths": "ragged", "cells_per_rank_step": cells}
    return out if rank == 0 else None


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--steps", type=int, default=6)
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--t-text", type=int, default=200)
    ap.add_argument("--t-mel", type=int, default=1000)
    a = ap.parse_args()
    res = run(steps=a.steps, batch=a.batch, t_text=a.t_text, t_mel=a.t_mel)
    if res is not None:
        print(json.dumps(res))
    if int(os.environ.get("WORLD_SIZE", "1")) > 1:
        import torch.distributed as dist

        dist.barrier()
        dist.destroy_process_group()
