#!/usr/bin/env python
"""bench.py -- MAS cells/s (logp + DP + backtrack) on B200, next to the reference's CPU path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload c2|c1|c3|c4]

One "step" = one pass of the hot path (models.py:362-382: log-likelihood matrix -> maximum_path
-> dense path + durations) over one synthetic LJSpeech-shaped batch.  The default workload is the
configuration BASELINE.json quotes the metric on for one GPU, configs[1]: fused logp+MAS for
Glow-TTS base (80 mel channels), B=32, T_text=200, T_mel=1000, fp32, full lengths.  With N GPUs
every rank processes its own batch of that shape (utterance sharding, no data-path collective):
weak scaling, value = cells of all ranks / max-over-ranks device time.

Printed JSON line (rank 0): the driver contract + ``roofline`` (dominant kernel, algorithmic bytes
/ live CUDA-event time, vs MEASURED_PEAKS.json), ``cpu_baseline`` (the reference's CPU path timed
on this box's host cores on a bounded sample), ``e2e`` (same metric through the public API from
pinned HOST buffers, H2D + D2H inside the timed region), ``clocks`` and ``gpu_launches``.

``--impl reference`` times the reference's own CPU implementation of the same step (its torch
logp program on CPU + monotonic_align.maximum_path around its compiled OpenMP Cython kernel,
oracle/_ref) and prints the same line with ``"impl": "reference"``.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time
from pathlib import Path

REPO = Path(__file__).resolve().parent
sys.path.insert(0, str(REPO))

import numpy as np  # noqa: E402
import torch  # noqa: E402

import __graft_entry__ as entry  # noqa: E402

WORKLOADS = {
    # name: (B per GPU, D, T_text, T_mel, description)
    "c1": (32, 80, 200, 1000, "C1 standalone maximum_path B=32 T_text=200 T_mel=1000 fp32"),
    "c2": (32, 80, 200, 1000, "C2 fused logp+MAS, Glow-TTS base (80 mels) B=32 T_text=200 T_mel=1000 fp32"),
    "c3": (256, 80, 400, 2000, "C3 batch sweep B=256 T_text=400 T_mel=2000 fp32 (per GPU)"),
    "c4": (8, 80, 1024, 8192, "C4 long-form B=8 T_text=1024 T_mel=8192 fp32"),
}
SEED = 1234  # the reference's config.seed (config.py:66)
L2_BYTES = 126 * 2**20
METRIC = "MAS cells/sec (logp+DP+backtrack)"
UNIT = "cells/s"


# ----------------------------------------------------------------------------------------------
# synthetic inputs
# ----------------------------------------------------------------------------------------------
def synth_inputs(B, D, T_x, T_y, seed, mean_only=False):
    """x_m ~ N(0,1), x_logs = 0.3 N(0,1) - 0.5 (general case; zeros when mean_only),
    z = x_m[:, :, y*T_x/T_y] + exp(x_logs) N(0,1) ("trained-like"), full lengths."""
    g = torch.Generator().manual_seed(seed)
    x_m = torch.randn(B, D, T_x, generator=g)
    x_logs = torch.zeros(B, D, T_x) if mean_only else 0.3 * torch.randn(B, D, T_x, generator=g) - 0.5
    idx = (torch.arange(T_y) * T_x) // T_y
    z = x_m[:, :, idx] + torch.exp(x_logs[:, :, idx]) * torch.randn(B, D, T_y, generator=g)
    x_len = torch.full((B,), T_x, dtype=torch.int32)
    y_len = torch.full((B,), T_y, dtype=torch.int32)
    return x_m, x_logs, z, x_len, y_len


def measured_peaks():
    p = REPO / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""

    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
             "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu_index = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits", "-lms", "100",
                 "-i", str(self.gpu_index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.lines:
            f = [s.strip() for s in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                smax.append(float(f[2]))
            except ValueError:
                continue
            for name, val in zip(names, f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None,
                "sm_max_mhz": max(smax) if smax else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# ----------------------------------------------------------------------------------------------
# the reference arm (CPU)
# ----------------------------------------------------------------------------------------------
def reference_setup(B, D, T_x, T_y, mean_only=False):
    oracle = entry.load_oracle()
    core = oracle.reference_core("omp")
    kind = "reference"
    if core is None:  # oracle/_ref never built: fall back to the C port of the same kernel
        threads = oracle.host_threads()
        kernel = lambda p, v, tx, ty: oracle.maximum_path_c(p, v, tx, ty, threads=threads)  # noqa: E731
        kind = "port"
    else:
        kernel = core.maximum_path_c
    cores = oracle.host_threads()
    os.environ.setdefault("OMP_NUM_THREADS", str(cores))
    torch.set_num_threads(cores)
    x_m, x_logs, z, x_len, y_len = synth_inputs(B, D, T_x, T_y, SEED + 1, mean_only)   # zeros when mean_only, as models.py:139 does
    x_mask = (torch.arange(T_x)[None] < x_len[:, None]).float()
    z_mask = (torch.arange(T_y)[None] < y_len[:, None]).float()
    attn_mask = x_mask[:, :, None] * z_mask[:, None, :]              # models.py:337 (squeezed)

    def step():
        return oracle.reference_step(x_m, x_logs, z, attn_mask, kernel=kernel)

    return step, kind, cores


def time_reference(step, steps, warmup):
    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    return (time.perf_counter() - t0) / steps


def run_reference(args):
    B, D, T_x, T_y, desc = WORKLOADS[args.workload]
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    step, kind, cores = reference_setup(B, D, T_x, T_y, args.mean_only)
    sec = time_reference(step, args.steps, args.warmup)
    cells = B * T_x * T_y
    value = cells / sec
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": desc, "lengths": "full", "mean_only": args.mean_only, "host": "CPU only: torch logp program (models.py:363-376) "
                   "+ monotonic_align.maximum_path around the reference's OpenMP Cython kernel"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind,
                         "sample": f"{args.steps} full batches of the workload (B={B})"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


# ----------------------------------------------------------------------------------------------
# our arm (GPU)
# ----------------------------------------------------------------------------------------------
def run_ours(args):
    B, D, T_x, T_y, desc = WORKLOADS[args.workload]
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    assert torch.cuda.is_available(), "bench.py (ours) needs a CUDA device; there is no CPU fallback"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist  # noqa: WPS433

        dist.init_process_group("nccl", device_id=dev)
    pkg = entry.load_package()
    assert pkg._lib.load().mas_b200_device_ok() == 0, "device is not sm_100 (B200)"
    standalone = args.workload == "c1"

    mean_only = args.mean_only          # x_logs == 0 (config.py:52 default): passed as None, not copied

    def drop_logs(t):
        return (t[0], None, t[2], t[3], t[4]) if mean_only else t

    cells = B * T_x * T_y
    in_bytes = 4 * B * D * (T_y + (1 if mean_only else 2) * T_x) + 8 * B
    out_bytes = 4 * cells + 4 * B * T_x
    # ---- resident inputs, rotated so that consecutive steps never hit L2 ----
    per_set = in_bytes + out_bytes + pkg._lib.load().mas_b200_fused_workspace_bytes(B, D, T_x, T_y)
    n_sets = max(3, int(2.5 * L2_BYTES // per_set) + 1)
    sets = []
    for i in range(n_sets):
        x_m, x_logs, z, x_len, y_len = synth_inputs(B, D, T_x, T_y, SEED + 1 + rank * 1000 + i, mean_only)
        sets.append(drop_logs(tuple(t.to(dev) for t in (x_m, x_logs, z, x_len, y_len))))
    logp_sets = None
    if standalone:
        logp_sets = [(pkg.log_likelihood_matrix(s[0], s[1], s[2]), s[3], s[4]) for s in sets]

    def step(i):
        s = sets[i % n_sets]
        if standalone:
            lp, tx, ty = logp_sets[i % n_sets]
            return pkg.maximum_path_from_lengths(lp, tx, ty, want_durations=True)
        return pkg.fused_maximum_path(s[0], s[1], s[2], s[3], s[4])

    launches_per_step = 1                        # kernel (1) alone | the single fused launch (plus a 4 KB flag memset)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def timed_graph(fn, steps, first):
        """K back-to-back steps captured in ONE CUDA graph (the path is launch-bound from Python
        otherwise: a step is tens of microseconds), replayed between two events on the launching stream."""
        graph = torch.cuda.CUDAGraph()
        keep = []
        with torch.cuda.graph(graph):
            for i in range(steps):
                keep.append(fn(first + i))
        graph.replay()                                   # untimed: instantiate / upload
        torch.cuda.synchronize()
        return graph, keep

    for i in range(args.warmup):
        step(i)
    barrier()
    graph, keep = timed_graph(step, args.steps, args.warmup)
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    graph.replay()
    ev1.record()
    barrier()
    dev_ms = ev0.elapsed_time(ev1)
    # nvidia-smi samples every 100 ms and the timed replay lasts a few: keep the SAME load running for
    # ~0.6 s more so that the clocks / throttle reasons reported are the ones this load runs at
    extra = min(5000, int(600.0 / max(dev_ms, 0.05)) + 1)
    for _ in range(extra):
        graph.replay()
    torch.cuda.synchronize()
    clocks = sampler.stop()
    clocks["sampled"] = f"the timed replay and {extra} identical replays right after it (nvidia-smi -lms 100)"
    del graph, keep

    # ---- dominant kernel alone: kernel (1) on materialised scores, rotated buffers ----
    if logp_sets is None:
        logp_sets = [(pkg.log_likelihood_matrix(s[0], s[1], s[2]), s[3], s[4]) for s in sets]

    def k1_step(i):
        lp, tx, ty = logp_sets[i % len(logp_sets)]
        return pkg.maximum_path_from_lengths(lp, tx, ty, want_durations=True)

    for i in range(3):
        k1_step(i)
    torch.cuda.synchronize()
    graph, keep = timed_graph(k1_step, args.steps, 0)
    k0, k1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    k0.record()
    graph.replay()
    k1.record()
    torch.cuda.synchronize()
    kern_ms = k0.elapsed_time(k1) / args.steps
    del graph, keep

    # ---- end to end through the public API from pinned host buffers ----
    # Every step copies ITS inputs host->device from pinned memory, runs the public call and copies
    # the dense path + durations device->host; the caller then reads them.  Steps alternate between
    # two streams (double-buffered pinned results), the way a host loop that feeds the GPU would be
    # written: step i's D2H overlaps step i+1's H2D and kernels (separate copy engines).
    n_lanes = 3
    host = [drop_logs(tuple(t.pin_memory() for t in synth_inputs(B, D, T_x, T_y, SEED + 77 + rank * 1000 + i, mean_only)))
            for i in range(n_lanes)]
    host_out = [torch.empty((B, T_x, T_y), dtype=torch.float32).pin_memory() for _ in range(n_lanes)]
    host_dur = [torch.empty((B, T_x), dtype=torch.int32).pin_memory() for _ in range(n_lanes)]
    lanes = [torch.cuda.Stream(device=dev) for _ in range(n_lanes)]
    done_ev = [torch.cuda.Event() for _ in range(n_lanes)]
    checksum = [0]

    def e2e_run(nsteps, start_ev, dense=True):
        for st in lanes:
            st.wait_event(start_ev)
        for i in range(nsteps):
            k = i % n_lanes
            if i >= n_lanes:
                done_ev[k].synchronize()                 # the caller reads step i-3's result before its buffers are reused
                checksum[0] += int(host_dur[k][0, 0])
            with torch.cuda.stream(lanes[k]):
                x_m, x_logs, z, x_len, y_len = host[k]
                d = [None if t is None else t.to(dev, non_blocking=True) for t in (x_m, x_logs, z, x_len, y_len)]
                path, dur = pkg.fused_maximum_path(*d)
                if dense:
                    host_out[k].copy_(path, non_blocking=True)
                host_dur[k].copy_(dur, non_blocking=True)
                done_ev[k].record(lanes[k])
        for k in range(n_lanes):
            done_ev[k].synchronize()
            checksum[0] += int(host_dur[k][0, 0])
            torch.cuda.current_stream().wait_stream(lanes[k])

    warm_ev = torch.cuda.Event()
    warm_ev.record()
    e2e_run(max(4, args.warmup), warm_ev)
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e2e_steps = max(8, args.steps // 2)
    e0.record()
    e2e_run(e2e_steps, e0)
    e1.record()
    barrier()
    e2e_ms = e0.elapsed_time(e1)
    # the same loop when only the compact result leaves the device (integer durations: everything the
    # dense path says; the path itself stays on the GPU for its consumers, as in the training step)
    c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    c0.record()
    e2e_run(e2e_steps, c0, dense=False)
    c1.record()
    barrier()
    e2e_compact_ms = c0.elapsed_time(c1)

    times = torch.tensor([dev_ms, e2e_ms, e2e_compact_ms], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
    dev_ms, e2e_ms, e2e_compact_ms = times.tolist()
    value = world * cells * args.steps / (dev_ms * 1e-3)
    e2e_value = world * cells * e2e_steps / (e2e_ms * 1e-3)

    if rank == 0:
        peak, peak_src = measured_peaks()
        traffic_all = {}
        tfile = REPO / "profiles" / "traffic.json"
        if tfile.exists():
            traffic_all = json.loads(tfile.read_text())
        # kernel (1) alone: read fp32 scores + write fp32 path = 8 B/cell
        k1_bytes = 8 * cells
        k1 = {"kernel": "mas_path_systolic (kernel 1: DP + backtrack + dense path)", "bound": "hbm",
              "achieved": k1_bytes / (kern_ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
              "frac": k1_bytes / (kern_ms * 1e-3) / 1e9 / peak, "traffic": traffic_all.get("c1"),
              "algorithmic_bytes_per_launch": k1_bytes, "kernel_ms": kern_ms}
        if standalone:
            roofline = dict(k1, peak_source=peak_src)
        else:
            # the step IS one launch of the fused kernel: inputs (z, x_m, x_logs) + dense path + durations
            step_ms = dev_ms / args.steps
            f_bytes = in_bytes - 8 * B + out_bytes
            flops = (160 if mean_only else 320) * cells
            fp32_peak = 148 * 128 * 2 * 1.965e9 / 1e12            # CUDA-core FMA peak, derived
            roofline = {"kernel": "mas_fused (kernel 2: FFMA producers + sweep CTAs, one launch)", "bound": "hbm",
                        "achieved": f_bytes / (step_ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                        "frac": f_bytes / (step_ms * 1e-3) / 1e9 / peak, "traffic": traffic_all.get("fused_c2"),
                        "algorithmic_bytes_per_launch": f_bytes, "kernel_ms": step_ms, "peak_source": peak_src,
                        "note": "bounded by the FP32 FMA pipe and the sweep's dependent chain, not by HBM (DESIGN.md 4-5)",
                        "fp32_fma": {"achieved_tflops": flops / (step_ms * 1e-3) / 1e12, "peak_tflops_derived": fp32_peak,
                                     "frac": flops / (step_ms * 1e-3) / 1e12 / fp32_peak, "flops_per_launch": flops},
                        # what the unfused pipeline must move for the same work: write logp, read logp, write the
                        # path (12 B/cell) + the inputs.  The compulsory bytes alone cannot reach 50 % of
                        # the HBM roof: 2x their transfer time leaves room for 1/(2 x 6.1 us) x 2.05 GFLOP = 168 TFLOP/s,
                        # 2.3x the CUDA-core FP32 peak.
                        "hbm_equivalent": {"bytes_per_launch": 12 * cells + (in_bytes - 8 * B),
                                           "achieved": (12 * cells + in_bytes - 8 * B) / (step_ms * 1e-3) / 1e9,
                                           "frac": (12 * cells + in_bytes - 8 * B) / (step_ms * 1e-3) / 1e9 / peak},
                        "kernel1_alone": k1}
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": dev_ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": desc, "per_gpu_batch": B, "global_batch": B * world, "lengths": "full",
                       "channels": D, "mean_only": mean_only, "parallelism": f"utterance-sharded x{world}, no collective",
                       "l2": f"inputs/outputs rotated over {n_sets} buffer sets ({n_sets * per_set / 2**20:.0f} MiB > L2)",
                       "kernels_per_step": launches_per_step,
                       "launch": f"{args.steps} steps captured in one CUDA graph, one replay timed"},
            "roofline": roofline,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": in_bytes, "d2h_bytes_per_step": out_bytes,
                    "steps": e2e_steps, "ms_per_step": e2e_ms / e2e_steps,
                    "pipeline": "3 streams, triple-buffered pinned results; each step: H2D inputs, fused call, D2H dense path + durations",
                    "durations_only": {"value": world * cells * e2e_steps / (e2e_compact_ms * 1e-3), "unit": UNIT,
                                       "d2h_bytes_per_step": B * T_x * 4, "ms_per_step": e2e_compact_ms / e2e_steps,
                                       "what": "same loop, the dense path stays on the device (its consumers run there); "
                                               "only the integer durations are read back"}},
            "clocks": clocks,
            "gpu_launches": launches_per_step * args.steps,
        }
        if world == 1 and not args.no_cpu_baseline:
            ref_step, kind, cores = reference_setup(B, D, T_x, T_y, mean_only)
            ref_step()
            reps, t0 = 0, time.perf_counter()
            while reps < 3 or (time.perf_counter() - t0 < args.cpu_seconds and reps < 400):
                ref_step()
                reps += 1
            sec = (time.perf_counter() - t0) / reps
            line["cpu_baseline"] = {"value": cells / sec, "unit": UNIT, "cores": cores, "kind": kind,
                                    "sample": f"{reps} full batches (B={B}) of the same workload on the host, "
                                              f"{sec * 1e3:.1f} ms each"}
            # the reference as it runs in training -- logp by torch ON THE GPU, then its
            # own maximum_path on the CUDA tensors (device sync, 2 D2H, Cython OpenMP, 1 H2D)
            try:
                oracle = entry.load_oracle()
                core = oracle.reference_core("omp")
                kern = core.maximum_path_c if core is not None else None
                xg = [t.to(dev) for t in synth_inputs(B, D, T_x, T_y, SEED + 5, mean_only)]
                xmask = (torch.arange(T_x, device=dev)[None] < xg[3][:, None]).float()
                zmask = (torch.arange(T_y, device=dev)[None] < xg[4][:, None]).float()
                amask = xmask[:, :, None] * zmask[:, None, :]
                for _ in range(2):
                    oracle.reference_step(xg[0], xg[1], xg[2], amask, kernel=kern)
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                nrep = 10
                for _ in range(nrep):
                    oracle.reference_step(xg[0], xg[1], xg[2], amask, kernel=kern)
                torch.cuda.synchronize()
                sec2 = (time.perf_counter() - t0) / nrep
                line["cpu_baseline"]["reference_in_training"] = {
                    "value": cells / sec2, "unit": UNIT, "ms_per_step": sec2 * 1e3,
                    "what": "models.py:362-382 as the reference runs it on this box: torch logp on the GPU + "
                            "monotonic_align.maximum_path (sync, 2 D2H, OpenMP Cython, 1 H2D)"}
            except Exception as exc:  # noqa: BLE001
                line["cpu_baseline"]["reference_in_training"] = {"error": str(exc)[:200]}
        print(json.dumps(line))
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--workload", choices=sorted(WORKLOADS), default="c2")
    ap.add_argument("--cpu-seconds", type=float, default=10.0, help="CPU baseline sample budget")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--mean-only", action="store_true",
                    help="x_logs == 0 (the reference's default ModelConfig.mean_only): the contraction halves")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
