#!/usr/bin/env python
"""bench.py -- MAS cells/s (logp + DP + backtrack) on B200, next to the reference's CPU path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload c2|c1|c3|c4]

One "step" = one pass of the hot path (models.py:362-382: log-likelihood matrix -> maximum_path
-> dense path + durations) over one synthetic LJSpeech-shaped batch.  The default workload is the
configuration BASELINE.json quotes the metric on for one GPU, configs[1]: fused logp+MAS for
Glow-TTS base (80 mel channels), B=32, T_text=200, T_mel=1000, fp32, full lengths.  With N GPUs
every rank processes its own batch of that shape (utterance sharding, no data-path collective):
weak scaling, value = cells of all ranks / max-over-ranks device time.

Printed JSON line (rank 0): the driver contract + ``roofline`` (dominant kernel; SURVEY 8d:
max(algorithmic bytes / t / measured HBM peak, flops / t / derived FP32 peak), both components kept),
``cpu_baseline`` (the reference's CPU path timed on this box's host cores on a bounded sample, plus
``kernel_only``: its Cython maximum_path alone, serial as built and OpenMP), ``e2e`` (same metric
through the public API from pinned HOST buffers, H2D + D2H inside the timed region), ``clocks``,
``gpu_launches`` and three blocks for the other configurations BASELINE.json lists:
``configs`` (c3, c4: fused entry and kernel (1) alone, N = 1), ``c3_strong`` (B=256, 400 x 2000 SPLIT
over the N GPUs with the package's sharding helpers: full lengths / ragged lengths, balanced and
contiguous) and ``c5`` (the reference's multi-speaker training step with its own maximum_path and
with the module swapped, DistributedDataParallel for N > 1: profiles/c5_train_step.py).

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
# shared by both arms
# ----------------------------------------------------------------------------------------------
def config_for(workload, world, mean_only, lengths="full"):
    """The workload's description: the SAME keys and values for both arms (the driver compares them)."""
    B, D, T_x, T_y, desc = WORKLOADS[workload]
    return {"workload": desc, "per_gpu_batch": B, "global_batch": B * world, "T_text": T_x, "T_mel": T_y, "channels": D,
            "lengths": lengths, "mean_only": bool(mean_only), "parallelism": f"utterance-sharded x{world}, no collective"}


def set_host_threads(cores):
    """All host cores for the reference's OpenMP kernel and torch's CPU ops.  torchrun exports
    OMP_NUM_THREADS=1 and libgomp has read it by now (torch is imported): set the team size through
    the OpenMP API as well."""
    os.environ["OMP_NUM_THREADS"] = str(cores)
    torch.set_num_threads(cores)
    try:
        import ctypes

        ctypes.CDLL("libgomp.so.1").omp_set_num_threads(int(cores))
    except OSError:
        pass


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
    set_host_threads(cores)
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


def cython_kernel_only(shapes, budget_s=12.0):
    """BASELINE.md 3(i) / north_star: the reference's Cython `maximum_path_c` ALONE (core.pyx:40-45) on
    prepared C-contiguous arrays, a fresh `values` copy per repetition (the kernel clobbers it), in the
    flavour the reference's setup.py builds (serial: no -fopenmp, monotonic_align/setup.py:9-13) and
    with OpenMP on all host cores.  Best and median of up to 15 repetitions after 3 warm-ups, bounded
    by `budget_s` seconds in total."""
    oracle = entry.load_oracle()
    cores = oracle.host_threads()
    set_host_threads(cores)
    out = {"cores": cores, "what": "reference Cython maximum_path_c alone (core.pyx:40-45), values = 10 N(0,1) - 100, full lengths"}
    t_start = time.perf_counter()
    for name, (B, T_x, T_y) in shapes.items():
        rng = np.random.default_rng(SEED)
        values = (10 * rng.standard_normal((B, T_x, T_y)) - 100).astype(np.float32)
        t_xs, t_ys = np.full(B, T_x, np.int32), np.full(B, T_y, np.int32)
        paths = np.zeros((B, T_x, T_y), np.int32)
        row = {"B": B, "T_text": T_x, "T_mel": T_y}
        for flavour in ("omp", "serial"):
            core = oracle.reference_core(flavour)
            if core is None:
                row[flavour] = {"unavailable": "oracle/_ref not built"}
                continue
            ts = []
            for rep in range(18):
                v = values.copy()
                paths[:] = 0
                t0 = time.perf_counter()
                core.maximum_path_c(paths, v, t_xs, t_ys)
                dt = time.perf_counter() - t0
                if rep >= 3:
                    ts.append(dt)
                if len(ts) >= 3 and time.perf_counter() - t_start > budget_s * (list(shapes).index(name) + (1 if flavour == "serial" else 0.5)) / len(shapes):
                    break
            cells = B * T_x * T_y
            row[flavour] = {"best_ms": min(ts) * 1e3, "median_ms": statistics.median(ts) * 1e3, "reps": len(ts),
                            "cells_per_s": cells / min(ts), "threads": cores if flavour == "omp" else 1}
        out[name] = row
    return out


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
        "config": config_for(args.workload, args.gpus, args.mean_only),
        "method": "CPU only, rank 0: the reference's torch logp program (models.py:363-376) + monotonic_align.maximum_path "
                  "around its own compiled OpenMP Cython kernel (oracle/_ref), all host threads",
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
class Timer:
    """Steps captured in ONE CUDA graph (a step is tens of microseconds: launched from Python the
    launches would be the thing measured), replayed several times, each replay between two events
    on the launching stream and bracketed by a barrier + synchronize; median and best replay."""

    def __init__(self, dist, dev):
        self.dist, self.dev = dist, dev

    def barrier(self):
        if self.dist is not None:
            self.dist.barrier()
        torch.cuda.synchronize()

    def run(self, fn, steps, replays, first=0):
        graph = torch.cuda.CUDAGraph()
        keep = []
        with torch.cuda.graph(graph):
            for i in range(steps):
                keep.append(fn(first + i))
        graph.replay()                                   # untimed: instantiate / upload
        times = []
        for _ in range(replays):
            self.barrier()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            graph.replay()
            e1.record()
            self.barrier()
            times.append(e0.elapsed_time(e1))
        t = torch.tensor(times, dtype=torch.float64, device=self.dev)
        if self.dist is not None:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)       # per replay: the slowest rank
        times = t.tolist()
        return graph, keep, statistics.median(times), min(times), times


def numa_local_affinity(dev):
    """Pin this process to the CPUs of the NUMA node the GPU hangs off (sysfs), so that the pinned host
    buffers of the end-to-end loop are allocated there.  Returns what was found; "restore" = the affinity
    to put back afterwards (the CPU baseline uses every core)."""
    info = {"numa_node": None, "cpus": None}
    try:
        props = torch.cuda.get_device_properties(dev)
        bdf = f"{props.pci_domain_id:04x}:{props.pci_bus_id:02x}:{props.pci_device_id:02x}.0"
        node = int((Path("/sys/bus/pci/devices") / bdf / "numa_node").read_text())
        info["pci"] = bdf
        if node < 0:
            return info
        cpus = set()
        for part in (Path(f"/sys/devices/system/node/node{node}/cpulist")).read_text().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        allowed = os.sched_getaffinity(0)
        use = cpus & allowed
        info["numa_node"] = node
        if use:
            info["restore"] = allowed
            os.sched_setaffinity(0, use)
            info["cpus"] = len(use)
    except (OSError, ValueError, AttributeError):
        pass
    return info


def fp32_peak_tflops(dev, clocks):
    """CUDA-core FMA peak, derived: SMs x 128 lanes x 2 flop x the SM clock's maximum."""
    props = torch.cuda.get_device_properties(dev)
    mhz = clocks.get("sm_max_mhz") or getattr(props, "clock_rate", 1965000) / 1e3
    return props.multi_processor_count * 128 * 2 * mhz * 1e6 / 1e12, props.multi_processor_count, mhz


def fused_roofline(cells, in_bytes, out_bytes, B, mean_only, step_ms, peak, fp32_peak):
    """achieved_fused = max(bytes / t / HBM peak, flops / t / FP32 peak); both components kept."""
    f_bytes = in_bytes - 8 * B + out_bytes
    flops = (160 if mean_only else 320) * cells
    hbm = f_bytes / (step_ms * 1e-3) / 1e9
    tfl = flops / (step_ms * 1e-3) / 1e12
    by_flops = tfl / fp32_peak > hbm / peak
    return {"bound": "fp32_fma" if by_flops else "hbm",
            "achieved": tfl if by_flops else hbm, "peak": fp32_peak if by_flops else peak, "unit": "TFLOP/s" if by_flops else "GB/s",
            "frac": max(tfl / fp32_peak, hbm / peak),
            "hbm": {"achieved": hbm, "peak": peak, "unit": "GB/s", "frac": hbm / peak, "algorithmic_bytes_per_launch": f_bytes},
            "fp32_fma": {"achieved": tfl, "peak": fp32_peak, "unit": "TFLOP/s", "frac": tfl / fp32_peak, "flops_per_launch": flops},
            # what the unfused pipeline must move for the same work: write logp, read logp, write the path
            # (12 B/cell) + the inputs
            "hbm_equivalent": {"bytes_per_launch": 12 * cells + (in_bytes - 8 * B),
                               "achieved": (12 * cells + in_bytes - 8 * B) / (step_ms * 1e-3) / 1e9,
                               "frac": (12 * cells + in_bytes - 8 * B) / (step_ms * 1e-3) / 1e9 / peak},
            "kernel_ms": step_ms}


def ragged_lengths_torch(B, T_x, T_y, seed):
    """SURVEY 8d `ragged` mode: t_x ~ U[T_x/2, T_x], t_y ~ t_x T_y/T_x U[0.8, 1.2], even, sorted by t_x
    descending (dataset.py:79-81), element 0 full size."""
    g = torch.Generator().manual_seed(seed)
    tx = torch.randint(T_x // 2, T_x + 1, (B,), generator=g)
    tx[0] = T_x
    tx, _ = torch.sort(tx, descending=True)
    ty = (tx.float() * (T_y / T_x) * (0.8 + 0.4 * torch.rand(B, generator=g))).round().long()
    ty = torch.minimum(torch.maximum(ty, tx), torch.tensor(T_y)) // 2 * 2
    ty = torch.maximum(ty, tx)
    ty[0] = T_y
    return tx.to(torch.int32), ty.to(torch.int32)


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
    lib = pkg._lib.load()
    assert lib.mas_b200_device_ok() == 0, "device is not sm_100 (B200)"
    timer = Timer(dist, dev)
    standalone = args.workload == "c1"
    mean_only = args.mean_only          # x_logs == 0 (config.py:52 default): passed as None, not copied
    replays = 5

    def drop_logs(t):
        return (t[0], None, t[2], t[3], t[4]) if mean_only else t

    cells = B * T_x * T_y
    in_bytes = 4 * B * D * (T_y + (1 if mean_only else 2) * T_x) + 8 * B
    out_bytes = 4 * cells + 4 * B * T_x
    # ---- resident inputs, rotated so that consecutive steps never hit L2 ----
    per_set = in_bytes + out_bytes
    n_sets = max(3, int(2.5 * L2_BYTES // per_set) + 1)
    sets = []
    for i in range(n_sets):
        x_m, x_logs, z, x_len, y_len = synth_inputs(B, D, T_x, T_y, SEED + 1 + rank * 1000 + i, mean_only)
        sets.append(drop_logs(tuple(t.to(dev) for t in (x_m, x_logs, z, x_len, y_len))))
    logp_sets = [(pkg.log_likelihood_matrix(s[0], s[1], s[2]), s[3], s[4]) for s in sets]

    def k1_step(i):
        lp, tx, ty = logp_sets[i % n_sets]
        return pkg.maximum_path_from_lengths(lp, tx, ty, want_durations=True)

    def step(i):
        if standalone:
            return k1_step(i)
        s = sets[i % n_sets]
        return pkg.fused_maximum_path(s[0], s[1], s[2], s[3], s[4])

    # how the fused entry runs this shape: its single launch, or (by its cost estimate) the two kernels
    launches_per_step = 1
    if not standalone:
        geom = np.zeros(12, np.int32)
        props = torch.cuda.get_device_properties(dev)
        lib.mas_b200_debug_fused_geom(B, D, T_x, T_y, 232448 - 1024, props.multi_processor_count, geom.ctypes.data)

    for i in range(args.warmup):
        step(i)
    timer.barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    graph, keep, dev_ms, best_ms, all_ms = timer.run(step, args.steps, replays, first=args.warmup)
    # nvidia-smi samples every 100 ms and a replay lasts a few: keep the SAME load running for ~0.6 s
    # more so that the clocks / throttle reasons reported are the ones this load runs at
    extra = min(5000, int(600.0 / max(dev_ms, 0.05)) + 1)
    for _ in range(extra):
        graph.replay()
    torch.cuda.synchronize()
    clocks = sampler.stop()
    clocks["sampled"] = f"the {replays} timed replays and {extra} identical replays right after them (nvidia-smi -lms 100)"
    del graph, keep

    # ---- kernel (1) alone on materialised scores (the C1 workload), same method ----
    for i in range(3):
        k1_step(i)
    graph, keep, k1_ms, k1_best, _ = timer.run(k1_step, args.steps, replays)
    kern_ms = k1_ms / args.steps
    del graph, keep

    # ---- end to end through the public API from pinned host buffers ----
    # Every step copies ITS inputs host->device from pinned memory, runs the public call and copies
    # the dense path + durations device->host; the caller then reads them.  Steps rotate over three
    # streams (triple-buffered pinned results), the way a host loop that feeds the GPU would be
    # written: step i's D2H overlaps step i+1's H2D and kernels (separate copy engines).
    n_lanes = 3
    # pinned buffers first-touched from the CPUs of the GPU's own NUMA node (N > 1: eight ranks' copies
    # otherwise cross the socket interconnect on their way to the PCIe root)
    host_info = numa_local_affinity(dev)
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
    timer.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e2e_steps = max(8, args.steps // 2)
    e0.record()
    e2e_run(e2e_steps, e0)
    e1.record()
    timer.barrier()
    e2e_ms = e0.elapsed_time(e1)
    # the same loop when only the compact result leaves the device (integer durations: everything the
    # dense path says; the path itself stays on the GPU for its consumers, as in the training step)
    c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    c0.record()
    e2e_run(e2e_steps, c0, dense=False)
    c1.record()
    timer.barrier()
    e2e_compact_ms = c0.elapsed_time(c1)
    times = torch.tensor([e2e_ms, e2e_compact_ms], dtype=torch.float64, device=dev)
    if dist is not None:
        dist.all_reduce(times, op=dist.ReduceOp.MAX)
    e2e_ms, e2e_compact_ms = times.tolist()
    del host_out, host_dur, host
    if host_info.get("restore") is not None:
        os.sched_setaffinity(0, host_info.pop("restore"))
    value = world * cells * args.steps / (dev_ms * 1e-3)
    e2e_value = world * cells * e2e_steps / (e2e_ms * 1e-3)

    peak, peak_src = measured_peaks()
    fp32_peak, n_sms, sm_mhz = fp32_peak_tflops(dev, clocks)
    del sets, logp_sets
    torch.cuda.empty_cache()

    # ---- the other configurations BASELINE.json lists (bounded: a few replays each) ----
    def measure_shape(Bc, Tx, Ty, x_len=None, y_len=None, steps=3, reps=3, seed=SEED + 500, want_k1=True):
        """The fused entry (general x_logs) and kernel (1) alone on a batch of this shape held by THIS rank."""
        g = torch.Generator().manual_seed(seed + rank)
        x_m = torch.randn(Bc, D, Tx, generator=g).to(dev)
        x_logs = (0.3 * torch.randn(Bc, D, Tx, generator=g) - 0.5).to(dev)
        z = torch.randn(Bc, D, Ty, generator=g).to(dev)
        xl = (torch.full((Bc,), Tx, dtype=torch.int32) if x_len is None else x_len).to(dev)
        yl = (torch.full((Bc,), Ty, dtype=torch.int32) if y_len is None else y_len).to(dev)
        res = {}
        assert Bc > 0, "more ranks than utterances"
        f = lambda i: pkg.fused_maximum_path(x_m, x_logs, z, xl, yl)   # noqa: E731
        f(0)
        gr, kp, med, best, _ = timer.run(f, steps, reps)
        res["fused_ms"], res["fused_best_ms"] = med / steps, best / steps
        del gr, kp
        if want_k1:
            lp = pkg.log_likelihood_matrix(x_m, x_logs, z)
            k = lambda i: pkg.maximum_path_from_lengths(lp, xl, yl, want_durations=True)   # noqa: E731
            k(0)
            gr, kp, med, best, _ = timer.run(k, steps, reps)
            res["k1_ms"], res["k1_best_ms"] = med / steps, best / steps
            del gr, kp, lp
        torch.cuda.empty_cache()
        return res

    def sub_line(name, Bc, Tx, Ty):
        r = measure_shape(Bc, Tx, Ty)
        c = Bc * Tx * Ty
        ib = 4 * Bc * D * (Ty + 2 * Tx) + 8 * Bc
        ob = 4 * c + 4 * Bc * Tx
        return {"workload": WORKLOADS[name][4], "cells": c,
                "fused": {"ms_per_step": r["fused_ms"], "best_ms": r["fused_best_ms"], "cells_per_s": c / (r["fused_ms"] * 1e-3),
                          "roofline": fused_roofline(c, ib, ob, Bc, False, r["fused_ms"], peak, fp32_peak)},
                "kernel1": {"ms_per_step": r["k1_ms"], "best_ms": r["k1_best_ms"], "cells_per_s": c / (r["k1_ms"] * 1e-3),
                            "roofline": {"bound": "hbm", "achieved": 8 * c / (r["k1_ms"] * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
                                         "frac": 8 * c / (r["k1_ms"] * 1e-3) / 1e9 / peak, "algorithmic_bytes_per_launch": 8 * c}}}

    configs = None
    if world == 1 and not args.no_configs:
        configs = {"method": "fused entry (general x_logs) and kernel (1) alone on materialised scores; 3 steps per CUDA graph, "
                             "median of 3 replays; every step writes its own dense path (> L2 for c3/c4)",
                   "c3": sub_line("c3", 256, 400, 2000), "c4": sub_line("c4", 8, 1024, 8192)}

    # ---- C3 as BASELINE.json words it: B=256, 400 x 2000 SPLIT over the N GPUs (strong scaling) ----
    c3_strong = None
    if not args.no_configs:
        from glow_tts_train_b200 import sharding  # noqa: WPS433  (the package's own sharding helpers)

        Bg, Tx3, Ty3 = 256, 400, 2000
        mine = sharding.contiguous_shard(Bg, world, rank)
        full = measure_shape(len(mine), Tx3, Ty3, want_k1=False)
        tx, ty = ragged_lengths_torch(Bg, Tx3, Ty3, SEED + 3)
        costs = (tx.double() * ty.double()).tolist()
        shards = sharding.balanced_shards(costs, world)
        idx = torch.tensor(shards[rank], dtype=torch.long)
        rag = measure_shape(len(idx), Tx3, Ty3, tx[idx], ty[idx], want_k1=False)
        cidx = torch.tensor(list(mine), dtype=torch.long)
        rag_contig = measure_shape(len(cidx), Tx3, Ty3, tx[cidx], ty[cidx], want_k1=False)
        cg = Bg * Tx3 * Ty3
        valid = float((tx.double() * ty.double()).sum())
        c3_strong = {"workload": "C3 B=256 T_text=400 T_mel=2000 split by utterance over the GPUs, fused entry, no collective",
                     "n_gpus": world, "per_gpu_batch": len(mine),
                     "full": {"ms_per_step": full["fused_ms"], "cells_per_s": cg / (full["fused_ms"] * 1e-3),
                              "sharding": "contiguous_shard"},
                     "ragged": {"valid_cells_fraction": valid / cg,
                                "balanced_shards": {"ms_per_step": rag["fused_ms"], "padded_cells_per_s": cg / (rag["fused_ms"] * 1e-3),
                                                    "shard_sizes": [len(sh) for sh in shards]},
                                "contiguous_shard": {"ms_per_step": rag_contig["fused_ms"],
                                                     "padded_cells_per_s": cg / (rag_contig["fused_ms"] * 1e-3)}},
                     "note": "times are the slowest rank's (max over ranks per replay); efficiency vs the 1-GPU line's c3_strong"}

    # ---- C5: the reference's training step with the module swapped (profiles/c5_train_step.py) ----
    c5 = None
    if not args.no_c5:
        try:
            sys.path.insert(0, str(REPO / "profiles"))
            import c5_train_step  # noqa: WPS433

            c5 = c5_train_step.run(steps=4, warmup=2)
        except Exception as exc:  # noqa: BLE001
            c5 = {"error": repr(exc)[:300]}

    if rank == 0:
        traffic_all = {}
        tfile = REPO / "profiles" / "traffic.json"
        if tfile.exists():
            traffic_all = json.loads(tfile.read_text())
        # kernel (1) alone: read fp32 scores + write fp32 path = 8 B/cell
        k1_bytes = 8 * cells
        k1 = {"kernel": "mas_path_systolic_kernel (kernel 1: DP + backtrack + dense path)", "bound": "hbm",
              "achieved": k1_bytes / (kern_ms * 1e-3) / 1e9, "peak": peak, "unit": "GB/s",
              "frac": k1_bytes / (kern_ms * 1e-3) / 1e9 / peak, "traffic": traffic_all.get("c1"),
              "algorithmic_bytes_per_launch": k1_bytes, "kernel_ms": kern_ms, "best_kernel_ms": k1_best / args.steps}
        step_ms = dev_ms / args.steps
        if standalone:
            roofline = dict(k1, peak_source=peak_src)
        else:
            roofline = fused_roofline(cells, in_bytes, out_bytes, B, mean_only, step_ms, peak, fp32_peak)
            roofline.update({"kernel": "mas_fused_kernel (kernel 2: one cluster of CTAs per utterance; FFMA teams + one sweep warp per CTA, "
                                       "scores in shared memory)",
                             "traffic": traffic_all.get("fused_c2"), "peak_source": peak_src,
                             "fp32_peak_derived_from": {"sms": n_sms, "lanes_per_sm": 128, "sm_max_mhz": sm_mhz},
                             "note": "bounded by the FP32 FMA pipe and the sweep's dependent chain, not by HBM (DESIGN.md 5-6)",
                             "kernel1_alone": k1})
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": step_ms, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": config_for(args.workload, world, mean_only),
            "method": {"timing": f"{args.steps} steps captured in one CUDA graph; {replays} replays, each between two events and a "
                                 "barrier + synchronize; value = the MEDIAN replay (max over ranks per replay)",
                       "replays_ms": all_ms, "best_ms_per_step": best_ms / args.steps,
                       "l2": f"inputs/outputs rotated over {n_sets} buffer sets ({n_sets * per_set / 2**20:.0f} MiB > L2)",
                       "kernels_per_step": launches_per_step,
                       "fused_geometry": None if standalone else dict(zip(
                           ["ctas_per_utterance", "tokens_per_lane", "tokens_per_cta", "teams", "warps_per_team", "column_groups",
                            "frames_per_chunk", "ring_boxes", "bits_in_smem", "smem_bytes", "ffma_warps", "ring_rows"], geom.tolist()))},
            "roofline": roofline,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": in_bytes, "d2h_bytes_per_step": out_bytes,
                    "steps": e2e_steps, "ms_per_step": e2e_ms / e2e_steps,
                    "h2d_gb_per_s_all_gpus": world * in_bytes * e2e_steps / (e2e_ms * 1e-3) / 1e9,
                    "d2h_gb_per_s_all_gpus": world * out_bytes * e2e_steps / (e2e_ms * 1e-3) / 1e9,
                    "pipeline": "3 streams, triple-buffered pinned results; each step: H2D inputs, fused call, D2H dense path + durations",
                    "host": {k: v for k, v in host_info.items() if k != "restore"},
                    "durations_only": {"value": world * cells * e2e_steps / (e2e_compact_ms * 1e-3), "unit": UNIT,
                                       "d2h_bytes_per_step": B * T_x * 4, "ms_per_step": e2e_compact_ms / e2e_steps,
                                       "what": "same loop, the dense path stays on the device (its consumers run there); "
                                               "only the integer durations are read back"}},
            "clocks": clocks,
            "gpu_launches": launches_per_step * args.steps,
        }
        if configs is not None:
            line["configs"] = configs
        if c3_strong is not None:
            line["c3_strong"] = c3_strong
        if c5 is not None:
            line["c5"] = c5
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
            try:
                line["cpu_baseline"]["kernel_only"] = cython_kernel_only(
                    {"c1": (32, 200, 1000), "c3_shard_of_8": (32, 400, 2000), "c4": (2, 1024, 8192)})
            except Exception as exc:  # noqa: BLE001
                line["cpu_baseline"]["kernel_only"] = {"error": repr(exc)[:200]}
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
    ap.add_argument("--no-configs", action="store_true", help="skip the c3/c4 sub-lines and the strong-scaling C3 block")
    ap.add_argument("--no-c5", action="store_true", help="skip the training-step block (the reference's train_step with the module swapped)")
    ap.add_argument("--mean-only", action="store_true",
                    help="x_logs == 0 (the reference's default ModelConfig.mean_only): the contraction halves")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
