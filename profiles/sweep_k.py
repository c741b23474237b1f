"""Kernel (1) time vs CTAs-per-utterance (cluster size):  python profiles/sweep_k.py B T_x T_y [K ...]
Times a CUDA graph of 20 launches over rotating inputs (> L2) with events; prints us per launch and
achieved GB/s on the algorithmic 8 B/cell."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import __graft_entry__ as entry  # noqa: E402

pkg = entry.load_package()
lib = pkg._lib.load()
B, T_x, T_y = (int(a) for a in sys.argv[1:4])
Ks = [int(a) for a in sys.argv[4:]] or [0, 1, 2, 4, 8]
dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(1)
cells = B * T_x * T_y
nset = max(3, int(2.5 * 126 * 2**20 // (8 * cells)) + 1)
vals = [(10 * torch.randn(B, T_x, T_y, generator=g) - 100).to(dev) for _ in range(nset)]
tx = torch.full((B,), T_x, dtype=torch.int32, device=dev)
ty = torch.full((B,), T_y, dtype=torch.int32, device=dev)
steps = 20
for K in Ks:
    if K and T_x // K < 32:
        continue
    lib.mas_b200_debug_force_cluster(K)
    try:
        for i in range(3):
            pkg.maximum_path_from_lengths(vals[i % nset], tx, ty)
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        keep = []
        with torch.cuda.graph(graph):
            for i in range(steps):
                keep.append(pkg.maximum_path_from_lengths(vals[i % nset], tx, ty))
        graph.replay()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        graph.replay()
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 1e3 / steps
        print(f"B={B} T_x={T_x} T_y={T_y} K={K or 'auto'}: {us:9.1f} us  {cells / us / 1e3:8.1f} Gcells/s  {8 * cells / us / 1e3:8.1f} GB/s")
        del graph, keep
    except RuntimeError as e:
        print(f"K={K}: {e}")
lib.mas_b200_debug_force_cluster(0)
