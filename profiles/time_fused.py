"""Time the fused entry (single launch) vs the two kernels back to back, events around eager launches
and around a CUDA graph of N steps:  python profiles/time_fused.py [B T_x T_y] [--ragged]
(--ragged: LJSpeech-like lengths as in SURVEY.md 8d -- t_x ~ U[T_x/2, T_x], t_y ~ t_x T_y/T_x U[0.8, 1.2],
sorted by t_x descending, element 0 full size; times are per padded batch)"""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import __graft_entry__ as entry  # noqa: E402

pkg = entry.load_package()
lib = pkg._lib.load()
args = [a for a in sys.argv[1:] if not a.startswith("--")]
ragged = "--ragged" in sys.argv
B, T_x, T_y = (int(a) for a in args[:3]) if len(args) >= 3 else (32, 200, 1000)
D = 80
dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(1)
sets = []
for i in range(6):
    x_m = torch.randn(B, D, T_x, generator=g).to(dev)
    x_logs = (0.3 * torch.randn(B, D, T_x, generator=g) - 0.5).to(dev)
    z = torch.randn(B, D, T_y, generator=g).to(dev)
    sets.append((x_m, x_logs, z))
xl = torch.full((B,), T_x, dtype=torch.int32, device=dev)
yl = torch.full((B,), T_y, dtype=torch.int32, device=dev)
if ragged:
    tx = torch.randint(T_x // 2, T_x + 1, (B,), generator=g)
    tx[0] = T_x
    tx, _ = torch.sort(tx, descending=True)
    ty = (tx.float() * (T_y / T_x) * (0.8 + 0.4 * torch.rand(B, generator=g))).round().long()
    ty = torch.minimum(torch.maximum(ty, tx), torch.tensor(T_y)) // 2 * 2
    ty[0] = T_y
    xl, yl = tx.to(torch.int32).to(dev), torch.maximum(ty, tx).to(torch.int32).to(dev)
    print(f"ragged lengths: valid cells {float((xl.float() * yl.float()).sum()) / (B * T_x * T_y):.1%} of the padded batch")


def run(i):
    s = sets[i % len(sets)]
    return pkg.fused_maximum_path(s[0], s[1], s[2], xl, yl)


for mode in (0, 1):
    lib.mas_b200_debug_force_unfused(mode)
    for i in range(3):
        run(i)
    torch.cuda.synchronize()
    ts = []
    for i in range(10):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        run(i)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    graph = torch.cuda.CUDAGraph()
    keep = []
    with torch.cuda.graph(graph):
        for i in range(10):
            keep.append(run(i))
    graph.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    graph.replay()
    e1.record()
    torch.cuda.synchronize()
    print(f"{'two launches' if mode else 'single launch':14s}: eager median {sorted(ts)[5]:8.1f} us (min {min(ts):.1f}), graph of 10: {e0.elapsed_time(e1) * 100:8.1f} us/step")
    del graph, keep
lib.mas_b200_debug_force_unfused(0)
