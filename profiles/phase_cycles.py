"""Per-warp phase timing of kernel (1) from the clock64() stamps of the profiling hook.

    python profiles/phase_cycles.py [B T_x T_y]

Prints, per DP warp (median over utterances): sweep cycles, cycles spent waiting for the previous
warp's boundary / the next warp's ring slot / TMA, and the CTA-level phases (backtrack, output).
"""
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import __graft_entry__ as entry  # noqa: E402

pkg = entry.load_package()
lib = pkg._lib.load()
B, T_x, T_y = (int(a) for a in sys.argv[1:4]) if len(sys.argv) >= 4 else (32, 200, 1000)
dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(1)
vals = [(10 * torch.randn(B, T_x, T_y, generator=g) - 100).to(dev) for _ in range(6)]
tx = torch.full((B,), T_x, dtype=torch.int32, device=dev)
ty = torch.full((B,), T_y, dtype=torch.int32, device=dev)
for v in vals:
    pkg.maximum_path_from_lengths(v, tx, ty)
torch.cuda.synchronize()
K = int(sys.argv[4]) if len(sys.argv) > 4 else 0          # CTAs per utterance (0 = heuristic)
lib.mas_b200_debug_force_cluster(K)
buf = torch.zeros(B * 8, 16, 16, dtype=torch.int64, device=dev)
lib.mas_b200_debug_set_cycle_buffer(buf.data_ptr())
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
pkg.maximum_path_from_lengths(vals[0], tx, ty)
e1.record()
torch.cuda.synchronize()
lib.mas_b200_debug_set_cycle_buffer(None)
d = buf.cpu().numpy()
d = d[d[:, :, 0].any(axis=1)]                                # CTAs that ran
print(f"shape B={B} T_x={T_x} T_y={T_y}; event time {e0.elapsed_time(e1) * 1e3:.1f} us")
t0 = d[:, :, 0].copy()
t0[t0 == 0] = np.iinfo(np.int64).max
start = t0.min(axis=1)                                      # per CTA
for w in range(16):
    if not d[:, w, 1].any():
        continue
    sweep = np.median(d[:, w, 1] - d[:, w, 0])
    print(f"warp {w:2d}: start +{np.median(d[:, w, 0] - start):8.0f}  sweep/fill {sweep:9.0f} cyc   "
          f"wait prev {np.median(d[:, w, 2]):8.0f}  next {np.median(d[:, w, 3]):8.0f}  tma {np.median(d[:, w, 4]):8.0f}"
          f"  | blocks {np.median(d[:, w, 10]):4.0f} compute {np.median(d[:, w, 8]):8.0f} of which inside sweep_block {np.median(d[:, w, 9]):7.0f}")
print(f"{len(d)} CTAs ({len(d) // B} per utterance)")
print(f"CTA: barrier1 at +{np.median(d[:, 0, 5] - start):.0f}, backtrack {np.median(d[:, 0, 6] - d[:, 0, 5]):.0f} cyc, "
      f"end at +{np.median(d[:, 0, 7] - start):.0f} cyc")
