"""Timeline of the single-launch kernel from globaltimer stamps (profiling hook):
when do the producers finish each chunk, when do the sweep CTAs start / finish?
    python profiles/fused_timeline.py [B T_x T_y]"""
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import __graft_entry__ as entry  # noqa: E402

pkg = entry.load_package()
lib = pkg._lib.load()
B, T_x, T_y = (int(a) for a in sys.argv[1:4]) if len(sys.argv) >= 4 else (32, 200, 1000)
D = 80
dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(1)
x_m = torch.randn(B, D, T_x, generator=g).to(dev)
x_logs = (0.3 * torch.randn(B, D, T_x, generator=g) - 0.5).to(dev)
z = torch.randn(B, D, T_y, generator=g).to(dev)
xl = torch.full((B,), T_x, dtype=torch.int32, device=dev)
yl = torch.full((B,), T_y, dtype=torch.int32, device=dev)
for _ in range(3):
    pkg.fused_maximum_path(x_m, x_logs, z, xl, yl)
torch.cuda.synchronize()
ncta = 2048
buf = torch.zeros(B * 16 + ncta, 16, dtype=torch.int64, device=dev)
lib.mas_b200_debug_set_cycle_buffer(buf.data_ptr())
pkg.fused_maximum_path(x_m, x_logs, z, xl, yl)
torch.cuda.synchronize()
lib.mas_b200_debug_set_cycle_buffer(None)
d = buf.cpu().numpy()
dp = d[: B * 16].reshape(B, 16, 16)
prod = d[B * 16:]
prod = prod[prod[:, 0] > 0]
t0 = min(prod[:, 0].min(), dp[:, 0, 11].min())
print(f"{len(prod)} producer CTAs, {B} sweep CTAs; times in us from the first stamp")
names = {8: "token side staged", 9: "contraction #1 done", 10: "contraction #2 done", 11: "contraction #3 done",
         12: "chunk #1 operands visible", 13: "chunk #2 operands visible", 14: "chunk #3 operands visible"}
start = prod[:, 0]
print(f"  producers start    {np.median(start - t0) / 1e3:7.1f}")
for k in sorted(names):
    col = prod[:, k]
    ok = col > 0
    if ok.any():
        print(f"  producers: {names[k]:28s} {np.median(col[ok] - t0) / 1e3:7.1f}")
cyc = (prod[:, 6] - prod[:, 5]).astype(float)
ns = (prod[:, 9] - prod[:, 12]).astype(float)
print(f"  producers: contraction #1 = {np.median(cyc):.0f} cycles in {np.median(ns) / 1e3:.2f} us -> {np.median(cyc / ns):.3f} GHz")
for k in range(1, 5):
    col = prod[:, k]
    col = col[col > 0]
    if len(col):
        print(f"  producers: chunk #{k} of each CTA stored at {np.median(col - t0) / 1e3:7.1f} (min {(col.min() - t0) / 1e3:.1f}, max {(col.max() - t0) / 1e3:.1f})")
print(f"  sweep CTAs start   {np.median(dp[:, 0, 11] - t0) / 1e3:7.1f}")
print(f"  sweep done         {np.median(dp[:, 0, 12] - t0) / 1e3:7.1f} (max {(dp[:, 0, 12].max() - t0) / 1e3:.1f})")
print(f"  CTA done           {np.median(dp[:, 0, 13] - t0) / 1e3:7.1f} (max {(dp[:, 0, 13].max() - t0) / 1e3:.1f})")
for w in range(4):
    print(f"  sweep warp {w}: wait prev {np.median(dp[:, w, 2]):8.0f} cyc, wait tma/flags {np.median(dp[:, w, 4]):8.0f} cyc, sweep {np.median(dp[:, w, 8]):8.0f} cyc, blocks {np.median(dp[:, w, 10]):.0f}")
