"""Run the log-likelihood kernel alone (for ncu / timing):  python profiles/run_logp.py [B D T_x T_y] [--mean-only]"""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import __graft_entry__ as entry  # noqa: E402

pkg = entry.load_package()
args = [a for a in sys.argv[1:] if not a.startswith("--")]
mean_only = "--mean-only" in sys.argv
B, D, T_x, T_y = (int(a) for a in args[:4]) if len(args) >= 4 else (32, 80, 200, 1000)
dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(1)
sets = []
for _ in range(5):      # rotate > L2 worth of operands + outputs
    x_m = torch.randn(B, D, T_x, generator=g).to(dev)
    x_logs = None if mean_only else (0.3 * torch.randn(B, D, T_x, generator=g)).to(dev)
    z = torch.randn(B, D, T_y, generator=g).to(dev)
    sets.append((x_m, x_logs, z))
for s in sets:
    out = pkg.log_likelihood_matrix(*s)
torch.cuda.synchronize()
ev = [torch.cuda.Event(enable_timing=True) for _ in range(21)]
ev[0].record()
for i in range(20):
    out = pkg.log_likelihood_matrix(*sets[i % 5])
    ev[i + 1].record()
torch.cuda.synchronize()
ts = sorted(ev[i].elapsed_time(ev[i + 1]) * 1e3 for i in range(20))
cells = B * T_x * T_y
print(f"logp kernel alone B={B} D={D} T_x={T_x} T_y={T_y} mean_only={mean_only}: median {ts[10]:.1f} us  min {ts[0]:.1f} us"
      f"  ({cells * D * 2 / ts[10] / 1e6:.2f} TFFMA/s of the two contractions)")
