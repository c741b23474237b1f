"""Run kernel (1) alone a few times on synthetic scores (for ncu):  python profiles/run_kernel1.py B T_x T_y [reps]"""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import __graft_entry__ as entry  # noqa: E402

pkg = entry.load_package()
B, T_x, T_y = (int(a) for a in sys.argv[1:4])
reps = int(sys.argv[4]) if len(sys.argv) > 4 else 4
dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(1)
vals = [(10 * torch.randn(B, T_x, T_y, generator=g) - 100).to(dev) for _ in range(3)]
tx = torch.full((B,), T_x, dtype=torch.int32, device=dev)
ty = torch.full((B,), T_y, dtype=torch.int32, device=dev)
for i in range(reps):
    out = pkg.maximum_path_from_lengths(vals[i % 3], tx, ty)
torch.cuda.synchronize()
print("ok", float(out.sum()))
