"""One shape through the forced single launch of kernel (2), checked against kernel (1) on the
materialised scores; prints ok / mismatch / the CUDA error and the elapsed time of the call.
    python profiles/probe_shape.py B T_x T_y [t_x t_y]"""
import sys
import time
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import __graft_entry__ as entry  # noqa: E402

pkg = entry.load_package()
lib = pkg._lib.load()
B, T_x, T_y = (int(a) for a in sys.argv[1:4])
tx, ty = (int(sys.argv[4]), int(sys.argv[5])) if len(sys.argv) > 5 else (T_x, T_y)
dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(3)
x_m = torch.randn(B, 80, T_x, generator=g).to(dev)
x_logs = (0.3 * torch.randn(B, 80, T_x, generator=g) - 0.5).to(dev)
z = torch.randn(B, 80, T_y, generator=g).to(dev)
xl = torch.full((B,), tx, dtype=torch.int32, device=dev)
yl = torch.full((B,), ty, dtype=torch.int32, device=dev)
logp = pkg.log_likelihood_matrix(x_m, x_logs, z)
want = pkg.maximum_path_from_lengths(logp, xl, yl)
torch.cuda.synchronize()
lib.mas_b200_debug_force_unfused(2)
t0 = time.time()
try:
    path, dur = pkg.fused_maximum_path(x_m, x_logs, z, xl, yl)
    torch.cuda.synchronize()
    print(f"B={B} {T_x}x{T_y} ({tx},{ty}):", "ok" if torch.equal(path, want) else "MISMATCH", f"{time.time() - t0:.2f} s")
except Exception as e:  # noqa: BLE001
    print(f"B={B} {T_x}x{T_y} ({tx},{ty}): {str(e).splitlines()[0]} after {time.time() - t0:.2f} s")
