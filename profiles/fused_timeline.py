"""Timeline of the single-launch kernel (2) from globaltimer stamps (profiling hook): per CTA of a
cluster -- operands staged, each FFMA team done, sweep done, backtrack done, output done.
    python profiles/fused_timeline.py [B T_x T_y] [--mean-only] [--random]"""
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import __graft_entry__ as entry  # noqa: E402

pkg = entry.load_package()
lib = pkg._lib.load()
args = [a for a in sys.argv[1:] if not a.startswith("--")]
B, T_x, T_y = (int(a) for a in args[:3]) if len(args) >= 3 else (32, 200, 1000)
D = 80
dev = torch.device("cuda:0")
g = torch.Generator().manual_seed(1)
x_m = torch.randn(B, D, T_x, generator=g).to(dev)
x_logs = None if "--mean-only" in sys.argv else (0.3 * torch.randn(B, D, T_x, generator=g) - 0.5).to(dev)
if "--random" in sys.argv:        # unstructured scores: long walks in the block maps and the backtrack
    z = torch.randn(B, D, T_y, generator=g).to(dev)
else:                             # "trained-like" (what bench.py times): z follows the token it belongs to
    idx = (torch.arange(T_y) * T_x) // T_y
    xl_cpu = torch.zeros(B, D, T_x) if x_logs is None else x_logs.cpu()
    z = (x_m.cpu()[:, :, idx] + torch.exp(xl_cpu[:, :, idx]) * torch.randn(B, D, T_y, generator=g)).to(dev)
xl = torch.full((B,), T_x, dtype=torch.int32, device=dev)
yl = torch.full((B,), T_y, dtype=torch.int32, device=dev)
for _ in range(3):
    pkg.fused_maximum_path(x_m, x_logs, z, xl, yl)
torch.cuda.synchronize()
buf = torch.zeros(2048, 32, dtype=torch.int64, device=dev)
lib.mas_b200_debug_set_cycle_buffer(buf.data_ptr())
pkg.fused_maximum_path(x_m, x_logs, z, xl, yl)
torch.cuda.synchronize()
lib.mas_b200_debug_set_cycle_buffer(None)
d = buf.cpu().numpy()
d = d[d[:, 0] > 0]
out = np.zeros(12, np.int32)
props = torch.cuda.get_device_properties(0)
lib.mas_b200_debug_fused_geom(B, D, T_x, T_y, 232448 - 1024, props.multi_processor_count, out.ctypes.data)
K = int(out[0])
print(f"{len(d)} CTAs, clusters of {K}; geometry {out.tolist()}; times in us from the first stamp (last utterance of each CTA)")
t0 = d[:, 0].min()
for c in range(K):
    r = d[c::K]
    def med(k):
        col = r[:, k]
        col = col[col > 0]
        return (np.median(col - t0) / 1e3, (col.max() - t0) / 1e3) if len(col) else (float("nan"), float("nan"))
    teams = " ".join(f"{med(8 + t)[0]:6.1f}" for t in range(5) if (r[:, 8 + t] > 0).any())
    print(f"  CTA {c}: start {med(0)[0]:5.1f} | raw operands {med(26)[0]:5.1f}, element-wise {med(27)[0]:5.1f}, all {med(1)[0]:5.1f} | teams done {teams} | sweep done {med(4)[0]:6.1f} (max {med(4)[1]:6.1f})"
          f" | zero fill waited {med(14)[0]:6.1f}, maps (warp 1) {med(3)[0]:6.1f} (last builder {med(15)[0]:6.1f}), CTA barrier {med(2)[0]:6.1f} | backtrack {med(5)[0]:6.1f} (handed down {med(13)[0]:6.1f}) -> {med(6)[0]:6.1f} | output {med(7)[0]:6.1f} (max {med(7)[1]:6.1f})")
    cyc = lambda k: float(np.median(r[:, k]))
    nb = max(cyc(21), 1.0)
    print(f"         sweep warp, cycles per 32-frame block ({nb:.0f} blocks): waits chunks {cyc(16) / nb:6.0f}, boundary {cyc(17) / nb:6.0f}, credit {cyc(18) / nb:6.0f};"
          f" sweep {cyc(19) / nb:6.0f}, bits {cyc(22) / nb:5.0f}, boundary hand-back {cyc(23) / nb:5.0f}, consumed + zero fill {cyc(24) / nb:5.0f}; all {cyc(20) / nb:6.0f}; last eight blocks: sweep {cyc(28) / 8:6.0f}, all {cyc(29) / 8:6.0f}")
    print(f"         backtrack thread: entry block walk {cyc(30):6.0f} cycles, {cyc(25):.0f} blocks composed (+ the exit block walked) {cyc(31):6.0f} cycles")
