"""A few small batches through kernel (2)'s single launch, for compute-sanitizer:
    compute-sanitizer --tool memcheck  python profiles/sanitize_fused.py
    compute-sanitizer --tool racecheck python profiles/sanitize_fused.py
(forced single launch: multi-CTA clusters, several rounds per cluster, few channels / mean_only,
non-finite input -> redo)."""
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
import __graft_entry__ as entry  # noqa: E402

pkg = entry.load_package()
from conftest import ragged_lengths  # noqa: E402
from test_fused_gpu import synth_prior, to_dev  # noqa: E402

lib = pkg._lib.load()
lib.mas_b200_debug_force_unfused(2)
rng = np.random.default_rng(3)
for (B, D, T_x, T_y, mean_only, poison) in [(3, 80, 200, 520, False, False), (5, 16, 96, 400, True, False), (150, 20, 24, 64, True, False),
                                            (2, 80, 300, 640, False, True), (2, 40, 1024, 1280, False, False)]:
    t_x, t_y = ragged_lengths(rng, B, T_x, T_y)
    x_m, x_logs, z = synth_prior(rng, B, D, T_x, T_y, t_x, t_y, mean_only)
    if poison:
        z[0, 3, 17] = np.nan
    args = (to_dev(x_m), to_dev(x_logs), to_dev(z), torch.from_numpy(t_x), torch.from_numpy(t_y))
    path, dur, tok = pkg.fused_maximum_path(*args, want_frame_token=True)
    torch.cuda.synchronize()
    print((B, D, T_x, T_y, mean_only, poison), int(dur.sum()), int(t_y.sum()))
lib.mas_b200_debug_force_unfused(0)
