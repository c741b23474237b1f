"""Random shapes, lengths and score distributions (dense, exact ties, NaN / inf sprinkled in) through
kernel (1) for a few minutes, every path compared bit for bit with the oracle:
  python profiles/fuzz_path.py [seconds]          (B200, round 1: 504 shapes in 120 s, 0 failures)"""
import sys
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests"))
import numpy as np  # noqa: E402
import torch  # noqa: E402

import __graft_entry__ as entry  # noqa: E402

pkg = entry.load_package()
oracle = entry.load_oracle()
from conftest import ragged_lengths  # noqa: E402

budget = float(sys.argv[1]) if len(sys.argv) > 1 else 120.0
import os
rng = np.random.default_rng(int(os.environ.get("FUZZ_SEED", "4242")))
t0 = time.time()
n = bad = 0
while time.time() - t0 < budget:
    B = int(rng.integers(1, 40))
    T_x = int(rng.integers(1, 700))
    T_y = int(rng.integers(T_x, 3000))
    if rng.random() < 0.7:
        T_y = (T_y + 3) // 4 * 4
    kind = rng.integers(0, 4)
    if kind == 0:
        v = (10 * rng.standard_normal((B, T_x, T_y)) - 100).astype(np.float32)
    elif kind == 1:
        v = -rng.integers(0, 3, (B, T_x, T_y)).astype(np.float32)          # exact ties everywhere
    elif kind == 2:
        v = rng.standard_normal((B, T_x, T_y)).astype(np.float32)            # positive scores too
    else:
        v = (10 * rng.standard_normal((B, T_x, T_y)) - 100).astype(np.float32)
        k = max(1, v.size // 5000)
        idx = rng.integers(0, v.size, k)
        v.reshape(-1)[idx] = rng.choice(np.array([np.nan, np.inf, -np.inf, 3e38, -3e38], np.float32), k)
    t_x, t_y = ragged_lengths(rng, B, T_x, T_y)
    if rng.random() < 0.3:
        t_x[:], t_y[:] = T_x, T_y
    got, dur = pkg.maximum_path_from_lengths(torch.from_numpy(v).cuda(), torch.from_numpy(t_x).cuda(), torch.from_numpy(t_y).cuda(),
                                             want_durations=True)
    want = oracle.maximum_path(v, t_x, t_y)
    ok = np.array_equal(got.cpu().numpy().astype(np.int32), want) and np.array_equal(dur.cpu().numpy(), want.sum(-1))
    n += 1
    if not ok:
        bad += 1
        print("FAIL", B, T_x, T_y, int(kind), flush=True)
print(f"{n} shapes, {bad} failures")
