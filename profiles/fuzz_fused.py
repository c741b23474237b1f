"""Random shapes, lengths and channel counts through the fused entry for a few minutes, every result
checked against the oracle (path and durations bit-exact on our scores, scores within 1e-5 of fp64):
  python profiles/fuzz_fused.py [seconds]        (B200, round 1: 1 863 shapes in 150 s, 0 failures)
Round 2: every shape goes through the single launch (forced) AND through whatever the entry's estimate
picks; batches up to 400 utterances (several rounds per cluster), tokens up to 1024 (long slices, bits
in the workspace), some batches with NaN / inf in z (the literal redo)."""
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
from test_fused_gpu import synth_prior, to_dev  # noqa: E402

budget = float(sys.argv[1]) if len(sys.argv) > 1 else 150.0
import os
verbose = os.environ.get('FUZZ_VERBOSE') is not None
only = int(os.environ['FUZZ_ONLY']) if 'FUZZ_ONLY' in os.environ else None
repeat = int(os.environ.get('FUZZ_REPEAT', '1'))
rng = np.random.default_rng(int(os.environ.get("FUZZ_SEED", "20261018")))
t0 = time.time(); n = 0; bad = 0
while time.time() - t0 < budget:
    big = rng.random() < 0.15
    B = int(rng.integers(1, 400 if not big else 6)); T_x = int(rng.integers(1, 320 if not big else 1025))
    if B > 60: T_x = min(T_x, 120)
    T_y = int(rng.integers(T_x, max(T_x + 1, (1300 if B <= 60 else 400) if not big else 3000)))
    if rng.random() < 0.7: T_y = (T_y + 3) // 4 * 4
    D = 80 if rng.random() < 0.8 else int(rng.integers(1, 100))
    mean_only = bool(rng.random() < 0.4)
    t_x, t_y = ragged_lengths(rng, B, T_x, T_y)
    if rng.random() < 0.3: t_x[:] = T_x; t_y[:] = T_y
    x_m, x_logs, z = synth_prior(rng, B, D, T_x, T_y, t_x, t_y, mean_only)
    args = (to_dev(x_m), to_dev(x_logs), to_dev(z), torch.from_numpy(t_x), torch.from_numpy(t_y))
    poison = torch.full((pkg._lib.load().mas_b200_fused_workspace_bytes(B, D, T_x, T_y) // 4 + 64,), float('nan'), device='cuda:0'); del poison
    if rng.random() < 0.1:
        zz = args[2].clone()
        for b in range(0, B, 3):
            zz[b, int(rng.integers(0, D)), int(rng.integers(0, max(1, t_y[b])))] = float(rng.choice([np.nan, np.inf, -np.inf]))
        args = (args[0], args[1], zz, args[3], args[4])
    lib = pkg._lib.load()
    if only is not None and n != only:      # FUZZ_ONLY=i: draw everything, run only shape i (FUZZ_REPEAT times)
        n += 1
        continue
    if verbose:     # FUZZ_VERBOSE=1: the shape before it runs and a synchronize after every call (a kernel fault names its shape)
        print(n, 'B', B, 'D', D, 'T_x', T_x, 'T_y', T_y, 'mean_only', mean_only, 't_x', t_x[:6].tolist(), 't_y', t_y[:6].tolist(), flush=True)
    lib.mas_b200_debug_force_unfused(2)
    for rep_i in range(repeat - 1):
        pkg.fused_maximum_path(*args, want_frame_token=True)
        torch.cuda.synchronize()
        print('  repeat', rep_i, 'ok', flush=True)
    path, dur, tok = pkg.fused_maximum_path(*args, want_frame_token=True)
    if verbose: torch.cuda.synchronize(); print('  forced single launch ok', flush=True)
    lib.mas_b200_debug_force_unfused(0)
    path0, dur0, tok0 = pkg.fused_maximum_path(*args, want_frame_token=True)
    if verbose: torch.cuda.synchronize(); print('  entry ok', flush=True)
    logp = pkg.log_likelihood_matrix(*args[:3])
    k1 = pkg.maximum_path_from_lengths(logp, to_dev(t_x), to_dev(t_y))
    ok = torch.equal(k1, path) and torch.equal(path0, path) and torch.equal(dur0, dur) and torch.equal(tok0, tok)
    want = oracle.maximum_path(logp.cpu().numpy(), t_x, t_y)
    ok = ok and np.array_equal(path.cpu().numpy().astype(np.int32), want) and np.array_equal(dur.cpu().numpy(), want.sum(-1))
    ref64 = oracle.logp_f64(x_m, x_logs, args[2].cpu().numpy())
    fin = np.isfinite(ref64)
    rel = np.max(np.abs(logp.cpu().numpy()[fin] - ref64[fin]) / np.maximum(np.abs(ref64[fin]), 1.0)) if fin.any() else 0.0
    # (with a handful of channels the terms cancel to scores near zero: the formula's own conditioning)
    ok = ok and rel < (1e-5 if D >= 16 else 1e-4)
    n += 1
    if not ok:
        bad += 1; print('FAIL', B, D, T_x, T_y, mean_only, rel, flush=True)
    if only is not None:
        break
print(f'{n} shapes, {bad} failures')
