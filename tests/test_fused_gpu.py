"""Kernel (2) and the materialised log-likelihood matrix on a B200: scores within 1e-5 relative of
the fp64 formula (north_star's tolerance), path identical to kernel (1) on our own scores and to
the oracle on them, durations identical to the reference's except at documented near-ties."""
from __future__ import annotations

import zlib

import numpy as np
import pytest
import torch

from conftest import ragged_lengths

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
LOGP_RTOL = 1e-5      # BASELINE.json north_star: "logp within 1e-5 relative error"
EPS32 = float(np.finfo(np.float32).eps)


@pytest.fixture(params=["single launch", "by estimate"])
def fused_mode(request, pkg):
    """The fused entry picks between its single launch (a cluster of CTAs per utterance) and the two
    kernels back to back by a cost estimate; the tests run every case through the single launch
    (mas_b200_debug_force_unfused(2)) and through whatever the estimate picks."""
    lib = pkg._lib.load()
    lib.mas_b200_debug_force_unfused(2 if request.param == "single launch" else 0)
    yield request.param
    lib.mas_b200_debug_force_unfused(0)


def near_tie_report(ref64, p, want, t_x, t_y):
    """Frames where the path on our fp32 scores (`p`) and the path on the fp64 scores rounded to fp32
    (`want`) sit on different tokens, with the proof that each is a near-tie: going back from the
    last frame, two paths that agree on (token x, frame y) part exactly where one keeps the token and
    the other advances from x-1, i.e. where the reference compares V[x, y-1] with V[x-1, y-1]
    (core.pyx:34).  In the fp64 recurrence on the fp64 scores that pair must be closer than the
    fp32 rounding of the running scores the kernels compare.  Returns [(b, frame, gap, bound)]."""
    out = []
    for b in range(p.shape[0]):
        if np.array_equal(p[b], want[b]):
            continue
        tx, ty = int(t_x[b]), int(t_y[b])
        L = ref64[b, :tx, :ty].astype(np.float32).astype(np.float64)   # the values both sweeps start from
        V = np.full((tx, ty), -1e9)
        prev = np.full(tx, -1e9)
        for y in range(ty):
            adv = np.concatenate(([0.0 if y == 0 else -1e9], prev[:-1]))
            cur = np.maximum(prev, adv) + L[:, y]
            cur[np.arange(tx) > y] = -1e9
            cur[np.arange(tx) < tx + y - ty] = -1e9
            V[:, y] = cur
            prev = cur
        ra, rb = p[b, :tx, :ty].argmax(0), want[b, :tx, :ty].argmax(0)
        for y in range(ty - 1, 0, -1):
            if ra[y] == rb[y] and ra[y - 1] != rb[y - 1]:
                x = int(ra[y])
                gap = abs(V[x, y - 1] - V[x - 1, y - 1])
                # both candidates carry the rounding of y fp32 additions (0.5 ulp each, random walk) of
                # scores of magnitude |V|; 8 eps |V| covers that with a wide margin and is still 1e-6 relative
                bound = 8 * EPS32 * max(abs(V[x, y - 1]), abs(V[x - 1, y - 1]))
                out.append((b, y, gap, bound))
    return out


def synth_prior(rng, B, D, T_x, T_y, t_x, t_y, mean_only, trained_like=True):
    """SURVEY.md 8d synthetic inputs: x_m ~ N(0,1); x_logs = 0 (mean_only) or 0.3 N(0,1) - 0.5;
    z = x_m[:, :, y*T_x/T_y] + exp(x_logs) N(0,1) ("trained-like") or N(0,1); all pre-masked."""
    xmask = (np.arange(T_x)[None] < t_x[:, None]).astype(np.float32)[:, None]
    ymask = (np.arange(T_y)[None] < t_y[:, None]).astype(np.float32)[:, None]
    x_m = rng.standard_normal((B, D, T_x)).astype(np.float32) * xmask
    x_logs = None if mean_only else ((0.3 * rng.standard_normal((B, D, T_x)) - 0.5).astype(np.float32) * xmask)
    noise = rng.standard_normal((B, D, T_y)).astype(np.float32)
    if trained_like:
        z = np.empty((B, D, T_y), np.float32)
        for b in range(B):
            idx = np.minimum((np.arange(T_y) * t_x[b]) // max(int(t_y[b]), 1), t_x[b] - 1)
            scale = 1.0 if mean_only else np.exp(x_logs[b][:, idx])
            z[b] = x_m[b][:, idx] + scale * noise[b]
    else:
        z = noise
    return x_m, x_logs, z * ymask


def to_dev(a):
    return None if a is None else torch.from_numpy(a).to(DEV)


def test_logp_golden_from_reference_model(pkg, oracle, model_golden, fused_mode):
    g = model_golden
    x_logs = None if bool(g["mean_only"]) else g["x_logs"]
    got = pkg.log_likelihood_matrix(to_dev(g["x_m"]), to_dev(x_logs), to_dev(g["z"])).cpu().numpy()
    ref64 = oracle.logp_f64(g["x_m"], g["x_logs"], g["z"])
    assert np.max(np.abs(got - ref64) / np.abs(ref64)) < LOGP_RTOL
    assert np.max(np.abs(got - g["logp"]) / np.abs(g["logp"])) < LOGP_RTOL       # the reference's own fp32 logp
    # explicit zeros take the general contraction, None the mean_only one: same scores within tolerance
    if bool(g["mean_only"]):
        got2 = pkg.log_likelihood_matrix(to_dev(g["x_m"]), to_dev(g["x_logs"]), to_dev(g["z"])).cpu().numpy()
        assert np.max(np.abs(got - got2) / np.abs(got2)) < LOGP_RTOL
    path, dur = pkg.fused_maximum_path(to_dev(g["x_m"]), to_dev(x_logs), to_dev(g["z"]),
                                       torch.from_numpy(g["x_len"]), torch.from_numpy(g["y_len"]))
    assert np.array_equal(path.cpu().numpy().astype(np.int8), g["path"])           # the reference model's attn
    assert np.array_equal(dur.cpu().numpy(), g["path"].sum(-1))


@pytest.mark.parametrize("mean_only", [True, False])
@pytest.mark.parametrize("shape", [(3, 80, 5, 9), (2, 80, 33, 130), (4, 80, 70, 300), (2, 64, 64, 64),
                                   (2, 17, 40, 90), (8, 80, 200, 1000)])
def test_fused_parity(pkg, oracle, shape, mean_only, fused_mode):
    B, D, T_x, T_y = shape
    rng = np.random.default_rng(zlib.crc32(repr((shape, mean_only)).encode()))
    t_x, t_y = ragged_lengths(rng, B, T_x, T_y)
    x_m, x_logs, z = synth_prior(rng, B, D, T_x, T_y, t_x, t_y, mean_only)
    logp = pkg.log_likelihood_matrix(to_dev(x_m), to_dev(x_logs), to_dev(z))
    ref64 = oracle.logp_f64(x_m, x_logs, z)
    rel = np.max(np.abs(logp.cpu().numpy() - ref64) / np.abs(ref64))
    assert rel < LOGP_RTOL, rel
    path, dur, tok = pkg.fused_maximum_path(to_dev(x_m), to_dev(x_logs), to_dev(z), torch.from_numpy(t_x),
                                            torch.from_numpy(t_y), want_frame_token=True)
    p = path.cpu().numpy().astype(np.int32)
    # (a) identical to kernel (1) on our own materialised scores, and to the oracle on them: bit-exact
    k1 = pkg.maximum_path_from_lengths(logp, to_dev(t_x), to_dev(t_y))
    assert torch.equal(k1, path)
    assert np.array_equal(p, oracle.maximum_path(logp.cpu().numpy(), t_x, t_y))
    assert np.array_equal(dur.cpu().numpy(), p.sum(-1))
    # (b) against the fp64 scores rounded to fp32: identical durations except at near-ties -- and every
    # place where the two paths part is PROVEN to be one: the two candidates of that decision differ,
    # in the fp64 recurrence, by less than the fp32 rounding of the running scores
    want = oracle.maximum_path(ref64.astype(np.float32), t_x, t_y)
    ties = near_tie_report(ref64, p, want, t_x, t_y)
    if ties:
        print(f"near-ties {shape} mean_only={mean_only}: " + ", ".join(f"b={b} frame={y} gap={g:.3g} (bound {bd:.3g})" for b, y, g, bd in ties))
    for b, y, gap, bound in ties:
        assert gap <= bound, (b, y, gap, bound)
    if not ties:
        assert np.array_equal(p, want)


def test_fused_is_length_robust(pkg, oracle, fused_mode):
    """Garbage beyond the valid lengths must not leak into the path."""
    rng = np.random.default_rng(21)
    B, D, T_x, T_y = 3, 80, 30, 100
    t_x, t_y = np.array([30, 19, 7], np.int32), np.array([100, 64, 30], np.int32)
    x_m, x_logs, z = synth_prior(rng, B, D, T_x, T_y, t_x, t_y, False)
    a = pkg.fused_maximum_path(to_dev(x_m), to_dev(x_logs), to_dev(z), torch.from_numpy(t_x), torch.from_numpy(t_y))[0]
    for b in range(B):
        x_m[b, :, t_x[b]:] = 1e3
        x_logs[b, :, t_x[b]:] = -3.0
        z[b, :, t_y[b]:] = -1e3
    c = pkg.fused_maximum_path(to_dev(x_m), to_dev(x_logs), to_dev(z), torch.from_numpy(t_x), torch.from_numpy(t_y))[0]
    assert torch.equal(a, c)


@pytest.mark.parametrize("mean_only", [False, True])
@pytest.mark.parametrize("shape", [(4, 80, 200, 1000), (3, 80, 64, 256), (2, 40, 300, 640), (40, 80, 96, 320),
                                   (160, 80, 40, 96),      # more utterances than clusters: several rounds per cluster
                                   (3, 80, 1024, 1536),    # 128-token slices over 8 CTAs, direction bits in the workspace
                                   # a score ring of 5 boxes under three teams of 48-frame chunks: the first round of chunks
                                   # spans more boxes than the ring has (the fuzzer's find: the teams' first store used to
                                   # wait for ring room BEFORE the barrier the sweep -- which frees the room -- waits at)
                                   (1, 80, 824, 2496),
                                   (5, 33, 10, 52),
                                   # few channels: the contraction outruns the sweep (back-pressure on the score ring) and a
                                   # chunk is fewer z panels than the pipeline holds (a fast warp is a whole chunk ahead)
                                   (32, 22, 95, 956), (11, 16, 269, 1604), (28, 17, 77, 1540)])
def test_single_launch_equals_two_launches(pkg, oracle, shape, mean_only):
    """The single launch (scores produced and consumed in shared memory) and the two kernels back to
    back contract the same scores and run the same recurrence: identical path, durations and
    frame->token map, bit for bit."""
    lib = pkg._lib.load()
    B, D, T_x, T_y = shape
    rng = np.random.default_rng(zlib.crc32(repr(shape).encode()))
    t_x, t_y = ragged_lengths(rng, B, T_x, T_y)
    x_m, x_logs, z = synth_prior(rng, B, D, T_x, T_y, t_x, t_y, mean_only)
    args = (to_dev(x_m), to_dev(x_logs), to_dev(z), torch.from_numpy(t_x), torch.from_numpy(t_y))
    lib.mas_b200_debug_force_unfused(2)
    one = pkg.fused_maximum_path(*args, want_frame_token=True)
    torch.cuda.synchronize()
    lib.mas_b200_debug_force_unfused(1)
    try:
        two = pkg.fused_maximum_path(*args, want_frame_token=True)
        torch.cuda.synchronize()
    finally:
        lib.mas_b200_debug_force_unfused(0)
    for a, b in zip(one, two):
        assert torch.equal(a, b)
    logp = pkg.log_likelihood_matrix(*args[:3]).cpu().numpy()
    assert np.array_equal(one[0].cpu().numpy().astype(np.int32), oracle.maximum_path(logp, t_x, t_y))


@pytest.mark.parametrize("shape", [
    (32, 80, 200, 1000),    # C2: 4 dedicated CTAs per utterance + spares taking one chunk each
    (160, 80, 40, 96),      # more token tiles than SMs: a whole pass, then a partial one
    (5, 80, 300, 640),      # two token tiles per utterance
    (3, 80, 257, 264),      # tile of 132 tokens, chunk count that does not divide
    (2, 96, 50, 120),       # more than 80 channels: the panelled generic path
    (2, 80, 50, 122),       # T_y % 4 != 0: the generic path
    (1, 80, 7, 8),
    (2, 8, 200, 64),        # few channels: the staging scratch must not depend on D
    (3, 5, 120, 256),
])
def test_logp_every_unit_dealt_once(pkg, oracle, shape):
    """Every (utterance, token tile, chunk) unit must be produced exactly once whatever the deal of
    units to persistent CTAs: the output buffer is pre-filled with NaN, the scores are checked
    against the fp64 formula."""
    B, D, T_x, T_y = shape
    rng = np.random.default_rng(zlib.crc32(repr(shape).encode()))
    t_x, t_y = np.full(B, T_x, np.int32), np.full(B, T_y, np.int32)
    for mean_only in (False, True):
        x_m, x_logs, z = synth_prior(rng, B, D, T_x, T_y, t_x, t_y, mean_only, trained_like=False)
        out = torch.full((B, T_x, T_y), float("nan"), device=DEV)
        got = pkg.log_likelihood_matrix(to_dev(x_m), to_dev(x_logs), to_dev(z), out=out).cpu().numpy()
        assert np.isfinite(got).all()
        ref64 = oracle.logp_f64(x_m, x_logs, z)
        # (with a handful of channels a score can come close to zero: relative to max(|ref|, 1) there)
        rel = np.max(np.abs(got - ref64) / np.maximum(np.abs(ref64), 1.0))
        assert rel < LOGP_RTOL, (mean_only, rel)


def test_single_launch_ignores_garbage_in_the_workspace(pkg, oracle, fused_mode):
    """Only cells inside the reference's band are ever contracted; whatever the workspace held
    before (here: NaN, left in the caching allocator's block) must not reach the path."""
    lib = pkg._lib.load()
    B, D, T_x, T_y = 6, 80, 120, 520
    rng = np.random.default_rng(77)
    t_x, t_y = ragged_lengths(rng, B, T_x, T_y)
    t_x[1], t_y[1] = 40, 80           # a short utterance: whole chunks beyond its last frame
    t_x[2], t_y[2] = 1, 1
    x_m, x_logs, z = synth_prior(rng, B, D, T_x, T_y, t_x, t_y, False)
    args = (to_dev(x_m), to_dev(x_logs), to_dev(z), torch.from_numpy(t_x), torch.from_numpy(t_y))
    for _ in range(2):
        poison = torch.full((lib.mas_b200_fused_workspace_bytes(B, D, T_x, T_y) // 4 + 64,), float("nan"), device=DEV)
        torch.cuda.synchronize()
        del poison
        path, dur = pkg.fused_maximum_path(*args)
    logp = pkg.log_likelihood_matrix(*args[:3]).cpu().numpy()
    assert np.array_equal(path.cpu().numpy().astype(np.int32), oracle.maximum_path(logp, t_x, t_y))
    assert np.array_equal(dur.cpu().numpy(), path.cpu().numpy().sum(-1).astype(np.int32))


def test_fused_random_shapes(pkg, oracle, fused_mode):
    """A seeded slice of profiles/fuzz_fused.py: random batch sizes, lengths, channel counts and
    (un)aligned frame counts through every path the fused entry can take."""
    rng = np.random.default_rng(20261018)
    lib = pkg._lib.load()
    for _ in range(24):
        B, T_x = int(rng.integers(1, 50)), int(rng.integers(1, 320))
        T_y = int(rng.integers(T_x, 1300))
        if rng.random() < 0.7:
            T_y = (T_y + 3) // 4 * 4
        D = 80 if rng.random() < 0.8 else int(rng.integers(1, 100))
        mean_only = bool(rng.random() < 0.4)
        t_x, t_y = ragged_lengths(rng, B, T_x, T_y)
        if rng.random() < 0.3:
            t_x[:], t_y[:] = T_x, T_y
        x_m, x_logs, z = synth_prior(rng, B, D, T_x, T_y, t_x, t_y, mean_only)
        args = (to_dev(x_m), to_dev(x_logs), to_dev(z), torch.from_numpy(t_x), torch.from_numpy(t_y))
        poison = torch.full((lib.mas_b200_fused_workspace_bytes(B, D, T_x, T_y) // 4 + 64,), float("nan"), device=DEV)
        del poison
        path, dur = pkg.fused_maximum_path(*args)
        logp = pkg.log_likelihood_matrix(*args[:3])
        want = oracle.maximum_path(logp.cpu().numpy(), t_x, t_y)
        assert np.array_equal(path.cpu().numpy().astype(np.int32), want), (B, D, T_x, T_y, mean_only)
        assert np.array_equal(dur.cpu().numpy(), want.sum(-1)), (B, D, T_x, T_y, mean_only)
        ref64 = oracle.logp_f64(x_m, x_logs, z)
        rel = np.max(np.abs(logp.cpu().numpy() - ref64) / np.maximum(np.abs(ref64), 1.0))
        assert rel < LOGP_RTOL, (B, D, T_x, T_y, mean_only, rel)


@pytest.mark.parametrize("mean_only", [False, True])
@pytest.mark.parametrize("shape", [
    (32, 80, 200, 1000),     # C2, the benchmarked launch: 32 clusters of 4 CTAs, one round
    (64, 80, 400, 2000),     # a C3 shard: 100-token slices, bits in the workspace, two rounds
    (4, 80, 1024, 8192),     # C4: 128-token slices over 8 CTAs
])
def test_fused_at_the_benchmark_shapes(pkg, oracle, shape, mean_only):
    """The exact shapes bench.py times, through the fused entry, every way it can run them: bit-equal
    to kernel (1) on our materialised scores and to the oracle on them."""
    lib = pkg._lib.load()
    B, D, T_x, T_y = shape
    rng = np.random.default_rng(zlib.crc32(repr((shape, mean_only)).encode()))
    t_x, t_y = ragged_lengths(rng, B, T_x, T_y)
    t_x[: B // 2], t_y[: B // 2] = T_x, T_y          # half the batch full length (bench.py's `full` mode), half ragged
    x_m, x_logs, z = synth_prior(rng, B, D, T_x, T_y, t_x, t_y, mean_only)
    args = (to_dev(x_m), to_dev(x_logs), to_dev(z), torch.from_numpy(t_x), torch.from_numpy(t_y))
    logp = pkg.log_likelihood_matrix(*args[:3])
    k1, k1_dur = pkg.maximum_path_from_lengths(logp, to_dev(t_x), to_dev(t_y), want_durations=True)
    want = oracle.maximum_path(logp.cpu().numpy(), t_x, t_y, threads=8)
    assert np.array_equal(k1.cpu().numpy().astype(np.int32), want)
    del logp
    try:
        for mode in (2, 0, 1):
            lib.mas_b200_debug_force_unfused(mode)
            path, dur = pkg.fused_maximum_path(*args)
            assert torch.equal(path, k1), mode
            assert torch.equal(dur, k1_dur), mode
            del path
    finally:
        lib.mas_b200_debug_force_unfused(0)


@pytest.mark.parametrize("shape", [(6, 80, 200, 1000), (3, 80, 70, 300), (150, 80, 40, 96)])
def test_fused_non_finite_scores_follow_the_reference_compare(pkg, oracle, shape, fused_mode):
    """NaN / inf in z make scores non-finite; the sign trick of the fast sweep is not the reference's
    compare there (core.c:2697-2708), so the single launch recomputes such an utterance literally.
    Whatever path it takes must equal the oracle's on the same (non-finite) scores."""
    B, D, T_x, T_y = shape
    rng = np.random.default_rng(zlib.crc32(repr(shape).encode()))
    t_x, t_y = ragged_lengths(rng, B, T_x, T_y)
    x_m, x_logs, z = synth_prior(rng, B, D, T_x, T_y, t_x, t_y, False)
    for b in range(0, B, 2):                      # every other utterance gets a few poisoned frames
        ys = rng.integers(0, t_y[b], 3)
        z[b, rng.integers(0, D), ys[0]] = np.nan
        z[b, rng.integers(0, D), ys[1]] = np.inf
        z[b, :, ys[2]] = -np.inf
    args = (to_dev(x_m), to_dev(x_logs), to_dev(z), torch.from_numpy(t_x), torch.from_numpy(t_y))
    path, dur = pkg.fused_maximum_path(*args)
    logp = pkg.log_likelihood_matrix(*args[:3]).cpu().numpy()
    want = oracle.maximum_path(logp, t_x, t_y)
    assert np.array_equal(path.cpu().numpy().astype(np.int32), want)
    assert np.array_equal(dur.cpu().numpy(), want.sum(-1))
