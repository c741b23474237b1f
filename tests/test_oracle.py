"""The oracle against everything that pins it: the golden vectors the reference produced, the
reference's own compiled Cython kernel (when oracle/_ref exists), an independent numpy
restatement, and the structural properties of a monotonic alignment."""
from __future__ import annotations

import numpy as np
import pytest
from hypothesis import given, settings, strategies as st

from conftest import prefix_mask, ragged_lengths


def test_golden_kat(oracle, mas_kat):
    assert len(mas_kat) >= 15
    for name, value, t_x, t_y, want in mas_kat:
        got = oracle.maximum_path(value, t_x, t_y)
        assert np.array_equal(got, want), name
        got_np = oracle.maximum_path_numpy(value, t_x, t_y)
        assert np.array_equal(got_np, want), name + " (numpy restatement)"


def test_survey_known_answers(oracle):
    # all-ties input, full and padded
    v = np.zeros((1, 5, 12), np.float32)
    assert oracle.maximum_path(v, [5], [12]).sum(-1).tolist() == [[1, 1, 1, 1, 8]]
    p = oracle.maximum_path(v, [3], [7])
    assert p.sum(-1).tolist() == [[1, 1, 5, 0, 0]]
    assert p[:, 3:, :].sum() == 0 and p[:, :, 7:].sum() == 0


@pytest.mark.parametrize("flavour", ["serial", "omp"])
def test_against_reference_cython(oracle, flavour):
    core = oracle.reference_core(flavour)
    if core is None:
        pytest.skip("oracle/_ref not built (no /root/reference at build time)")
    rng = np.random.default_rng(7)
    for it in range(120):
        B = 4
        T_x = int(rng.integers(1, 40))
        T_y = int(rng.integers(T_x, 140))
        t_x = rng.integers(1, T_x + 1, B).astype(np.int32)
        t_y = np.array([rng.integers(t, T_y + 1) for t in t_x], np.int32)
        kind = it % 3
        if kind == 0:
            value = (10 * rng.standard_normal((B, T_x, T_y)) - 100).astype(np.float32)
        elif kind == 1:
            value = -rng.integers(0, 3, (B, T_x, T_y)).astype(np.float32)
        else:
            value = (-3e8 * rng.random((B, T_x, T_y))).astype(np.float32)   # drives scores below -1e9
        want = oracle.maximum_path(value, t_x, t_y, flavour=flavour)
        assert np.array_equal(oracle.maximum_path(value, t_x, t_y), want)
        assert np.array_equal(oracle.maximum_path(value, t_x, t_y, threads=4), want)


def test_reference_boundary_matches_kernel(oracle):
    """The restated marshalling (__init__.py:6-21) around the C oracle == around the Cython."""
    import torch

    rng = np.random.default_rng(3)
    B, T_x, T_y = 3, 17, 60
    t_x, t_y = ragged_lengths(rng, B, T_x, T_y)
    value = torch.from_numpy((10 * rng.standard_normal((B, T_x, T_y)) - 100).astype(np.float32))
    mask = torch.from_numpy(prefix_mask(t_x, t_y, T_x, T_y))
    a = oracle.reference_boundary(value, mask)
    assert a.dtype == value.dtype and a.shape == value.shape
    assert np.array_equal(a.numpy().astype(np.int32), oracle.maximum_path(value.numpy(), t_x, t_y))
    core = oracle.reference_core("serial")
    if core is not None:
        b = oracle.reference_boundary(value, mask, kernel=core.maximum_path_c)
        assert torch.equal(a, b)
    # half precision in -> half precision out (dtype contract of __init__.py:13,21)
    h = oracle.reference_boundary(value.half(), mask.half())
    assert h.dtype == torch.float16


def test_lengths_from_mask(oracle):
    rng = np.random.default_rng(5)
    t_x, t_y = ragged_lengths(rng, 6, 23, 90)
    a, b = oracle.lengths_from_mask(prefix_mask(t_x, t_y, 23, 90))
    assert np.array_equal(a, t_x) and np.array_equal(b, t_y)


def _check_alignment_properties(path, t_x, t_y):
    B, T_x, T_y = path.shape
    for b in range(B):
        p = path[b]
        tx, ty = int(t_x[b]), int(t_y[b])
        assert p[tx:, :].sum() == 0 and p[:, ty:].sum() == 0
        assert (p[:, :ty].sum(0) == 1).all()                    # one token per valid frame
        rows = p[:, :ty].argmax(0)
        assert rows[0] == 0 and rows[-1] == tx - 1              # starts on token 0, ends on the last
        steps = np.diff(rows)
        assert ((steps == 0) | (steps == 1)).all()              # monotone, never skips a token


@settings(max_examples=60, deadline=None)
@given(st.integers(1, 24), st.integers(0, 60), st.integers(0, 2**31 - 1), st.booleans())
def test_alignment_properties(oracle, T_x, extra, seed, ties):
    rng = np.random.default_rng(seed)
    T_y = T_x + extra
    B = 3
    t_x = rng.integers(1, T_x + 1, B).astype(np.int32)
    t_y = np.array([rng.integers(t, T_y + 1) for t in t_x], np.int32)
    if ties:
        value = -rng.integers(0, 2, (B, T_x, T_y)).astype(np.float32)
    else:
        value = (5 * rng.standard_normal((B, T_x, T_y)) - 50).astype(np.float32)
    path = oracle.maximum_path(value, t_x, t_y)
    _check_alignment_properties(path, t_x, t_y)
    assert np.array_equal(path, oracle.maximum_path_numpy(value, t_x, t_y))


def test_optimality_small(oracle):
    """On tiny cases the path's score equals the brute-force maximum over all monotone paths."""
    from itertools import combinations

    rng = np.random.default_rng(11)
    for _ in range(20):
        t_x, t_y = int(rng.integers(1, 5)), int(rng.integers(5, 9))
        value = rng.standard_normal((1, t_x, t_y)).astype(np.float64).astype(np.float32)
        path = oracle.maximum_path(value, [t_x], [t_y])[0]
        got = float((value[0].astype(np.float64) * path).sum())
        best = -np.inf
        for cuts in combinations(range(1, t_y), t_x - 1):   # frames where the token index advances
            rows = np.zeros(t_y, int)
            for c in cuts:
                rows[c:] += 1
            best = max(best, float(value[0].astype(np.float64)[rows, np.arange(t_y)].sum()))
        assert got >= best - 1e-4


def test_logp_against_reference_model(oracle, model_golden):
    g = model_golden
    want = g["logp"].astype(np.float64)
    l64 = oracle.logp_f64(g["x_m"], g["x_logs"], g["z"])
    l32 = oracle.logp_f32(g["x_m"], g["x_logs"], g["z"])
    # tolerance of north_star: 1e-5 relative.  The reference's own fp32 matmul sits ~2e-7 from fp64.
    assert np.max(np.abs(want - l64) / np.abs(l64)) < 1e-5
    assert np.max(np.abs(l32 - l64) / np.abs(l64)) < 1e-5
    if bool(g["mean_only"]):
        assert np.array_equal(oracle.logp_f64(g["x_m"], None, g["z"]), l64)
    t_x, t_y = oracle.lengths_from_mask(g["attn_mask"])
    assert np.array_equal(t_x, g["x_len"]) and np.array_equal(t_y, g["y_len"])
    path = oracle.maximum_path(g["logp"] * g["attn_mask"], t_x, t_y)
    assert np.array_equal(path, g["path"].astype(np.int32))
    # the durations the model derives from it (models.py:393)
    logw = np.log(1e-8 + path.sum(-1)) * (np.arange(path.shape[1])[None] < t_x[:, None])
    assert np.allclose(logw, g["logw_"][:, 0], atol=1e-6)
