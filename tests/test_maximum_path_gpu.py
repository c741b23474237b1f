"""Parity of kernel (1) with the oracle -- bit-exact -- through the C ABI on a B200."""
from __future__ import annotations

import ctypes

import zlib

import numpy as np
import pytest
import torch

from conftest import prefix_mask, ragged_lengths

pytestmark = pytest.mark.gpu
DEV = "cuda:0"


def run_gpu(pkg, value, t_x, t_y, via_mask=True, **kw):
    v = torch.from_numpy(np.ascontiguousarray(value, np.float32)).to(DEV)
    B, T_x, T_y = v.shape
    if via_mask:
        mask = torch.from_numpy(prefix_mask(t_x, t_y, T_x, T_y)).to(DEV)
        out = pkg.maximum_path_from_lengths(v, mask=mask, **kw)
    else:
        out = pkg.maximum_path_from_lengths(v, torch.as_tensor(t_x, dtype=torch.int32, device=DEV),
                                            torch.as_tensor(t_y, dtype=torch.int32, device=DEV), **kw)
    return out


def as_i32(t):
    return t.float().cpu().numpy().astype(np.int32)


def test_golden_kat(pkg, mas_kat):
    for name, value, t_x, t_y, want in mas_kat:
        for via_mask in (True, False):
            got = as_i32(run_gpu(pkg, value, t_x, t_y, via_mask))
            assert np.array_equal(got, want), (name, via_mask)


def test_drop_in_signature_and_contract(pkg, oracle):
    rng = np.random.default_rng(0)
    B, T_x, T_y = 5, 33, 140
    t_x, t_y = ragged_lengths(rng, B, T_x, T_y)
    value = torch.from_numpy((10 * rng.standard_normal((B, T_x, T_y)) - 100).astype(np.float32)).to(DEV)
    keep = value.clone()
    # the mask exactly as models.py:334-337 builds it and :379 passes it: a squeezed [b,1,t,t'] view
    x_mask = torch.from_numpy((np.arange(T_x)[None] < t_x[:, None]).astype(np.float32)).to(DEV).unsqueeze(1)
    z_mask = torch.from_numpy((np.arange(T_y)[None] < t_y[:, None]).astype(np.float32)).to(DEV).unsqueeze(1)
    attn_mask = torch.unsqueeze(x_mask, -1) * torch.unsqueeze(z_mask, 2)
    path = pkg.monotonic_align.maximum_path(value, attn_mask.squeeze(1))
    assert path.shape == value.shape and path.dtype == value.dtype and path.device == value.device
    assert torch.equal(value, keep), "value must not be clobbered"
    want = oracle.maximum_path(keep.cpu().numpy(), t_x, t_y)
    assert np.array_equal(as_i32(path), want)
    # the reference wrapper itself, end to end on the same tensors
    ref = oracle.reference_boundary(value, attn_mask.squeeze(1))
    assert torch.equal(ref, path)
    # a broadcast (stride-0) mask view works too
    bmask = x_mask.squeeze(1)[:, :, None].expand(B, T_x, T_y) * z_mask.squeeze(1)[:, None, :]
    assert torch.equal(pkg.monotonic_align.maximum_path(value, bmask), path)
    # other dtypes: computed in fp32 like the reference (.astype(np.float32)), returned in value.dtype
    for dt in (torch.float16, torch.bfloat16, torch.float64):
        v = value.to(dt)
        got = pkg.monotonic_align.maximum_path(v, attn_mask.squeeze(1).to(dt))
        assert got.dtype == dt
        want_dt = oracle.maximum_path(v.float().cpu().numpy(), t_x, t_y)
        assert np.array_equal(as_i32(got), want_dt)
    # non-prefix mask: the reference's value * mask semantics, by default (verified on the device)
    holes = attn_mask.squeeze(1).clone()
    holes[:, 3:9, 10:30] = 0
    holes[:, 0, :] = attn_mask.squeeze(1)[:, 0, :]
    holes[:, :, 0] = attn_mask.squeeze(1)[:, :, 0]
    got = pkg.monotonic_align.maximum_path(value, holes)
    assert torch.equal(got, oracle.reference_boundary(value, holes))


@pytest.mark.parametrize("shape", [(6, 40, 160), (4, 37, 150), (3, 200, 1000)])
def test_any_mask_matches_the_reference_wrapper(pkg, oracle, shape):
    """monotonic_align/__init__.py:11,18-19 for masks that are NOT the prefix masks of models.py:334-337:
    interior zeros, fractional entries, a hole in column 0 (shortens t_x), mixed with clean
    utterances in the same batch -- the device-side check flags exactly the utterances whose
    value * mask differs, and the result equals the reference wrapper run on the same tensors.
    Shapes cover the TMA kernel (T_y % 4 == 0) and the generic one (150 frames)."""
    B, T_x, T_y = shape
    rng = np.random.default_rng(zlib.crc32(repr(shape).encode()))
    t_x, t_y = ragged_lengths(rng, B, T_x, T_y)
    value = torch.from_numpy((10 * rng.standard_normal(shape) - 100).astype(np.float32)).to(DEV)
    mask = torch.from_numpy(prefix_mask(t_x, t_y, T_x, T_y)).to(DEV)
    m = mask.clone()
    m[0, 2:5, 7:20] = 0                                  # interior zeros
    if B > 1:
        m[1, 1:, 3:9] = 0.5                              # fractional entries (scores are halved there)
    if B > 2:
        m[2, int(t_x[2]) - 1, :] = 0                     # last valid token masked out everywhere: t_x shrinks by one
    # utterances 3.. stay clean prefix masks
    got = pkg.monotonic_align.maximum_path(value, m)
    want = oracle.reference_boundary(value, m)
    assert torch.equal(got, want)
    # and the clean batch still takes the fast path with identical results
    assert torch.equal(pkg.monotonic_align.maximum_path(value, mask), oracle.reference_boundary(value, mask))


def test_cpu_tensors_are_staged_not_computed_on_host(pkg, oracle):
    rng = np.random.default_rng(1)
    B, T_x, T_y = 3, 20, 77
    t_x, t_y = ragged_lengths(rng, B, T_x, T_y)
    value = torch.from_numpy((10 * rng.standard_normal((B, T_x, T_y)) - 100).astype(np.float32)).pin_memory()
    mask = torch.from_numpy(prefix_mask(t_x, t_y, T_x, T_y))
    path = pkg.monotonic_align.maximum_path(value, mask)
    assert not path.is_cuda and path.dtype == torch.float32
    assert np.array_equal(as_i32(path), oracle.maximum_path(value.numpy(), t_x, t_y))


@pytest.mark.parametrize("kind", ["randn", "int_ties", "below_neg", "zeros"])
@pytest.mark.parametrize("shape", [(1, 1, 1), (2, 1, 17), (3, 7, 7), (4, 31, 33), (3, 32, 64), (2, 33, 65),
                                   (5, 64, 200), (3, 100, 101), (2, 129, 515), (2, 200, 1000), (1, 257, 999),
                                   (1, 512, 1301)])
def test_parity_ragged(pkg, oracle, shape, kind):
    B, T_x, T_y = shape
    rng = np.random.default_rng(zlib.crc32(repr((shape, kind)).encode()))
    if kind == "randn":
        value = (10 * rng.standard_normal(shape) - 100).astype(np.float32)
    elif kind == "int_ties":
        value = -rng.integers(0, 3, shape).astype(np.float32)
    elif kind == "below_neg":
        value = (-3e8 * rng.random(shape)).astype(np.float32)
    else:
        value = np.zeros(shape, np.float32)
    t_x = rng.integers(1, T_x + 1, B).astype(np.int32)
    t_y = np.array([rng.integers(t, T_y + 1) for t in t_x], np.int32)
    t_x[0], t_y[0] = T_x, T_y
    want = oracle.maximum_path(value, t_x, t_y)
    path, dur, tok = run_gpu(pkg, value, t_x, t_y, via_mask=bool(B % 2), want_durations=True, want_frame_token=True)
    assert np.array_equal(as_i32(path), want)
    assert np.array_equal(dur.cpu().numpy(), want.sum(-1))
    tok = tok.cpu().numpy()
    for b in range(B):
        assert np.array_equal(tok[b, :t_y[b]], want[b, :, :t_y[b]].argmax(0))
        assert (tok[b, t_y[b]:] == -1).all()


def test_degenerate_lengths(pkg, oracle):
    """Defined behaviour where the reference has none."""
    rng = np.random.default_rng(2)
    B, T_x, T_y = 4, 9, 20
    value = (10 * rng.standard_normal((B, T_x, T_y)) - 100).astype(np.float32)
    t_x = np.array([0, 9, 5, 9], np.int32)
    t_y = np.array([0, 4, 5, 20], np.int32)   # empty; t_x > t_y; square; full
    path = as_i32(run_gpu(pkg, value, t_x, t_y, via_mask=False))
    assert path[0].sum() == 0
    assert np.array_equal(path[1, :4, :4], np.eye(4, dtype=np.int32)) and path[1].sum() == 4
    assert np.array_equal(path[2, :5, :5], np.eye(5, dtype=np.int32)) and path[2].sum() == 5
    assert np.array_equal(path[3], oracle.maximum_path(value[3:4], [9], [20])[0])


def test_nan_scores_follow_the_reference_compare(pkg, oracle):
    rng = np.random.default_rng(4)
    value = (10 * rng.standard_normal((2, 12, 40)) - 100).astype(np.float32)
    value[0, 3, 7] = np.nan
    value[1, 0, 0] = np.nan
    t_x, t_y = np.array([12, 10], np.int32), np.array([40, 33], np.int32)
    want = oracle.maximum_path(value, t_x, t_y)
    assert np.array_equal(as_i32(run_gpu(pkg, value, t_x, t_y)), want)


def test_strided_value_and_stream(pkg, oracle):
    rng = np.random.default_rng(5)
    B, T_x, T_y = 3, 40, 128
    big = torch.from_numpy((10 * rng.standard_normal((B, T_x + 5, T_y + 16)) - 100).astype(np.float32)).to(DEV)
    view = big[:, 2:2 + T_x, 8:8 + T_y]          # token stride != T_y, frame stride 1
    t_x, t_y = ragged_lengths(rng, B, T_x, T_y)
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        path = pkg.maximum_path_from_lengths(view, torch.as_tensor(t_x, device=DEV), torch.as_tensor(t_y, device=DEV))
    s.synchronize()
    assert np.array_equal(as_i32(path), oracle.maximum_path(view.cpu().numpy(), t_x, t_y))


def test_host_entry_matches_reference_layout(pkg, oracle):
    """mas_b200_maximum_path_host_i32 takes exactly maximum_path_c's buffers (core.pyx:40)."""
    lib = pkg._lib.load()
    rng = np.random.default_rng(6)
    B, T_x, T_y = 4, 50, 210
    t_x, t_y = ragged_lengths(rng, B, T_x, T_y)
    values = (10 * rng.standard_normal((B, T_x, T_y)) - 100).astype(np.float32)
    keep = values.copy()
    paths = np.full((B, T_x, T_y), 7, np.int32)
    rc = lib.mas_b200_maximum_path_host_i32(paths.ctypes.data, values.ctypes.data, t_x.ctypes.data, t_y.ctypes.data,
                                            B, T_x, T_y, ctypes.c_float(-1e9), 0)
    assert rc == 0
    assert np.array_equal(values, keep)
    assert np.array_equal(paths, oracle.maximum_path(values, t_x, t_y))


def test_graph_capturable(pkg, oracle):
    rng = np.random.default_rng(8)
    B, T_x, T_y = 2, 24, 90
    t_x, t_y = ragged_lengths(rng, B, T_x, T_y)
    value = torch.from_numpy((10 * rng.standard_normal((B, T_x, T_y)) - 100).astype(np.float32)).to(DEV)
    tx_d, ty_d = torch.as_tensor(t_x, device=DEV), torch.as_tensor(t_y, device=DEV)
    pkg.maximum_path_from_lengths(value, tx_d, ty_d)     # warm-up (allocator, module load)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        path = pkg.maximum_path_from_lengths(value, tx_d, ty_d)
    value.copy_(torch.from_numpy((10 * rng.standard_normal((B, T_x, T_y)) - 100).astype(np.float32)))
    g.replay()
    torch.cuda.synchronize()
    assert np.array_equal(as_i32(path), oracle.maximum_path(value.cpu().numpy(), t_x, t_y))


@pytest.mark.parametrize("shape", [(32, 200, 1000), (8, 400, 2000), (2, 1024, 8192)])
def test_full_size_parity_and_properties(pkg, oracle, shape):
    """BASELINE.json sizes: the C oracle still finishes in seconds, so compare outright, and check
    the size-independent properties (one token per frame, monotone, durations sum to t_y)."""
    B, T_x, T_y = shape
    rng = np.random.default_rng(1234)
    value = (10 * rng.standard_normal(shape) - 100).astype(np.float32)
    t_x, t_y = ragged_lengths(rng, B, T_x, T_y)
    path, dur = run_gpu(pkg, value, t_x, t_y, want_durations=True)
    p = as_i32(path)
    assert np.array_equal(dur.cpu().numpy().sum(-1), t_y)
    for b in range(B):
        rows = p[b, :, :t_y[b]].argmax(0)
        assert (p[b, :, :t_y[b]].sum(0) == 1).all() and rows[0] == 0 and rows[-1] == t_x[b] - 1
        assert set(np.diff(rows).tolist()) <= {0, 1}
        assert p[b, t_x[b]:].sum() == 0 and p[b, :, t_y[b]:].sum() == 0
    want = oracle.maximum_path(value, t_x, t_y, threads=oracle.host_threads())
    assert np.array_equal(p, want)


@pytest.mark.parametrize("K", [1, 2, 4, 8])
@pytest.mark.parametrize("shape", [(3, 64, 256), (4, 200, 1000), (2, 257, 640), (2, 512, 1024), (1, 1024, 2048)])
def test_cluster_sharded_tokens(pkg, oracle, shape, K):
    """Kernel (1) with the utterance's tokens sharded over a thread-block cluster of K CTAs (forced
    through the testing hook): boundary scores, progress and the backtrack hand-over travel over
    distributed shared memory.  Same bits as one CTA, same bits as the oracle."""
    lib = pkg._lib.load()
    B, T_x, T_y = shape
    if T_x // K < 32:
        pytest.skip("fewer than 32 tokens per CTA")
    rng = np.random.default_rng(zlib.crc32(repr((shape, K)).encode()))
    value = (10 * rng.standard_normal(shape) - 100).astype(np.float32)
    t_x, t_y = ragged_lengths(rng, B, T_x, T_y)
    if B > 1:
        t_x[-1], t_y[-1] = max(1, T_x // 8), max(T_x // 8, T_y // 5) // 4 * 4    # leaves upper CTAs without tokens
        value[-1, 3, 17] = np.nan                                                 # and takes the exact-redo path
    want = oracle.maximum_path(value, t_x, t_y)
    lib.mas_b200_debug_force_cluster(K)
    try:
        path, dur, tok = run_gpu(pkg, value, t_x, t_y, via_mask=False, want_durations=True, want_frame_token=True)
        torch.cuda.synchronize()
    finally:
        lib.mas_b200_debug_force_cluster(0)
    assert np.array_equal(as_i32(path), want)
    assert np.array_equal(dur.cpu().numpy(), want.sum(-1))
    tok = tok.cpu().numpy()
    for b in range(B):
        assert np.array_equal(tok[b, :t_y[b]], want[b, :, :t_y[b]].argmax(0))
