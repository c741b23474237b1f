"""oracle -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

CPU restatement of the reference's alignment hot path (``monotonic_align.maximum_path`` and the
log-likelihood matrix of ``FlowGenerator.forward``).  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may
import this package; the product (``glow-tts-train_b200/``) never does.

Parity status: PINNED against the reference's own compiled Cython kernel (``oracle/_ref``, built
by ``oracle/build_ref.py`` from ``/root/reference`` where it lies) and against the golden vectors
under ``tests/golden/`` that the reference produced (``tests/golden/make_golden.py``).

Layers:
  * ``mas_oracle.c`` via ctypes  -> :func:`maximum_path_c`, :func:`logp_f64`, :func:`logp_f32`
  * numpy restatements           -> :func:`maximum_path_numpy` (tiny cases; independent of the C)
  * the reference's kernel       -> :func:`reference_core` (``oracle/_ref/{serial,omp}``)
  * the reference's marshalling  -> :func:`reference_boundary`
    (restates ``glow_tts_train/monotonic_align/__init__.py:6-21`` around either kernel)
"""
from __future__ import annotations

import ctypes
import importlib.util
import os
from pathlib import Path

import numpy as np

from . import build_ref

MAX_NEG_VAL = -1e9  # core.pyx:40 default

_lib = None
_ref_cores: dict = {}


def build(force: bool = False) -> dict:
    """Compile the C oracle (always) and the reference kernel (when /root/reference exists)."""
    global _lib
    info = build_ref.build_all(force)
    if force:
        _lib = None
    return info


def _c() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        path = build_ref.build_c_oracle()
        lib = ctypes.CDLL(str(path))
        f32p = ctypes.POINTER(ctypes.c_float)
        i32p = ctypes.POINTER(ctypes.c_int32)
        f64p = ctypes.POINTER(ctypes.c_double)
        lib.mas_oracle_batch.argtypes = [i32p, f32p, i32p, i32p, ctypes.c_int, ctypes.c_int,
                                         ctypes.c_int, ctypes.c_float, ctypes.c_int]
        lib.mas_oracle_batch.restype = None
        lib.mas_oracle_lengths.argtypes = [f32p, ctypes.c_int, ctypes.c_int, ctypes.c_int, i32p, i32p]
        lib.mas_oracle_lengths.restype = None
        lib.mas_oracle_logp_f64.argtypes = [f32p, f32p, f32p, f64p] + [ctypes.c_int] * 4
        lib.mas_oracle_logp_f64.restype = ctypes.c_int
        lib.mas_oracle_logp_f32.argtypes = [f32p, f32p, f32p, f32p] + [ctypes.c_int] * 4
        lib.mas_oracle_logp_f32.restype = ctypes.c_int
        lib.mas_oracle_has_openmp.restype = ctypes.c_int
        _lib = lib
    return _lib


def _ptr(a: np.ndarray, ctype):
    return a.ctypes.data_as(ctypes.POINTER(ctype))


def host_threads() -> int:
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:  # pragma: no cover
        return os.cpu_count() or 1


# --------------------------------------------------------------------------------------------
# DP + backtrack
# --------------------------------------------------------------------------------------------
def maximum_path_c(paths: np.ndarray, values: np.ndarray, t_xs: np.ndarray, t_ys: np.ndarray,
                   max_neg_val: float = MAX_NEG_VAL, threads: int = 1) -> None:
    """Same contract as the reference's ``maximum_path_c`` (core.pyx:40-45): ``paths`` int32
    [B,T_x,T_y] pre-zeroed, ``values`` fp32 [B,T_x,T_y] CLOBBERED into cumulative scores,
    ``t_xs``/``t_ys`` int32 [B].  ``threads>1`` = OpenMP over utterances."""
    for a, dt in ((paths, np.int32), (values, np.float32), (t_xs, np.int32), (t_ys, np.int32)):
        if a.dtype != dt or not a.flags.c_contiguous:
            raise ValueError("buffer dtype/contiguity mismatch (the reference's memoryviews raise too)")
    B, T_x, T_y = values.shape
    _c().mas_oracle_batch(_ptr(paths, ctypes.c_int32), _ptr(values, ctypes.c_float),
                          _ptr(t_xs, ctypes.c_int32), _ptr(t_ys, ctypes.c_int32),
                          B, T_x, T_y, ctypes.c_float(max_neg_val), int(threads))


def lengths_from_mask(mask: np.ndarray):
    """t_x, t_y as the reference wrapper derives them (__init__.py:18-19)."""
    mask = np.ascontiguousarray(mask, dtype=np.float32)
    B, T_x, T_y = mask.shape
    t_xs = np.empty(B, np.int32)
    t_ys = np.empty(B, np.int32)
    _c().mas_oracle_lengths(_ptr(mask, ctypes.c_float), B, T_x, T_y,
                            _ptr(t_xs, ctypes.c_int32), _ptr(t_ys, ctypes.c_int32))
    return t_xs, t_ys


def maximum_path_numpy(value: np.ndarray, t_xs, t_ys, max_neg_val: float = MAX_NEG_VAL) -> np.ndarray:
    """Pure-numpy/Python restatement of core.pyx:9-35 for SMALL cases -- written independently of
    mas_oracle.c (rolling score column + explicit move table) so the two can check each other."""
    value = np.asarray(value, dtype=np.float32)
    B, T_x, T_y = value.shape
    neg = np.float32(max_neg_val)
    out = np.zeros((B, T_x, T_y), np.int32)
    for b in range(B):
        t_x, t_y = int(t_xs[b]), int(t_ys[b])
        score = np.full(t_x, neg, np.float32)          # scores of frame-1, per token
        moved = np.zeros((t_x, t_y), bool)             # True: best predecessor is token-1
        for frame in range(t_y):
            new = score.copy()
            lo, hi = max(0, t_x + frame - t_y), min(t_x, frame + 1)
            for tok in range(lo, hi):
                stay = neg if tok == frame else score[tok]
                if tok == 0:
                    adv = np.float32(0.0) if frame == 0 else neg
                else:
                    adv = score[tok - 1]
                take_adv = bool(adv > stay)
                moved[tok, frame] = take_adv
                new[tok] = np.float32((adv if take_adv else stay) + value[b, tok, frame])
            score = new
        tok = t_x - 1
        for frame in range(t_y - 1, -1, -1):
            out[b, tok, frame] = 1
            if tok != 0 and (tok == frame or moved[tok, frame]):
                tok -= 1
    return out


# --------------------------------------------------------------------------------------------
# The reference's own compiled kernel + its marshalling
# --------------------------------------------------------------------------------------------
def reference_core(flavour: str = "serial"):
    """The reference's compiled ``core`` module (``maximum_path_c``) from oracle/_ref, or None when
    it was never built (no /root/reference at build time)."""
    if flavour not in _ref_cores:
        build_ref.build_reference()
        so = build_ref.ref_so(flavour)
        if not so.exists():
            _ref_cores[flavour] = None
        else:
            spec = importlib.util.spec_from_file_location("core", str(so))
            mod = importlib.util.module_from_spec(spec)
            spec.loader.exec_module(mod)
            _ref_cores[flavour] = mod
    return _ref_cores[flavour]


def reference_boundary(value, mask, kernel=None):
    """What ``monotonic_align.maximum_path(value, mask)`` does in the reference
    (glow_tts_train/monotonic_align/__init__.py:6-21), step for step, around ``kernel``
    (default: the C oracle; pass ``reference_core(...).maximum_path_c`` for the real thing).
    torch tensors in, torch tensor out (same dtype/device as ``value``)."""
    import torch

    if kernel is None:
        kernel = maximum_path_c
    masked = value * mask                                              # __init__.py:11
    device, dtype = masked.device, masked.dtype                        # :12-13
    scores = masked.detach().cpu().numpy().astype(np.float32)          # :14
    path = np.zeros_like(scores).astype(np.int32)                      # :15
    mask_np = mask.detach().cpu().numpy()                              # :16
    t_x_max = mask_np.sum(1)[:, 0].astype(np.int32)                    # :18
    t_y_max = mask_np.sum(2)[:, 0].astype(np.int32)                    # :19
    kernel(path, scores, t_x_max, t_y_max)                             # :20
    return torch.from_numpy(path).to(device=device, dtype=dtype)       # :21


def maximum_path(value: np.ndarray, t_xs, t_ys, threads: int = 1, flavour: str | None = None) -> np.ndarray:
    """Convenience: dense int32 path for fp32 ``value`` [B,T_x,T_y] without clobbering it.
    ``flavour`` None -> C oracle; 'serial' / 'omp' -> the reference's Cython build."""
    scores = np.array(value, dtype=np.float32, order="C", copy=True)
    path = np.zeros(scores.shape, np.int32)
    t_xs = np.ascontiguousarray(t_xs, dtype=np.int32)
    t_ys = np.ascontiguousarray(t_ys, dtype=np.int32)
    if flavour is None:
        maximum_path_c(path, scores, t_xs, t_ys, threads=threads)
    else:
        core = reference_core(flavour)
        if core is None:
            raise RuntimeError(f"oracle/_ref/{flavour} was not built")
        core.maximum_path_c(path, scores, t_xs, t_ys)
    return path


# --------------------------------------------------------------------------------------------
# log-likelihood matrix (models.py:363-376)
# --------------------------------------------------------------------------------------------
def _logp(fn, out_dtype, x_m, x_logs, z):
    x_m = np.ascontiguousarray(x_m, np.float32)
    z = np.ascontiguousarray(z, np.float32)
    B, D, T_x = x_m.shape
    T_y = z.shape[2]
    if x_logs is not None:
        x_logs = np.ascontiguousarray(x_logs, np.float32)
    out = np.empty((B, T_x, T_y), out_dtype)
    ctype = ctypes.c_double if out_dtype == np.float64 else ctypes.c_float
    rc = fn(_ptr(x_m, ctypes.c_float),
            _ptr(x_logs, ctypes.c_float) if x_logs is not None else None,
            _ptr(z, ctypes.c_float), _ptr(out, ctype), B, D, T_x, T_y)
    if rc != 0:
        raise ValueError("mas_oracle_logp: unsupported channel count")
    return out


def logp_f64(x_m, x_logs, z) -> np.ndarray:
    """[B,T_x,T_y] float64 log-likelihood matrix, models.py:363-376 evaluated in fp64."""
    return _logp(_c().mas_oracle_logp_f64, np.float64, x_m, x_logs, z)


def logp_f32(x_m, x_logs, z) -> np.ndarray:
    """Same in fp32 arithmetic, ascending-channel contraction, reference term order."""
    return _logp(_c().mas_oracle_logp_f32, np.float32, x_m, x_logs, z)


def durations_from_path(path: np.ndarray) -> np.ndarray:
    """Per-token frame counts: row sums of the path (models.py:393 takes log(1e-8 + this))."""
    return np.asarray(path).sum(-1).astype(np.int32)


def logp_torch(x_m, x_logs, z):
    """The reference's own tensor program for the log-likelihood matrix, models.py:363-376, term by
    term with torch ops (on whatever device the inputs live on; the CPU baseline runs it on CPU)."""
    import math

    import torch

    with torch.no_grad():                                                         # models.py:362
        x_s_sq_r = torch.exp(-2 * x_logs)                                         # :363
        logp1 = torch.sum(-0.5 * math.log(2 * math.pi) - x_logs, [1]).unsqueeze(-1)   # :364-366
        logp2 = torch.matmul(x_s_sq_r.transpose(1, 2), -0.5 * (z ** 2))           # :367-369
        logp3 = torch.matmul((x_m * x_s_sq_r).transpose(1, 2), z)                 # :370-372
        logp4 = torch.sum(-0.5 * (x_m ** 2) * x_s_sq_r, [1]).unsqueeze(-1)        # :373-375
        return logp1 + logp2 + logp3 + logp4                                      # :376


def reference_step(x_m, x_logs, z, attn_mask, kernel=None):
    """models.py:362-382 as the reference runs it: logp_torch -> maximum_path(logp, mask)."""
    return reference_boundary(logp_torch(x_m, x_logs, z), attn_mask, kernel=kernel)
