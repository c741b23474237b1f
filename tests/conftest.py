"""Test configuration.  ``-m "not gpu"`` runs on the CPU-only build box (oracle vs golden vectors,
host logic, ABI surface); ``-m gpu`` are the parity tests proper and call the sm_100a kernels
through the C ABI on a B200."""
from __future__ import annotations

import sys
from pathlib import Path

import numpy as np
import pytest

REPO = Path(__file__).resolve().parent.parent
if str(REPO) not in sys.path:
    sys.path.insert(0, str(REPO))

import __graft_entry__ as entry  # noqa: E402

GOLDEN = Path(__file__).resolve().parent / "golden"


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")


def pytest_collection_modifyitems(config, items):
    try:
        import torch

        have_cuda = torch.cuda.is_available()
    except Exception:  # pragma: no cover
        have_cuda = False
    if have_cuda:
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def oracle():
    mod = entry.load_oracle()
    mod.build()
    return mod


@pytest.fixture(scope="session")
def pkg():
    builder = entry._load_file(entry.PKG_NAME + "_build", entry.PKG_DIR / "build.py")
    builder.build()
    return entry.load_package()


@pytest.fixture(scope="session")
def mas_kat():
    data = np.load(GOLDEN / "mas_kat.npz")
    cases = []
    for i, name in enumerate(data["names"]):
        cases.append((str(name), data[f"value_{i}"], data[f"t_x_{i}"], data[f"t_y_{i}"],
                      data[f"path_{i}"].astype(np.int32)))
    return cases


@pytest.fixture(scope="session", params=["model_meanonly", "model_general"])
def model_golden(request):
    return dict(np.load(GOLDEN / f"{request.param}.npz"))


def prefix_mask(t_xs, t_ys, T_x, T_y):
    """attn_mask as models.py:334-337 builds it."""
    xm = (np.arange(T_x)[None, :] < np.asarray(t_xs)[:, None]).astype(np.float32)
    ym = (np.arange(T_y)[None, :] < np.asarray(t_ys)[:, None]).astype(np.float32)
    return xm[:, :, None] * ym[:, None, :]


def ragged_lengths(rng, B, T_x, T_y):
    """LJSpeech-like ragged lengths: t_x ~ U[T_x/2, T_x], t_y ~ t_x * ratio * U[0.8,1.2],
    even, clipped to [t_x, T_y], sorted by t_x descending, element 0 full-size."""
    t_x = rng.integers(max(1, T_x // 2), T_x + 1, B)
    t_x[0] = T_x
    t_x = np.sort(t_x)[::-1].copy()
    ratio = T_y / T_x
    t_y = np.round(t_x * ratio * rng.uniform(0.8, 1.2, B)).astype(np.int64)
    t_y = (t_y // 2) * 2
    t_y = np.clip(t_y, t_x + (t_x % 2), T_y)
    t_y[0] = T_y
    return t_x.astype(np.int32), t_y.astype(np.int32)
