"""Regenerate tests/golden/*.npz by running the REFERENCE itself (build container only).

    python tests/golden/make_golden.py

Needs /root/reference (read-only) and oracle/_ref (``python oracle/build_ref.py``).  Nothing from
the reference is copied: its package is imported from where it lies, with

  * ``glow_tts_train.monotonic_align`` given a second search path (oracle/_ref/serial) so that its
    ``from .core import maximum_path_c`` (monotonic_align/__init__.py:3) finds the compiled kernel,
  * a 3-line ``dataclasses_json`` stub (not installed here; config.py:8 only needs the mixin name).

Outputs (committed, small):
  mas_kat.npz            known-answer vectors for maximum_path produced by the reference's
                         monotonic_align.maximum_path (Cython kernel + its own wrapper)
  model_meanonly.npz     FlowGenerator.forward on CPU, default config (mean_only=True): the tensors
  model_general.npz      entering/leaving models.py:362-382 (x_m, x_logs, z, logp, attn_mask, attn),
                         mean_only=False for the general case
"""
from __future__ import annotations

import importlib.util
import sys
import types
from pathlib import Path

import numpy as np
import torch

HERE = Path(__file__).resolve().parent
REPO = HERE.parent.parent
REF = Path("/root/reference")
sys.path.insert(0, str(REPO))

from oracle import build_ref  # noqa: E402


def import_reference():
    build_ref.build_reference()
    so_dir = build_ref.ref_so("serial").parent
    stub = types.ModuleType("dataclasses_json")

    class DataClassJsonMixin:  # config.py:8
        pass

    stub.DataClassJsonMixin = DataClassJsonMixin
    sys.modules["dataclasses_json"] = stub

    pkg_dir = REF / "glow_tts_train"
    spec = importlib.util.spec_from_file_location(
        "glow_tts_train", pkg_dir / "__init__.py", submodule_search_locations=[str(pkg_dir)])
    pkg = importlib.util.module_from_spec(spec)
    sys.modules["glow_tts_train"] = pkg
    spec.loader.exec_module(pkg)

    ma_dir = pkg_dir / "monotonic_align"
    spec = importlib.util.spec_from_file_location(
        "glow_tts_train.monotonic_align", ma_dir / "__init__.py",
        submodule_search_locations=[str(ma_dir), str(so_dir)])
    ma = importlib.util.module_from_spec(spec)
    sys.modules["glow_tts_train.monotonic_align"] = ma
    spec.loader.exec_module(ma)
    pkg.monotonic_align = ma
    return pkg, ma


def prefix_mask(t_xs, t_ys, T_x, T_y):
    """attn_mask exactly as models.py:334-337 builds it (outer product of two prefix masks)."""
    x_mask = (torch.arange(T_x)[None, :] < torch.as_tensor(t_xs)[:, None]).float()
    y_mask = (torch.arange(T_y)[None, :] < torch.as_tensor(t_ys)[:, None]).float()
    return x_mask[:, :, None] * y_mask[:, None, :]


def make_mas_kat(ma):
    rng = np.random.default_rng(1234)
    cases = []

    def add(name, value, t_xs, t_ys):
        value = np.asarray(value, np.float32)
        B, T_x, T_y = value.shape
        mask = prefix_mask(t_xs, t_ys, T_x, T_y)
        path = ma.maximum_path(torch.from_numpy(value), mask).numpy()
        cases.append((name, value, np.asarray(t_xs, np.int32), np.asarray(t_ys, np.int32),
                      path.astype(np.int8)))

    add("zeros_full", np.zeros((1, 5, 12)), [5], [12])
    add("zeros_ragged", np.zeros((1, 5, 12)), [3], [7])                      # row 2
    add("single_token", 5 * rng.standard_normal((2, 4, 9)) - 50, [1, 1], [9, 1])
    add("square", 5 * rng.standard_normal((2, 6, 6)) - 50, [6, 4], [6, 4])   # t_x == t_y: forced diagonal
    add("one_frame", rng.standard_normal((1, 1, 1)), [1], [1])
    for i in range(6):
        T_x = int(rng.integers(2, 24))
        T_y = int(rng.integers(T_x, 70))
        B = 3
        t_xs = rng.integers(1, T_x + 1, B)
        t_xs[0] = T_x
        t_ys = np.array([rng.integers(t, T_y + 1) for t in t_xs])
        t_ys[0] = T_y
        add(f"randn_{i}", 10 * rng.standard_normal((B, T_x, T_y)) - 100, t_xs, t_ys)
        add(f"int_ties_{i}", -rng.integers(0, 3, (B, T_x, T_y)).astype(np.float32), t_xs, t_ys)
    # scores so negative that -1e9 stops acting as -infinity
    add("below_neg", (-4e8 * rng.random((2, 5, 14))).astype(np.float32), [5, 3], [14, 9])
    # a 40x130 case that crosses a 32-frame word and a 32-token warp boundary of the CUDA kernel
    add("wide", 10 * rng.standard_normal((2, 40, 130)) - 100, [40, 33], [130, 97])

    out = {"names": np.array([c[0] for c in cases])}
    for i, (_, value, t_xs, t_ys, path) in enumerate(cases):
        out[f"value_{i}"] = value
        out[f"t_x_{i}"] = t_xs
        out[f"t_y_{i}"] = t_ys
        out[f"path_{i}"] = path
    np.savez_compressed(HERE / "mas_kat.npz", **out)
    print("mas_kat.npz:", len(cases), "cases")


def make_model_golden(pkg, ma, mean_only: bool, name: str):
    from glow_tts_train import models
    from glow_tts_train.config import TrainingConfig

    torch.manual_seed(1234)
    config = TrainingConfig()
    config.model.num_symbols = 40
    config.model.mean_only = mean_only
    # a small but structurally complete network: the hot path only sees 80-channel outputs
    config.model.hidden_channels = 32
    config.model.hidden_channels_enc = 32
    config.model.hidden_channels_dec = 32
    config.model.filter_channels = 64
    config.model.filter_channels_dp = 32
    config.model.n_blocks_dec = 2
    config.model.n_layers_enc = 2
    model, _ = models.setup_model(config, create_optimizer=False, use_cuda=False)
    model.eval()

    B, T_x, T_y = 3, 14, 52
    x_len = torch.tensor([14, 11, 6])
    y_len = torch.tensor([52, 41, 25])
    x = torch.randint(1, 40, (B, T_x))
    for b in range(B):
        x[b, x_len[b]:] = 0
    y = torch.randn(B, 80, T_y)

    captured = {}
    real = ma.maximum_path

    def recorder(value, mask):
        captured["logp"] = value.detach().clone()
        captured["mask"] = mask.detach().clone()
        out = real(value, mask)
        captured["path"] = out.detach().clone()
        return out

    models.monotonic_align.maximum_path = recorder
    enc_out = {}
    hook = model.encoder.register_forward_hook(lambda m, i, o: enc_out.update(x_m=o[0], x_logs=o[1]))
    dec_out = {}
    hook2 = model.decoder.register_forward_hook(lambda m, i, o: dec_out.update(z=o[0]))
    with torch.no_grad():
        (z, z_m, z_logs, logdet, z_mask), (x_m, x_logs, x_mask), (attn, logw, logw_) = model(
            x, x_len, y, y_len)
    hook.remove()
    hook2.remove()
    models.monotonic_align.maximum_path = real

    np.savez_compressed(
        HERE / name,
        x_m=enc_out["x_m"].numpy(), x_logs=enc_out["x_logs"].numpy(), z=dec_out["z"].numpy(),
        logp=captured["logp"].numpy(), attn_mask=captured["mask"].numpy(),
        path=captured["path"].numpy().astype(np.int8),
        x_len=x_len.numpy().astype(np.int32),
        y_len=((y_len // 2) * 2).numpy().astype(np.int32),   # models.py:405 floors to n_sqz
        logw_=logw_.numpy(), z_m=z_m.numpy(), z_logs=z_logs.numpy(),
        mean_only=np.array(mean_only),
    )
    print(name, "logp", tuple(captured["logp"].shape), "durations", captured["path"].sum(-1)[0][:6].tolist())


if __name__ == "__main__":
    if not REF.exists():
        sys.exit("needs /root/reference (build container)")
    pkg, ma = import_reference()
    make_mas_kat(ma)
    make_model_golden(pkg, ma, True, "model_meanonly.npz")
    make_model_golden(pkg, ma, False, "model_general.npz")
