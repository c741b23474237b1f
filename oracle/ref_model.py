"""The reference's own model and training step, imported for MODEL-LEVEL checks -- TEST
INFRASTRUCTURE, NOT PRODUCT CODE (used by tests/, bench.py's `c5` block and profiles/ only).

``import_reference()`` returns the reference's ``glow_tts_train`` package (unpacked into a temporary
directory from oracle/_ref/refpkg.zip, where oracle/build_ref.py packs it unmodified in the build
container; the archive travels to the GPU box with the snapshot).  ``swap_monotonic_align(pkg, module)`` is the two-line drop-in INTEGRATION.md describes:
``glow_tts_train.models`` looks ``monotonic_align`` up as a module global at call time
(models.py:9, :379), so replacing that global replaces the path -- nothing else of the reference
changes.  ``synthetic_batch`` makes LJSpeech-shaped inputs for ``FlowGenerator.forward`` /
``train_step`` (train.py:91-162); ``make_model`` builds the reference's model + optimizer through its
own ``setup_model`` (models.py:420-470).
"""
from __future__ import annotations

import importlib
import sys
from pathlib import Path

from . import build_ref

_pkg = None


def import_reference():
    """The reference package, or None when it was never staged (no /root/reference at build time)."""
    global _pkg
    if _pkg is not None:
        return _pkg
    archive = build_ref.stage_reference_package()
    if archive is None:
        return None
    import atexit
    import shutil
    import tempfile
    import zipfile

    root = tempfile.mkdtemp(prefix="mas_refpkg_")   # (the compiled kernel cannot be imported from inside a zip)
    atexit.register(shutil.rmtree, root, ignore_errors=True)
    with zipfile.ZipFile(archive) as z:
        z.extractall(root)
    sys.path.insert(0, root)                        # also makes the dataclasses_json stub importable
    _pkg = importlib.import_module("glow_tts_train")
    importlib.import_module("glow_tts_train.models")
    importlib.import_module("glow_tts_train.train")
    return _pkg


def swap_monotonic_align(pkg, module):
    """models.py:9 `from . import monotonic_align` -> `module`; returns the previous module."""
    models = sys.modules[pkg.__name__ + ".models"]
    previous = models.monotonic_align
    models.monotonic_align = module
    return previous


def make_model(pkg, *, mean_only=True, n_speakers=1, gin_channels=0, num_symbols=100, seed=1234, device="cuda"):
    """The reference's FlowGenerator + Noam/Adam optimizer via its own setup_model, default Glow-TTS base
    sizes (config.py:36-61); `n_speakers > 1, gin_channels = 256` is BASELINE.json configs[4]."""
    import torch

    config_mod = sys.modules[pkg.__name__ + ".config"]
    models = sys.modules[pkg.__name__ + ".models"]
    config = config_mod.TrainingConfig()
    config.model.num_symbols = num_symbols
    config.model.mean_only = mean_only
    config.model.n_speakers = n_speakers
    config.model.gin_channels = gin_channels
    torch.manual_seed(seed)
    model, optimizer = models.setup_model(config, use_cuda=(str(device) != "cpu"))
    return config, model, optimizer


def synthetic_batch(B, T_x, T_y, *, num_symbols=100, n_speakers=1, mel_channels=80, seed=0, ragged=True, device="cuda"):
    """(x, x_lengths, y, y_lengths, speaker_ids) shaped like PhonemeMelCollate's output
    (dataset.py:77-116): phoneme ids [B,T_x] int64 zero-padded, mels [B,80,T_y] fp32 zero-padded,
    lengths sorted descending (dataset.py:79-81), element 0 full size."""
    import torch

    g = torch.Generator().manual_seed(seed)
    if ragged:
        x_len = torch.randint(max(1, T_x // 2), T_x + 1, (B,), generator=g)
        x_len[0] = T_x
        x_len, _ = torch.sort(x_len, descending=True)
        y_len = (x_len.float() * (T_y / T_x) * (0.8 + 0.4 * torch.rand(B, generator=g))).round().long()
        y_len = torch.minimum(torch.maximum(y_len, x_len + 2), torch.tensor(T_y)) // 2 * 2
        y_len[0] = T_y
    else:
        x_len = torch.full((B,), T_x)
        y_len = torch.full((B,), T_y)
    x = torch.randint(1, num_symbols, (B, T_x), generator=g)
    y = torch.randn(B, mel_channels, T_y, generator=g)
    x = x * (torch.arange(T_x)[None] < x_len[:, None])
    y = y * (torch.arange(T_y)[None, None] < y_len[:, None, None])
    spk = torch.randint(0, n_speakers, (B,), generator=g) if n_speakers > 1 else None
    to = lambda t: None if t is None else t.to(device)  # noqa: E731
    return to(x), to(x_len), to(y), to(y_len), to(spk)
