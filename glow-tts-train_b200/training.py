"""The training step AROUND the alignment path without its host synchronisations (SURVEY.md 8f ranks
2-3).  Once ``maximum_path`` stays on the GPU, what serialises the reference's step is its own
bookkeeping:

  glow_tts_train/utils.py:118-132   clip_grad_value_: ``.item()`` on every gradient's norm -- one
                                    device sync per parameter tensor and step (~300 of them)
  glow_tts_train/train.py:131, 149  ``loss_g.item()`` twice per step
  glow_tts_train/utils.py:26-28     duration_loss on a target that models.py:393 builds from the
                                    dense path

``clip_grad_value_`` / ``duration_loss`` / ``train_step`` keep the reference's names, arguments and
results; nothing here reads a value back from the device inside the step (``train_step`` reads the
epoch's mean loss once, at the end, like the reference logs it).  CUDA only, like the rest of the
package.
"""
from __future__ import annotations

import ctypes
import logging
import typing

import torch
from torch.autograd.function import once_differentiable

from . import _lib

_LOGGER = logging.getLogger("glow_tts_train_b200.training")
_CHUNK = 1 << 16          # floats per clipping chunk: every SM gets work, the table stays a few KB


class _GradTable:
    """Device table of gradient chunks for one set of gradient tensors (rebuilt only when a gradient's
    address or size changes: optimizers that zero gradients in place keep both)."""

    def __init__(self):
        self.key = None
        self.ptrs = self.counts = self.ws = self.norm = None
        self.nchunks = 0

    def update(self, grads, device):
        key = tuple((g.data_ptr(), g.numel()) for g in grads)
        if key == self.key:
            return
        ptrs, counts = [], []
        for g in grads:
            base, n = g.data_ptr(), g.numel()
            for off in range(0, n, _CHUNK):
                ptrs.append(base + 4 * off)
                counts.append(min(_CHUNK, n - off))
        lib = _lib.load()
        self.nchunks = len(ptrs)
        # (a host->device copy of a few KB, enqueued on the current stream; no synchronisation)
        self.ptrs = torch.tensor(ptrs, dtype=torch.int64).to(device, non_blocking=True)
        self.counts = torch.tensor(counts, dtype=torch.int32).to(device, non_blocking=True)
        self.ws = torch.empty(max(1, lib.mas_b200_clip_grad_workspace_bytes(self.nchunks)), dtype=torch.uint8, device=device)
        self.key = key


_tables: typing.Dict[int, _GradTable] = {}


def _is_dense(t: torch.Tensor) -> bool:
    """True when the tensor's elements occupy numel() consecutive floats starting at data_ptr() (in any
    order): positive strides that nest without gaps or overlap."""
    if t.is_contiguous():
        return True
    dims = sorted((st, sz) for st, sz in zip(t.stride(), t.shape) if sz > 1)
    expect = 1
    for st, sz in dims:
        if st != expect:
            return False
        expect *= sz
    return True


def clip_grad_value_(parameters, clip_value, norm_type=2):
    """``clip_grad_value_(parameters, clip_value)`` of the reference (utils.py:118-132): clamps every
    gradient to [-clip_value, clip_value] in place and returns the total 2-norm of the gradients BEFORE
    clamping -- as a 0-d CUDA tensor instead of a Python float, because producing the float is what
    costs the reference one device synchronisation per parameter tensor."""
    if float(norm_type) != 2.0:
        raise ValueError("only the 2-norm the reference uses (utils.py:118) is implemented")
    if isinstance(parameters, torch.Tensor):
        parameters = [parameters]
    grads = [p.grad for p in parameters if p.grad is not None]
    if not grads:
        return torch.zeros((), dtype=torch.float32)
    dev = grads[0].device
    flat, copies = [], []
    for g in grads:
        if not g.is_cuda or g.device != dev:
            raise RuntimeError("clip_grad_value_ needs every gradient on one CUDA device: there is no CPU implementation")
        if g.dtype != torch.float32:
            raise TypeError("gradients must be float32 tensors")
        if _is_dense(g):
            flat.append(g)          # a permuted but gap-free layout (e.g. a transposed weight): clamping and the sum of
        else:                       # squares do not care about the order, its numel() floats are one flat range
            c = g.contiguous()      # (rare) a gradient with gaps: through a contiguous copy
            copies.append((g, c))
            flat.append(c)
    lib = _lib.load()
    table = _tables.setdefault(dev.index or 0, _GradTable())
    table.update(flat, dev)
    norm = torch.empty(1, dtype=torch.float32, device=dev)
    with torch.cuda.device(dev):
        rc = lib.mas_b200_clip_grad_value_f32(table.ptrs.data_ptr(), table.counts.data_ptr(), table.nchunks, ctypes.c_float(float(clip_value)),
                                              table.ws.data_ptr(), table.ws.numel(), norm.data_ptr(),
                                              torch.cuda.current_stream(dev).cuda_stream)
    _lib.check(rc, "mas_b200_clip_grad_value_f32")
    for g, c in copies:
        g.copy_(c)
    return norm[0]


class _DurationLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logw, durations, x_len):
        lib = _lib.load()
        B, T_x = durations.shape
        flat = logw.reshape(B, T_x).contiguous()
        out = torch.empty(2, dtype=torch.float32, device=flat.device)
        with torch.cuda.device(flat.device):
            rc = lib.mas_b200_duration_loss_f32(flat.data_ptr(), durations.data_ptr(), x_len.data_ptr(), out.data_ptr(), B, T_x,
                                                torch.cuda.current_stream(flat.device).cuda_stream)
        _lib.check(rc, "mas_b200_duration_loss_f32")
        ctx.save_for_backward(flat, durations, x_len, out)
        ctx.shape = logw.shape
        return out[0]

    @staticmethod
    @once_differentiable
    def backward(ctx, g):
        lib = _lib.load()
        flat, durations, x_len, out = ctx.saved_tensors
        B, T_x = durations.shape
        scale = (g.float() * out[1]).reshape(1).contiguous()           # stays on the device
        dlogw = torch.empty_like(flat)
        with torch.cuda.device(flat.device):
            rc = lib.mas_b200_duration_loss_backward_f32(flat.data_ptr(), durations.data_ptr(), x_len.data_ptr(), scale.data_ptr(),
                                                         dlogw.data_ptr(), B, T_x, torch.cuda.current_stream(flat.device).cuda_stream)
        _lib.check(rc, "mas_b200_duration_loss_backward_f32")
        return dlogw.reshape(ctx.shape), None, None


def duration_loss(logw, durations, x_lengths):
    """``duration_loss(logw, logw_, lengths)`` of the reference (utils.py:26-28, train.py:125) with the
    target taken from the INTEGER ``durations`` [B, T_x] the alignment kernels emit instead of
    ``logw_ = log(1e-8 + attn.sum(-1)) * x_mask`` (models.py:393), which is formed on the fly.
    ``logw`` fp32 [B, 1, T_x] or [B, T_x] (masked, as the duration predictor returns it).  Scalar fp32,
    differentiable w.r.t. ``logw``."""
    if not logw.is_cuda:
        raise RuntimeError("logw must be a CUDA tensor: there is no CPU implementation")
    if durations.dtype != torch.int32 or durations.dim() != 2:
        raise TypeError("durations must be an int32 [B, T_x] tensor (as fused_maximum_path returns it)")
    B, T_x = durations.shape
    if logw.dtype != torch.float32 or logw.numel() != B * T_x:
        raise TypeError("logw must be float32 with B * T_x elements")
    if B == 0 or T_x == 0:
        raise ValueError("duration_loss needs a non-empty batch")
    x_len = x_lengths.to(device=logw.device, dtype=torch.int32).contiguous()
    return _DurationLoss.apply(logw, durations.contiguous(), x_len)


def train_step(global_step, epoch, model, optimizer, config, train_loader, fp16_run, scaler=None):
    """The reference's ``train_step`` (train.py:91-162) -- same arguments, same return value, same
    arithmetic -- with the two host synchronisations per step (``loss_g.item()``, train.py:131 and :149)
    and the ~300 of ``clip_grad_value_`` gone: losses are accumulated on the device and read ONCE when
    the epoch's mean is logged.  The model is the caller's (the reference's FlowGenerator with
    ``monotonic_align`` swapped); its losses come from its own package."""
    import importlib

    pkg = type(model.module if hasattr(model, "module") else model).__module__.rsplit(".", 1)[0]
    utils = importlib.import_module(pkg + ".utils")                  # mle_loss, duration_loss, to_gpu (utils.py)
    from torch.amp import autocast

    steps_per_epoch = len(train_loader)
    loss_sum, count = None, 0
    model.train()
    for batch_idx, (x, x_lengths, y, y_lengths, speaker_ids) in enumerate(train_loader):
        x, x_lengths = utils.to_gpu(x), utils.to_gpu(x_lengths)
        y, y_lengths = utils.to_gpu(y), utils.to_gpu(y_lengths)
        if speaker_ids is not None:
            speaker_ids = utils.to_gpu(speaker_ids)
        optimizer.zero_grad()
        with autocast("cuda", enabled=fp16_run):
            (z, z_m, z_logs, logdet, z_mask), _, (_attn, logw, logw_) = model(x, x_lengths, y, y_lengths, g=speaker_ids)
            l_mle = utils.mle_loss(z, z_m, z_logs, logdet, z_mask)
            l_length = utils.duration_loss(logw, logw_, x_lengths)
            loss_g = l_mle + l_length
        loss_sum = loss_g.detach() if loss_sum is None else loss_sum + loss_g.detach()
        count += 1
        if fp16_run:
            assert scaler is not None
            scaler.scale(loss_g).backward()
            scaler.unscale_(optimizer._optim)  # noqa: SLF001 (train.py:140)
            clip_grad_value_(model.parameters(), config.grad_clip)
            scaler.step(optimizer._optim)  # noqa: SLF001
            scaler.update()
        else:
            loss_g.backward()
            clip_grad_value_(model.parameters(), config.grad_clip)
            optimizer.step()
        global_step += 1
        _LOGGER.debug("step %s/%s of epoch %s enqueued", batch_idx + 1, steps_per_epoch, epoch)
    if count:
        _LOGGER.info("Avg. Loss for epoch %s: %s (global step=%s)", epoch, float(loss_sum) / count, global_step)
    return global_step
