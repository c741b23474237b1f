"""Host API of the alignment hot path (Python/PyTorch above the C ABI).

PyTorch is used for device memory, streams and dtype plumbing only; every arithmetic step runs in
the sm_100a kernels of libmas_b200.so.  All entry points are asynchronous on the current torch CUDA
stream of the inputs' device and never synchronise the host.

Reference call sites replaced (rhasspy/glow-tts-train):
  glow_tts_train/models.py:362-376   -> log_likelihood_matrix
  glow_tts_train/models.py:362-382   -> fused_maximum_path        (+ :393 durations)
  glow_tts_train/monotonic_align/core.pyx:40-45 -> maximum_path_from_lengths
"""
from __future__ import annotations

import torch
from torch.autograd.function import once_differentiable

from . import _lib


def _require_cuda(t: torch.Tensor, name: str) -> None:
    if not t.is_cuda:
        raise RuntimeError(f"{name} must be a CUDA tensor: the alignment path has no CPU implementation")


def _ptr(t):
    return None if t is None else t.data_ptr()


def _stream(device) -> int:
    return torch.cuda.current_stream(device).cuda_stream


def _workspace(nbytes: int, device) -> torch.Tensor:
    return torch.empty(max(int(nbytes), 1), dtype=torch.uint8, device=device)


def maximum_path_from_lengths(value, t_x=None, t_y=None, *, mask=None, want_durations=False,
                              want_frame_token=False, max_neg_val=_lib.MAX_NEG_VAL):
    """Kernel (1).  ``value`` fp32 CUDA [B,T_x,T_y] (frame stride 1); lengths either as int32 CUDA
    vectors ``t_x``/``t_y`` or read on the device from ``mask`` (fp32 CUDA, any strides) the way
    monotonic_align/__init__.py:18-19 does.  Returns ``path`` fp32 [B,T_x,T_y] (and optionally
    int32 ``durations`` [B,T_x], ``frame_token`` [B,T_y])."""
    lib = _lib.load()
    _require_cuda(value, "value")
    if value.dtype != torch.float32 or value.dim() != 3:
        raise TypeError("value must be a float32 [B, T_x, T_y] tensor")
    if value.stride(2) != 1 and value.numel() > 0:
        value = value.contiguous()
    B, T_x, T_y = value.shape
    dev = value.device
    if (t_x is None) != (t_y is None):
        raise ValueError("pass both t_x and t_y, or neither (and a mask)")
    if t_x is None:
        if mask is None:
            raise ValueError("lengths or mask required")
        _require_cuda(mask, "mask")
        if mask.dtype != torch.float32 or tuple(mask.shape) != (B, T_x, T_y) or mask.device != dev:
            raise TypeError("mask must be float32 with value's shape and device")
        ms = mask.stride()
    else:
        for name, t in (("t_x", t_x), ("t_y", t_y)):
            if t.dtype != torch.int32 or t.device != dev or t.shape != (B,) or not t.is_contiguous():
                raise TypeError(f"{name} must be a contiguous int32 [B] tensor on value's device")
        ms = (0, 0, 0)
        mask = None
    path = torch.empty((B, T_x, T_y), dtype=torch.float32, device=dev)
    durations = torch.empty((B, T_x), dtype=torch.int32, device=dev) if want_durations else None
    frame_token = torch.empty((B, T_y), dtype=torch.int32, device=dev) if want_frame_token else None
    if path.numel() == 0:
        return _pack(path, durations, frame_token)
    ws = _workspace(lib.mas_b200_workspace_bytes(B, T_x, T_y), dev)
    with torch.cuda.device(dev):
        rc = lib.mas_b200_maximum_path_f32(
            value.data_ptr(), value.stride(0), value.stride(1), _ptr(t_x), _ptr(t_y),
            _ptr(mask), ms[0], ms[1], ms[2], path.data_ptr(), _ptr(durations), _ptr(frame_token),
            ws.data_ptr(), ws.numel(), B, T_x, T_y, max_neg_val, _stream(dev))
    _lib.check(rc, "mas_b200_maximum_path_f32")
    return _pack(path, durations, frame_token)


def _pack(path, durations, frame_token):
    extra = tuple(t for t in (durations, frame_token) if t is not None)
    return (path, *extra) if extra else path


def _check_prior(x_m, x_logs, z):
    for name, t in (("x_m", x_m), ("z", z)) + ((("x_logs", x_logs),) if x_logs is not None else ()):
        _require_cuda(t, name)
        if t.dtype != torch.float32 or t.dim() != 3:
            raise TypeError(f"{name} must be a float32 [B, D, T] tensor")
    B, D, T_x = x_m.shape
    if z.shape[0] != B or z.shape[1] != D or (x_logs is not None and x_logs.shape != x_m.shape):
        raise ValueError("x_m / x_logs [B,D,T_x] and z [B,D,T_y] disagree")
    return B, D, T_x, z.shape[2]


def log_likelihood_matrix(x_m, x_logs, z, out=None):
    """models.py:362-376 on the GPU: ``x_m``/``x_logs`` [B,D,T_x] (``x_logs=None`` == zeros, the
    mean_only configuration), ``z`` [B,D,T_y] -> logp fp32 [B,T_x,T_y] (written into ``out`` when
    given: contiguous fp32 of that shape on the same device)."""
    lib = _lib.load()
    B, D, T_x, T_y = _check_prior(x_m, x_logs, z)
    x_m, z = x_m.contiguous(), z.contiguous()
    x_logs = x_logs.contiguous() if x_logs is not None else None
    if out is None:
        out = torch.empty((B, T_x, T_y), dtype=torch.float32, device=x_m.device)
    elif (out.dtype != torch.float32 or out.device != x_m.device or tuple(out.shape) != (B, T_x, T_y)
          or not out.is_contiguous()):
        raise ValueError("out must be a contiguous float32 [B, T_x, T_y] tensor on the inputs' device")
    if out.numel() == 0:
        return out
    with torch.cuda.device(x_m.device):
        rc = lib.mas_b200_logp_f32(x_m.data_ptr(), _ptr(x_logs), z.data_ptr(), out.data_ptr(),
                                   B, D, T_x, T_y, _stream(x_m.device))
    _lib.check(rc, "mas_b200_logp_f32")
    return out


def fused_maximum_path(x_m, x_logs, z, x_lengths, y_lengths, *, want_durations=True,
                       want_frame_token=False, max_neg_val=_lib.MAX_NEG_VAL):
    """Kernel (2): models.py:362-382 in one call, the score matrix never leaves the chip.
    ``x_lengths``/``y_lengths`` are the integer lengths the prefix masks of models.py:334-337
    encode (``y_lengths`` already floored to n_sqz, models.py:405).  Returns ``path`` fp32
    [B,T_x,T_y] (+ int32 ``durations`` [B,T_x], ``frame_token`` [B,T_y] on request)."""
    lib = _lib.load()
    B, D, T_x, T_y = _check_prior(x_m, x_logs, z)
    dev = x_m.device
    x_m, z = x_m.contiguous(), z.contiguous()
    x_logs = x_logs.contiguous() if x_logs is not None else None
    x_len = x_lengths.to(device=dev, dtype=torch.int32).contiguous()
    y_len = y_lengths.to(device=dev, dtype=torch.int32).contiguous()
    if x_len.shape != (B,) or y_len.shape != (B,):
        raise ValueError("x_lengths / y_lengths must have shape [B]")
    path = torch.empty((B, T_x, T_y), dtype=torch.float32, device=dev)
    durations = torch.empty((B, T_x), dtype=torch.int32, device=dev) if want_durations else None
    frame_token = torch.empty((B, T_y), dtype=torch.int32, device=dev) if want_frame_token else None
    if path.numel() == 0:
        return _pack(path, durations, frame_token)
    ws = _workspace(lib.mas_b200_fused_workspace_bytes(B, D, T_x, T_y), dev)
    with torch.cuda.device(dev):
        rc = lib.mas_b200_fused_maximum_path_f32(
            x_m.data_ptr(), _ptr(x_logs), z.data_ptr(), x_len.data_ptr(), y_len.data_ptr(),
            path.data_ptr(), _ptr(durations), _ptr(frame_token), ws.data_ptr(), ws.numel(),
            B, D, T_x, T_y, max_neg_val, _stream(dev))
    _lib.check(rc, "mas_b200_fused_maximum_path_f32")
    return _pack(path, durations, frame_token)


# ------------------------------------------------------------------------------------------------
# the path's consumers: models.py:383-393 without the dense path
# ------------------------------------------------------------------------------------------------
class _ExpandPrior(torch.autograd.Function):
    """z[b,d,y] = x[b,d,frame_token[b,y]] -- what ``(attn^T @ x^T)^T`` computes (models.py:383-392),
    as a gather; the backward is the segmented sum over each token's run of frames."""

    @staticmethod
    def forward(ctx, x, frame_token, durations):
        lib = _lib.load()
        _require_cuda(x, "x")
        if x.dtype != torch.float32 or x.dim() != 3:
            raise TypeError("x must be a float32 [B, D, T_x] tensor")
        x = x.contiguous()
        B, D, T_x = x.shape
        if (frame_token.dim() != 2 or frame_token.shape[0] != B or tuple(durations.shape) != (B, T_x)
                or frame_token.device != x.device or durations.device != x.device):
            raise ValueError("frame_token must be [B, T_y] and durations [B, T_x], both on x's device")
        T_y = frame_token.shape[1]
        z = torch.empty((B, D, T_y), dtype=torch.float32, device=x.device)
        with torch.cuda.device(x.device):
            rc = lib.mas_b200_expand_prior_f32(x.data_ptr(), frame_token.data_ptr(), z.data_ptr(), B, D, T_x, T_y,
                                               _stream(x.device))
        _lib.check(rc, "mas_b200_expand_prior_f32")
        ctx.save_for_backward(durations)
        ctx.shape = (B, D, T_x, T_y)
        return z

    @staticmethod
    @once_differentiable
    def backward(ctx, dz):
        lib = _lib.load()
        (durations,) = ctx.saved_tensors
        B, D, T_x, T_y = ctx.shape
        dz = dz.contiguous().float()
        dx = torch.empty((B, D, T_x), dtype=torch.float32, device=dz.device)
        with torch.cuda.device(dz.device):
            rc = lib.mas_b200_expand_prior_backward_f32(dz.data_ptr(), durations.data_ptr(), dx.data_ptr(), B, D, T_x, T_y,
                                                        _stream(dz.device))
        _lib.check(rc, "mas_b200_expand_prior_backward_f32")
        return dx, None, None


def expand_prior(x, frame_token, durations):
    """Frame-level prior from the token-level one: ``z_m = expand_prior(x_m, frame_token, durations)``
    replaces ``torch.matmul(attn.squeeze(1).transpose(1, 2), x_m.transpose(1, 2)).transpose(1, 2)``
    (models.py:383-387; same for ``x_logs``, :388-392).  Differentiable w.r.t. ``x``."""
    if frame_token.dtype != torch.int32 or durations.dtype != torch.int32:
        raise TypeError("frame_token and durations must be int32 (as fused_maximum_path returns them)")
    return _ExpandPrior.apply(x, frame_token.contiguous(), durations.contiguous())


def log_durations(durations, x_lengths):
    """``logw_ = torch.log(1e-8 + torch.sum(attn, -1)) * x_mask`` (models.py:393) from the integer
    durations: fp32 [B, 1, T_x]."""
    lib = _lib.load()
    _require_cuda(durations, "durations")
    if durations.dtype != torch.int32 or durations.dim() != 2:
        raise TypeError("durations must be an int32 [B, T_x] tensor (as the alignment kernels return it)")
    B, T_x = durations.shape
    durations = durations.contiguous()                      # (bound to a name: alive until the launch is enqueued)
    x_len = x_lengths.to(device=durations.device, dtype=torch.int32).contiguous()
    if x_len.shape != (B,):
        raise ValueError("x_lengths must have shape [B]")
    out = torch.empty((B, 1, T_x), dtype=torch.float32, device=durations.device)
    with torch.cuda.device(durations.device):
        rc = lib.mas_b200_log_durations_f32(durations.data_ptr(), x_len.data_ptr(), out.data_ptr(), B, T_x,
                                            _stream(durations.device))
    _lib.check(rc, "mas_b200_log_durations_f32")
    return out


class _AlignedMleLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, z, x_m, x_logs, logdet, frame_token, durations, y_len):
        lib = _lib.load()
        B, D, T_y = z.shape
        T_x = x_m.shape[2]
        out = torch.empty(2, dtype=torch.float32, device=z.device)
        ws = torch.empty(lib.mas_b200_mle_loss_workspace_bytes(B, T_y), dtype=torch.uint8, device=z.device)
        with torch.cuda.device(z.device):
            rc = lib.mas_b200_mle_loss_f32(z.data_ptr(), x_m.data_ptr(), _ptr(x_logs), frame_token.data_ptr(), _ptr(logdet),
                                           y_len.data_ptr(), out.data_ptr(), ws.data_ptr(), ws.numel(), B, D, T_x, T_y,
                                           _stream(z.device))
        _lib.check(rc, "mas_b200_mle_loss_f32")
        ctx.save_for_backward(z, x_m, x_logs if x_logs is not None else z.new_empty(0), frame_token, durations, out)
        ctx.has_logs = x_logs is not None
        ctx.logdet_shape = None if logdet is None else logdet.shape
        return out[0]

    @staticmethod
    @once_differentiable
    def backward(ctx, g):
        lib = _lib.load()
        z, x_m, x_logs, frame_token, durations, out = ctx.saved_tensors
        x_logs = x_logs if ctx.has_logs else None
        B, D, T_y = z.shape
        T_x = x_m.shape[2]
        need_z, need_m, need_logs, need_logdet = ctx.needs_input_grad[:4]
        scale = (g.float() * out[1]).reshape(1).contiguous()            # stays on the device: no host sync
        dz = torch.empty_like(z) if need_z else None
        want_tokens = need_m or (need_logs and x_logs is not None)
        dx_m = torch.empty_like(x_m) if want_tokens else None
        dx_logs = torch.empty_like(x_logs) if (need_logs and x_logs is not None) else None
        if need_z or want_tokens:
            with torch.cuda.device(z.device):
                rc = lib.mas_b200_mle_loss_backward_f32(z.data_ptr(), x_m.data_ptr(), _ptr(x_logs), frame_token.data_ptr(),
                                                        durations.data_ptr(), scale.data_ptr(), _ptr(dz), _ptr(dx_m), _ptr(dx_logs),
                                                        B, D, T_x, T_y, _stream(z.device))
            _lib.check(rc, "mas_b200_mle_loss_backward_f32")
        dlogdet = (-scale).expand(ctx.logdet_shape).clone() if (need_logdet and ctx.logdet_shape is not None) else None
        return dz, (dx_m if need_m else None), dx_logs, dlogdet, None, None, None


def aligned_mle_loss(z, x_m, x_logs, logdet, frame_token, durations, y_lengths):
    """``mle_loss(z, z_m, z_logs, logdet, z_mask)`` (utils.py:14-23, train.py:124) for the aligned prior,
    from the token-level ``x_m`` / ``x_logs`` [B,D,T_x] and the ``frame_token`` / ``durations`` the
    alignment kernels emit -- ``z_m`` / ``z_logs`` (models.py:383-392) are never built.  ``z`` [B,D,T_y]
    (already masked, as the decoder returns it), ``logdet`` [B] or None, ``x_logs`` None == zeros
    (mean_only), ``y_lengths`` the frame counts behind ``z_mask``.  Scalar fp32, differentiable w.r.t.
    ``z``, ``x_m``, ``x_logs`` and ``logdet``."""
    B, D, T_x, T_y = _check_prior(x_m, x_logs, z)
    if frame_token.dtype != torch.int32 or durations.dtype != torch.int32:
        raise TypeError("frame_token and durations must be int32 (as fused_maximum_path returns them)")
    if tuple(frame_token.shape) != (B, T_y) or tuple(durations.shape) != (B, T_x):
        raise ValueError("frame_token must be [B, T_y] and durations [B, T_x]")
    if B == 0 or D == 0 or T_x == 0 or T_y == 0:
        raise ValueError("aligned_mle_loss needs a non-empty batch")
    if logdet is not None:
        _require_cuda(logdet, "logdet")
        if logdet.numel() != B:
            raise ValueError("logdet must have one entry per utterance")
        logdet = logdet.float().contiguous()
    y_len = y_lengths.to(device=z.device, dtype=torch.int32).contiguous()
    return _AlignedMleLoss.apply(z.contiguous(), x_m.contiguous(), None if x_logs is None else x_logs.contiguous(), logdet,
                                 frame_token.contiguous(), durations.contiguous(), y_len)


def generate_path(duration, mask):
    """``generate_path(duration, mask)`` of the reference (utils.py:99-115, the inference branch at
    models.py:340): ``duration`` [B, T_x] (ceil-ed frame counts, any float/int dtype), ``mask``
    [B, T_x, T_y] (a strided view is fine) -> dense path [B, T_x, T_y] in ``mask``'s dtype."""
    lib = _lib.load()
    _require_cuda(duration, "duration")
    _require_cuda(mask, "mask")
    if duration.dim() != 2 or mask.dim() != 3 or tuple(mask.shape[:2]) != tuple(duration.shape):
        raise ValueError("duration must be [B, T_x] and mask [B, T_x, T_y]")
    B, T_x, T_y = mask.shape
    m32 = mask if mask.dtype == torch.float32 else mask.float()
    d32 = duration.float().contiguous()
    out = torch.empty((B, T_x, T_y), dtype=torch.float32, device=mask.device)
    if out.numel():
        with torch.cuda.device(mask.device):
            rc = lib.mas_b200_generate_path_f32(d32.data_ptr(), m32.data_ptr(), m32.stride(0), m32.stride(1), m32.stride(2),
                                                out.data_ptr(), B, T_x, T_y, _stream(mask.device))
        _lib.check(rc, "mas_b200_generate_path_f32")
    return out if mask.dtype == torch.float32 else out.to(mask.dtype)
