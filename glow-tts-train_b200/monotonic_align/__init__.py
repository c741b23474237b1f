"""Drop-in replacement for ``glow_tts_train.monotonic_align`` (reference:
glow_tts_train/monotonic_align/__init__.py:6-21).

    from glow_tts_train_b200 import monotonic_align
    attn = monotonic_align.maximum_path(logp, attn_mask.squeeze(1))      # models.py:378-382

Same name, same positional signature, same return contract: a new ``[b, t_x, t_y]`` tensor of
zeros and ones with ``value``'s dtype and device.  What differs is where it runs: the reference
synchronises the device, copies ``value*mask`` and ``mask`` to the host, runs the Cython DP and
copies a dense int32 path back; here everything stays on the GPU, asynchronous on the current
stream.  There is no CPU implementation: without a CUDA device this raises.
"""
from __future__ import annotations

import torch

from .. import alignment

__all__ = ["maximum_path"]


def maximum_path(value, mask):
    """value: [b, t_x, t_y] scores, mask: [b, t_x, t_y] (models.py:334-337 builds a prefix mask).

    Same results as the reference for ANY mask: the valid sizes come from the mask's first column
    and first row (__init__.py:18-19), read on the device; the product ``value * mask``
    (__init__.py:11) is an identity on every cell the algorithm touches when the mask is all ones
    on its valid rectangle -- the kernel entry verifies exactly that on the device (one pass over
    the mask) and computes any other utterance from ``value * mask`` literally.  No host
    synchronisation either way.

    CPU tensors are accepted for parity with the reference's device-agnostic signature: they are
    staged through the current CUDA device (pinned if they are pinned) and the result is copied
    back -- the compute still happens on the GPU.
    """
    if value.dim() != 3 or mask.dim() != 3:
        raise ValueError("value and mask must be [b, t_x, t_y]")
    if not torch.cuda.is_available():
        raise RuntimeError("monotonic_align.maximum_path needs a CUDA device (sm_100a); there is no CPU fallback")
    device, dtype = value.device, value.dtype              # __init__.py:12-13
    on_host = not value.is_cuda
    work_dev = torch.device("cuda", torch.cuda.current_device()) if on_host else device

    scores = value.detach()
    mask = mask.detach()
    if on_host:
        # the product on the host side of the copy (the full mask never crosses PCIe), lengths from
        # the (small) first row / column
        scores = scores * mask                              # __init__.py:11
        t_x = mask[:, :, 0].sum(1).to(torch.int32).to(work_dev, non_blocking=True)   # __init__.py:18
        t_y = mask[:, 0, :].sum(1).to(torch.int32).to(work_dev, non_blocking=True)   # __init__.py:19
        scores = scores.to(work_dev, dtype=torch.float32, non_blocking=True)         # __init__.py:14
        path = alignment.maximum_path_from_lengths(scores, t_x, t_y)
        out = torch.empty(path.shape, dtype=dtype, pin_memory=value.is_pinned())
        out.copy_(path, non_blocking=False)                                            # __init__.py:21
        return out

    if tuple(mask.shape) != tuple(scores.shape):
        mask = mask.expand_as(scores)
    if scores.dtype != torch.float32:
        # the reference forms value * mask in value's dtype BEFORE .astype(np.float32) (__init__.py:11,14):
        # for half-precision scores keep that order
        scores = (scores * mask.to(device=device, dtype=scores.dtype)).to(torch.float32)
    mask = mask.to(device=device, dtype=torch.float32)      # a no-op for the fp32 view models.py:379 passes
    path = alignment.maximum_path_from_lengths(scores, mask=mask)
    return path if dtype == torch.float32 else path.to(dtype)   # __init__.py:21
