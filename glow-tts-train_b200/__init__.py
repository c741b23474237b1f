"""glow-tts-train_b200 -- B200-native (sm_100a) alignment hot path of Glow-TTS training.

The directory name is not a Python identifier; import it through ``__graft_entry__.load_package()``
(registers it as ``glow_tts_train_b200``) or put this directory's parent on ``sys.path`` and use
``importlib``.  Contents:

  monotonic_align   drop-in for ``glow_tts_train.monotonic_align`` (``maximum_path(value, mask)``)
  alignment         the wider host API: fused logp+MAS, materialised logp, durations, the path's consumers
  training          the step around the path without host syncs: clip_grad_value_, duration_loss, train_step
  bucketing         length-bucketed batch sampler (padding is what the path pays for)
  sharding          utterance sharding across the GPUs of a box
  _lib              ctypes binding of libmas_b200.so (the C ABI in include/mas_b200.h)
  build             nvcc recipe for the library
  csrc/             the CUDA kernels and the C-ABI layer

There is no CPU fallback: without the compiled library or without a CUDA device every compute
entry point raises.
"""
from . import _lib, alignment, bucketing, monotonic_align, sharding, training  # noqa: F401
from .alignment import (  # noqa: F401
    aligned_mle_loss,
    expand_prior,
    fused_maximum_path,
    generate_path,
    log_durations,
    log_likelihood_matrix,
    maximum_path_from_lengths,
)
from .bucketing import LengthBucketBatchSampler  # noqa: F401
from .monotonic_align import maximum_path  # noqa: F401
from .training import clip_grad_value_, duration_loss, train_step  # noqa: F401

__all__ = [
    "maximum_path",
    "maximum_path_from_lengths",
    "fused_maximum_path",
    "log_likelihood_matrix",
    "expand_prior",
    "log_durations",
    "generate_path",
    "aligned_mle_loss",
    "duration_loss",
    "clip_grad_value_",
    "train_step",
    "LengthBucketBatchSampler",
    "monotonic_align",
    "alignment",
]
