"""ctypes binding of libmas_b200.so -- one Python callable per symbol of include/mas_b200.h.

Fails loudly: a missing library raises ImportError with the build command; a non-zero status from
an entry point raises RuntimeError.  Nothing here computes anything on the host.
"""
from __future__ import annotations

import ctypes
from pathlib import Path

PKG_DIR = Path(__file__).resolve().parent
LIB_PATH = PKG_DIR / "libmas_b200.so"

MAS_OK = 0
MAX_NEG_VAL = -1e9  # core.pyx:40 default

# every symbol include/mas_b200.h declares (tests check the .so exports exactly these)
EXPORTED_SYMBOLS = (
    "mas_b200_abi_version",
    "mas_b200_status_string",
    "mas_b200_last_cuda_error",
    "mas_b200_device_ok",
    "mas_b200_debug_set_cycle_buffer",
    "mas_b200_debug_force_cluster",
    "mas_b200_debug_force_unfused",
    "mas_b200_debug_tile_shape",
    "mas_b200_debug_path_plan",
    "mas_b200_debug_deal",
    "mas_b200_debug_fused_geom",
    "mas_b200_workspace_bytes",
    "mas_b200_fused_workspace_bytes",
    "mas_b200_maximum_path_f32",
    "mas_b200_logp_f32",
    "mas_b200_fused_maximum_path_f32",
    "mas_b200_duration_loss_f32",
    "mas_b200_duration_loss_backward_f32",
    "mas_b200_clip_grad_workspace_bytes",
    "mas_b200_clip_grad_value_f32",
    "mas_b200_maximum_path_host_i32",
    "mas_b200_shutdown",
    "mas_b200_expand_prior_f32",
    "mas_b200_expand_prior_backward_f32",
    "mas_b200_log_durations_f32",
    "mas_b200_generate_path_f32",
    "mas_b200_mle_loss_workspace_bytes",
    "mas_b200_mle_loss_f32",
    "mas_b200_mle_loss_backward_f32",
)

_lib = None

_vp, _i64, _i32, _f32, _sz = ctypes.c_void_p, ctypes.c_int64, ctypes.c_int, ctypes.c_float, ctypes.c_size_t


def load() -> ctypes.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python {PKG_DIR / 'build.py'}` "
            "(nvcc, sm_100a).  There is no CPU fallback."
        )
    lib = ctypes.CDLL(str(LIB_PATH))
    lib.mas_b200_abi_version.restype = _i32
    lib.mas_b200_status_string.restype = ctypes.c_char_p
    lib.mas_b200_status_string.argtypes = [_i32]
    lib.mas_b200_last_cuda_error.restype = _i32
    lib.mas_b200_device_ok.restype = _i32
    lib.mas_b200_debug_set_cycle_buffer.restype = None
    lib.mas_b200_debug_set_cycle_buffer.argtypes = [_vp]
    lib.mas_b200_debug_force_cluster.restype = None
    lib.mas_b200_debug_force_cluster.argtypes = [_i32]
    lib.mas_b200_debug_tile_shape.restype = _i32
    lib.mas_b200_debug_tile_shape.argtypes = [_i32, _i32, _vp]
    lib.mas_b200_debug_path_plan.restype = _i32
    lib.mas_b200_debug_path_plan.argtypes = [_i32, _i32, _i32, _i32, _i32, _vp]
    lib.mas_b200_debug_deal.restype = _i32
    lib.mas_b200_debug_deal.argtypes = [_i32, _i32, _i32, _vp, _vp]
    lib.mas_b200_duration_loss_f32.restype = _i32
    lib.mas_b200_duration_loss_f32.argtypes = [_vp, _vp, _vp, _vp, _i32, _i32, _vp]
    lib.mas_b200_duration_loss_backward_f32.restype = _i32
    lib.mas_b200_duration_loss_backward_f32.argtypes = [_vp, _vp, _vp, _vp, _vp, _i32, _i32, _vp]
    lib.mas_b200_clip_grad_workspace_bytes.restype = _sz
    lib.mas_b200_clip_grad_workspace_bytes.argtypes = [_i32]
    lib.mas_b200_clip_grad_value_f32.restype = _i32
    lib.mas_b200_clip_grad_value_f32.argtypes = [_vp, _vp, _i32, ctypes.c_float, _vp, _sz, _vp, _vp]
    lib.mas_b200_debug_fused_geom.restype = _i32
    lib.mas_b200_debug_fused_geom.argtypes = [_i32, _i32, _i32, _i32, _i32, _i32, _vp]
    lib.mas_b200_debug_force_unfused.restype = None
    lib.mas_b200_debug_force_unfused.argtypes = [_i32]
    lib.mas_b200_workspace_bytes.restype = _sz
    lib.mas_b200_workspace_bytes.argtypes = [_i32, _i32, _i32]
    lib.mas_b200_fused_workspace_bytes.restype = _sz
    lib.mas_b200_fused_workspace_bytes.argtypes = [_i32, _i32, _i32, _i32]
    lib.mas_b200_maximum_path_f32.restype = _i32
    lib.mas_b200_maximum_path_f32.argtypes = [
        _vp, _i64, _i64,            # value, stride_b, stride_x
        _vp, _vp,                   # t_x, t_y
        _vp, _i64, _i64, _i64,      # mask + strides
        _vp, _vp, _vp,              # path, durations, frame_token
        _vp, _sz,                   # workspace
        _i32, _i32, _i32, _f32, _vp,
    ]
    lib.mas_b200_logp_f32.restype = _i32
    lib.mas_b200_logp_f32.argtypes = [_vp, _vp, _vp, _vp, _i32, _i32, _i32, _i32, _vp]
    lib.mas_b200_fused_maximum_path_f32.restype = _i32
    lib.mas_b200_fused_maximum_path_f32.argtypes = [
        _vp, _vp, _vp, _vp, _vp,    # x_m, x_logs, z, x_len, y_len
        _vp, _vp, _vp,              # path, durations, frame_token
        _vp, _sz,
        _i32, _i32, _i32, _i32, _f32, _vp,
    ]
    lib.mas_b200_expand_prior_f32.restype = _i32
    lib.mas_b200_expand_prior_f32.argtypes = [_vp, _vp, _vp, _i32, _i32, _i32, _i32, _vp]
    lib.mas_b200_expand_prior_backward_f32.restype = _i32
    lib.mas_b200_expand_prior_backward_f32.argtypes = [_vp, _vp, _vp, _i32, _i32, _i32, _i32, _vp]
    lib.mas_b200_log_durations_f32.restype = _i32
    lib.mas_b200_log_durations_f32.argtypes = [_vp, _vp, _vp, _i32, _i32, _vp]
    lib.mas_b200_generate_path_f32.restype = _i32
    lib.mas_b200_generate_path_f32.argtypes = [_vp, _vp, _i64, _i64, _i64, _vp, _i32, _i32, _i32, _vp]
    lib.mas_b200_mle_loss_workspace_bytes.restype = ctypes.c_size_t
    lib.mas_b200_mle_loss_workspace_bytes.argtypes = [_i32, _i32]
    lib.mas_b200_mle_loss_f32.restype = _i32
    lib.mas_b200_mle_loss_f32.argtypes = [_vp] * 8 + [ctypes.c_size_t] + [_i32] * 4 + [_vp]
    lib.mas_b200_mle_loss_backward_f32.restype = _i32
    lib.mas_b200_mle_loss_backward_f32.argtypes = [_vp] * 9 + [_i32] * 4 + [_vp]
    lib.mas_b200_maximum_path_host_i32.restype = _i32
    lib.mas_b200_maximum_path_host_i32.argtypes = [_vp, _vp, _vp, _vp, _i32, _i32, _i32, _f32, _i32]
    lib.mas_b200_shutdown.restype = None
    lib.mas_b200_shutdown.argtypes = []
    _lib = lib
    return lib


def check(status: int, what: str) -> None:
    if status != MAS_OK:
        lib = load()
        msg = lib.mas_b200_status_string(status).decode()
        detail = f" (cudaError {lib.mas_b200_last_cuda_error()})" if status == 5 else ""
        raise RuntimeError(f"{what}: {msg}{detail}")
