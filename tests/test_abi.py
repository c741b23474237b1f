"""The C-ABI library without a GPU: it loads, exports exactly what include/mas_b200.h declares, and
its argument validation answers without touching CUDA."""
from __future__ import annotations

import ctypes
import re
import subprocess
from pathlib import Path

import pytest

REPO = Path(__file__).resolve().parent.parent
HEADER = REPO / "include" / "mas_b200.h"


def declared_symbols():
    text = re.sub(r"/\*.*?\*/", "", HEADER.read_text(), flags=re.S)
    return sorted(set(re.findall(r"\b(mas_b200_\w+)\s*\(", text)))


def test_header_and_binding_agree(pkg):
    assert sorted(pkg._lib.EXPORTED_SYMBOLS) == declared_symbols()


def test_library_exports_every_declared_symbol(pkg):
    lib = pkg._lib.load()
    out = subprocess.run(["nm", "-D", "--defined-only", str(pkg._lib.LIB_PATH)], capture_output=True, text=True, check=True).stdout
    exported = set(re.findall(r"\bT (mas_b200_\w+)", out))
    assert exported == set(declared_symbols())
    for sym in declared_symbols():
        assert getattr(lib, sym) is not None


def test_library_is_sm100a_only(pkg):
    out = subprocess.run(["cuobjdump", "--list-elf", str(pkg._lib.LIB_PATH)], capture_output=True, text=True)
    if out.returncode != 0:
        pytest.skip("cuobjdump unavailable")
    archs = set(re.findall(r"sm_\d+a?", out.stdout))
    assert archs == {"sm_100a"}, archs


def test_version_and_status_strings(pkg):
    lib = pkg._lib.load()
    assert lib.mas_b200_abi_version() == 2
    assert lib.mas_b200_status_string(0) == b"ok"
    assert lib.mas_b200_status_string(3) == b"workspace too small"
    assert lib.mas_b200_status_string(99) == b"unknown status"


def test_workspace_queries(pkg):
    lib = pkg._lib.load()
    assert lib.mas_b200_workspace_bytes(0, 10, 10) == 0 or lib.mas_b200_workspace_bytes(0, 10, 10) % 256 == 0
    a = lib.mas_b200_workspace_bytes(32, 200, 1000)
    assert a % 256 == 0
    assert lib.mas_b200_fused_workspace_bytes(32, 80, 200, 1000) >= a
    assert lib.mas_b200_workspace_bytes(1, 4096, 10) == 0      # beyond MAS_B200_MAX_TOKENS
    assert lib.mas_b200_workspace_bytes(-1, 4, 10) == 0


def test_argument_validation_without_gpu(pkg):
    lib = pkg._lib.load()
    null = None
    one = ctypes.c_void_p(256)   # never dereferenced: validation fails first
    # negative sizes
    assert lib.mas_b200_maximum_path_f32(one, 0, 0, null, null, one, 0, 0, 0, one, null, null, null, 0, -1, 4, 4, -1e9, null) == 1
    # too many tokens
    assert lib.mas_b200_maximum_path_f32(one, 0, 0, null, null, one, 0, 0, 0, one, null, null, null, 0, 1, 5000, 6000, -1e9, null) == 2
    # empty batch is a no-op
    assert lib.mas_b200_maximum_path_f32(null, 0, 0, null, null, null, 0, 0, 0, null, null, null, null, 0, 0, 4, 4, -1e9, null) == 0
    # null value
    assert lib.mas_b200_maximum_path_f32(null, 16, 4, null, null, one, 0, 0, 0, one, null, null, null, 0, 1, 4, 4, -1e9, null) == 1
    # t_x without t_y
    assert lib.mas_b200_maximum_path_f32(one, 16, 4, one, null, null, 0, 0, 0, one, null, null, null, 0, 1, 4, 4, -1e9, null) == 1
    # neither lengths nor mask
    assert lib.mas_b200_maximum_path_f32(one, 16, 4, null, null, null, 0, 0, 0, one, null, null, null, 0, 1, 4, 4, -1e9, null) == 1
    # token stride smaller than a row
    assert lib.mas_b200_maximum_path_f32(one, 16, 2, one, one, null, 0, 0, 0, one, null, null, null, 0, 1, 4, 4, -1e9, null) == 1
    assert lib.mas_b200_logp_f32(null, null, one, one, 1, 80, 4, 4, null) == 1
    assert lib.mas_b200_logp_f32(one, null, one, one, 1, 1000, 4, 4, null) == 2
    assert lib.mas_b200_fused_maximum_path_f32(one, null, one, null, one, one, null, null, one, 1 << 30, 1, 80, 4, 4, -1e9, null) == 1
    # host entry: inconsistent lengths are rejected before any CUDA call
    tx = (ctypes.c_int32 * 1)(5)
    ty = (ctypes.c_int32 * 1)(3)
    buf = (ctypes.c_float * 64)()
    out = (ctypes.c_int32 * 64)()
    assert lib.mas_b200_maximum_path_host_i32(out, buf, tx, ty, 1, 8, 8, -1e9, 0) == 6
