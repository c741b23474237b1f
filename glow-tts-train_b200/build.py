"""Build libmas_b200.so (the C-ABI library of include/mas_b200.h) in-tree with nvcc for sm_100a.

    python glow-tts-train_b200/build.py [--force] [--verbose]

nvcc cross-compiles without a GPU; the resulting .so sits next to this file so that it travels to
the GPU box with the repository snapshot.  The CUDA runtime is linked as a SHARED library
(`-cudart shared`; libcudart.so.12, found through the process -- torch has loaded its copy by the
time the binding loads this library -- or through the rpath /usr/local/cuda/lib64): the static
runtime would embed every runtime entry-point name in the product, including batch-copy calls the
library never makes.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from pathlib import Path

PKG_DIR = Path(__file__).resolve().parent
REPO = PKG_DIR.parent
CSRC = PKG_DIR / "csrc"
INCLUDE = REPO / "include"
LIB_PATH = PKG_DIR / "libmas_b200.so"

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC",
    # bit-exactness: no fast-math, no flush-to-zero, IEEE div/sqrt (these are the defaults; spelled
    # out so nobody "optimises" them away).  FMA contraction is left on: the DP has no a*b+c
    # pattern, and the logp contraction spells its FMAs out (fmaf / fma.rn.f32x2).
    "--ftz=false", "--prec-div=true", "--prec-sqrt=true",
]
OBJ_DIR = PKG_DIR / "build"          # git-ignored object files, one per translation unit


def find_nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and Path(cand).exists():
            return cand
    raise RuntimeError("nvcc not found (set NVCC=/path/to/nvcc)")


def sources() -> list[Path]:
    return sorted(CSRC.glob("*.cu"))


def needs_build() -> bool:
    if not LIB_PATH.exists():
        return True
    t = LIB_PATH.stat().st_mtime
    deps = list(CSRC.glob("*.cu")) + list(CSRC.glob("*.cuh")) + list(INCLUDE.glob("*.h")) + [Path(__file__)]
    return any(d.stat().st_mtime > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> Path:
    if not force and not needs_build():
        return LIB_PATH
    from concurrent.futures import ThreadPoolExecutor

    nvcc = find_nvcc()
    OBJ_DIR.mkdir(exist_ok=True)
    common = [nvcc, *NVCC_FLAGS, f"-I{INCLUDE}", f"-I{CSRC}"] + (["-Xptxas", "-v"] if verbose else [])
    headers = list(CSRC.glob("*.cuh")) + list(INCLUDE.glob("*.h")) + [Path(__file__)]
    newest_header = max(h.stat().st_mtime for h in headers)

    def compile_one(src: Path):
        obj = OBJ_DIR / (src.stem + ".o")
        cmd = common + ["-c", "-o", str(obj), str(src)]
        # a translation unit is rebuilt when it or any header changed (the template-heavy ones take minutes)
        if not force and not verbose and obj.exists() and obj.stat().st_mtime > max(src.stat().st_mtime, newest_header):
            return obj, cmd, subprocess.CompletedProcess(cmd, 0, "", "")
        return obj, cmd, subprocess.run(cmd, capture_output=True, text=True)

    # the translation units are independent and the template-heavy ones take a minute each: in parallel
    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as pool:
        results = list(pool.map(compile_one, sources()))
    objs = []
    for obj, cmd, proc in results:
        if verbose or proc.returncode != 0:
            sys.stderr.write(proc.stdout + proc.stderr)
        if proc.returncode != 0:
            raise RuntimeError(f"nvcc failed ({proc.returncode}): {' '.join(cmd)}")
        objs.append(str(obj))
    tmp = LIB_PATH.with_suffix(".so.tmp")
    link = [nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-cudart", "shared",
            "-Xlinker", "-rpath=/usr/local/cuda/lib64", "-o", str(tmp), *objs]
    proc = subprocess.run(link, capture_output=True, text=True)
    if verbose or proc.returncode != 0:
        sys.stderr.write(proc.stdout + proc.stderr)
    if proc.returncode != 0:
        tmp.unlink(missing_ok=True)
        raise RuntimeError(f"nvcc link failed ({proc.returncode}): {' '.join(link)}")
    tmp.replace(LIB_PATH)
    return LIB_PATH


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
