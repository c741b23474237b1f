"""Build recipe for the oracle -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Two things are built, both only ever used as checkers / reported CPU baselines:

1. ``oracle/_build/libmas_oracle.so`` -- the plain-C restatement in ``oracle/mas_oracle.c``
   (gcc, ``-O2 -ffp-contract=off -fopenmp``).  Always buildable.

2. ``oracle/_ref/{serial,omp}/core.<abi>.so`` -- the REFERENCE's own kernel,
   ``/root/reference/glow_tts_train/monotonic_align/core.pyx``, cythonized and compiled from where
   it lies (no source is copied into this repository; the generated C goes to ``oracle/_ref/``
   which is git-ignored but travels to the GPU box).  Two flavours, as SURVEY.md 7.1 describes:
   ``serial`` is what the reference's ``monotonic_align/setup.py:9-13`` produces (no ``-fopenmp``
   so ``prange`` is a plain loop) and ``omp`` adds ``-fopenmp`` (what north_star calls "the
   OpenMP Cython").  Built only when ``/root/reference`` exists (i.e. in the build container);
   on the GPU box the prebuilt files are used.

3. ``oracle/_ref/refpkg.zip`` -- the reference's Python package (the modules the model and its
   training step need, unmodified) + its compiled OpenMP kernel, packed for the model-level checks on
   the GPU box (``stage_reference_package``; git-ignored like the rest of ``oracle/_ref``).

The shipped ``core.c`` of the reference does not compile on Python 3.12 (it includes the removed
``longintrepr.h``), so ``core.pyx`` is re-cythonized with the installed Cython, passing
``legacy_implicit_noexcept=True`` to keep Cython-0.29 ``nogil`` call semantics.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
import sysconfig
from pathlib import Path

HERE = Path(__file__).resolve().parent
BUILD_DIR = HERE / "_build"
REF_DIR = HERE / "_ref"
REFERENCE_ROOT = Path(os.environ.get("MAS_REFERENCE_ROOT", "/root/reference"))
REFERENCE_PYX = REFERENCE_ROOT / "glow_tts_train" / "monotonic_align" / "core.pyx"
EXT_SUFFIX = sysconfig.get_config_var("EXT_SUFFIX")
GCC = "/usr/bin/gcc" if Path("/usr/bin/gcc").exists() else "gcc"


def _newer(target: Path, *sources: Path) -> bool:
    if not target.exists():
        return False
    t = target.stat().st_mtime
    return all(s.exists() and s.stat().st_mtime <= t for s in sources)


def build_c_oracle(force: bool = False) -> Path:
    src = HERE / "mas_oracle.c"
    out = BUILD_DIR / "libmas_oracle.so"
    if not force and _newer(out, src):
        return out
    BUILD_DIR.mkdir(exist_ok=True)
    cmd = [GCC, "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-fopenmp", "-Wall",
           "-o", str(out), str(src), "-lm"]
    subprocess.run(cmd, check=True)
    return out


def ref_so(flavour: str) -> Path:
    return REF_DIR / flavour / f"core{EXT_SUFFIX}"


def build_reference(force: bool = False) -> dict:
    """Compile the reference's core.pyx (from /root/reference) into oracle/_ref/.  Returns
    {flavour: path} for what exists afterwards; silently keeps prebuilt files when the reference
    tree is absent (the GPU box)."""
    result = {}
    have_src = REFERENCE_PYX.exists()
    for flavour, omp in (("serial", False), ("omp", True)):
        out = ref_so(flavour)
        if have_src and (force or not _newer(out, REFERENCE_PYX)):
            import numpy  # noqa: WPS433 (build-time only)

            out.parent.mkdir(parents=True, exist_ok=True)
            c_file = out.parent / "core.c"
            subprocess.run(
                [sys.executable, "-m", "cython", "-3", "-X", "legacy_implicit_noexcept=True",
                 str(REFERENCE_PYX), "-o", str(c_file)],
                check=True,
            )
            cmd = [GCC, "-O3", "-fPIC", "-shared", "-fwrapv", "-fno-strict-aliasing",
                   "-DNPY_NO_DEPRECATED_API=NPY_1_7_API_VERSION",
                   f"-I{sysconfig.get_paths()['include']}", f"-I{numpy.get_include()}"]
            if omp:
                cmd.append("-fopenmp")
            cmd += ["-o", str(out), str(c_file)]
            subprocess.run(cmd, check=True)
            c_file.unlink()  # generated C embeds the .pyx text as comments: do not keep it around
        if out.exists():
            result[flavour] = out
    return result


REF_PKG_ZIP = REF_DIR / "refpkg.zip"     # the reference's Python package, staged as ONE archive for the GPU box

# The modules the model-level checks import (SURVEY.md 8c "model-level oracle", 8d C5); the CLI, dataset,
# export and inference modules are not needed and stay where they are.
_REF_MODULES = ("__init__.py", "attentions.py", "checkpoint.py", "config.py", "layers.py", "models.py", "optimize.py",
                "train.py", "utils.py", "monotonic_align/__init__.py")


def stage_reference_package(force: bool = False):
    """the model-level checks run the reference's OWN FlowGenerator and
    train_step with `monotonic_align` swapped, on the GPU box -- where /root/reference does not
    exist.  So the build container packs the modules they need, unmodified, together with the
    reference's compiled OpenMP kernel and a two-line `dataclasses_json` stub (config.py:8 only needs
    the mixin's name) into ONE archive under the git-ignored oracle/_ref/ (it travels with the snapshot
    like the compiled kernel; nothing of it enters the repository's history or its working tree as
    source).  oracle/ref_model.py unpacks it into a temporary directory when a test imports it.
    Returns the archive's path, or None when neither it nor /root/reference exists."""
    import zipfile

    src = REFERENCE_ROOT / "glow_tts_train"
    so = ref_so("omp")
    stale = REF_PKG_ZIP.exists() and so.exists() and REF_PKG_ZIP.stat().st_mtime < so.stat().st_mtime
    if src.exists() and (force or stale or not REF_PKG_ZIP.exists()):
        REF_DIR.mkdir(parents=True, exist_ok=True)
        tmp = REF_PKG_ZIP.with_suffix(".zip.tmp")
        with zipfile.ZipFile(tmp, "w", zipfile.ZIP_DEFLATED) as z:
            for m in _REF_MODULES:
                z.write(src / m, f"glow_tts_train/{m}")
            if so.exists():
                z.write(so, f"glow_tts_train/monotonic_align/{so.name}")
            z.writestr("dataclasses_json/__init__.py", "class DataClassJsonMixin:\n    pass\n")
        tmp.replace(REF_PKG_ZIP)
    return REF_PKG_ZIP if REF_PKG_ZIP.exists() else None


def build_all(force: bool = False) -> dict:
    info = {"c_oracle": build_c_oracle(force)}
    info.update(build_reference(force))
    pkg = stage_reference_package(force)
    if pkg is not None:
        info["reference_package"] = pkg
    return info


def clean() -> None:
    for d in (BUILD_DIR, REF_DIR):
        shutil.rmtree(d, ignore_errors=True)


if __name__ == "__main__":
    if "--clean" in sys.argv:
        clean()
    for key, path in build_all(force="--force" in sys.argv).items():
        print(f"{key}: {path}")
