"""Build the CUDA extension in-tree: ``mettagrid_b200/libmettagrid_b200.so`` (sm_100a only)."""

from __future__ import annotations

import os
import shutil
import subprocess
from pathlib import Path

PKG = Path(__file__).resolve().parent
CSRC = PKG / "csrc"
LIB = PKG / "libmettagrid_b200.so"
SOURCES = ["mg_kernels.cu", "mg_fast.cu", "mg_gridobs.cu", "mg_vecenv.cu", "mg_capi.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "--fmad=false",  # the reference's float math is unfused mul/add
    "-Xcompiler", "-fPIC",
]  # fmt: skip


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and Path(cand).exists():
            return cand
    raise RuntimeError("nvcc not found: the CUDA extension cannot be built")


def needs_build() -> bool:
    if not LIB.exists():
        return True
    deps = list(CSRC.glob("*.cu")) + list(CSRC.glob("*.cuh")) + list((PKG.parent / "include").glob("*.h"))
    return LIB.stat().st_mtime < max(p.stat().st_mtime for p in deps)


def build_native(force: bool = False, verbose: bool = False, out: Path | None = None, extra_flags: list[str] | None = None) -> Path:
    """Build the library.  `out` / `extra_flags` produce an A/B variant (e.g. -DMG_PHASE_TIMING) next to the product
    without touching it: python -m mettagrid_b200.build --variant NAME -DFLAG ... -> variants/lib_NAME.so"""
    if out is None and not force and not needs_build():
        return LIB
    nvcc = _nvcc()
    objs = []
    build_dir = PKG / "build" if out is None else PKG / "build" / Path(out).stem
    build_dir.mkdir(parents=True, exist_ok=True)
    procs = []
    for src in SOURCES:
        obj = build_dir / (Path(src).stem + ".o")
        cmd = [nvcc, *NVCC_FLAGS, *(extra_flags or []), "-c", str(CSRC / src), "-o", str(obj)]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        procs.append((cmd, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(str(obj))
    for cmd, p in procs:
        log, _ = p.communicate(timeout=1200)
        if verbose and log:
            print(log)
        if p.returncode != 0:
            raise RuntimeError(f"nvcc failed: {' '.join(cmd)}\n{log}")
    target = LIB if out is None else Path(out)
    target.parent.mkdir(parents=True, exist_ok=True)
    subprocess.check_call([nvcc, "-shared", "-o", str(target), *objs])
    if out is not None:  # a variant's objects are of no further use (and would travel to the GPU box with every snapshot)
        shutil.rmtree(build_dir, ignore_errors=True)
    return target


if __name__ == "__main__":
    import sys

    if "--variant" in sys.argv:
        i = sys.argv.index("--variant")
        name, flags = sys.argv[i + 1], [a for a in sys.argv[i + 2 :] if a != "-v"]
        print(build_native(verbose="-v" in sys.argv, out=PKG.parent / "variants" / f"lib_{name}.so", extra_flags=flags))
    else:
        print(build_native(force="--force" in sys.argv, verbose="-v" in sys.argv))
