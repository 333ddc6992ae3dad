"""In-tree build of libb200spec.so (sm_100a only).

    python -m audio_tabs_b200.build [--force] [--verbose]

nvcc cross-compiles without a GPU.  The translation units compile in parallel; the shared
library lands in ``audio_tabs_b200/lib/libb200spec.so`` (git-ignored, shipped to the GPU box by
gpurun).  Nothing is JIT-compiled at import time: a missing library is a hard error in _ffi.py.
"""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

PKG = Path(__file__).resolve().parent
CSRC = PKG / "csrc"
LIBDIR = PKG / "lib"
OBJDIR = PKG / "build"
LIB = LIBDIR / "libb200spec.so"

SOURCES = ["b200spec.cu", "front_f1024.cu", "front_f2048.cu", "front_f4096.cu", "front_f8192.cu", "front_multi.cu"]
NVCC_FLAGS = [
    "-O3", "-std=c++17", "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo",
    "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden", "--expt-relaxed-constexpr",
    "-DB200SPEC_BUILD",
] + os.environ.get("B200SPEC_EXTRA_NVCC_FLAGS", "").split()   # e.g. -DB2_GROUPS=3 for tuning experiments


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and Path(cand).exists():
            return cand
    raise RuntimeError("nvcc not found; libb200spec.so cannot be built")


def _fingerprint() -> str:
    h = hashlib.sha256()
    for p in sorted(list(CSRC.glob("*.cu")) + list(CSRC.glob("*.cuh")) +
                    [PKG.parent / "include" / "b200spec.h", Path(__file__)]):
        h.update(p.name.encode())
        h.update(p.read_bytes())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def build(force: bool = False, verbose: bool = False) -> Path:
    stamp = LIBDIR / "libb200spec.stamp"
    fp = _fingerprint()
    if not force and LIB.exists() and stamp.exists() and stamp.read_text().strip() == fp:
        return LIB
    nvcc = _nvcc()
    LIBDIR.mkdir(exist_ok=True)
    OBJDIR.mkdir(exist_ok=True)
    extra = ["-Xptxas", "-v"] if verbose else []

    def compile_one(src: str) -> Path:
        obj = OBJDIR / (Path(src).stem + ".o")
        cmd = [nvcc, *NVCC_FLAGS, *extra, "-c", str(CSRC / src), "-o", str(obj)]
        res = subprocess.run(cmd, capture_output=True, text=True)
        if verbose or res.returncode != 0:
            sys.stderr.write(f"$ {' '.join(cmd)}\n{res.stdout}{res.stderr}\n")
        if res.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src}")
        return obj

    with ThreadPoolExecutor(max_workers=min(len(SOURCES), os.cpu_count() or 1)) as ex:
        objs = list(ex.map(compile_one, SOURCES))
    cmd = [nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", str(LIB), *map(str, objs)]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError("link of libb200spec.so failed")
    stamp.write_text(fp)
    return LIB


if __name__ == "__main__":
    out = build(force="--force" in sys.argv, verbose="--verbose" in sys.argv)
    print(out)
