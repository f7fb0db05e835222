"""In-tree build of libtoist_b200.so (sm_100a only) and of the oracle's small C helpers.

`python -m toist_b200.build` or `__graft_entry__.build()`.  nvcc cross-compiles without a GPU; the resulting .so is
git-ignored but travels to the GPU box with the repo snapshot.
"""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
CSRC = ROOT / "toist_b200" / "csrc"
BUILD = ROOT / "build"
LIB = ROOT / "toist_b200" / "libtoist_b200.so"

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "--expt-relaxed-constexpr",
    "-Xcompiler", "-fPIC",
    "-Xptxas", "-v",
    "-I", str(ROOT / "include"),
    "-I", str(CSRC),
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and Path(cand).exists():
            return cand
    raise RuntimeError("nvcc not found: toist_b200 has no CPU fallback and cannot be built without the CUDA toolkit")


def _stamp(src: Path) -> str:
    h = hashlib.sha1()
    h.update(src.read_bytes())
    for hdr in sorted(list(CSRC.glob("*.cuh")) + list(CSRC.glob("*.h")) + [ROOT / "include" / "toist_b200.h"]):
        h.update(hdr.read_bytes())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def _compile_one(nvcc: str, src: Path, verbose: bool) -> Path:
    obj = BUILD / (src.stem + ".o")
    stamp = BUILD / (src.stem + ".stamp")
    want = _stamp(src)
    if obj.exists() and stamp.exists() and stamp.read_text() == want:
        return obj
    cmd = [nvcc, *NVCC_FLAGS, "-c", str(src), "-o", str(obj)]
    res = subprocess.run(cmd, capture_output=True, text=True)
    (BUILD / (src.stem + ".ptxas.log")).write_text(res.stderr)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError(f"nvcc failed on {src.name}")
    if verbose:
        print(f"[toist_b200.build] compiled {src.name}")
    stamp.write_text(want)
    return obj


def build_library(verbose: bool = True) -> Path:
    BUILD.mkdir(exist_ok=True)
    nvcc = _nvcc()
    srcs = sorted(CSRC.glob("*.cu")) + sorted(CSRC.glob("*.cpp"))
    with ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        objs = list(ex.map(lambda s: _compile_one(nvcc, s, verbose), srcs))
    newest = max(o.stat().st_mtime for o in objs)
    if LIB.exists() and LIB.stat().st_mtime >= newest:
        return LIB
    cmd = [nvcc, "-shared", "-cudart", "static", "-gencode", "arch=compute_100a,code=sm_100a",
           "-o", str(LIB), *map(str, objs)]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
        raise RuntimeError("link of libtoist_b200.so failed")
    if verbose:
        print(f"[toist_b200.build] linked {LIB}")
    return LIB


if __name__ == "__main__":
    build_library()
