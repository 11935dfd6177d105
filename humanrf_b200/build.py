"""Builds libhumanrf_b200.so (sm_100a) in-tree with nvcc.  No torch involvement: the library is
a plain C-ABI shared object (include/humanrf_b200.h)."""
from __future__ import annotations

import hashlib
import os
import subprocess
import sys
from pathlib import Path

PKG = Path(__file__).resolve().parent
CSRC = PKG / "csrc"
LIB = PKG / "libhumanrf_b200.so"
STAMP = PKG / ".libhumanrf_b200.stamp"
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "-shared", "--expt-relaxed-constexpr", "-Xptxas", "-v",
]


def _sources():
    return sorted(CSRC.glob("*.cu"))


def _digest() -> str:
    h = hashlib.sha256()
    for f in sorted(list(CSRC.glob("*.cu")) + list(CSRC.glob("*.cuh")) + [PKG.parent / "include" / "humanrf_b200.h"]):
        h.update(f.name.encode())
        h.update(f.read_bytes())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def nvcc_path() -> str:
    for c in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if c and (os.path.isabs(c) and os.path.exists(c) or not os.path.isabs(c)):
            return c
    return "nvcc"


def build_library(force: bool = False, verbose: bool = False) -> Path:
    dig = _digest()
    if not force and LIB.exists() and STAMP.exists() and STAMP.read_text().strip() == dig:
        return LIB
    cmd = [nvcc_path(), *NVCC_FLAGS, "-o", str(LIB), *map(str, _sources())]
    res = subprocess.run(cmd, capture_output=True, text=True)
    log = res.stdout + res.stderr
    (PKG / "build.log").write_text(log)
    if res.returncode != 0:
        sys.stderr.write(log)
        raise RuntimeError("nvcc failed building libhumanrf_b200.so (see humanrf_b200/build.log)")
    if verbose:
        print(log)
    STAMP.write_text(dig)
    return LIB


if __name__ == "__main__":
    print(build_library(force="--force" in sys.argv, verbose=True))
