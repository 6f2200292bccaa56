"""Builds libssq_b200.so (sm_100a CUDA kernels + C-ABI) in-tree with nvcc.

The library has no torch dependency: plain pointers in, kernels enqueued on the stream handed in.
Run as `python -m shiftedscalequantization_b200.build` or through `__graft_entry__.build()`.
"""
from __future__ import annotations

import hashlib
import os
import subprocess
import sys
from pathlib import Path

PKG_DIR = Path(__file__).resolve().parent
CSRC = PKG_DIR / "csrc"
LIB_PATH = CSRC / "libssq_b200.so"
STAMP = CSRC / ".libssq_b200.stamp"
SOURCES = ["fq_affine.cu", "fq_adaround.cu", "fq_shift.cu", "scale_search.cu", "recon_loss.cu", "loop.cu", "export_codes.cu", "exchange.cu"]
HEADERS = [CSRC / "ssq_common.cuh", CSRC / "ssq_fastdiv.h", CSRC / "ssq_slab_plan.h", PKG_DIR.parent / "include" / "ssq_b200.h"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "--shared", "-Xcompiler", "-fPIC",
    "-Xptxas", "-v",
]


def _nvcc() -> str:
    cand = os.environ.get("NVCC") or "/usr/local/cuda/bin/nvcc"
    return cand if Path(cand).exists() else "nvcc"


def _digest() -> str:
    h = hashlib.sha256()
    for f in [CSRC / s for s in SOURCES] + HEADERS:
        h.update(f.read_bytes())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def build(force: bool = False, verbose: bool = False) -> Path:
    """Compile every kernel for sm_100a into one shared library; returns its path."""
    digest = _digest()
    if not force and LIB_PATH.exists() and STAMP.exists() and STAMP.read_text().strip() == digest:
        return LIB_PATH
    cmd = [_nvcc(), *NVCC_FLAGS, "-o", str(LIB_PATH), *[str(CSRC / s) for s in SOURCES]]
    proc = subprocess.run(cmd, capture_output=True, text=True)
    log = proc.stdout + proc.stderr
    (CSRC / "build.log").write_text(" ".join(cmd) + "\n" + log)
    if proc.returncode != 0:
        sys.stderr.write(log)
        raise RuntimeError(f"nvcc failed ({proc.returncode}); see {CSRC / 'build.log'}")
    if verbose:
        print(log)
    STAMP.write_text(digest)
    return LIB_PATH


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
