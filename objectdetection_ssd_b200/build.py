"""Build libssdhead.so in-tree with nvcc for sm_100a (no torch headers: the library is a plain C ABI).

``python -m objectdetection_ssd_b200.build`` or ``__graft_entry__.build()``.  nvcc cross-compiles
without a GPU; the resulting .so is git-ignored but travels to the GPU box with the tree.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG_DIR, "csrc")
LIB_PATH = os.path.join(PKG_DIR, "libssdhead.so")
SOURCES = ["api.cu", "match.cu", "loss.cu", "detect.cu", "evalmap.cu", "host_ctx.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-lineinfo", "-std=c++17",
    "-fmad=false",                      # IoU / encode must not contract into FMAs (bit-exact indices)
    "-Xcompiler", "-fPIC", "-shared",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: libssdhead.so cannot be built (there is no CPU fallback)")


def sources():
    return [os.path.join(CSRC, s) for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]


def is_stale() -> bool:
    if not os.path.exists(LIB_PATH):
        return True
    t = os.path.getmtime(LIB_PATH)
    deps = sources() + [os.path.join(CSRC, "common.cuh"),
                        os.path.join(os.path.dirname(PKG_DIR), "include", "ssdhead.h")]
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def source_hash() -> str:
    """sha256 (first 16 hex digits) over the kernel sources, the shared headers and the compile flags: identifies what
    a libssdhead.so was built from.  profiles/traffic.json records it next to the ncu numbers, and bench.py quotes
    those numbers only for a library built from the same sources."""
    import hashlib
    h = hashlib.sha256()
    deps = sources() + [os.path.join(CSRC, "common.cuh"), os.path.join(os.path.dirname(PKG_DIR), "include", "ssdhead.h")]
    for path in deps:
        h.update(os.path.basename(path).encode())
        with open(path, "rb") as f:
            h.update(f.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()[:16]


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not is_stale():
        return LIB_PATH
    extra = os.environ.get("SSDHEAD_NVCC_EXTRA", "").split()        # developer builds (e.g. -DSSDHEAD_PHASE_TIMES)
    cmd = [_nvcc()] + NVCC_FLAGS + extra + (["-Xptxas", "-v"] if verbose else []) + ["-o", LIB_PATH] + sources()
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + res.stdout + res.stderr)
    if verbose:
        sys.stderr.write(res.stderr)
    return LIB_PATH


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
