"""Builds the in-tree C-ABI shared library with nvcc for sm_100a (no JIT cache: the .so must
travel with the tree).  `python -m distributedmandelbrot_b200.build [--force]`."""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG, "csrc")
SO = os.path.join(PKG, "libmandelbrot_b200.so")
STAMP = SO + ".srchash"
SOURCES = ["mandelbrot_capi.cu", "tile_wsq.cpp", "host_frame.cpp"]
HEADERS = ["escape_kernels.cuh", "escape_refill.cuh", "chunk_codec.cuh", os.path.join("..", "..", "include", "mandelbrot_b200.h")]

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "-fmad=false",                      # never contract a*b+c (parity; the kernels also use *_rn intrinsics)
    "-prec-div=true", "-prec-sqrt=true", "-ftz=false",
    "-Xcompiler", "-fPIC,-O2,-ffp-contract=off,-Wall",
    "-shared", "-cudart", "shared",
]


def nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def have_nvcc() -> bool:
    try:
        nvcc()
        return True
    except RuntimeError:
        return False


def source_hash() -> str:
    """sha256 over the sources, headers and flags the library is built from."""
    import hashlib
    h = hashlib.sha256(" ".join(NVCC_FLAGS).encode())
    for d in SOURCES + HEADERS:
        with open(os.path.join(CSRC, d), "rb") as f:
            h.update(d.encode() + b"\0" + f.read())
    return h.hexdigest()


def stale() -> bool:
    """True when the .so was not built from the sources as they are now.  Content hash in a sidecar
    file, not mtimes: the tree is copied to the GPU box and copies do not keep file times."""
    if not os.path.exists(SO):
        return True
    try:
        with open(STAMP) as f:
            return f.read().strip() != source_hash()
    except OSError:
        return True


def build(force: bool = False, verbose: bool = False) -> str:
    if not (force or stale()):
        return SO
    tmp = f"{SO}.{os.getpid()}.tmp"          # several ranks may get here at once: build aside, rename atomically
    cmd = [nvcc()] + NVCC_FLAGS + (["-Xptxas", "-v"] if verbose else []) + \
          ["-o", tmp] + [os.path.join(CSRC, s) for s in SOURCES] + ["-lrt"]
    env = dict(os.environ)
    env.pop("CC", None)   # $CC in this image points at a wrapper gcc nvcc must not pick up
    env.pop("CXX", None)
    res = subprocess.run(cmd, env=env, capture_output=True, text=True)
    if verbose or res.returncode != 0:
        sys.stderr.write(res.stdout + res.stderr)
    if res.returncode != 0:
        if os.path.exists(tmp):
            os.remove(tmp)
        raise RuntimeError("nvcc failed: " + " ".join(cmd))
    os.replace(tmp, SO)
    with open(STAMP + ".tmp", "w") as f:
        f.write(source_hash() + "\n")
    os.replace(STAMP + ".tmp", STAMP)
    return SO


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv or "--verbose" in sys.argv))
