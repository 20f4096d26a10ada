"""ctypes loader for the CPU oracle (oracle/libmandel_oracle.so) + a numpy restatement.

TEST INFRASTRUCTURE ONLY.  Importable from tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs -- never from distributedmandelbrot_b200/ (the product).

The C oracle restates DistributedMandelbrotWorkerCUDA/DistributedMandelbrotWorkerCUDA.py
:19-37 (coordinates), :39-68 (iteration), :70-100 (geometry + uint8 epilogue); see
mandel_oracle.h for the evaluation rule and how parity is pinned (tests/golden/).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libmandel_oracle.so")
_lib = None


def build(force: bool = False) -> str:
    """Compile the oracle with its Makefile (gcc -O2 -ffp-contract=off)."""
    srcs = [os.path.join(_HERE, f) for f in ("mandel_oracle.c", "mandel_oracle_simd.c", "mandel_oracle.h", "Makefile")]
    stale = (not os.path.exists(_SO)) or os.path.getmtime(_SO) < max(os.path.getmtime(f) for f in srcs)
    if force or stale:
        subprocess.check_call(["make", "-s", "-C", _HERE, "-B", "libmandel_oracle.so"])
    return _SO


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        L = C.CDLL(_SO)
        d, u32, i32, u64, sz = C.c_double, C.c_uint32, C.c_int32, C.c_uint64, C.c_size_t
        pd, pi32, pu8 = C.POINTER(C.c_double), C.POINTER(C.c_int32), C.POINTER(C.c_uint8)
        L.mbo_chunk_geometry.argtypes = [u32, u32, u32, pd, pd, pd]
        L.mbo_chunk_geometry.restype = None
        L.mbo_linspace.argtypes = [d, d, u32, pd]
        L.mbo_linspace.restype = None
        L.mbo_escape_f64.argtypes = [d, d, i32]
        L.mbo_escape_f64.restype = i32
        L.mbo_escape_f32.argtypes = [C.c_float, C.c_float, i32]
        L.mbo_escape_f32.restype = i32
        L.mbo_count_to_u8.argtypes = [i32, i32]
        L.mbo_count_to_u8.restype = C.c_uint8
        L.mbo_render_rect.argtypes = [d, d, d, u32, i32, u32, u32, u32, u32, pi32, sz, C.c_int, C.c_int]
        L.mbo_render_rect.restype = C.c_int
        L.mbo_process_workload.argtypes = [u32, i32, u32, u32, u32, pu8, pi32, C.c_int]
        L.mbo_process_workload.restype = C.c_int
        L.mbo_render_frame.argtypes = [u32, u32, u32, u32, u32, i32, C.c_int, C.c_int, pi32, C.POINTER(u64)]
        L.mbo_render_frame.restype = C.c_int
        L.mbo_sticky_violations.argtypes = [pd, pd, sz, i32, C.c_int, d]
        L.mbo_sticky_violations.restype = u64
        L.mbo_sticky_violations_f32.argtypes = [pd, pd, sz, i32, C.c_int, C.c_float]
        L.mbo_sticky_violations_f32.restype = u64
        L.mbo_counts_to_u8.argtypes = [pi32, pu8, sz, i32]
        L.mbo_counts_to_u8.restype = None
        L.mbo_useful_iterations.argtypes = [pi32, sz, i32]
        L.mbo_useful_iterations.restype = u64
        L.mbo_fnv1a_words.argtypes = [pi32, sz]
        L.mbo_fnv1a_words.restype = u64
        L.mbo_max_threads.argtypes = []
        L.mbo_max_threads.restype = C.c_int
        L.mbo_simd_render_rect.argtypes = L.mbo_render_rect.argtypes
        L.mbo_simd_render_rect.restype = C.c_int
        L.mbo_simd_render_frame.argtypes = L.mbo_render_frame.argtypes
        L.mbo_simd_render_frame.restype = C.c_int
        L.mbo_simd_width.argtypes = []
        L.mbo_simd_width.restype = C.c_int
        _lib = L
    return _lib


def _p(a: np.ndarray, ct):
    return a.ctypes.data_as(C.POINTER(ct))


def chunk_geometry(level: int, index_real: int, index_imag: int):
    """(.py:75-78) -> (range, start_r, start_i) as Python floats."""
    r, sr, si = C.c_double(), C.c_double(), C.c_double()
    lib().mbo_chunk_geometry(level, index_real, index_imag, C.byref(r), C.byref(sr), C.byref(si))
    return r.value, sr.value, si.value


def linspace(start: float, stop: float, d: int) -> np.ndarray:
    out = np.empty(d, dtype=np.float64)
    lib().mbo_linspace(start, stop, d, _p(out, C.c_double))
    return out


def escape_f64(r: float, i: float, mrd: int) -> int:
    return int(lib().mbo_escape_f64(r, i, mrd))


def render_rect(start_r, start_i, rng, d, mrd, x0=0, y0=0, w=None, h=None, dtype="f64", threads=0,
                simd=False) -> np.ndarray:
    """int32 counts [h, w] for the sub-rectangle of the chunk-level linspace grid.
    simd=True: the data-parallel build of the same loop (mandel_oracle_simd.c), bit-identical."""
    w = d - x0 if w is None else w
    h = d - y0 if h is None else h
    out = np.empty((h, w), dtype=np.int32)
    fn = lib().mbo_simd_render_rect if simd else lib().mbo_render_rect
    rc = fn(start_r, start_i, rng, d, mrd, x0, y0, w, h, _p(out, C.c_int32), w, {"f64": 0, "f32": 1}[dtype], threads)
    if rc != 0:
        raise ValueError(f"mbo_render_rect failed rc={rc}")
    return out


def render_chunk_counts(level, mrd, index_real, index_imag, d=4096, dtype="f64", threads=0,
                        x0=0, y0=0, w=None, h=None, simd=False) -> np.ndarray:
    rng, sr, si = chunk_geometry(level, index_real, index_imag)
    return render_rect(sr, si, rng, d, mrd, x0, y0, w, h, dtype, threads, simd)


def simd_width() -> int:
    return int(lib().mbo_simd_width())


def process_workload(level, mrd, index_real, index_imag, d=4096, threads=0, want_counts=False):
    """(.py:70-100) -> uint8[d*d] (and int32 counts[d*d] if want_counts)."""
    out = np.empty(d * d, dtype=np.uint8)
    counts = np.empty(d * d, dtype=np.int32) if want_counts else None
    rc = lib().mbo_process_workload(level, mrd, index_real, index_imag, d, _p(out, C.c_uint8),
                                    _p(counts, C.c_int32) if want_counts else None, threads)
    if rc != 0:
        raise ValueError(f"mbo_process_workload failed rc={rc}")
    return (out, counts) if want_counts else out


def render_frame(level, ir0, ii0, blocks, d, mrd, dtype="f64", threads=0, want_counts=True, simd=False):
    """Whole frame on all cores -> (counts[blocks*blocks, d, d] or None, useful pixel-iterations)."""
    out = np.empty((blocks * blocks, d, d), dtype=np.int32) if want_counts else None
    useful = C.c_uint64()
    rc = (lib().mbo_simd_render_frame if simd else lib().mbo_render_frame)(level, ir0, ii0, blocks, d, mrd, {"f64": 0, "f32": 1}[dtype], threads,
                                _p(out, C.c_int32) if want_counts else None, C.byref(useful))
    if rc != 0:
        raise ValueError(f"mbo_render_frame failed rc={rc}")
    return out, int(useful.value)


def sticky_violations(r, i, mrd: int, extra: int = 7, band: float = 2.0 ** -36) -> int:
    """Points where the computed |z|^2 falls below 4 (or overflows) within `extra` updates of first
    reaching 4; the CUDA kernel's deferred escape test relies on this never happening."""
    r = np.ascontiguousarray(r, dtype=np.float64).ravel()
    i = np.ascontiguousarray(i, dtype=np.float64).ravel()
    return int(lib().mbo_sticky_violations(_p(r, C.c_double), _p(i, C.c_double), r.size, mrd, extra, band))


def sticky_violations_f32(r, i, mrd: int, extra: int = 15, band: float = 2.0 ** -9) -> int:
    r = np.ascontiguousarray(r, dtype=np.float64).ravel()
    i = np.ascontiguousarray(i, dtype=np.float64).ravel()
    return int(lib().mbo_sticky_violations_f32(_p(r, C.c_double), _p(i, C.c_double), r.size, mrd, extra, band))


def counts_to_u8(counts: np.ndarray, mrd: int) -> np.ndarray:
    c = np.ascontiguousarray(counts, dtype=np.int32).ravel()
    out = np.empty(c.size, dtype=np.uint8)
    lib().mbo_counts_to_u8(_p(c, C.c_int32), _p(out, C.c_uint8), c.size, mrd)
    return out.reshape(np.shape(counts))


def useful_iterations(counts: np.ndarray, mrd: int) -> int:
    c = np.ascontiguousarray(counts, dtype=np.int32).ravel()
    return int(lib().mbo_useful_iterations(_p(c, C.c_int32), c.size, mrd))


def fnv1a_words(counts: np.ndarray) -> int:
    c = np.ascontiguousarray(counts, dtype=np.int32).ravel()
    return int(lib().mbo_fnv1a_words(_p(c, C.c_int32), c.size))


def max_threads() -> int:
    return int(lib().mbo_max_threads())


# --------------------------------------------------------------------------------------
# numpy restatement (independent of the C code; used to cross-check the C oracle on small
# cases).  Vectorised over pixels; each numpy ufunc call is one individually-rounded IEEE op.
# --------------------------------------------------------------------------------------
def np_counts(r: np.ndarray, i: np.ndarray, mrd: int, dtype=np.float64) -> np.ndarray:
    """(.py:39-68) over broadcast arrays r, i -> int32 counts."""
    c0 = np.asarray(r, dtype=np.float64).astype(dtype)
    c1 = np.asarray(i, dtype=np.float64).astype(dtype)
    c0, c1 = np.broadcast_arrays(c0, c1)
    z0, z1 = c0.copy(), c1.copy()
    out = np.zeros(c0.shape, dtype=np.int32)
    live = np.ones(c0.shape, dtype=bool)
    two, four = dtype(2), dtype(4)
    with np.errstate(over="ignore", invalid="ignore"):
        for n in range(1, mrd):
            sq0 = z0 * z0 - z1 * z1
            sq1 = (two * z0) * z1
            z0 = sq0 + c0
            z1 = sq1 + c1
            m = z0 * z0 + z1 * z1
            esc = live & (m >= four)
            out[esc] = n
            live &= ~esc
            if not live.any():
                break
    return out


def np_process_counts(level, mrd, index_real, index_imag, d, dtype=np.float64) -> np.ndarray:
    """(.py:70-94) with numpy's own linspace; int32 counts [d, d] (row = imag index)."""
    chunk_range = (2 - (-2)) / level
    start_r = -2 + (chunk_range * index_real)
    start_i = -2 + (chunk_range * index_imag)
    r = np.linspace(start=start_r, stop=start_r + chunk_range, num=d)
    i = np.linspace(start=start_i, stop=start_i + chunk_range, num=d)
    return np_counts(r[None, :], i[:, None], mrd, dtype)


def np_counts_to_u8(counts: np.ndarray, mrd: int) -> np.ndarray:
    """(.py:96-98) literally."""
    out = (counts.astype(np.float64) * 256) / mrd
    with np.errstate(invalid="ignore"):
        return np.ceil(out).astype(np.uint8)
