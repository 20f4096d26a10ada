"""ctypes binding of the C-ABI in include/mandelbrot_b200.h (libmandelbrot_b200.so).

There is no fallback of any kind here: if the shared library is missing it is built with nvcc
(distributedmandelbrot_b200.build); if that fails, or a compute entry point is called without
a usable CUDA device, a NativeError is raised.  The CPU oracle under oracle/ is test
infrastructure and is never imported from this package.
"""
from __future__ import annotations

import ctypes as C
import os
import threading

from . import build as _build

MB_OK = 0
MB_ERR_INVALID, MB_ERR_CUDA, MB_ERR_NOMEM, MB_ERR_STATE, MB_ERR_SYS = -1, -2, -3, -4, -5
DTYPE_F64, DTYPE_F32 = 0, 1
DTYPES = {"f64": DTYPE_F64, "f32": DTYPE_F32, DTYPE_F64: DTYPE_F64, DTYPE_F32: DTYPE_F32}
CHUNK_DEFINITION = 4096   # DataChunk.dataChunkRange (DataChunk.cs:20) ; .py:80
TILE_PERIOD_SHORTCUT = 1  # mb_tile.flags bit 0: exact orbit-periodicity shortcut (bit-identical counts)
TILE_LANE_REFILL = 2      # mb_tile.flags bit 1: persistent warps with lane refill (bit-identical counts; for deep zooms)


class NativeError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"libmandelbrot_b200: {msg} (status {code})")
        self.code = code


class InvalidArgument(NativeError, ValueError):
    """MB_ERR_INVALID -- mirrors the ArgumentException of the DataChunk ctor (DataChunk.cs:99-106)."""


class Tile(C.Structure):
    """mb_tile (include/mandelbrot_b200.h): a sub-rectangle of one DataChunk's pixel grid."""
    _fields_ = [
        ("start_r", C.c_double), ("start_i", C.c_double),
        ("stop_r", C.c_double), ("stop_i", C.c_double),
        ("step_r", C.c_double), ("step_i", C.c_double),
        ("definition", C.c_uint32), ("mrd", C.c_uint32),
        ("x0", C.c_uint32), ("y0", C.c_uint32), ("w", C.c_uint32), ("h", C.c_uint32),
        ("out_offset", C.c_uint64),
        ("out_pitch", C.c_uint32), ("flags", C.c_uint32),
    ]

    def copy(self) -> "Tile":
        t = Tile()
        C.memmove(C.byref(t), C.byref(self), C.sizeof(Tile))
        return t

    def sub(self, x0: int, y0: int, w: int, h: int, out_offset: int, out_pitch: int | None = None) -> "Tile":
        """Sub-rectangle of this (whole-chunk) tile; indexes the same chunk-level linspace."""
        t = self.copy()
        t.x0, t.y0, t.w, t.h = x0, y0, w, h
        t.out_offset = out_offset
        t.out_pitch = w if out_pitch is None else out_pitch
        return t


assert C.sizeof(Tile) == 88


class ChunkInfo(C.Structure):
    """mb_chunk_info: result record of the device serializer for one chunk."""
    _fields_ = [("out_len", C.c_uint64), ("runs", C.c_uint32), ("codec", C.c_uint32),
                ("entry_type", C.c_uint32), ("first_value", C.c_uint32), ("reserved", C.c_uint32 * 2)]


assert C.sizeof(ChunkInfo) == 32

_lib = None
_lock = threading.RLock()

# name -> (restype, argtypes); every symbol include/mandelbrot_b200.h declares
_u32, _i32, _u64, _sz, _vp, _int = C.c_uint32, C.c_int32, C.c_uint64, C.c_size_t, C.c_void_p, C.c_int
_PT = C.POINTER(Tile)
SIGNATURES = {
    "mb_init": (_int, [_int]),
    "mb_shutdown": (None, []),
    "mb_last_error": (C.c_char_p, []),
    "mb_abi_version": (_int, []),
    "mb_device_count": (_int, []),
    "mb_sm_count": (_int, []),
    "mb_set_period_shortcut": (_int, [_int]),
    "mb_set_tile_flags": (_int, [_u32]),
    "mb_set_executed_counter": (_int, [_vp]),
    "mb_frame_last_stats": (_int, [C.POINTER(C.c_double), C.POINTER(C.c_double)]),
    "mb_hostframe_create": (_int, [C.c_char_p, _u64, _u32, _u32, C.POINTER(_vp)]),
    "mb_hostframe_open": (_int, [C.c_char_p, C.POINTER(_vp)]),
    "mb_hostframe_data": (_vp, [_vp]),
    "mb_hostframe_bytes": (_u64, [_vp]),
    "mb_hostframe_units": (_u32, [_vp]),
    "mb_hostframe_groups": (_u32, [_vp]),
    "mb_hostframe_reset": (_int, [_vp]),
    "mb_hostframe_reset_active": (_int, [_vp, _u32]),
    "mb_hostframe_pin": (_int, [_vp]),
    "mb_hostframe_unpin": (_int, [_vp]),
    "mb_hostframe_mark_done": (_int, [_vp, _u32, _u32, C.POINTER(_u32), _u32, C.POINTER(_u32)]),
    "mb_hostframe_group_left": (C.c_int64, [_vp, _u32]),
    "mb_hostframe_missing": (_int, [_vp, C.POINTER(_u32), _u32, C.POINTER(_u32)]),
    "mb_hostframe_close": (_int, [_vp]),
    "mb_hostframe_unlink": (_int, [C.c_char_p]),
    "mb_frame_render_u8_host": (_int, [_vp, _u32, _u32, _PT, _u32, _u32, _u32, _vp, _int, C.POINTER(_u32), _u32,
                                       C.POINTER(_u32), C.POINTER(_u32), C.POINTER(_u32), _u32, C.POINTER(_u32)]),
    "mb_frame_render_u8_host_units": (_int, [C.POINTER(_u32), _u32, _PT, _u32, _u32, _vp, _int, C.POINTER(_u32), _u32,
                                             C.POINTER(_u32)]),
    "mb_chunk_geometry": (_int, [_u32, _u32, _u32, _u32, _u32, _PT]),
    "mb_tile_coord": (C.c_double, [_PT, _int, _u32]),
    "mb_render_counts": (_int, [_PT, _u32, _vp, _int, _vp]),
    "mb_calc_mb_value_dev": (_int, [_vp, _vp, _sz, _u32, _int, _vp, _vp]),
    "mb_calc_mb_value": (_int, [_vp, _vp, _sz, _u32, _int, _vp]),
    "mb_render_u8": (_int, [_PT, _u32, _vp, _int, _vp]),
    "mb_counts_to_u8": (_int, [_vp, _vp, _sz, _u32, _vp]),
    "mb_useful_iterations": (_int, [_vp, _sz, _u32, _vp, _vp]),
    "mb_classify_u8": (_int, [_vp, _sz, _vp, _vp]),
    "mb_serialize_out_stride": (_sz, [_sz]),
    "mb_serialize_workspace_bytes": (_sz, [_sz, _u32]),
    "mb_serialize_chunks": (_int, [_vp, _sz, _u32, _sz, _vp, _sz, _vp, _sz, _vp, _vp]),
    "mb_serialize_chunk_host": (_int, [_vp, _sz, _vp, _sz, C.POINTER(_u64)]),
    "mb_process_workload_serialized": (_int, [_u32, _u32, _u32, _u32, _u32, _vp, _sz, C.POINTER(_u64), C.POINTER(_i32)]),
    "mb_host_alloc": (_int, [_sz, C.POINTER(_vp)]),
    "mb_host_free": (_int, [_vp]),
    "mb_process_workload_u8": (_int, [_u32, _u32, _u32, _u32, _u32, _vp]),
    "mb_process_workload_counts": (_int, [_u32, _u32, _u32, _u32, _u32, _int, _vp]),
    "mb_process_workloads_u8": (_int, [_vp, _u32, _u32, _vp, _sz]),
    "mb_wsq_create": (_int, [C.c_char_p, _u32, _u32, C.POINTER(_vp)]),
    "mb_wsq_open": (_int, [C.c_char_p, C.POINTER(_vp)]),
    "mb_wsq_reset": (_int, [_vp, _u32]),
    "mb_wsq_pop": (_int, [_vp, _u32, _u32, C.POINTER(_u32), C.POINTER(_u32), C.POINTER(_u32)]),
    "mb_wsq_remaining": (C.c_int64, [_vp]),
    "mb_wsq_close": (_int, [_vp]),
    "mb_frame_render_stolen": (_int, [_vp, _u32, _u32, _PT, _u32, _u32, _u32, _vp, _int, _vp, C.POINTER(_u32), _u32,
                                      C.POINTER(_u32), C.POINTER(_u32)]),
    "mb_wsq_unlink": (_int, [C.c_char_p]),
    "mb_peer_alloc": (_int, [_sz, C.POINTER(_vp), _vp]),
    "mb_peer_open": (_int, [_vp, C.POINTER(_vp)]),
    "mb_peer_close": (_int, [_vp]),
    "mb_peer_free": (_int, [_vp]),
    "mb_measure_fma_peak": (_int, [_int, _int, C.POINTER(C.c_double), C.POINTER(C.c_float)]),
    "mb_measure_pipe_rate": (_int, [_int, _int, _int, C.POINTER(C.c_double), C.POINTER(C.c_float)]),
}


def library_path() -> str:
    return _build.SO


def lib() -> C.CDLL:
    """Load (building first if needed) the shared library and declare every signature."""
    global _lib
    with _lock:
        if _lib is None:
            if not os.path.exists(_build.SO):
                _build.build()
            elif _build.stale() and _build.have_nvcc():
                _build.build()               # edited csrc/ since the last build: never load a stale binary
            L = C.CDLL(_build.SO)
            for name, (res, args) in SIGNATURES.items():
                fn = getattr(L, name)   # AttributeError here = header/library mismatch: fail loudly
                fn.restype = res
                fn.argtypes = args
            if L.mb_abi_version() != 1:
                raise NativeError(MB_ERR_STATE, "ABI version mismatch")
            _lib = L
    return _lib


def check(rc: int) -> int:
    if rc < 0:
        msg = lib().mb_last_error().decode("utf-8", "replace")
        raise (InvalidArgument if rc == MB_ERR_INVALID else NativeError)(rc, msg)
    return rc


_inited = threading.local()


def init(device: int = 0) -> None:
    """mb_init: bind the calling thread to `device`.  Raises NativeError without a CUDA device."""
    check(lib().mb_init(int(device)))
    _inited.device = int(device)


def ensure_init(device: int | None = None) -> int:
    cur = getattr(_inited, "device", None)
    if device is None:
        device = cur if cur is not None else int(os.environ.get("LOCAL_RANK", "0"))
    if cur != device:
        init(device)
    return device


def set_period_shortcut(enabled: bool) -> bool:
    """Switch the one-call entry points (process_workload & co) of THIS thread to the exact
    orbit-periodicity shortcut; returns the previous setting.  Counts are identical either way."""
    return bool(lib().mb_set_period_shortcut(1 if enabled else 0))


def set_tile_flags(flags: int) -> int:
    """Flag word (TILE_PERIOD_SHORTCUT | TILE_LANE_REFILL) of this thread's one-call entry points; -> previous."""
    return check(lib().mb_set_tile_flags(int(flags)))


def chunk_geometry(level: int, index_real: int, index_imag: int, definition: int = CHUNK_DEFINITION,
                   mrd: int = 0, shortcut: bool = False, refill: bool = False) -> Tile:
    """Whole-chunk tile descriptor (.py:75-78 + linspace parameters of .py:24-32).  Host only.
    shortcut=True sets MB_TILE_PERIOD_SHORTCUT on the tile (and on every sub-tile cut from it)."""
    for name, v in (("level", level), ("index_real", index_real), ("index_imag", index_imag),
                    ("definition", definition), ("mrd", mrd)):
        if not (0 <= int(v) <= 0xFFFFFFFF):
            raise InvalidArgument(MB_ERR_INVALID, f"{name}={v} is not a uint32")
    t = Tile()
    check(lib().mb_chunk_geometry(int(level), int(index_real), int(index_imag), int(definition), int(mrd), C.byref(t)))
    t.flags = (TILE_PERIOD_SHORTCUT if shortcut else 0) | (TILE_LANE_REFILL if refill else 0)
    return t


def tile_array(tiles) -> C.Array:
    arr = (Tile * len(tiles))()
    for k, t in enumerate(tiles):
        C.memmove(C.byref(arr, k * C.sizeof(Tile)), C.byref(t), C.sizeof(Tile))
    return arr


def measure_pipe_rate(dtype="f64", mode: int = 0, iters: int = 4096):
    """(tera lane-ops/s, ms) of the FMA pipe on the bound device (roofline denominator)."""
    ensure_init()
    rate, ms = C.c_double(), C.c_float()
    check(lib().mb_measure_pipe_rate(DTYPES[dtype], mode, iters, C.byref(rate), C.byref(ms)))
    return rate.value, ms.value


class PinnedPool:
    """Recycled page-locked host buffers handed out as numpy arrays.

    `take(n)` returns a fresh-looking, writable uint8 array of n bytes that the caller owns like
    any numpy array; when the array (and every view of it) is garbage-collected the page-locked
    block goes back to the pool instead of being freed, so steady-state calls neither allocate nor
    pin memory.  At most `keep` idle blocks per size are retained, and at most `max_bytes` are
    page-locked and handed out at any time: a caller that hoards results (a whole level in a list)
    gets ordinary pageable arrays beyond that instead of exhausting lockable memory."""

    def __init__(self, alloc=None, free=None, keep: int = 4, max_bytes: int = 1 << 30):
        self._alloc, self._free, self._keep, self._max = alloc, free, keep, max_bytes
        self.outstanding = 0                     # page-locked bytes currently owned by callers
        self._idle = {}                          # nbytes -> [address, ...]
        self._lock = threading.RLock()           # re-entrant: _give_back may run from __del__ inside take() (cyclic GC)
        self.allocated = 0

    def _raw_alloc(self, n: int) -> int:
        if self._alloc is not None:
            return self._alloc(n)
        p = _vp()
        check(lib().mb_host_alloc(n, C.byref(p)))
        return p.value

    def _raw_free(self, addr: int) -> None:
        if self._free is not None:
            self._free(addr)
        else:
            lib().mb_host_free(_vp(addr))

    def take(self, n: int):
        import numpy as np
        with self._lock:
            if self.outstanding + n > self._max:
                return np.empty(n, dtype=np.uint8)       # over the cap: pageable memory, still correct, slower copy
            self.outstanding += n
            idle = self._idle.get(n)
            addr = idle.pop() if idle else None
        if addr is None:
            addr = self._raw_alloc(n)
            with self._lock:
                self.allocated += 1
        buf = (C.c_uint8 * n).from_address(addr)
        buf._lease = _Lease(self, n, addr)       # dies with the last array/view over `buf`
        return np.frombuffer(buf, dtype=np.uint8)

    def _give_back(self, n: int, addr: int) -> None:
        with self._lock:
            self.outstanding -= n
            idle = self._idle.setdefault(n, [])
            if len(idle) < self._keep:
                idle.append(addr)
                return
        self._raw_free(addr)


class _Lease:
    __slots__ = ("pool", "n", "addr")

    def __init__(self, pool, n, addr):
        self.pool, self.n, self.addr = pool, n, addr

    def __del__(self):
        try:
            self.pool._give_back(self.n, self.addr)
        except Exception:  # noqa: BLE001 -- interpreter shutdown
            pass


class TileQueue:
    """mb_wsq: work-stealing tile-index ranges shared by the GPU processes of one box.

    In-box replacement for the Distributer's lease dispatch (Distributer.cs:335-392)."""

    def __init__(self, name: str, n_ranks: int | None = None, total: int | None = None, create: bool = False):
        self.name = name.encode()
        self._h = _vp()
        self._owner = create
        if create:
            check(lib().mb_wsq_create(self.name, n_ranks, total, C.byref(self._h)))
        else:
            check(lib().mb_wsq_open(self.name, C.byref(self._h)))

    def reset(self, total: int) -> None:
        check(lib().mb_wsq_reset(self._h, total))

    def pop(self, rank: int, want: int = 1):
        """-> (first, got, stolen) or None when the queue is drained."""
        first, got, stolen = _u32(), _u32(), _u32()
        rc = check(lib().mb_wsq_pop(self._h, rank, want, C.byref(first), C.byref(got), C.byref(stolen)))
        if rc == 0:
            return None
        return first.value, got.value, bool(stolen.value)

    def remaining(self) -> int:
        return int(check(lib().mb_wsq_remaining(self._h)))

    def close(self) -> None:
        if self._h:
            lib().mb_wsq_close(self._h)
            self._h = _vp()
            if self._owner:
                lib().mb_wsq_unlink(self.name)

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()


class HostFrame:
    """mb_hostframe: the uint8 payloads of one frame in POSIX shared memory, page-locked by every GPU
    process of the box, with per-band / per-chunk completion counters (csrc/host_frame.cpp)."""

    def __init__(self, name: str, payload_bytes: int | None = None, n_units: int | None = None,
                 units_per_group: int | None = None, create: bool = False):
        self.name = name.encode()
        self._h = _vp()
        self._owner = create
        self._pinned = False
        if create:
            check(lib().mb_hostframe_create(self.name, payload_bytes, n_units, units_per_group, C.byref(self._h)))
        else:
            check(lib().mb_hostframe_open(self.name, C.byref(self._h)))
        self.nbytes = int(lib().mb_hostframe_bytes(self._h))
        self.n_units = int(lib().mb_hostframe_units(self._h))
        self.n_groups = int(lib().mb_hostframe_groups(self._h))
        self._array = None

    @property
    def array(self):
        """numpy uint8 view of the payload area (valid until close())."""
        import numpy as np
        if self._array is None:
            ptr = lib().mb_hostframe_data(self._h)
            self._array = np.ctypeslib.as_array((C.c_uint8 * self.nbytes).from_address(ptr))
        return self._array

    def pin(self) -> None:
        check(lib().mb_hostframe_pin(self._h))
        self._pinned = True

    def reset(self, active_units: int | None = None) -> None:
        """Re-arm for another pass; with active_units only units [0, active_units) take part."""
        if active_units is None:
            check(lib().mb_hostframe_reset(self._h))
        else:
            check(lib().mb_hostframe_reset_active(self._h, int(active_units)))

    def mark_done(self, first: int, count: int = 1):
        """Publish units [first, first+count); returns the groups (chunks) this completed."""
        cap = count
        buf, n = (_u32 * cap)(), _u32(0)
        check(lib().mb_hostframe_mark_done(self._h, first, count, buf, cap, C.byref(n)))
        return [buf[k] for k in range(min(n.value, cap))]

    def group_left(self, group: int) -> int:
        return int(check(lib().mb_hostframe_group_left(self._h, group)))

    def missing(self):
        n = _u32()
        check(lib().mb_hostframe_missing(self._h, None, 0, C.byref(n)))
        if n.value == 0:
            return []
        buf = (_u32 * n.value)()
        check(lib().mb_hostframe_missing(self._h, buf, n.value, C.byref(n)))
        return list(buf[:n.value])

    def close(self) -> None:
        if self._h:
            self._array = None
            if self._pinned:
                lib().mb_hostframe_unpin(self._h)
                self._pinned = False
            lib().mb_hostframe_close(self._h)
            self._h = _vp()
            if self._owner:
                lib().mb_hostframe_unlink(self.name)
