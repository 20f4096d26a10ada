"""In-box frame renderer: a B x B block of DataChunks rendered on the GPUs of one box.

A frame is what the reference's Distributer iterates over for one level setting
(Distributer.TryGetNextNeededWorkload, Distributer.cs:335-353): chunks (level, indexReal,
indexImag) with the same mrd.  Here the chunks of a frame are cut into row bands (the unit of
dispatch), the bands are handed to one process per GPU by the shared-memory work-stealing queue
(csrc/tile_wsq.cpp -- the in-box replacement for the socket dispatch), every band is rendered by
the sm_100a escape-time kernel straight into its final position of a chunk-contiguous int32
image, and ONE batched NCCL send/recv moves the bands to rank 0 (the only collective: pixels,
hence bands, are independent -- no halo, no reduction).

Output layout (also the wire/disk shape of the reference, one DataChunk after another):
    image[chunk][row][col], chunk = bi * B + br  for chunk indices (ir0 + br, ii0 + bi),
    row = imag index, col = real index inside the chunk (p = row*d + col, .py:34-35).
"""
from __future__ import annotations

import ctypes as C
import dataclasses
import os
from typing import Callable, List, Optional, Sequence, Tuple

from . import _native as nat


@dataclasses.dataclass(frozen=True)
class FrameSpec:
    """B x B chunks at `level`, chunk indices [ir0, ir0+B) x [ii0, ii0+B)."""
    name: str
    level: int
    blocks: int
    definition: int
    ir0: int
    ii0: int
    mrd: int
    dtype: str = "f64"
    band_rows: int = 256          # rows per dispatch unit (a band is contiguous in the chunk layout)
    shortcut: bool = False        # MB_TILE_PERIOD_SHORTCUT on every band (bit-identical counts, fewer executed updates)
    refill: bool = False          # MB_TILE_LANE_REFILL on every band (persistent warps; for the deep zooms)

    @property
    def n_chunks(self) -> int:
        return self.blocks * self.blocks

    @property
    def pixels(self) -> int:
        return self.n_chunks * self.definition * self.definition

    @property
    def bands_per_chunk(self) -> int:
        return (self.definition + self.band_rows - 1) // self.band_rows

    @property
    def n_bands(self) -> int:
        return self.n_chunks * self.bands_per_chunk

    def chunk_indices(self) -> List[Tuple[int, int]]:
        return [(self.ir0 + br, self.ii0 + bi) for bi in range(self.blocks) for br in range(self.blocks)]

    def scaled(self, **kw) -> "FrameSpec":
        return dataclasses.replace(self, **kw)


# BASELINE.json configs pinned to concrete inputs (BASELINE.md section 5 / SURVEY.md 8d)
CONFIGS = {
    "cfg1": FrameSpec("cfg1: single 512x512 DataChunk, mrd=256", 4, 1, 512, 1, 2, 256, "f64", 64),
    "cfg2": FrameSpec("cfg2: 4096x4096 whole plane, mrd=1024, one 256x256 chunk per SM-sized tile", 16, 16, 256, 0, 0, 1024, "f64", 256),
    "cfg3": FrameSpec("cfg3: 16384x16384 whole plane, mrd=4096, FP32", 4, 4, 4096, 0, 0, 4096, "f32", 256),
    "cfg4": FrameSpec("cfg4: 65536x65536 zoom at -0.75+0.1i, mrd=10000", 800, 16, 4096, 242, 412, 10000, "f64", 256),
    "cfg5": FrameSpec("cfg5: 32768x32768 seahorse zoom at -0.743643+0.131825i, mrd=50000", 4000000, 8, 4096, 1256353, 2131821, 50000, "f64", 64),
}


def frame_bands(spec: FrameSpec):
    """ctypes array of band tiles in dispatch order + each band's (element offset, element count)."""
    d = spec.definition
    tiles, extents = [], []
    for c, (ir, ii) in enumerate(spec.chunk_indices()):
        chunk = nat.chunk_geometry(spec.level, ir, ii, d, spec.mrd, shortcut=spec.shortcut, refill=spec.refill)
        for y0 in range(0, d, spec.band_rows):
            h = min(spec.band_rows, d - y0)
            off = c * d * d + y0 * d
            tiles.append(chunk.sub(0, y0, d, h, off, d))
            extents.append((off, h * d))
    return nat.tile_array(tiles), extents


def coalesce(ranges: Sequence[Tuple[int, int]]) -> List[Tuple[int, int]]:
    """Merge adjacent (first, count) band ranges."""
    out: List[Tuple[int, int]] = []
    for first, cnt in sorted(ranges):
        if out and out[-1][0] + out[-1][1] == first:
            out[-1] = (out[-1][0], out[-1][1] + cnt)
        else:
            out.append((first, cnt))
    return out


class FrameRenderer:
    """Renders `spec` on this process's GPU (single GPU) or, with a TileQueue and a
    torch.distributed process group, on all GPUs of the box with a final gather to rank 0."""

    def __init__(self, spec: FrameSpec, device: int = 0, rank: int = 0, world: int = 1,
                 launch: Optional[Callable] = None, out=None, bands_per_claim: int = 4, depth: int = 3):
        import torch
        self.torch = torch
        self.spec, self.rank, self.world = spec, rank, world
        self.tiles, self.extents = frame_bands(spec)
        self.bands_per_claim, self.depth = bands_per_claim, depth
        self._tile_size = C.sizeof(nat.Tile)
        if launch is None:
            nat.ensure_init(device)
            self.device = torch.device("cuda", device)
            self._launch = self._launch_cuda
        else:                                   # test hook (CPU / gloo): launch(first, count, out)
            self.device = torch.device("cpu")
            self._launch = launch
        self.out = out if out is not None else torch.empty(spec.pixels, dtype=torch.int32, device=self.device)
        self.claimed: List[Tuple[int, int]] = []
        self.steals = 0               # of the last render_stolen call
        self.steals_total = 0         # accumulated over calls
        self.launches = 0

    # -- launching ---------------------------------------------------------------------------
    def _launch_cuda(self, first: int, count: int, out) -> None:
        torch = self.torch
        ptr = C.cast(C.byref(self.tiles, first * self._tile_size), C.POINTER(nat.Tile))
        nat.check(nat.lib().mb_render_counts(ptr, count, out.data_ptr(), nat.DTYPES[self.spec.dtype],
                                             torch.cuda.current_stream(self.device).cuda_stream))

    def render_all(self) -> None:
        """Single GPU, no queue: every band in ONE launch (grid.y = band index)."""
        self._launch(0, self.spec.n_bands, self.out)
        self.claimed = [(0, self.spec.n_bands)]
        self.launches += 1

    def render_stolen(self, queue: "nat.TileQueue") -> None:
        """Claim bands from the work-stealing queue until it is drained; at most `depth`
        launches in flight so that ranks which run ahead cannot hoard the queue."""
        torch = self.torch
        self.claimed, self.steals = [], 0
        inflight = []
        cuda = self.device.type == "cuda"
        if cuda:
            # the same loop in native code (mb_frame_render_stolen): ~10 us instead of ~60 us per claim
            cap = self.spec.n_bands
            buf = (C.c_uint32 * (2 * cap))()
            n_claims, n_stolen = C.c_uint32(), C.c_uint32()
            nat.check(nat.lib().mb_frame_render_stolen(
                queue._h, self.rank, self.world, self.tiles, self.spec.n_bands, self.bands_per_claim, self.depth,
                self.out.data_ptr(), nat.DTYPES[self.spec.dtype], torch.cuda.current_stream(self.device).cuda_stream,
                buf, cap, C.byref(n_claims), C.byref(n_stolen)))
            self.claimed = coalesce([(buf[2 * k], buf[2 * k + 1]) for k in range(n_claims.value)])
            self.steals = n_stolen.value
            self.steals_total += n_stolen.value
            self.launches += n_claims.value
            return
        while True:
            # guided self-scheduling: full-size claims while plenty is left, single bands at the end
            want = max(1, min(self.bands_per_claim, queue.remaining() // (self.world * 8)))
            got = queue.pop(self.rank, want)
            if got is None:
                break
            first, cnt, stolen = got
            self.steals += int(stolen)
            self.steals_total += int(stolen)
            self._launch(first, cnt, self.out)
            self.launches += 1
            self.claimed.append((first, cnt))
            if cuda:
                ev = torch.cuda.Event()
                ev.record()
                inflight.append(ev)
                if len(inflight) >= self.depth:
                    inflight.pop(0).synchronize()
        self.claimed = coalesce(self.claimed)

    # -- the one collective -------------------------------------------------------------------
    def gather_to_root(self, group=None) -> None:
        """All ranks learn who rendered what (tiny all_gather_object), then one batch of
        point-to-point transfers places every band at its final offset on rank 0 (grouped
        ncclSend/ncclRecv over NVLink; ownership is dynamic, so a fixed-shape gather would not fit)."""
        import torch.distributed as dist
        if self.world == 1:
            return
        owned = [None] * self.world
        dist.all_gather_object(owned, self.claimed, group=group)
        ops = []
        for src in range(1, self.world):
            for first, cnt in owned[src]:
                lo = self.extents[first][0]
                hi = self.extents[first + cnt - 1][0] + self.extents[first + cnt - 1][1]
                view = self.out[lo:hi]
                if self.rank == 0:
                    ops.append(dist.P2POp(dist.irecv, view, src, group=group))
                elif self.rank == src:
                    ops.append(dist.P2POp(dist.isend, view, 0, group=group))
        if ops:
            for req in dist.batch_isend_irecv(ops):
                req.wait()

    # -- metric -------------------------------------------------------------------------------
    def useful_iterations(self) -> int:
        """sum over the frame of (n ? n : mrd-1), reduced on the device."""
        torch = self.torch
        acc = torch.zeros(1, dtype=torch.int64, device=self.device)
        nat.check(nat.lib().mb_useful_iterations(self.out.data_ptr(), self.out.numel(), self.spec.mrd,
                                                 acc.data_ptr(), torch.cuda.current_stream(self.device).cuda_stream))
        return int(acc.item())


def probe_chunk_costs(spec: FrameSpec, samples: int = 48, device: int = 0):
    """Cheap cost estimate per chunk: render every chunk's window at `samples` x `samples` (one launch,
    ~2 ms for 256 chunks) and sum the useful iterations.  Used to hand out the expensive chunks first
    (longest-processing-time-first) when whole chunks are the unit of dispatch, e.g. behind the
    worker API where a DataChunk cannot be split; the Distributer's own order (Distributer.cs:335-353)
    ignores cost and leaves one 50 000-iteration chunk as the tail."""
    import torch
    dev = torch.device("cuda", nat.ensure_init(device))
    n = samples * samples
    tiles = []
    for c, (ir, ii) in enumerate(spec.chunk_indices()):
        t = nat.chunk_geometry(spec.level, ir, ii, samples, spec.mrd)
        t.out_offset = c * n
        tiles.append(t)
    out = torch.empty(spec.n_chunks * n, dtype=torch.int32, device=dev)
    nat.check(nat.lib().mb_render_counts(nat.tile_array(tiles), len(tiles), out.data_ptr(), nat.DTYPES[spec.dtype],
                                         torch.cuda.current_stream(dev).cuda_stream))
    counts = out.view(spec.n_chunks, n)
    never = max(spec.mrd - 1, 0)
    return torch.where(counts == 0, torch.full_like(counts, never), counts).sum(dim=1, dtype=torch.int64).cpu().tolist()


class _RawCudaArray:
    """Zero-copy view of a raw device pointer for torch.as_tensor (CUDA array interface)."""

    def __init__(self, ptr: int, n: int, typestr: str = "<i4"):
        self.__cuda_array_interface__ = {"shape": (n,), "typestr": typestr, "data": (ptr, False), "version": 2}


def shared_frame_buffer(spec: FrameSpec, rank: int, device: int, group=None):
    """Frame buffer that lives on rank 0 and is mapped into every other rank of the box through
    CUDA IPC (peer memory over NVLink): each rank's kernel stores straight into rank 0's HBM, so the
    final gather disappears.  Returns (tensor over the shared buffer, closer)."""
    import torch
    import torch.distributed as dist
    nat.ensure_init(device)
    lib = nat.lib()
    ptr = C.c_void_p()
    handle = (C.c_ubyte * 64)()
    if rank == 0:
        nat.check(lib.mb_peer_alloc(spec.pixels * 4, C.byref(ptr), handle))
    box = [bytes(handle)]
    dist.broadcast_object_list(box, src=0, group=group)
    if rank != 0:
        h = (C.c_ubyte * 64).from_buffer_copy(box[0])
        nat.check(lib.mb_peer_open(h, C.byref(ptr)))
    t = torch.as_tensor(_RawCudaArray(ptr.value, spec.pixels), device=torch.device("cuda", device))
    assert t.data_ptr() == ptr.value

    def close():
        (lib.mb_peer_free if rank == 0 else lib.mb_peer_close)(ptr)

    return t, close


def queue_name() -> str:
    return f"/mbq_{os.environ.get('MASTER_PORT', '0')}_{os.getppid()}"


# ---- end-to-end in-box path: band-granular dispatch into ONE shared pinned host frame ----------------

def workload_bands(workloads, definition: int, band_rows: int, shortcut: bool = False, refill: bool = False):
    """Row-band tiles for a list of (level, mrd, index_real, index_imag) workloads.  Band k of chunk c
    is unit c * bands_per_chunk + k of the host frame; out_offset is its BYTE offset in the frame of
    uint8 payloads (chunk c's payload = bytes [c*d*d, (c+1)*d*d), exactly process_workload's output)."""
    d = int(definition)
    tiles = []
    for c, (level, mrd, ir, ii) in enumerate(workloads):
        chunk = nat.chunk_geometry(level, ir, ii, d, mrd, shortcut=shortcut, refill=refill)
        for y0 in range(0, d, band_rows):
            h = min(band_rows, d - y0)
            tiles.append(chunk.sub(0, y0, d, h, c * d * d + y0 * d, d))
    return nat.tile_array(tiles), (d + band_rows - 1) // band_rows


class InboxWorker:
    """process_workloads for ALL GPUs of one box (SURVEY.md 8e + f4): the same uint8 payloads, one
    per (level, mrd, index_real, index_imag), but a chunk is no longer the unit of work.

    Behind the reference's worker API a DataChunk cannot be split, so one 50 000-iteration chunk
    (0.3 s) bounds an 8-GPU frame.  Here every chunk is cut into row bands, the bands of all chunks go
    through the shared work-stealing queue (csrc/tile_wsq.cpp), and every rank copies its bands
    device->host straight to their final offsets of ONE page-locked shm frame
    (csrc/host_frame.cpp): no gather, no second copy; a payload is complete when its band counter
    hits zero, and `consume(workload, payload)` is called for it while other chunks still render.
    Bands claimed by a rank that died are found by `missing()` and rendered again by `repair()`.

    One instance per process; rank 0 creates the segments, `barrier` (e.g. torch.distributed.barrier)
    separates create from open.  `launch(first, count, frame_array)` is a CPU test hook that replaces
    the CUDA dispatch loop."""

    def __init__(self, workloads, definition: int = nat.CHUNK_DEFINITION, band_rows: int = 256, rank: int = 0,
                 world: int = 1, device: Optional[int] = None, name: Optional[str] = None, barrier: Optional[Callable] = None,
                 shortcut: bool = False, dtype: str = "f64", bands_per_claim: int = 4, depth: int = 6,
                 launch: Optional[Callable] = None, refill: bool = False):
        self.workloads = [tuple(int(v) for v in w) for w in workloads]
        self.capacity = len(self.workloads)          # the shm frame is sized for this many chunks
        self.definition, self.band_rows = int(definition), int(band_rows)
        self.rank, self.world, self.dtype = rank, world, dtype
        self._flags = (shortcut, refill)
        self.bands_per_claim, self.depth = bands_per_claim, depth
        self._barrier = barrier if barrier is not None else (lambda: None)
        self._launch = launch
        if launch is None:
            nat.ensure_init(device)
        self.tiles, self.bands_per_chunk = workload_bands(self.workloads, self.definition, self.band_rows, shortcut, refill)
        self.n_bands = len(self.tiles)
        self.chunk_bytes = self.definition * self.definition
        name = name or queue_name()
        qn, fn = name + "_bq", name + "_hf"
        if rank == 0:
            self.queue = nat.TileQueue(qn, world, self.n_bands, create=True)
            self.frame = nat.HostFrame(fn, len(self.workloads) * self.chunk_bytes, self.n_bands, self.bands_per_chunk, create=True)
        self._barrier()
        if rank != 0:
            self.queue = nat.TileQueue(qn)
            self.frame = nat.HostFrame(fn)
        if launch is None:
            self.frame.pin()                 # cudaHostRegister on this rank's mapping: D2H at PCIe speed
        self.steals = self.claims = 0
        self.completed: List[int] = []

    def rearm(self) -> None:
        """Prepare another pass over the same workloads (collective: every rank calls it)."""
        self._barrier()
        if self.rank == 0:
            self.queue.reset(self.n_bands)
            self.frame.reset(self.n_bands)
        self._barrier()

    def retarget(self, workloads) -> None:
        """Point the same shm frame and queue at a NEW list of at most `capacity` workloads (collective;
        every rank passes the same list) -- e.g. the next batch of leases from a Distributer."""
        workloads = [tuple(int(v) for v in w) for w in workloads]
        if not 0 < len(workloads) <= self.capacity:
            raise ValueError(f"between 1 and {self.capacity} workloads per batch")
        self.workloads = workloads
        self.tiles, self.bands_per_chunk = workload_bands(workloads, self.definition, self.band_rows, *self._flags)
        self.n_bands = len(self.tiles)
        self.rearm()

    def payload(self, k: int):
        """uint8[definition^2] view of workload k's payload in the shared frame."""
        return self.frame.array[k * self.chunk_bytes:(k + 1) * self.chunk_bytes]

    def render(self, consume: Optional[Callable] = None) -> List[int]:
        """Claim bands until the queue is drained; returns when every band THIS rank claimed has landed
        in host memory.  -> indices of the workloads whose last band this rank delivered; `consume` is
        called for each of them (the frame is complete on all ranks only after the next barrier)."""
        if self._launch is not None:
            done = self._render_hook()
        else:
            cap = self.n_bands
            claimed = (C.c_uint32 * (2 * cap))()
            comp = (C.c_uint32 * len(self.workloads))()
            n_claims, n_stolen, n_comp = C.c_uint32(), C.c_uint32(), C.c_uint32()
            nat.check(nat.lib().mb_frame_render_u8_host(
                self.queue._h, self.rank, self.world, self.tiles, self.n_bands, self.bands_per_claim, self.depth,
                self.frame._h, nat.DTYPES[self.dtype], claimed, cap, C.byref(n_claims), C.byref(n_stolen),
                comp, len(self.workloads), C.byref(n_comp)))
            self.claims += n_claims.value
            self.steals += n_stolen.value
            done = [comp[k] for k in range(n_comp.value)]
        self.completed = done
        if consume is not None:
            for g in done:
                consume(self.workloads[g], self.payload(g))
        return done

    def _render_hook(self) -> List[int]:
        done: List[int] = []
        while True:
            want = max(1, min(self.bands_per_claim, self.queue.remaining() // (self.world * 8)))
            got = self.queue.pop(self.rank, want)
            if got is None:
                return done
            first, cnt, stolen = got
            self.steals += int(stolen)
            self.claims += 1
            if self._launch(first, cnt, self.frame.array) is False:
                continue                     # test hook: a rank that "died" after claiming
            done += self.frame.mark_done(first, cnt)

    def missing(self) -> List[int]:
        """Bands no rank delivered (call after the barrier that follows render())."""
        return self.frame.missing()

    def repair(self, consume: Optional[Callable] = None) -> List[int]:
        """Render the bands `missing()` reports on this rank; -> workloads completed by it."""
        units = self.missing()
        if not units:
            return []
        if self._launch is not None:
            done: List[int] = []
            for u in units:
                self._launch(u, 1, self.frame.array)
                done += self.frame.mark_done(u, 1)
        else:
            arr = (C.c_uint32 * len(units))(*units)
            comp = (C.c_uint32 * len(self.workloads))()
            n_comp = C.c_uint32()
            nat.check(nat.lib().mb_frame_render_u8_host_units(arr, len(units), self.tiles, self.n_bands, self.depth, self.frame._h,
                                                              nat.DTYPES[self.dtype], comp, len(self.workloads), C.byref(n_comp)))
            done = [comp[k] for k in range(n_comp.value)]
        if consume is not None:
            for g in done:
                consume(self.workloads[g], self.payload(g))
        return done

    def close(self) -> None:
        self._barrier()
        self.frame.close()
        self.queue.close()
