"""Drop-in worker: the host side of the hot path, mirroring the reference worker module
DistributedMandelbrotWorkerCUDA/DistributedMandelbrotWorkerCUDA.py (".py" below).

Same module-level names, same wire protocol, same function boundary:

    process_workload(level, mrd, index_real, index_imag) -> np.ndarray[uint8, 4096*4096]   (.py:70-100)
    calc_mb_value(r, i, mrd, out=None) -> int32 counts                                     (.py:39-68)
    receive_workload(sock) -> (level, mrd, indexReal, indexImag)                           (.py:102-109)
    do_workload_single(addr, port) -> bool                                                 (.py:111-176)
    main()                                                                                 (.py:178-184)

What changed underneath: coordinates are generated inside the CUDA kernel from an 88-byte tile
descriptor (the reference builds two 128 MiB float64 arrays on the host and copies them to the
device, .py:82-88), the escape loop is a hand-written sm_100a kernel with strict (non
FMA-contracted) binary64 arithmetic, and the uint8 scaling of .py:96-98 is fused into the
kernel epilogue, so the only device->host traffic is the 16 MiB payload itself.

There is no CPU fallback: every compute function raises if the CUDA library or device is
missing.
"""
from __future__ import annotations

import ctypes as C
import os
import socket
from struct import pack, unpack

import numpy as np

from . import _native as nat

# .py:7-8 / DataChunk.cs:14-15
MIN_AXIS = -2
MAX_AXIS = 2

# .py:10-17 / Distributer.cs:30-45
REQUEST_CODE = 0x00
RESPONSE_CODE = 0x01

WORKLOAD_AVAILABLE_CODE = 0x10
WORKLOAD_NOT_AVAILABLE_CODE = 0x11

WORKLOAD_ACCEPT_CODE = 0x20
WORKLOAD_REJECT_CODE = 0x21

# DataChunk.dataChunkRange / dataChunkSize (DataChunk.cs:20,27); .py:80
DEFINITION = nat.CHUNK_DEFINITION
DATA_CHUNK_SIZE = DEFINITION * DEFINITION


def process_workload(level: int, mrd: int, index_real: int, index_imag: int, *,
                     definition: int = DEFINITION, out: np.ndarray | None = None) -> np.ndarray:
    """(.py:70-100) one DataChunk -> uint8[definition*definition], row-major, p = row*d + col.

    Synchronous; returns a fresh C-contiguous host array owned by the caller (or fills `out`).
    The array is page-locked memory from a recycling pool (the copy runs at PCIe speed and the
    block returns to the pool when the array is dropped); once a caller holds more than 1 GiB of
    such results, further ones come as ordinary pageable arrays.  Raises ValueError for level == 0 / index >= level (the checks of the
    DataChunk ctor, DataChunk.cs:99-106) and mrd == 0 (division by zero at .py:96).
    """
    nat.ensure_init()
    n = int(definition) * int(definition)
    if out is None:
        # a page-locked block from a recycling pool: the 16 MiB device->host copy runs at PCIe
        # speed, and the block returns to the pool when the caller drops the array
        out = _PAYLOADS.take(n)
    elif out.dtype != np.uint8 or out.size != n or not out.flags.c_contiguous:
        raise ValueError("out must be a C-contiguous uint8 array of definition*definition elements")
    _check_u32(level=level, mrd=mrd, index_real=index_real, index_imag=index_imag, definition=definition)
    nat.check(nat.lib().mb_process_workload_u8(level, mrd, index_real, index_imag, definition,
                                               out.ctypes.data_as(C.c_void_p)))
    return out


_PAYLOADS = nat.PinnedPool()
_SMALL_CHUNK = 1 << 20        # chunks up to 1 MiB go through the native list loop (mb_process_workloads_u8)
_SLAB_BYTES = 64 << 20        # ... in page-locked slabs of at most this many payload bytes
_LAUNCH_BYTES = 2 << 20       # ... about this much payload per launch: the last launch's copy is the exposed tail


def process_workload_counts(level: int, mrd: int, index_real: int, index_imag: int, *,
                            definition: int = DEFINITION, dtype: str = "f64") -> np.ndarray:
    """The int32 iteration counts calc_mb_value produces for the chunk, i.e. `out` at .py:94
    before the uint8 scaling.  dtype "f32" selects the binary32 variant (no upstream counterpart)."""
    nat.ensure_init()
    _check_u32(level=level, mrd=mrd, index_real=index_real, index_imag=index_imag, definition=definition)
    out = np.empty(int(definition) * int(definition), dtype=np.int32)
    nat.check(nat.lib().mb_process_workload_counts(level, mrd, index_real, index_imag, definition,
                                                   nat.DTYPES[dtype], out.ctypes.data_as(C.c_void_p)))
    return out


def process_workloads(workloads, *, definition: int = DEFINITION, consume=None, depth: int = 4, shortcut: bool = False,
                      refill: bool = False):
    """Batch form of process_workload for a list of (level, mrd, index_real, index_imag).

    Same result per workload as process_workload (uint8[definition^2] in HOST memory), but the
    chunks are pipelined: kernels alternate between two CUDA streams (the tail of chunk k overlaps
    the head of chunk k+1) and each payload's device->host copy into a pinned ring buffer runs on a
    third stream behind its kernel.  `consume(workload, payload)` is called in order with a view
    that is only valid during the call; without `consume` a list of owned arrays is returned.
    `workloads` may be any iterable and is consumed lazily (e.g. a generator that pops chunk
    indices from the in-box work-stealing queue, or one that asks a Distributer for the next lease).
    shortcut=True renders with the exact orbit-periodicity shortcut (MB_TILE_PERIOD_SHORTCUT):
    identical payloads, fewer executed updates inside the set; refill=True uses the persistent-warp
    kernel with lane refill (MB_TILE_LANE_REFILL, for deep zooms).
    """
    import itertools
    dev_index = nat.ensure_init()
    n = int(definition) * int(definition)
    results = [] if consume is None else None
    if n <= _SMALL_CHUNK:
        # small chunks (reduced definitions, e.g. the 256^2 tiles of BASELINE config 2): a frame lasts well
        # under a millisecond, so the whole loop runs in native code (mb_process_workloads_u8): geometry,
        # grouping into ~4 MiB launches and the pipelined device->host copies, one call per slab
        it = iter(workloads)
        cap = max(1, min(4096, _SLAB_BYTES // n))
        flags = (nat.TILE_PERIOD_SHORTCUT if shortcut else 0) | (nat.TILE_LANE_REFILL if refill else 0)
        while True:
            batch = list(itertools.islice(it, cap))
            if not batch:
                return results
            rec64 = np.asarray(batch, dtype=np.int64).reshape(-1, 4)          # one vectorised range check, not 4 per chunk
            if (rec64 < 0).any() or (rec64 > 0xFFFFFFFF).any():
                raise ValueError("workload fields must be uint32 (DistributerWorkload.cs:61-75)")
            _check_u32(definition=definition)
            rec = np.ascontiguousarray(rec64.astype(np.uint32))
            slab = _PAYLOADS.take(len(batch) * n)
            prev = nat.set_tile_flags(flags)
            try:
                nat.check(nat.lib().mb_process_workloads_u8(rec.ctypes.data_as(C.c_void_p), len(batch), int(definition),
                                                            slab.ctypes.data_as(C.c_void_p), _LAUNCH_BYTES))
            finally:
                nat.set_tile_flags(prev)
            rows = slab.reshape(len(batch), n)                               # one view per payload, made in C
            if consume is None:
                results.extend(row.copy() for row in rows)
            else:
                for w, payload in zip(batch, rows):
                    consume(w, payload)
    import torch
    device = torch.device("cuda", dev_index)
    depth = max(2, int(depth))
    # small chunks (reduced definitions) are grouped so that one launch and one copy carry ~4 MiB;
    # the reference's 4096^2 chunks (16 MiB) travel one per launch
    group = max(1, min(64, (4 << 20) // n))
    ring = _ring(device, n * group, depth)
    compute = ring["compute"]
    copy = ring["copy"]
    done = [None] * depth                    # copy-finished event per slot
    pending = [None] * depth                 # workloads whose payloads sit in the slot
    lib = nat.lib()

    def finish(slot):
        done[slot].synchronize()
        for j, w in enumerate(pending[slot]):
            payload = ring["host_np"][slot][j * n:(j + 1) * n]
            if consume is None:
                results.append(payload.copy())
            else:
                consume(w, payload)
        pending[slot] = None

    it = iter(workloads)                     # consumed lazily: the source may be a live queue
    k = 0
    while True:
        batch = list(itertools.islice(it, group))
        if not batch:
            break
        tiles = []
        for j, (level, mrd, ir, ii) in enumerate(batch):
            _check_u32(level=level, mrd=mrd, index_real=ir, index_imag=ii, definition=definition)
            t = nat.chunk_geometry(level, ir, ii, definition, mrd, shortcut=shortcut, refill=refill)
            t.out_offset = j * n
            tiles.append(t)
        slot = k % depth
        if pending[slot] is not None:
            finish(slot)                     # frees this slot's host and device buffers
        st = compute[k % len(compute)]
        nat.check(lib.mb_render_u8(nat.tile_array(tiles), len(tiles), ring["dev"][slot].data_ptr(), nat.DTYPE_F64, st.cuda_stream))
        ev = torch.cuda.Event()
        ev.record(st)
        copy.wait_event(ev)
        used = len(batch) * n
        with torch.cuda.stream(copy):
            ring["host"][slot][:used].copy_(ring["dev"][slot][:used], non_blocking=True)
        done[slot] = torch.cuda.Event()
        done[slot].record(copy)
        pending[slot] = batch
        k += 1
    for j in range(k, k + depth):            # drain in submission order
        if pending[j % depth] is not None:
            finish(j % depth)
    return results


_RINGS = {}


def _ring(device, n: int, depth: int):
    """Pinned host + device payload ring and the streams of process_workloads (cached per shape)."""
    import torch
    key = (device.index, n, depth)
    r = _RINGS.get(key)
    if r is None:
        _RINGS.clear()                       # one shape at a time: the buffers are 16 MiB each
        host = [torch.empty(n, dtype=torch.uint8).pin_memory() for _ in range(depth)]
        r = {"host": host, "host_np": [h.numpy() for h in host],
             "dev": [torch.empty(n, dtype=torch.uint8, device=device) for _ in range(depth)],
             "compute": [torch.cuda.Stream(device), torch.cuda.Stream(device)],
             "copy": torch.cuda.Stream(device)}
        _RINGS[key] = r
    return r


def calc_mb_value(r, i, mrd: int, out: np.ndarray | None = None) -> np.ndarray:
    """(.py:39-68) element-wise escape count over explicit coordinate arrays, like the
    reference's numba ufunc `int32(float64, float64, int32)`.  r and i broadcast together."""
    nat.ensure_init()
    r64 = np.asarray(r, dtype=np.float64)
    i64 = np.asarray(i, dtype=np.float64)
    r64, i64 = np.broadcast_arrays(r64, i64)
    shape = r64.shape
    r64 = np.ascontiguousarray(r64).ravel()
    i64 = np.ascontiguousarray(i64).ravel()
    if not (0 <= int(mrd) <= 0x7FFFFFFF):
        raise ValueError("mrd must fit int32 (.py:39)")
    if out is None:
        out = np.empty(shape, dtype=np.int32)
    elif out.dtype != np.int32 or out.shape != shape or not out.flags.c_contiguous:
        raise ValueError("out must be a C-contiguous int32 array of the broadcast shape")
    nat.check(nat.lib().mb_calc_mb_value(r64.ctypes.data_as(C.c_void_p), i64.ctypes.data_as(C.c_void_p),
                                         r64.size, int(mrd), nat.DTYPE_F64, out.ctypes.data_as(C.c_void_p)))
    return out


def _check_u32(**kw) -> None:
    for name, v in kw.items():
        if not (0 <= int(v) <= 0xFFFFFFFF):
            raise ValueError(f"{name}={v} is not a uint32 (DistributerWorkload.cs:61-75)")


# ---- wire protocol (.py:102-176 ; Distributer.cs:230-270,358-458 ; DistributerWorkload.cs:53-100)

def _recv_exact(sock: socket.socket, n: int) -> bytes:
    """The reference calls sock.recv(4) without a length loop (.py:104-107); loop here."""
    buf = bytearray()
    while len(buf) < n:
        part = sock.recv(n - len(buf))
        if not part:
            raise ConnectionError(f"connection closed after {len(buf)} of {n} bytes")
        buf += part
    return bytes(buf)


def receive_workload(sock: socket.socket):
    """(.py:102-109) four little-endian uint32: level, mrd, indexReal, indexImag
    (Workload.Send, DistributerWorkload.cs:53-77)."""
    level, mrd, indexReal, indexImag = unpack("<IIII", _recv_exact(sock, 16))
    return level, mrd, indexReal, indexImag


def request_workload(addr: str, port: int):
    """First half of do_workload_single (.py:115-134): returns the workload tuple, or None when
    the Distributer answers WORKLOAD_NOT_AVAILABLE (0x11)."""
    sock = socket.socket()
    try:
        sock.connect((addr, port))
        sock.sendall(pack("B", REQUEST_CODE))
        response = _recv_exact(sock, 1)[0]
        if response == WORKLOAD_AVAILABLE_CODE:
            return receive_workload(sock)
        if response == WORKLOAD_NOT_AVAILABLE_CODE:
            return None
        raise Exception("Unknown response code to request: " + str(response))
    finally:
        sock.close()   # never compute while connected: the server has a 100 ms receive timeout (Distributer.cs:17,196-202)


def return_workload(addr: str, port: int, workload, out: np.ndarray) -> bool:
    """Second half of do_workload_single (.py:148-172): True if the result was accepted and sent,
    False if the Distributer rejected it (0x21: unknown or timed-out lease) -- not an error."""
    sock = socket.socket()
    try:
        sock.connect((addr, port))
        sock.sendall(pack("B", RESPONSE_CODE) + pack("<IIII", *workload))
        response = _recv_exact(sock, 1)[0]
        if response == WORKLOAD_REJECT_CODE:
            return False
        if response != WORKLOAD_ACCEPT_CODE:
            raise Exception("Unknown response code to request: " + str(response))
        # exactly dataChunkSize bytes (Distributer.cs:415-416); sendall, the server reads once
        sock.sendall(memoryview(out).cast("B"))
        return True
    finally:
        sock.close()


def do_workload_single(addr: str, port: int, *, verbose: bool = True) -> bool:
    """(.py:111-176) request one workload, compute it, send it back.  False = no work left."""
    say = print if verbose else (lambda *a, **k: None)
    workload = request_workload(addr, port)
    if workload is None:
        say("No workload was available, ending program")
        return False
    say("Workload received:", workload)
    say("Starting calculation...")
    out = process_workload(workload[0], workload[1], workload[2], workload[3])
    say("Calculation complete")
    if return_workload(addr, port, workload, out):
        say("Sent response")
        say("Process complete")
    else:
        say("Response rejected")
    return True


def run_pipelined(addr: str, port: int, *, depth: int = 4, verbose: bool = True) -> int:
    """Same protocol as looping do_workload_single, but several leases in flight:
    while chunk k computes, chunk k+1.. are already leased and queued on the GPU and chunk k-1 is
    being sent back.  Each exchange still uses its own short connection (Distributer.cs:230-274)
    and nothing is held open during compute.  Returns the number of workloads processed."""
    say = print if verbose else (lambda *a, **k: None)
    count = 0

    def leases():
        while True:
            w = request_workload(addr, port)
            if w is None:
                say("No workload was available, ending program")
                return
            say("Workload received:", w)
            yield w

    def send_back(w, payload):
        nonlocal count
        count += 1
        say("Sent response" if return_workload(addr, port, w, payload) else "Response rejected", w)

    process_workloads(leases(), consume=send_back, depth=depth)
    return count


def run_inbox(addr: str, port: int, *, rank: int, world: int, device: int | None = None, barrier=None, share=None,
              batch: int | None = None, definition: int = DEFINITION, verbose: bool = True, launch=None,
              shortcut: bool = False, name: str | None = None) -> int:
    """ONE worker made of all GPUs of a box (SURVEY.md 8e + f4), behind an unmodified Distributer.

    Rank 0 leases a batch of workloads over the reference protocol (one short connection per lease,
    Distributer.cs:358-392) and shares the list; the chunks of the batch are cut into row bands and
    rendered by ALL ranks through the shared work-stealing queue into one page-locked shm frame
    (frame.InboxWorker), so an expensive chunk is finished by eight GPUs instead of occupying one; the
    rank that delivers a chunk's last band sends its 16 MiB payload back (its own connection,
    Distributer.cs:397-458).  With N independent reference workers a level's tail is one GPU grinding
    through the last deep chunk; here the tail is that chunk divided by N.

    barrier() / share(obj) are the two collectives needed (torch.distributed.barrier and an object
    broadcast from rank 0); `launch` is the CPU test hook of InboxWorker.  Returns the number of
    payloads this rank sent."""
    from .frame import InboxWorker
    say = print if verbose and rank == 0 else (lambda *a, **k: None)
    barrier = barrier if barrier is not None else (lambda: None)
    share = share if share is not None else (lambda obj: obj)
    batch = batch if batch is not None else max(2, 2 * world)
    sent = 0

    def lease_batch():
        got = []
        while len(got) < batch:
            w = request_workload(addr, port)
            if w is None:
                break
            got.append(w)
        return got

    def send_back(w, payload):
        nonlocal sent
        ok = return_workload(addr, port, w, payload)
        sent += 1
        if verbose:
            print("Sent response" if ok else "Response rejected", w, flush=True)

    iw = None
    while True:
        leases = share(lease_batch() if rank == 0 else None)
        if not leases:
            say("No workload was available, ending program")
            break
        say("Workloads received:", leases)
        if iw is None:
            # placeholders up to `batch`: the shm frame is sized once, later batches only retarget it
            first = list(leases) + [leases[-1]] * (batch - len(leases))
            iw = InboxWorker(first, definition=definition, rank=rank, world=world, device=device, barrier=barrier,
                             shortcut=shortcut, launch=launch, name=name)
        iw.retarget(leases)
        iw.render(consume=send_back)
        barrier()
        if rank == 0 and iw.missing():                 # bands a rank claimed and never delivered
            iw.repair(consume=send_back)
        barrier()
    if iw is not None:
        iw.close()
    return sent


def _inbox_main(addr: str, port: int, pipeline_unused: int = 0) -> None:
    """Entry point of one rank of `worker --gpus N --inbox` (environment set by the launcher)."""
    import torch.distributed as dist
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    dist.init_process_group("gloo", rank=rank, world_size=world)          # control plane only: two tiny collectives per batch

    def share(obj):
        box = [obj]
        dist.broadcast_object_list(box, src=0)
        return box[0]

    run_inbox(addr, port, rank=rank, world=world, device=rank, barrier=dist.barrier, share=share)
    dist.destroy_process_group()


def main(argv=None) -> None:
    """(.py:178-184) prompts `Server Addr> ` / `Server Port> ` like the reference; --addr/--port/
    --device skip the prompts."""
    import argparse
    ap = argparse.ArgumentParser(description="B200 worker for the DistributedMandelbrot Distributer")
    ap.add_argument("--addr")
    ap.add_argument("--port", type=int)
    ap.add_argument("--device", type=int, default=None)
    ap.add_argument("--pipeline", type=int, default=0, metavar="DEPTH",
                    help="keep DEPTH leases in flight (0 = one at a time, like the reference)")
    ap.add_argument("--gpus", type=int, default=0, metavar="N",
                    help="start one worker process per GPU for devices 0..N-1 (N independent workers, like N reference workers)")
    ap.add_argument("--inbox", action="store_true",
                    help="with --gpus N: the N GPUs act as ONE worker -- leases are rendered band by band by all of them "
                         "(frame.InboxWorker) instead of one chunk per GPU")
    ap.add_argument("--inbox-rank", type=int, default=None, help=argparse.SUPPRESS)
    args = ap.parse_args(argv)
    addr = args.addr if args.addr is not None else input("Server Addr> ")
    port = args.port if args.port is not None else int(input("Server Port> "))
    if args.inbox_rank is not None:
        _inbox_main(addr, port)
        return
    if args.gpus > 0 and args.inbox:
        import subprocess
        import sys
        env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT=str(29400 + os.getpid() % 2000), WORLD_SIZE=str(args.gpus))
        procs = [subprocess.Popen([sys.executable, "-m", "distributedmandelbrot_b200.worker", "--addr", addr, "--port", str(port),
                                   "--inbox-rank", str(k)], env=dict(env, RANK=str(k), LOCAL_RANK=str(k)))
                 for k in range(args.gpus)]
        codes = [p.wait() for p in procs]
        if any(codes):
            raise SystemExit(f"worker exit codes: {codes}")
        return
    if args.gpus > 0:
        # one process per GPU, exactly like running N reference workers against the same Distributer
        import subprocess
        import sys
        procs = [subprocess.Popen([sys.executable, "-m", "distributedmandelbrot_b200.worker", "--addr", addr,
                                   "--port", str(port), "--device", str(k), "--pipeline", str(args.pipeline)])
                 for k in range(args.gpus)]
        codes = [p.wait() for p in procs]
        if any(codes):
            raise SystemExit(f"worker exit codes: {codes}")
        return
    nat.ensure_init(args.device)
    if args.pipeline > 0:
        run_pipelined(addr, port, depth=max(2, args.pipeline))
        return
    while do_workload_single(addr, port):
        pass


if __name__ == "__main__":
    main()
