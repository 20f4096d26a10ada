"""In-box writer of the reference's on-disk format.

When the GPUs of one box render a whole level themselves there is no reason to push every
16 MiB payload through the Distributer's single-threaded TCP accept loop
(Distributer.cs:226-297, 397-458) only for the server to classify it (DataChunk.cs:82,87),
serialise it twice on one thread (DataChunk.cs:173-206) and append an index record
(DataStorage.cs:410-427).  This module writes the SAME files the server would have written --
`<parent>/Data/_index.dat` plus one file per Regular chunk -- so the unmodified DataServer/Viewer
can serve a level rendered entirely in-box, and a restarted Distributer sees those chunks as
completed (Distributer.cs:165-175).

Classification and serialisation happen on the device (mb_process_workload_serialized); only
the serialised bytes cross PCIe.  No CPU compute path exists here: without the CUDA library the
render call raises.

Formats:
  index record  u32 level, u32 indexReal, u32 indexImag, i32 type (0 Regular, 1 Never,
                2 Immediate) [, i32 filenameLength, ASCII filename]      DataStorage.cs:358-387
  chunk file    name "{level};{ir};{ii}" + decimal suffix if taken       DataStorage.cs:392-405
                content = 1 codec byte + raw payload or RLE records      DataChunkSerializer.cs
"""
from __future__ import annotations

import ctypes as C
import fcntl
import os
import struct

import numpy as np

from . import _native as nat

DATA_DIRECTORY_NAME = "Data"          # DataStorage.dataDirectoryName (DataStorage.cs:15)
INDEX_FILENAME = "_index.dat"         # DataStorage.indexFilename (DataStorage.cs:18)
ENTRY_REGULAR, ENTRY_NEVER, ENTRY_IMMEDIATE = 0, 1, 2   # IndexEntry.Type (DataStorage.cs:41-49)


class DataStore:
    """`<parent>/Data/` in the reference's layout; append-only like DataStorage.SaveDataChunk."""

    def __init__(self, parent: str):
        self.dir = os.path.join(parent, DATA_DIRECTORY_NAME)
        self.index_path = os.path.join(self.dir, INDEX_FILENAME)
        # SetUpDataDirectoryIfNeeded (DataStorage.cs:131-141)
        os.makedirs(self.dir, exist_ok=True)
        if not os.path.exists(self.index_path):
            open(self.index_path, "ab").close()

    def generate_filename(self, level: int, ir: int, ii: int) -> str:
        """GenerateDataChunkFilename (DataStorage.cs:392-405)."""
        base = f"{level};{ir};{ii}"
        if not os.path.exists(os.path.join(self.dir, base)):
            return base
        suffix = 0
        while os.path.exists(os.path.join(self.dir, base + str(suffix))):
            suffix += 1
        return base + str(suffix)

    def save(self, level: int, ir: int, ii: int, entry_type: int, serialized: bytes | memoryview | None) -> str | None:
        """SaveDataChunk (DataStorage.cs:410-427): append the index record, then write the chunk
        file for Regular entries.  The index append is serialised across PROCESSES with flock
        (the reference only needs an in-process lock, DataStorage.cs:415)."""
        filename = None
        with open(self.index_path, "ab") as idx:
            fcntl.flock(idx, fcntl.LOCK_EX)
            try:
                rec = struct.pack("<IIIi", level, ir, ii, entry_type)
                if entry_type == ENTRY_REGULAR:
                    filename = self.generate_filename(level, ir, ii)
                    name = filename.encode("ascii")
                    rec += struct.pack("<i", len(name)) + name
                    # reserve the name while we still hold the lock (another rank may pick the next suffix)
                    open(os.path.join(self.dir, filename), "ab").close()
                idx.write(rec)
                idx.flush()
            finally:
                fcntl.flock(idx, fcntl.LOCK_UN)
        if filename is not None:
            with open(os.path.join(self.dir, filename), "wb") as f:      # WriteDataToFile (DataStorage.cs:333-350)
                f.write(serialized)
        return filename

    def completed(self):
        """(level, ir, ii, type, filename) for every index record (ReadIndexEntry, DataStorage.cs:187-225)."""
        buf = open(self.index_path, "rb").read()
        pos, out = 0, []
        while pos + 16 <= len(buf):
            level, ir, ii, etype = struct.unpack_from("<IIIi", buf, pos)
            pos += 16
            name = None
            if etype == ENTRY_REGULAR:
                (ln,) = struct.unpack_from("<i", buf, pos)
                pos += 4
                name = buf[pos:pos + ln].decode("ascii")
                pos += ln
            out.append((level, ir, ii, etype, name))
        return out


def render_serialized(level: int, mrd: int, index_real: int, index_imag: int, *, definition: int = nat.CHUNK_DEFINITION,
                      out: np.ndarray | None = None):
    """process_workload + classify + serialise on the device -> (entry_type, memoryview of the
    serialised chunk).  The view aliases `out` (definition^2 + 1 bytes, allocated if omitted)."""
    nat.ensure_init()
    n = int(definition) * int(definition)
    if out is None:
        out = np.empty(n + 1, dtype=np.uint8)
    elif out.dtype != np.uint8 or out.size < n + 1 or not out.flags.c_contiguous:
        raise ValueError("out must be a C-contiguous uint8 array of at least definition^2 + 1 bytes")
    ln, et = C.c_uint64(), C.c_int32()
    nat.check(nat.lib().mb_process_workload_serialized(level, mrd, index_real, index_imag, definition,
                                                       out.ctypes.data_as(C.c_void_p), out.size, C.byref(ln), C.byref(et)))
    return int(et.value), memoryview(out)[:ln.value]


INFO_DTYPE = np.dtype([("out_len", "<u8"), ("runs", "<u4"), ("codec", "<u4"), ("entry_type", "<u4"),
                       ("first_value", "<u4"), ("reserved", "<u4", (2,))])     # mb_chunk_info
assert INFO_DTYPE.itemsize == C.sizeof(nat.ChunkInfo)


def render_serialized_batch(workloads, *, definition: int = nat.CHUNK_DEFINITION):
    """Render + classify + serialise a batch of (level, mrd, ir, ii) in ONE set of launches
    (one escape-kernel launch over all chunks, two serializer launches over all chunks).
    Returns (infos: structured numpy array, fetch(k) -> bytes of chunk k's serialised form)."""
    import torch
    dev = torch.device("cuda", nat.ensure_init())
    lib = nat.lib()
    workloads = list(workloads)
    nb = len(workloads)
    n = int(definition) * int(definition)
    stride = lib.mb_serialize_out_stride(n)
    wsb = lib.mb_serialize_workspace_bytes(n, nb)
    payload = torch.empty(nb * n, dtype=torch.uint8, device=dev)
    out = torch.empty(nb * stride, dtype=torch.uint8, device=dev)
    ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
    info = torch.empty(nb * INFO_DTYPE.itemsize, dtype=torch.uint8, device=dev)
    tiles = []
    for k, (level, mrd, ir, ii) in enumerate(workloads):
        t = nat.chunk_geometry(level, ir, ii, definition, mrd)
        t.out_offset = k * n
        tiles.append(t)
    st = torch.cuda.current_stream(dev).cuda_stream
    nat.check(lib.mb_render_u8(nat.tile_array(tiles), nb, payload.data_ptr(), nat.DTYPE_F64, st))
    nat.check(lib.mb_serialize_chunks(payload.data_ptr(), n, nb, n, out.data_ptr(), stride, ws.data_ptr(), wsb,
                                      info.data_ptr(), st))
    infos = info.cpu().numpy().view(INFO_DTYPE)

    def fetch(k: int) -> bytes:
        ln = int(infos["out_len"][k])
        return out[k * stride:k * stride + ln].cpu().numpy().tobytes()

    return infos, fetch


class BatchSerializer:
    """Double-buffered render + classify + serialise pipeline for the storage writer.

    `submit(workloads)` enqueues one escape-kernel launch and the serializer launches for a batch on
    the compute stream and returns at once; `collect(ticket)` waits for that batch's 32-byte info
    records, copies exactly the serialised bytes of every Regular chunk into page-locked host memory
    on a copy stream and returns them.  With two slots, batch k+1 renders while batch k crosses
    PCIe and is written to disk (DataStorage.WriteDataToFile, DataStorage.cs:333-350)."""

    def __init__(self, definition: int, batch: int, slots: int = 2):
        import torch
        self.torch = torch
        self.dev = torch.device("cuda", nat.ensure_init())
        self.lib = nat.lib()
        self.definition, self.batch = int(definition), max(1, int(batch))
        self.n = self.definition * self.definition
        self.stride = self.lib.mb_serialize_out_stride(self.n)
        self.wsb = self.lib.mb_serialize_workspace_bytes(self.n, self.batch)
        nb = self.batch
        self.slots = [{
            "payload": torch.empty(nb * self.n, dtype=torch.uint8, device=self.dev),
            "out": torch.empty(nb * self.stride, dtype=torch.uint8, device=self.dev),
            "ws": torch.empty(self.wsb, dtype=torch.uint8, device=self.dev),
            "info": torch.empty(nb * INFO_DTYPE.itemsize, dtype=torch.uint8, device=self.dev),
            "info_host": torch.empty(nb * INFO_DTYPE.itemsize, dtype=torch.uint8).pin_memory(),
            "host": torch.empty(nb * self.stride, dtype=torch.uint8).pin_memory(),
            "ready": None, "free": None,
        } for _ in range(max(1, slots))]
        self.compute = torch.cuda.Stream(self.dev)
        self.copy = torch.cuda.Stream(self.dev)
        self._next = 0

    def submit(self, workloads):
        torch = self.torch
        workloads = list(workloads)
        if not 0 < len(workloads) <= self.batch:
            raise ValueError("batch size out of range")
        slot = self.slots[self._next % len(self.slots)]
        self._next += 1
        if slot["free"] is not None:
            slot["free"].synchronize()            # the slot's previous device->host copies are done
        nb, n = len(workloads), self.n
        tiles = []
        for k, (level, mrd, ir, ii) in enumerate(workloads):
            t = nat.chunk_geometry(level, ir, ii, self.definition, mrd)
            t.out_offset = k * n
            tiles.append(t)
        st = self.compute.cuda_stream
        nat.check(self.lib.mb_render_u8(nat.tile_array(tiles), nb, slot["payload"].data_ptr(), nat.DTYPE_F64, st))
        nat.check(self.lib.mb_serialize_chunks(slot["payload"].data_ptr(), n, nb, n, slot["out"].data_ptr(), self.stride,
                                               slot["ws"].data_ptr(), self.wsb, slot["info"].data_ptr(), st))
        with torch.cuda.stream(self.compute):
            slot["info_host"][:nb * INFO_DTYPE.itemsize].copy_(slot["info"][:nb * INFO_DTYPE.itemsize], non_blocking=True)
        slot["ready"] = torch.cuda.Event()
        slot["ready"].record(self.compute)
        return slot, workloads

    def collect(self, ticket):
        """-> [(workload, entry_type, memoryview of the serialised bytes or None)], valid until the
        slot is reused (two submits later)."""
        torch = self.torch
        slot, workloads = ticket
        slot["ready"].synchronize()
        infos = slot["info_host"].numpy()[:len(workloads) * INFO_DTYPE.itemsize].view(INFO_DTYPE)
        self.copy.wait_event(slot["ready"])
        with torch.cuda.stream(self.copy):
            for k in range(len(workloads)):
                if int(infos["entry_type"][k]) == ENTRY_REGULAR:
                    lo, ln = k * self.stride, int(infos["out_len"][k])
                    slot["host"][lo:lo + ln].copy_(slot["out"][lo:lo + ln], non_blocking=True)
        slot["free"] = torch.cuda.Event()
        slot["free"].record(self.copy)
        slot["free"].synchronize()                # ONE wait per batch; the next batch is already rendering
        host = slot["host"].numpy()
        out = []
        for k, w in enumerate(workloads):
            et = int(infos["entry_type"][k])
            lo, ln = k * self.stride, int(infos["out_len"][k])
            out.append((w, et, memoryview(host)[lo:lo + ln] if et == ENTRY_REGULAR else None))
        return out


def render_level_to_store(store: DataStore, level: int, mrd: int, chunks=None, *, definition: int = nat.CHUNK_DEFINITION,
                          skip_completed: bool = True, batch: int = 16, serializer=None) -> int:
    """Render chunks of `level` (all level^2 of them unless `chunks` is given, e.g. an iterator
    that pops indices from the in-box work-stealing queue) straight into the store, `batch` chunks
    per set of launches.  Chunks the index already holds are skipped, like the Distributer's resume
    (Distributer.cs:165-175).

    `chunks` is consumed LAZILY, one batch at a time, and only after the previous batch has been
    submitted: a rank that draws from the shared queue claims work as fast as it renders it, so the
    queue (not a static split) balances the ranks.  Batches are double-buffered: while batch k's
    serialised bytes cross PCIe and are written to disk, batch k+1 is already rendering.
    `serializer` (submit/collect) is injectable for CPU tests."""
    import itertools
    have = {(l, r, i) for (l, r, i, _t, _f) in store.completed()} if skip_completed else set()
    if chunks is None:
        # Distributer.TryGetNextNeededWorkload order (Distributer.cs:335-353): indexReal outer, indexImag inner
        chunks = ((ir, ii) for ir in range(level) for ii in range(level))
    it = iter(chunks)
    ser = serializer if serializer is not None else BatchSerializer(definition, batch)
    batch = max(1, batch)
    done = 0

    def next_batch():
        part = []
        while len(part) < batch:
            got = list(itertools.islice(it, batch - len(part)))      # claims at most what this batch still needs
            if not got:
                break
            part += [(level, mrd, ir, ii) for (ir, ii) in got if (level, ir, ii) not in have]
        return part

    def write(ticket):
        nonlocal done
        for (lv, _m, ir, ii), etype, blob in ser.collect(ticket):
            store.save(lv, ir, ii, etype, blob)
            done += 1

    pending = None
    while True:
        part = next_batch()
        ticket = ser.submit(part) if part else None
        if pending is not None:
            write(pending)                      # overlaps with the batch just submitted
        pending = ticket
        if ticket is None:
            break
    return done


def queue_chunks(queue, rank: int, order, claim: int):
    """Chunk indices for this rank, drawn lazily from the shared work-stealing queue, `claim` at a time."""
    while (g := queue.pop(rank, claim)) is not None:
        yield from order[g[0]:g[0] + g[1]]


def parse_levels(text: str):
    """`-l level:mrd,level:mrd,...` exactly like the server's CLI (Program.cs:227-257)."""
    out = []
    for part in text.split(","):
        level, mrd = part.split(":")
        level, mrd = int(level), int(mrd)
        if level <= 0 or mrd <= 0:
            raise ValueError("levels and maximum recursion depths must be positive")
        out.append((level, mrd))
    return out


def main(argv=None) -> None:
    """In-box replacement for `DistributedMandelbrot -l ... -o DIR` + N workers: renders the level
    settings straight into DIR/Data.  Under torchrun the chunks of each level are handed to the
    ranks by the shared-memory work-stealing queue (one process per GPU); alone it uses GPU 0.

        python -m distributedmandelbrot_b200.storage -l 4:256,10:1024 -o /srv/mandelbrot/
        python -m torch.distributed.run --nproc-per-node 8 -m distributedmandelbrot_b200.storage -l 20:1024 -o DIR
    """
    import argparse
    ap = argparse.ArgumentParser(description=main.__doc__, formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument("-l", "--levels", required=True, help="level:mrd[,level:mrd...] (Program.cs:227-257)")
    ap.add_argument("-o", "--output", default=".", help="parent of the Data/ directory (Program.cs:384-398)")
    ap.add_argument("--definition", type=int, default=nat.CHUNK_DEFINITION)
    ap.add_argument("--batch", type=int, default=16)
    args = ap.parse_args(argv)
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    nat.ensure_init(local)
    store = DataStore(args.output)
    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    for k, (level, mrd) in enumerate(parse_levels(args.levels)):
        order = [(ir, ii) for ir in range(level) for ii in range(level)]     # Distributer.cs:335-353
        if world == 1:
            n = render_level_to_store(store, level, mrd, order, definition=args.definition, batch=args.batch)
        else:
            from .frame import queue_name
            qn = f"{queue_name()}_s{k}"
            q = nat.TileQueue(qn, world, len(order), create=True) if rank == 0 else None
            dist.barrier()
            if rank != 0:
                q = nat.TileQueue(qn)

            # small claims: a level's chunks differ in cost by orders of magnitude, and stealing can only move what is unclaimed
            n = render_level_to_store(store, level, mrd, queue_chunks(q, rank, order, max(1, args.batch // 4)),
                                      definition=args.definition, batch=args.batch)
            dist.barrier()
            q.close()
        print(f"[rank {rank}] level {level} mrd {mrd}: {n} chunk(s) written to {store.dir}", flush=True)
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
