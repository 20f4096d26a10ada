"""CPU restatement of the reference's DataChunk serialisation and storage records.

TEST INFRASTRUCTURE ONLY (like everything under oracle/): the checker for the device
serialiser (csrc/chunk_codec.cuh) and the storage writer (distributedmandelbrot_b200/storage.py).

Restates, byte for byte:
  RawSerializer   DistributedMandelbrot/DataChunkSerializer.cs:29-49   (code 0x00 + payload)
  RLESerializer   DistributedMandelbrot/DataChunkSerializer.cs:51-144  (code 0x01 + (u32 LE run, u8 value)*)
  DataChunk.Serialize / DeserializeData   DataChunk.cs:173-235        (shorter wins, Raw on ties)
  IndexEntry records   DataStorage.cs:187-225 (read), :358-387 (write), types :26-31
Pinned against the reference viewer's own decoder text (DistributedMandelbrotViewer.py:35-60,
executed by oracle/gen_golden.py) through the fixtures in tests/golden/.
"""
from __future__ import annotations

import struct

import numpy as np

RAW_CODE, RLE_CODE = 0x00, 0x01
ENTRY_REGULAR, ENTRY_NEVER, ENTRY_IMMEDIATE = 0, 1, 2      # DataStorage.cs:26-31


def rle_records(data: np.ndarray) -> bytes:
    """RLESerializer.SerializeData (DataChunkSerializer.cs:56-100): one record per maximal run."""
    d = np.ascontiguousarray(data, dtype=np.uint8).ravel()
    if d.size == 0:
        raise ValueError("No end run to write")            # DataChunkSerializer.cs:95-96
    starts = np.flatnonzero(np.r_[True, d[1:] != d[:-1]])
    lengths = np.diff(np.r_[starts, d.size]).astype("<u4")
    rec = np.zeros(starts.size, dtype=[("len", "<u4"), ("val", "u1")])
    rec["len"], rec["val"] = lengths, d[starts]
    return rec.tobytes()


def serialize(data: np.ndarray) -> bytes:
    """DataChunk.Serialize (DataChunk.cs:173-206): try Raw then RLE, keep the strictly shorter."""
    d = np.ascontiguousarray(data, dtype=np.uint8).ravel()
    raw = bytes([RAW_CODE]) + d.tobytes()
    rle = bytes([RLE_CODE]) + rle_records(d)
    return rle if len(rle) < len(raw) else raw


def deserialize(buf: bytes, n: int) -> np.ndarray:
    """DataChunk.DeserializeData (DataChunk.cs:214-235) + the two DeserializeData methods."""
    code, body = buf[0], memoryview(buf)[1:]
    if code == RAW_CODE:
        return np.frombuffer(body[:n], dtype=np.uint8).copy()
    if code != RLE_CODE:
        raise ValueError("No serializer found for chunk file")
    rec = np.frombuffer(body[:len(body) - len(body) % 5], dtype=[("len", "<u4"), ("val", "u1")])
    if (rec["len"] == 0).any():
        raise ValueError("Encountered run of length 0")
    out = np.repeat(rec["val"], rec["len"].astype(np.int64))
    if out.size > n:
        raise ValueError("Data exceeds chunk expected length")
    return out


def entry_type(data: np.ndarray) -> int:
    """IndexEntry.CreateIndexEntryForDataChunk (DataStorage.cs:71-84): Never is tested first."""
    d = np.asarray(data)
    if (d == 0).all():
        return ENTRY_NEVER
    if (d == 1).all():
        return ENTRY_IMMEDIATE
    return ENTRY_REGULAR


def index_entry(level: int, ir: int, ii: int, etype: int, filename: str | None = None) -> bytes:
    """WriteIndexEntry (DataStorage.cs:358-387): u32 level, u32 ir, u32 ii, i32 type,
    and for Regular entries i32 filename length + ASCII filename."""
    out = struct.pack("<IIIi", level, ir, ii, etype)
    if etype == ENTRY_REGULAR:
        name = filename.encode("ascii")
        out += struct.pack("<i", len(name)) + name
    return out


def parse_index(buf: bytes):
    """ReadIndexEntry (DataStorage.cs:187-225) over a whole _index.dat."""
    pos, out = 0, []
    while pos < len(buf):
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
