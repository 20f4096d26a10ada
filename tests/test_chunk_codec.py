"""DataChunk serialisation and the in-box storage writer (f3).

CPU: the restated serializer against fixtures whose bytes the reference viewer's own decoder
accepted (oracle/gen_golden.py), and the storage record format.  GPU (-m gpu): the device
serializer / classifier / storage writer against that restatement, byte for byte."""
import ctypes as C
import os
import struct

import numpy as np
import pytest


@pytest.fixture(scope="module")
def codec():
    import chunk_codec
    return chunk_codec


def test_restated_serializer_matches_golden(codec, golden_meta, golden_arrays):
    for rec in golden_meta["codec"]:
        payload = golden_arrays[rec["key"] + "_payload"]
        blob = codec.serialize(payload)
        assert blob == golden_arrays[rec["key"] + "_serialized"].tobytes(), rec
        assert blob[0] == rec["codec"] and len(blob) == rec["serialized_len"]
        assert np.array_equal(codec.deserialize(blob, payload.size), payload)


def test_known_encodings(codec):
    # an all-7 chunk RLE-encodes as 01 | 00 00 00 01 | 07
    assert codec.serialize(np.full(16_777_216, 7, np.uint8)) == bytes.fromhex("01" "00000001" "07")
    assert codec.serialize(np.array([5, 5, 5, 9], np.uint8)) == bytes([0, 5, 5, 5, 9])        # 2 runs * 5 > 4: raw
    assert codec.serialize(np.array([5] * 11 + [9] * 2, np.uint8)) == b"\x01" + struct.pack("<IB", 11, 5) + struct.pack("<IB", 2, 9)
    assert codec.serialize(np.repeat(np.array([1, 2, 3, 4], np.uint8), 5))[0] == 0            # tie -> Raw (DataChunk.cs:188)
    with pytest.raises(ValueError):
        codec.deserialize(b"\x02abc", 3)
    with pytest.raises(ValueError):
        codec.deserialize(b"\x01" + struct.pack("<IB", 0, 1), 3)


def test_storage_layout_and_records(tmp_path, codec):
    from distributedmandelbrot_b200.storage import DataStore, ENTRY_IMMEDIATE, ENTRY_NEVER, ENTRY_REGULAR
    st = DataStore(str(tmp_path))
    assert os.path.isdir(tmp_path / "Data") and os.path.getsize(tmp_path / "Data" / "_index.dat") == 0
    blob = codec.serialize(np.arange(64, dtype=np.uint8))
    assert st.save(4, 1, 2, ENTRY_REGULAR, blob) == "4;1;2"
    assert st.save(4, 0, 0, ENTRY_IMMEDIATE, None) is None
    assert st.save(4, 1, 1, ENTRY_NEVER, None) is None
    assert st.save(4, 1, 2, ENTRY_REGULAR, blob) == "4;1;20"       # GenerateDataChunkFilename suffix (DataStorage.cs:400-403)
    assert st.save(4, 1, 2, ENTRY_REGULAR, blob) == "4;1;21"
    raw = open(tmp_path / "Data" / "_index.dat", "rb").read()
    want = (codec.index_entry(4, 1, 2, 0, "4;1;2") + codec.index_entry(4, 0, 0, 2) + codec.index_entry(4, 1, 1, 1) +
            codec.index_entry(4, 1, 2, 0, "4;1;20") + codec.index_entry(4, 1, 2, 0, "4;1;21"))
    assert raw == want
    assert raw[:25] == bytes.fromhex("04000000" "01000000" "02000000" "00000000" "05000000") + b"4;1;2"
    assert codec.parse_index(raw) == st.completed()
    assert open(tmp_path / "Data" / "4;1;2", "rb").read() == blob
    # a second store on the same directory appends, it does not truncate (resume, Distributer.cs:165-175)
    st2 = DataStore(str(tmp_path))
    assert len(st2.completed()) == 5


# ---- GPU -----------------------------------------------------------------------------------------

def _serialize_gpu(native, payload):
    out = np.full(payload.size + 1 + 7, 0xEE, np.uint8)
    ln = C.c_uint64()
    native.check(native.lib().mb_serialize_chunk_host(payload.ctypes.data_as(C.c_void_p), payload.size,
                                                       out.ctypes.data_as(C.c_void_p), payload.size + 1, C.byref(ln)))
    assert (out[payload.size + 1:] == 0xEE).all(), "wrote past n + 1 bytes"
    return out[:ln.value].tobytes()


@pytest.mark.gpu
def test_device_serializer_golden_and_shapes(gpu, codec, golden_meta, golden_arrays):
    for rec in golden_meta["codec"]:
        payload = np.ascontiguousarray(golden_arrays[rec["key"] + "_payload"])
        assert _serialize_gpu(gpu, payload) == golden_arrays[rec["key"] + "_serialized"].tobytes(), rec
    rng = np.random.default_rng(3)
    # sizes around the 32-byte thread span, the 8 KiB piece and the 16 KiB tile of the single-pass kernel; run structures that
    # straddle tiles (long runs), that make RLE impossible late / early, and that sit exactly on tile seams
    for n in [1, 2, 15, 16, 17, 31, 32, 33, 4095, 4096, 4097, 8191, 8192, 8193, 16383, 16384, 16385, 24577, 32767, 32768, 32769, 100_003,
              300_000, 1_000_000]:
        for kind in ("runs", "noise", "const", "alt", "long", "dense_then_const", "const_then_noise", "seams"):
            if kind == "runs":
                p = np.repeat(rng.integers(0, 4, n // 3 + 2).astype(np.uint8), rng.integers(1, 12, n // 3 + 2))[:n]
            elif kind == "noise":
                p = rng.integers(0, 256, n).astype(np.uint8)
            elif kind == "const":
                p = np.full(n, 1, np.uint8)
            elif kind == "long":
                p = np.repeat(rng.integers(0, 256, n // 9000 + 2).astype(np.uint8), rng.integers(1, 30000, n // 9000 + 2))[:n]
                if p.size < n:
                    p = np.concatenate([p, np.full(n - p.size, 9, np.uint8)])
            elif kind == "dense_then_const":
                p = np.concatenate([rng.integers(0, 256, n // 10).astype(np.uint8), np.full(n - n // 10, 3, np.uint8)])
            elif kind == "const_then_noise":
                p = np.concatenate([np.full(n // 2, 3, np.uint8), rng.integers(0, 256, n - n // 2).astype(np.uint8)])
            elif kind == "seams":
                p = (np.arange(n) // 8192 % 2).astype(np.uint8)              # every run starts exactly on a tile boundary
            else:
                p = (np.arange(n) % 2).astype(np.uint8)
            p = np.ascontiguousarray(p)
            assert p.size == n
            assert _serialize_gpu(gpu, p) == codec.serialize(p), (n, kind)


@pytest.mark.gpu
def test_device_serializer_tiles_denser_than_the_list(gpu, codec):
    """Tiles with more run starts than the kernel's shared-memory list holds (emitted in two rounds) inside chunks
    that still end up RLE; same payloads as tests/test_codec_emulated.py."""
    rng = np.random.default_rng(9)
    p = np.full(200_000, 3, np.uint8)
    p[50_000:70_000] = rng.integers(0, 256, 20_000)
    q = np.full(114_688, 5, np.uint8)
    q[98_304:] = rng.integers(0, 256, 16_384)
    a = np.full(180_224, 9, np.uint8)
    a[32_768:49_152] = np.arange(16_384) % 2
    a[65_536:81_920] = np.arange(16_384) // 2 % 2
    big = np.full(16_777_216, 1, np.uint8)                       # a native-size chunk, one sixth of it noise
    big[5_000_000:7_700_000] = rng.integers(0, 256, 2_700_000)
    for x in (p, q, a, big):
        want = codec.serialize(x)
        assert want[0] == 1
        assert _serialize_gpu(gpu, np.ascontiguousarray(x)) == want


def _serialize_device_batch(native, chunks, payload_stride=None):
    """mb_serialize_chunks on device buffers (the batched C-ABI entry): list of equal-sized numpy payloads ->
    list of (bytes, info record); the payloads sit payload_stride apart (default: back to back)."""
    import torch
    lib = native.lib()
    n, nb = chunks[0].size, len(chunks)
    ps = payload_stride or n
    stride = lib.mb_serialize_out_stride(n)
    wsb = lib.mb_serialize_workspace_bytes(n, nb)
    src = torch.zeros(nb * ps, dtype=torch.uint8, device="cuda")
    for k, c in enumerate(chunks):
        src[k * ps:k * ps + n] = torch.from_numpy(c).cuda()
    out = torch.full((nb * stride,), 0xEE, dtype=torch.uint8, device="cuda")
    ws = torch.empty(wsb, dtype=torch.uint8, device="cuda")
    info = torch.zeros(nb * 32, dtype=torch.uint8, device="cuda")
    st = torch.cuda.current_stream().cuda_stream
    native.check(lib.mb_serialize_chunks(src.data_ptr(), n, nb, ps, out.data_ptr(), stride, ws.data_ptr(), wsb, info.data_ptr(), st))
    torch.cuda.synchronize()
    recs = info.cpu().numpy()
    res = []
    for k in range(nb):
        ln = int(recs[32 * k:32 * k + 8].view("<u8")[0])
        o = out[k * stride:(k + 1) * stride]
        assert bool((o[n + 1:] == 0xEE).all()), "wrote past n + 1 bytes"
        res.append((o[:ln].cpu().numpy().tobytes(), recs[32 * k:32 * k + 32].view("<u4")))
    return res


@pytest.mark.gpu
def test_device_serializer_unaligned_chunks_of_a_batch(gpu, codec):
    """n not a multiple of 16 with the payloads back to back: every chunk but the first starts off 16-byte alignment
    (the byte-wise load path on the device), and the second batch entry point argument, payload_stride > n."""
    rng = np.random.default_rng(11)
    n = 40_003
    chunks = [np.ascontiguousarray(np.repeat(rng.integers(0, 5, n // 4 + 2).astype(np.uint8), rng.integers(1, 9, n // 4 + 2))[:n]) for _ in range(5)]
    chunks[2] = rng.integers(0, 256, n).astype(np.uint8)                      # one Raw chunk among them
    for ps in (None, n + 13, n + 29):
        for (blob, info), c in zip(_serialize_device_batch(gpu, chunks, ps), chunks):
            assert blob == codec.serialize(c), ps
            assert int(info[2]) == 1 + int(np.count_nonzero(c[1:] != c[:-1]))


@pytest.mark.gpu
def test_device_serializer_more_than_65535_tiles(gpu, codec):
    """1 GiB + 5 bytes = 65 537 tiles of 16 KiB: the tile index is split over two grid dimensions (with padding
    blocks behind the last tile), positions and run lengths need 31 bits.  The expected bytes are written down by
    hand (a 1 GiB payload is not put through the numpy restatement)."""
    import struct
    import torch
    lib = gpu.lib()
    n = 65_536 * 16_384 + 5
    marks = [(0, 9), (16_383, 8), (700_000_001, 7), (n - 1, 6)]               # single bytes that differ from the background 5
    src = torch.full((n,), 5, dtype=torch.uint8, device="cuda")
    for pos, v in marks:
        src[pos] = v
    want, at = b"\x01", 0
    for pos, v in marks:
        if pos > at:
            want += struct.pack("<IB", pos - at, 5)
        want += struct.pack("<IB", 1, v)
        at = pos + 1
    assert at == n
    stride = lib.mb_serialize_out_stride(n)
    wsb = lib.mb_serialize_workspace_bytes(n, 1)
    out = torch.full((stride,), 0xEE, dtype=torch.uint8, device="cuda")
    ws = torch.empty(wsb, dtype=torch.uint8, device="cuda")
    info = torch.zeros(32, dtype=torch.uint8, device="cuda")
    gpu.check(lib.mb_serialize_chunks(src.data_ptr(), n, 1, n, out.data_ptr(), stride, ws.data_ptr(), wsb, info.data_ptr(),
                                      torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    rec = info.cpu().numpy()
    assert int(rec[:8].view("<u8")[0]) == len(want) and int(rec[8:12].view("<u4")[0]) == (len(want) - 1) // 5
    assert out[:len(want)].cpu().numpy().tobytes() == want
    assert bool((out[len(want):len(want) + 4096] == 0xEE).all())


@pytest.mark.gpu
def test_device_serializer_exactly_at_the_raw_rle_threshold(gpu, codec):
    # 5 * runs < n picks RLE, 5 * runs == n keeps Raw (DataChunk.cs:188)
    for runs, n in [(20, 100), (19, 100), (21, 100), (819, 4096), (820, 4096)]:
        lens = np.full(runs, n // runs)
        lens[: n - lens.sum()] += 1
        p = np.ascontiguousarray(np.repeat((np.arange(runs) % 251).astype(np.uint8), lens))
        assert p.size == n
        blob = _serialize_gpu(gpu, p)
        assert blob == codec.serialize(p)
        assert blob[0] == (1 if 5 * runs < n else 0)


@pytest.mark.gpu
def test_native_chunk_serialized_and_classified(gpu, codec, oracle):
    """Full 4096^2 DataChunks through render + classify + serialise on the device."""
    from distributedmandelbrot_b200.storage import render_serialized
    for (level, mrd, ir, ii, want_type) in [(4, 256, 1, 2, 0), (8, 64, 3, 4, 1), (8, 300, 0, 0, 2)]:
        etype, blob = render_serialized(level, mrd, ir, ii)
        payload = oracle.process_workload(level, mrd, ir, ii, 4096)
        assert etype == want_type == codec.entry_type(payload)
        assert bytes(blob) == codec.serialize(payload)
        assert np.array_equal(codec.deserialize(bytes(blob), 4096 * 4096), payload)      # round trip at full size


@pytest.mark.gpu
def test_batched_serializer_and_classification(gpu, codec, oracle):
    """Several chunks per set of launches; entry types fall out of the run counts."""
    from distributedmandelbrot_b200.storage import render_serialized_batch
    d = 160
    wl = [(8, 300, 0, 0), (8, 64, 3, 4), (10, 300, 3, 5), (4, 256, 1, 2), (6, 256, 0, 0), (3, 90, 1, 1), (16, 500, 5, 8)]
    infos, fetch = render_serialized_batch(wl, definition=d)
    for k, w in enumerate(wl):
        payload = oracle.process_workload(*w, d)
        want = codec.serialize(payload)
        assert int(infos["entry_type"][k]) == codec.entry_type(payload), w
        assert int(infos["out_len"][k]) == len(want) and int(infos["codec"][k]) == want[0]
        assert fetch(k) == want, w
    assert sorted(set(infos["entry_type"].tolist())) == [0, 1, 2]


@pytest.mark.gpu
def test_render_level_to_store(gpu, codec, oracle, tmp_path):
    from distributedmandelbrot_b200.storage import DataStore, render_level_to_store
    st = DataStore(str(tmp_path))
    level, mrd, d = 4, 40, 96
    assert render_level_to_store(st, level, mrd, definition=d, batch=5) == 16
    entries = st.completed()
    assert [(e[1], e[2]) for e in entries] == [(ir, ii) for ir in range(4) for ii in range(4)]   # Distributer order
    for (lv, ir, ii, etype, name) in entries:
        payload = oracle.process_workload(lv, mrd, ir, ii, d)
        assert etype == codec.entry_type(payload)
        if etype == 0:
            blob = open(os.path.join(st.dir, name), "rb").read()
            assert blob == codec.serialize(payload)
        else:
            assert name is None
    assert render_level_to_store(st, level, mrd, definition=d) == 0                                # resume: nothing left


def test_levels_cli_syntax():
    from distributedmandelbrot_b200.storage import parse_levels
    assert parse_levels("4:256,10:1024,20:1024") == [(4, 256), (10, 1024), (20, 1024)]      # launchSettings.json:5
    with pytest.raises(ValueError):
        parse_levels("4:0")


@pytest.mark.gpu
def test_storage_cli_single_gpu(gpu, codec, oracle, tmp_path):
    from distributedmandelbrot_b200 import storage
    storage.main(["-l", "2:30,3:20", "-o", str(tmp_path), "--definition", "64", "--batch", "3"])
    entries = storage.DataStore(str(tmp_path)).completed()
    assert [(e[0], e[1], e[2]) for e in entries] == [(2, r, i) for r in range(2) for i in range(2)] + \
                                                   [(3, r, i) for r in range(3) for i in range(3)]
    for (lv, ir, ii, etype, name) in entries:
        payload = oracle.process_workload(lv, 30 if lv == 2 else 20, ir, ii, 64)
        assert etype == codec.entry_type(payload)
        if etype == 0:
            assert open(tmp_path / "Data" / name, "rb").read() == codec.serialize(payload)
