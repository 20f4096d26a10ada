"""The device serializer's kernel text (csrc/chunk_codec.cuh) executed on the CPU (tests/cuda_emul: one fiber per
CUDA thread, CTAs in dispatch order) against the restated reference serializer, byte for byte.

This is what checks the index arithmetic of rle_tiles -- look-back over several windows, records that straddle
tiles, the aligned 16-byte word assembly and its ragged ends, the cut-off at n + 1 bytes -- in the container that
has no GPU; `-m gpu` repeats the same shapes on the real device (tests/test_chunk_codec.py).  The emulator is
test infrastructure, the product never builds it."""
import ctypes as C

import numpy as np
import pytest

from test_escape_emulated import build_emulator


@pytest.fixture(scope="module")
def emul():
    lib = C.CDLL(build_emulator())           # one shared object for the serializer's and the escape kernels' emulation
    lib.emul_serialize_chunks.argtypes = [C.c_void_p, C.c_size_t, C.c_uint32, C.c_size_t, C.c_void_p, C.c_size_t, C.c_void_p, C.c_int,
                                          C.c_int, C.c_uint]
    lib.emul_serialize_chunks.restype = C.c_int
    return lib


@pytest.fixture(scope="module")
def codec():
    import chunk_codec
    return chunk_codec


def _aligned(nbytes, offset):
    """uint8 view of nbytes whose address is `offset` past a 16-byte boundary."""
    buf = np.zeros(nbytes + 48, np.uint8)
    a = (-buf.ctypes.data) % 16 + offset
    return buf[a:a + nbytes]


def _serialize(emul, chunks, keep=1, in_off=0, out_off=0, spans=2, grid_y=65535):
    """chunks: list of equal-sized payloads -> list of (serialised bytes, info record)."""
    n, nb = chunks[0].size, len(chunks)
    pstride = (n + 15) // 16 * 16 + 16
    src = _aligned(nb * pstride, in_off)
    for k, p in enumerate(chunks):
        src[k * pstride:k * pstride + n] = p
    ostride = (n + 1 + 255) // 256 * 256
    out = _aligned(nb * ostride, out_off)
    out[:] = 0xEE
    info = np.zeros(nb * 32, np.uint8)
    assert emul.emul_serialize_chunks(src.ctypes.data, n, nb, pstride, out.ctypes.data, ostride, info.ctypes.data, keep,
                                      spans, grid_y) == 0
    res = []
    for k in range(nb):
        rec = info[32 * k:32 * k + 32]
        ln = int(rec[:8].view("<u8")[0])
        o = out[k * ostride:(k + 1) * ostride]
        assert (o[n + 1:] == 0xEE).all(), "wrote past n + 1 bytes"
        res.append((o[:ln].tobytes(), rec.view("<u4")))
    return res


def _payload(rng, n, kind):
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
        p = (np.arange(n) // 8192 % 2).astype(np.uint8)                  # every run starts exactly on a tile boundary
    elif kind == "pairs":
        p = (np.arange(n) // 2 % 2).astype(np.uint8)                     # 2.5 output bytes per input byte, RLE impossible
    else:
        p = (np.arange(n) % 2).astype(np.uint8)
    p = np.ascontiguousarray(p)
    assert p.size == n
    return p


KINDS = ("runs", "noise", "const", "alt", "long", "dense_then_const", "const_then_noise", "seams", "pairs")


def test_emulated_kernel_small_and_ragged(emul, codec):
    rng = np.random.default_rng(5)
    cases = [(n, ("runs", "noise", "const", "alt")) for n in (1, 2, 15, 16, 17, 31, 32, 33)]
    cases += [(n, KINDS) for n in (4097, 8191, 8192, 8193, 16383, 16384, 16385)]
    for n, kinds in cases:
        for kind in kinds:
            p = _payload(rng, n, kind)
            (blob, info), = _serialize(emul, [p])
            want = codec.serialize(p)
            assert blob == want, (n, kind)
            assert info[3] == want[0] and info[4] == codec.entry_type(p)          # codec, IndexEntry.Type


def test_emulated_kernel_golden_fixtures_and_the_gpu_suite_shapes(emul, codec, golden_meta, golden_arrays):
    """The serialised fixtures the reference viewer's decoder accepted (oracle/gen_golden.py), and the same sizes x
    run structures the device suite runs (tests/test_chunk_codec.py), through both tile geometries."""
    for rec in golden_meta["codec"]:
        payload = np.ascontiguousarray(golden_arrays[rec["key"] + "_payload"])
        assert _serialize(emul, [payload])[0][0] == golden_arrays[rec["key"] + "_serialized"].tobytes(), rec
    rng = np.random.default_rng(3)
    for n in [4095, 4096, 24_577, 32_767, 32_768, 32_769, 100_003, 300_000, 1_000_000]:
        for kind in KINDS:
            p = _payload(rng, n, kind)
            want = codec.serialize(p)
            assert _serialize(emul, [p], keep=1)[0][0] == want, (n, kind)
            assert _serialize(emul, [p], keep=0, spans=1)[0][0] == want, (n, kind)


def test_emulated_kernel_native_chunks(emul, codec, oracle):
    """Whole 4096 x 4096 DataChunks as the worker renders them (16 MiB, 1024 tiles): a regular chunk, a Never and an
    Immediate one, serialised as one batch; the regular one again with every inclusive prefix withheld."""
    wl = [(4, 256, 1, 2), (8, 64, 3, 4), (8, 300, 0, 0)]
    chunks = [oracle.process_workload(*w, 4096) for w in wl]
    want = [codec.serialize(c) for c in chunks]
    for (blob, info), c, w, ws in zip(_serialize(emul, chunks, keep=1), chunks, want, wl):
        assert blob == w, ws
        assert int(info[4]) == codec.entry_type(c) and int(info[3]) == w[0]
    assert [codec.entry_type(c) for c in chunks] == [0, 1, 2]
    assert _serialize(emul, [chunks[0]], keep=2)[0][0] == want[0]


def test_emulated_kernel_one_span_per_thread_and_3d_grid(emul, codec):
    """rle_tiles<1> (8 KiB tiles) and a grid whose tile index is split over y and z, with padding blocks."""
    rng = np.random.default_rng(8)
    for n, kinds in [(8193, KINDS), (40_000, ("runs", "noise", "long", "const_then_noise"))]:
        for kind in kinds:
            p = _payload(rng, n, kind)
            want = codec.serialize(p)
            assert _serialize(emul, [p], spans=1)[0][0] == want, (n, kind)
            assert _serialize(emul, [p], spans=2, grid_y=2)[0][0] == want, (n, kind)
    chunks = [_payload(rng, 50_001, "runs"), _payload(rng, 50_001, "noise"), _payload(rng, 50_001, "long")]
    for spans, gy in [(1, 3), (2, 3)]:
        for (blob, _), p in zip(_serialize(emul, chunks, spans=spans, grid_y=gy), chunks):
            assert blob == codec.serialize(p), (spans, gy)


def test_emulated_kernel_many_tiles_and_lookback_modes(emul, codec):
    """Several tiles: records straddling tile seams, runs longer than 65535 bytes, and the look-back summing
    aggregates (keep = 1: most tiles never publish an inclusive prefix; keep = 2: none but the first does) --
    over more than one window of 128 predecessors in the last case."""
    rng = np.random.default_rng(6)
    for n, kinds, keeps in [(40_961, KINDS, (1,)), (40_961, ("runs", "noise"), (0,)),
                            (100_003, ("runs", "long", "const", "const_then_noise"), (1,)), (300_000, ("runs",), (2,))]:
        for kind in kinds:
            p = _payload(rng, n, kind)
            want = codec.serialize(p)
            for keep in keeps:
                (blob, _), = _serialize(emul, [p], keep=keep)
                assert blob == want, (n, kind, keep)
    # one run of 70 000 bytes (length needs 17 bits, it ends several tiles after it started), then short runs
    p = np.concatenate([np.full(70_000, 7, np.uint8), _payload(rng, 12_345, "runs")])
    (blob, _), = _serialize(emul, [p], keep=1)
    assert blob == codec.serialize(p)
    # 141 tiles of 8 KiB, nobody but tile 0 inclusive: the last tiles walk back through two windows
    p = np.concatenate([_payload(rng, 1_100_000, "long"), _payload(rng, 50_000, "runs")])
    (blob, _), = _serialize(emul, [p], keep=2, spans=1)
    assert blob == codec.serialize(p)


def test_emulated_kernel_tiles_denser_than_the_list(emul, codec):
    """A 16 KiB tile with more than 8192 run starts is emitted piece by piece (two rounds over one list); the
    chunk as a whole still ends up RLE, so every byte of those rounds counts."""
    rng = np.random.default_rng(9)
    # noise island inside a constant chunk; it starts and ends inside tiles and covers whole tiles in between
    p = np.full(200_000, 3, np.uint8)
    p[50_000:70_000] = rng.integers(0, 256, 20_000)
    # ... and a chunk whose LAST tile is the dense one (the end-of-chunk sentinel goes into the second round)
    q = np.full(114_688, 5, np.uint8)
    q[98_304:] = rng.integers(0, 256, 16_384)
    # every byte a run start in one tile (both pieces hold exactly 8192 entries), alternating 2-byte runs elsewhere
    a = np.full(180_224, 9, np.uint8)
    a[32_768:49_152] = np.arange(16_384) % 2
    a[65_536:81_920] = np.arange(16_384) // 2 % 2
    for name, x in (("island", p), ("last", q), ("alt", a)):
        want = codec.serialize(x)
        assert want[0] == 1, name
        assert _serialize(emul, [x], keep=1)[0][0] == want, name
        assert _serialize(emul, [x], spans=1)[0][0] == want, name


def test_emulated_kernel_alignment_batches_and_threshold(emul, codec, oracle):
    rng = np.random.default_rng(7)
    # payload and output off 16-byte alignment: the byte-wise load path, every phase of the word assembly
    p = _payload(rng, 20_001, "runs")
    want = codec.serialize(p)
    for in_off, out_off in [(0, 1), (0, 7), (0, 15), (3, 0), (5, 11)]:
        (blob, _), = _serialize(emul, [p], in_off=in_off, out_off=out_off)
        assert blob == want, (in_off, out_off)
    # 5 * runs < n picks RLE, 5 * runs == n keeps Raw (DataChunk.cs:188)
    for runs, n in [(20, 100), (19, 100), (21, 100), (819, 4096), (820, 4096), (3276, 16384), (3277, 16384)]:
        lens = np.full(runs, n // runs)
        lens[: n - lens.sum()] += 1
        q = np.ascontiguousarray(np.repeat((np.arange(runs) % 251).astype(np.uint8), lens))
        (blob, info), = _serialize(emul, [q])
        assert blob == codec.serialize(q) and blob[0] == (1 if 5 * runs < n else 0), (runs, n)
    # a batch (blockIdx.y = chunk) of real payloads: never / immediate / regular chunks at a small definition
    d = 96
    wl = [(8, 300, 0, 0), (8, 64, 3, 4), (4, 256, 1, 2), (16, 500, 5, 8)]
    chunks = [oracle.process_workload(*w, d) for w in wl]
    for (blob, info), payload, w in zip(_serialize(emul, chunks), chunks, wl):
        want = codec.serialize(payload)
        assert blob == want, w
        assert int(info[4]) == codec.entry_type(payload)


def test_emulated_kernel_random_run_structures(emul, codec):
    """Property test (hypothesis): arbitrary mixtures of run lengths -- single bytes, short runs, runs of several
    tiles -- at arbitrary sizes, input / output alignments, tile geometries and look-back modes serialise to exactly
    the bytes of the restated reference serializer."""
    from hypothesis import given, settings, strategies as st

    run = st.one_of(st.integers(1, 3), st.integers(1, 3), st.integers(4, 40), st.integers(41, 3000), st.integers(8000, 40_000))
    segment = st.tuples(st.lists(run, min_size=1, max_size=400), st.integers(0, 2**32 - 1))

    @settings(max_examples=400, deadline=None, derandomize=True)
    @given(st.lists(segment, min_size=1, max_size=6), st.integers(0, 15), st.integers(0, 15), st.sampled_from([1, 2]), st.sampled_from([0, 1, 2]),
           st.integers(0, 70_000))
    def check(segments, in_off, out_off, spans, keep, cut):
        parts = []
        for lens, seed in segments:
            vals = np.random.default_rng(seed).integers(0, 256, len(lens)).astype(np.uint8)
            parts.append(np.repeat(vals, lens))
        p = np.concatenate(parts)
        p = np.ascontiguousarray(p[:max(1, min(p.size, cut if cut else p.size, 120_000))])
        (blob, info), = _serialize(emul, [p], keep=keep, in_off=in_off, out_off=out_off, spans=spans)
        want = codec.serialize(p)
        assert blob == want
        assert int(info[2]) == 1 + int(np.count_nonzero(p[1:] != p[:-1]))

    check()
