"""Parity tests proper: the sm_100a kernels, called through the C-ABI, against the golden
vectors (reference function text) and the CPU oracle.  Bit-exact: integer/index results."""
import ctypes as C
import hashlib

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _render(native, tiles, n_out, dtype="f64", u8=False):
    import torch
    arr = native.tile_array(tiles)
    out = torch.full((n_out,), 77 if u8 else -7, dtype=torch.uint8 if u8 else torch.int32, device="cuda")
    fn = native.lib().mb_render_u8 if u8 else native.lib().mb_render_counts
    native.check(fn(arr, len(tiles), out.data_ptr(), native.DTYPES[dtype], torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    return out.cpu().numpy()


def test_golden_small_counts_and_u8(gpu, golden_meta, golden_arrays):
    from distributedmandelbrot_b200 import worker
    for rec in golden_meta["small"]:
        d = rec["d"]
        args = (rec["level"], rec["mrd"], rec["ir"], rec["ii"])
        counts = worker.process_workload_counts(*args, definition=d)
        assert np.array_equal(counts.reshape(d, d), golden_arrays[rec["key"] + "_counts"]), rec
        u8 = worker.process_workload(*args, definition=d)
        assert u8.dtype == np.uint8 and u8.shape == (d * d,) and u8.flags.c_contiguous
        assert np.array_equal(u8.reshape(d, d), golden_arrays[rec["key"] + "_u8"]), rec


def test_golden_medium_hashes(gpu, oracle, golden_meta):
    from distributedmandelbrot_b200 import worker
    for rec in golden_meta["medium"]:
        d = rec["d"]
        args = (rec["level"], rec["mrd"], rec["ir"], rec["ii"])
        counts = worker.process_workload_counts(*args, definition=d)
        assert oracle.useful_iterations(counts, rec["mrd"]) == rec["useful"], rec
        assert f"{oracle.fnv1a_words(counts):016x}" == rec["fnv"], rec
        u8 = worker.process_workload(*args, definition=d)
        assert hashlib.sha256(u8.tobytes()).hexdigest() == rec["u8_sha256"], rec


def test_process_workload_native_chunk_is_reference_shape(gpu, golden_meta):
    """process_workload(level, mrd, ir, ii) with the reference's fixed 4096 definition (.py:80)."""
    from distributedmandelbrot_b200 import worker
    rec = [r for r in golden_meta["medium"] if r["d"] == 4096][0]
    out = worker.process_workload(rec["level"], rec["mrd"], rec["ir"], rec["ii"])
    assert out.shape == (4096 * 4096,) and out.dtype == np.uint8
    assert hashlib.sha256(out.tobytes()).hexdigest() == rec["u8_sha256"]


@pytest.mark.parametrize("case", [
    # level, ir, ii, d, mrd
    (1, 0, 0, 200, 500), (2, 1, 0, 129, 1000), (6, 2, 3, 256, 2000), (16, 4, 8, 256, 1024),
    (10, 3, 5, 300, 1024), (800, 250, 420, 160, 10000), (4000000, 1256353, 2131821, 96, 20000),
    (7, 6, 0, 2, 5), (5, 2, 2, 1, 100), (3, 0, 2, 31, 1), (3, 0, 2, 33, 2), (12, 5, 6, 100, 9),
])
def test_counts_vs_oracle(gpu, oracle, case):
    from distributedmandelbrot_b200 import worker
    level, ir, ii, d, mrd = case
    got = worker.process_workload_counts(level, mrd, ir, ii, definition=d).reshape(d, d)
    want = oracle.render_chunk_counts(level, mrd, ir, ii, d)
    assert np.array_equal(got, want), case


@pytest.mark.parametrize("seed", range(6))
def test_random_sub_rectangles_and_strides(gpu, oracle, seed):
    """Ragged tiles: odd origins/sizes, padded pitch, unaligned offsets (scalar-store fallback),
    several tiles of different shape in one launch, int32 and fused-uint8 outputs."""
    rng = np.random.default_rng(100 + seed)
    level = int(rng.integers(1, 30))
    ir, ii = int(rng.integers(0, level)), int(rng.integers(0, level))
    d = int(rng.integers(40, 200))
    mrd = int(rng.choice([2, 64, 300, 1500]))
    chunk = gpu.chunk_geometry(level, ir, ii, d, mrd)
    whole = oracle.render_chunk_counts(level, mrd, ir, ii, d)
    tiles, places, off = [], [], int(rng.integers(0, 5))
    for _ in range(int(rng.integers(1, 7))):
        w, h = int(rng.integers(1, d + 1)), int(rng.integers(1, d + 1))
        x0, y0 = int(rng.integers(0, d - w + 1)), int(rng.integers(0, d - h + 1))
        pitch = w + int(rng.integers(0, 6))
        tiles.append(chunk.sub(x0, y0, w, h, off, pitch))
        places.append((x0, y0, w, h, off, pitch))
        off += pitch * h + int(rng.integers(0, 9))
    for u8 in (False, True):
        if u8 and mrd == 0:
            continue
        out = _render(gpu, tiles, off + 8, u8=u8)
        fill = 77 if u8 else -7
        touched = np.zeros(off + 8, bool)
        for (x0, y0, w, h, o, pitch) in places:
            want = whole[y0:y0 + h, x0:x0 + w]
            if u8:
                want = oracle.counts_to_u8(want, mrd)
            idx = (o + np.arange(h)[:, None] * pitch + np.arange(w)[None, :])
            assert np.array_equal(out[idx], want), (seed, u8, x0, y0, w, h, o, pitch)
            touched[idx] = True
        assert (out[~touched] == fill).all(), "kernel wrote outside its tiles"


def test_bands_equal_whole_chunk(gpu, oracle):
    """Tiling invariance: a chunk rendered as row bands / column strips == rendered whole."""
    level, ir, ii, d, mrd = 10, 3, 5, 512, 1024
    chunk = gpu.chunk_geometry(level, ir, ii, d, mrd)
    whole = _render(gpu, [chunk], d * d).reshape(d, d)
    bands = [chunk.sub(0, y, d, 64, y * d, d) for y in range(0, d, 64)]
    assert np.array_equal(_render(gpu, bands, d * d).reshape(d, d), whole)
    strips = [chunk.sub(x, 0, 100 if x + 100 <= d else d - x, d, x, d) for x in range(0, d, 100)]
    assert np.array_equal(_render(gpu, strips, d * d).reshape(d, d), whole)
    assert np.array_equal(whole, oracle.render_chunk_counts(level, mrd, ir, ii, d))


def test_f32_variant_vs_oracle(gpu, oracle, golden_meta, golden_arrays):
    from distributedmandelbrot_b200 import worker
    for rec in golden_meta["f32"]:
        d = rec["d"]
        got = worker.process_workload_counts(rec["level"], rec["mrd"], rec["ir"], rec["ii"], definition=d, dtype="f32")
        assert np.array_equal(got.reshape(d, d), golden_arrays[rec["key"] + "_counts"]), rec
    for (level, ir, ii, d, mrd) in [(4, 1, 1, 256, 4096), (4, 2, 1, 300, 700), (1, 0, 0, 257, 333)]:
        got = worker.process_workload_counts(level, mrd, ir, ii, definition=d, dtype="f32").reshape(d, d)
        assert np.array_equal(got, oracle.render_chunk_counts(level, mrd, ir, ii, d, dtype="f32"))


def test_calc_mb_value_points(gpu, oracle):
    """The reference's ufunc boundary (.py:39-42,92): arbitrary coordinates, incl. |c| >= 2, the
    real axis, -2, the cusp 0.25, -0.75, and a cloud of points around the boundary."""
    from distributedmandelbrot_b200 import worker
    rng = np.random.default_rng(5)
    th = rng.uniform(0, 2 * np.pi, 3000)
    card_r = 0.5 * np.cos(th) - 0.25 * np.cos(2 * th)
    card_i = 0.5 * np.sin(th) - 0.25 * np.sin(2 * th)
    jit = rng.normal(0, 1e-3, (2, 3000))
    r = np.concatenate([[2.0, -2.0, 0.0, 0.25, -0.75, -0.75, 3.0, -2.0000000000000004, 1e-300, 0.0],
                        card_r + jit[0], rng.uniform(-2, 2, 2000)])
    i = np.concatenate([[2.0, 0.0, 0.0, 0.0, 0.0, 1e-9, 0.0, 0.0, 1e-300, 5e-324],
                        card_i + jit[1], rng.uniform(-2, 2, 2000)])
    for mrd in (1, 2, 50, 3000):
        got = worker.calc_mb_value(r, i, mrd)
        want = np.array([oracle.escape_f64(a, b, mrd) for a, b in zip(r, i)], np.int32)
        assert np.array_equal(got, want), mrd
    # broadcasting like a ufunc
    g = worker.calc_mb_value(np.linspace(-2, 1, 7)[None, :], np.linspace(-1, 1, 5)[:, None], 30)
    assert g.shape == (5, 7)


def test_every_mrd_around_the_block_boundaries(gpu, oracle):
    """The kernel tests every update for the first HEAD = 32 updates, then runs untested blocks of
    K = 8 (f64) / 16 (f32) with an exact replay, then a tested tail: sweep mrd across all of those
    boundaries on a tile that straddles the set boundary."""
    from distributedmandelbrot_b200 import worker
    level, ir, ii, d = 10, 3, 5, 64
    for dtype in ("f64", "f32"):
        for mrd in list(range(0, 84)) + [95, 96, 97, 127, 128, 129, 130, 255, 1000, 1001, 1003]:
            got = worker.process_workload_counts(level, mrd, ir, ii, definition=d, dtype=dtype).reshape(d, d)
            want = oracle.render_chunk_counts(level, mrd, ir, ii, d, dtype=dtype)
            assert np.array_equal(got, want), (dtype, mrd)


def test_points_on_the_radius_2_circle(gpu, oracle):
    """|c| ~= 2 is the one place where "escaped stays escaped" is not provable from rounding
    bounds; such lanes are flagged and take the test-every-update path.  Grid points like -2+0i or
    1.2+1.6i do occur in DataChunks."""
    from distributedmandelbrot_b200 import worker
    rng = np.random.default_rng(11)
    th = np.concatenate([rng.uniform(0, 2 * np.pi, 4000), np.linspace(np.pi - 1e-6, np.pi + 1e-6, 2001)])
    rad = 2.0 + np.concatenate([rng.choice([0.0, 1e-16, -1e-16, 4e-16, -4e-16, 1e-13, -1e-13, 1e-11, -1e-11, 2e-11, -2e-11, 1e-9, -1e-9], th.size)])
    r = np.concatenate([[-2.0, 2.0, 0.0, 0.0, 1.2, 1.6, -1.2, -1.6, 1.2, -2.0, -2.0, -1.9999999999999998],
                        rad * np.cos(th)])
    i = np.concatenate([[0.0, 0.0, 2.0, -2.0, 1.6, 1.2, 1.6, -1.2, -1.6, 5e-324, 1e-17, 0.0],
                        rad * np.sin(th)])
    for mrd in (2, 33, 41, 500, 5000):
        got = worker.calc_mb_value(r, i, mrd)
        want = np.array([oracle.escape_f64(a, b, mrd) for a, b in zip(r, i)], np.int32)
        assert np.array_equal(got, want), mrd
    # chunks whose corner/edge lies on the circle: level 1 (corners at +-2), level 5 (1.2 / 1.6 grid lines)
    for (level, ir, ii, d, mrd) in [(1, 0, 0, 129, 300), (2, 0, 1, 65, 300), (5, 4, 4, 101, 200), (5, 0, 4, 201, 200)]:
        got = worker.process_workload_counts(level, mrd, ir, ii, definition=d).reshape(d, d)
        assert np.array_equal(got, oracle.render_chunk_counts(level, mrd, ir, ii, d)), (level, ir, ii)


def test_escape_decided_by_the_last_unit_of_the_high_word(gpu, oracle):
    """tests/golden/threshold_points.json (oracle/threshold_points.c): first escape at a block end with
    4 <= |z|^2 < 4 + 2^-18, i.e. the kernel's integer compare `hi32(m) >= 0x40100000` holds with equality."""
    import json
    import os
    from distributedmandelbrot_b200 import worker
    pts = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "threshold_points.json")))["points"]
    r = np.array([float.fromhex(p["r"]) for p in pts])
    i = np.array([float.fromhex(p["i"]) for p in pts])
    n = np.array([p["n"] for p in pts], np.int32)
    assert np.array_equal(worker.calc_mb_value(r, i, 301), n)
    for mrd in (9, 25, 41, 49, 137):
        assert np.array_equal(worker.calc_mb_value(r, i, mrd), np.where(n < mrd, n, 0)), mrd


def test_calc_mb_value_non_finite_and_huge_inputs(gpu, oracle):
    """The ufunc boundary accepts any float64: inf, NaN and coordinates far outside the plane must
    give what the literal loop gives (e.g. NaN >= 4 is False forever -> 0)."""
    from distributedmandelbrot_b200 import worker
    # ... and tiny coordinates, where x*y goes subnormal and fl(xy)*2 != fl(2xy): the r = 1.5*2^-44, i = 2^-1030
    # family takes the literal (2x)*y path
    vals = [0.0, 1.0, -3.0, 4.0, 4.000001, 1e10, -1e154, 1e200, 1.7e308, np.inf, -np.inf, np.nan, 5e-324, 2.5,
            1.5 * 2.0 ** -44, 2.0 ** -1030, -3 * 2.0 ** -1040, 2.0 ** -501, 2.0 ** -499, 1.75 * 2.0 ** -520, -0.7, 2.0 ** -1022]
    r = np.array([a for a in vals for _ in vals])
    i = np.array([b for _ in vals for b in vals])
    with np.errstate(all="ignore"):
        for mrd in (2, 3, 40, 300):
            got = worker.calc_mb_value(r, i, mrd)
            want = np.array([oracle.escape_f64(a, b, mrd) for a, b in zip(r, i)], np.int32)
            assert np.array_equal(got, want), (mrd, np.flatnonzero(got != want)[:5])


def test_u8_epilogue_standalone_and_fused(gpu, golden_meta, golden_arrays):
    import torch
    for rec in golden_meta["u8"]:
        mrd = rec["mrd"]
        counts = torch.arange(mrd, dtype=torch.int32, device="cuda")
        out = torch.empty(mrd, dtype=torch.uint8, device="cuda")
        gpu.check(gpu.lib().mb_counts_to_u8(counts.data_ptr(), out.data_ptr(), mrd, mrd, None))
        torch.cuda.synchronize()
        assert np.array_equal(out.cpu().numpy(), golden_arrays[rec["key"]]), mrd
    # unaligned pointers take the scalar path
    counts = torch.arange(1001, dtype=torch.int32, device="cuda")[1:]
    out = torch.empty(1001, dtype=torch.uint8, device="cuda")[1:]
    gpu.check(gpu.lib().mb_counts_to_u8(counts.data_ptr(), out.data_ptr(), 1000, 1001, None))
    torch.cuda.synchronize()
    n = np.arange(1, 1001, dtype=np.int64)
    assert np.array_equal(out.cpu().numpy(), (((n * 256 + 1000) // 1001) & 0xFF).astype(np.uint8))


def test_useful_iterations_and_classify(gpu, oracle):
    import torch
    from distributedmandelbrot_b200 import worker
    counts = worker.process_workload_counts(10, 1024, 3, 5, definition=384)
    dev = torch.from_numpy(counts).cuda()
    acc = torch.zeros(1, dtype=torch.int64, device="cuda")
    gpu.check(gpu.lib().mb_useful_iterations(dev.data_ptr(), dev.numel(), 1024, acc.data_ptr(), None))
    assert int(acc.item()) == oracle.useful_iterations(counts, 1024)
    flags = torch.empty(2, dtype=torch.int32, device="cuda")
    for payload, want in [(np.zeros(5000, np.uint8), [1, 0]), (np.ones(5000, np.uint8), [0, 1]),
                          (np.r_[np.zeros(4999, np.uint8), np.uint8(1)], [0, 0])]:
        p = torch.from_numpy(payload).cuda()
        gpu.check(gpu.lib().mb_classify_u8(p.data_ptr(), p.numel(), flags.data_ptr(), None))
        assert flags.cpu().tolist() == want
    # the reference's own classification cases: an all-never and an all-immediate chunk
    never = worker.process_workload(8, 300, 3, 4, definition=128)         # inside the main cardioid
    assert not never.any()
    immediate = worker.process_workload(4, 100, 0, 0, definition=64)      # far corner: escapes at n <= 2
    assert immediate.min() >= 1


def test_invalid_arguments_raise(gpu):
    from distributedmandelbrot_b200 import worker
    for args in [(0, 10, 0, 0), (4, 10, 4, 0), (4, 10, 0, 4), (4, 0, 1, 1)]:
        with pytest.raises(ValueError):
            worker.process_workload(*args, definition=16)
    with pytest.raises(ValueError):
        worker.process_workload(4, 2**31, 1, 1, definition=16)
    with pytest.raises(ValueError):
        worker.process_workload(4, 10, 1, 1, definition=16, out=np.empty(5, np.uint8))


# ---- BASELINE sizes: size-independent properties ------------------------------------------------

def test_cfg2_full_frame_vs_oracle(gpu, oracle):
    """cfg2 (4096^2 whole plane, mrd 1024, 256 chunks of 256^2) is small enough for a full compare."""
    from distributedmandelbrot_b200.frame import CONFIGS, FrameRenderer
    spec = CONFIGS["cfg2"]
    fr = FrameRenderer(spec, device=0)
    fr.render_all()
    got = fr.out.cpu().numpy().reshape(spec.n_chunks, spec.definition, spec.definition)
    total = 0
    for c, (ir, ii) in enumerate(spec.chunk_indices()):
        want = oracle.render_chunk_counts(spec.level, spec.mrd, ir, ii, spec.definition)
        assert np.array_equal(got[c], want), (ir, ii)
        total += oracle.useful_iterations(want, spec.mrd)
    assert fr.useful_iterations() == total


def test_full_size_chunk_properties(gpu, oracle):
    """A native 4096^2 chunk of the cfg4 frame at mrd = 10000: determinism, tiling invariance,
    mrd-monotonicity (a pixel that escaped at n under mrd1 escapes at the same n under any larger
    mrd; one that did not must give 0 or >= mrd1), fused uint8 == epilogue(int32), plus an exact
    comparison of a strided sample of rows against the oracle."""
    import torch
    from distributedmandelbrot_b200.frame import CONFIGS
    spec = CONFIGS["cfg4"]
    d = spec.definition
    ir, ii = 250, 420                       # the chunk whose corner is the zoom centre -0.75+0.1i
    hi = gpu.chunk_geometry(spec.level, ir, ii, d, 10000)
    lo = gpu.chunk_geometry(spec.level, ir, ii, d, 700)
    a = _render(gpu, [hi], d * d)
    assert np.array_equal(a, _render(gpu, [hi], d * d))                                # idempotent
    bands = [hi.sub(0, y, d, 512, y * d, d) for y in range(0, d, 512)]
    assert np.array_equal(a, _render(gpu, bands, d * d))                               # tiling invariance
    b = _render(gpu, [lo], d * d)
    esc = b != 0
    assert np.array_equal(a[esc], b[esc])
    assert ((a[~esc] == 0) | (a[~esc] >= 700)).all()
    u8 = _render(gpu, [hi], d * d, u8=True)
    assert np.array_equal(u8, oracle.counts_to_u8(a, 10000))
    rows = list(range(0, d, 293))
    for y in rows:
        want = oracle.render_chunk_counts(spec.level, 10000, ir, ii, d, x0=0, y0=y, w=d, h=1)
        assert np.array_equal(a.reshape(d, d)[y], want[0]), y


def test_process_workloads_pipeline_equals_single_calls(gpu, oracle):
    """The batched/pipelined worker entry point returns exactly what process_workload returns."""
    from distributedmandelbrot_b200 import worker
    wl = [(10, 300, 3, 5), (4, 64, 0, 0), (16, 500, 5, 8), (10, 300, 4, 5), (8, 100, 3, 4), (3, 200, 1, 1), (2, 50, 1, 0)]
    d = 192
    got = worker.process_workloads(wl, definition=d, depth=3)
    assert len(got) == len(wl)
    for w, g in zip(wl, got):
        assert g.dtype == np.uint8 and g.shape == (d * d,)
        assert np.array_equal(g, worker.process_workload(*w, definition=d)), w
        assert np.array_equal(g, oracle.process_workload(*w, d)), w
    seen = []
    worker.process_workloads(wl[:2], definition=d, consume=lambda w, p: seen.append((w, int(p.sum()))))
    assert [s[0] for s in seen] == wl[:2] and seen[0][1] == int(got[0].sum())
    assert worker.process_workloads([], definition=d) == []


def test_full_size_f32_chunk_properties(gpu, oracle):
    """cfg3 (FP32, mrd 4096) at the native 4096^2 chunk size: determinism, tiling invariance,
    fused uint8 == epilogue(int32), and strided rows against the binary32 oracle."""
    from distributedmandelbrot_b200.frame import CONFIGS
    spec = CONFIGS["cfg3"]
    d, mrd = spec.definition, spec.mrd
    chunk = gpu.chunk_geometry(spec.level, 1, 1, d, mrd)
    a = _render(gpu, [chunk], d * d, dtype="f32")
    assert np.array_equal(a, _render(gpu, [chunk], d * d, dtype="f32"))
    bands = [chunk.sub(0, y, d, 256, y * d, d) for y in range(0, d, 256)]
    assert np.array_equal(a, _render(gpu, bands, d * d, dtype="f32"))
    assert np.array_equal(_render(gpu, [chunk], d * d, dtype="f32", u8=True), oracle.counts_to_u8(a, mrd))
    for y in range(0, d, 373):
        want = oracle.render_chunk_counts(spec.level, mrd, 1, 1, d, dtype="f32", x0=0, y0=y, w=d, h=1)
        assert np.array_equal(a.reshape(d, d)[y], want[0]), y


def test_cfg5_geometry_rows_vs_oracle(gpu, oracle):
    """cfg5 (level 4,000,000, mrd 50000): pixel pitch 2.4e-10, coordinates differ in the last few
    bits -- rows of a boundary chunk and of the frame's corner chunks against the oracle."""
    from distributedmandelbrot_b200.frame import CONFIGS
    spec = CONFIGS["cfg5"]
    d, mrd = spec.definition, spec.mrd
    for (ir, ii, rows) in [(spec.ir0 + 2, spec.ii0 + 5, (0, 1777, 4095)), (spec.ir0, spec.ii0, (4095,)),
                           (spec.ir0 + 7, spec.ii0 + 7, (0,))]:
        chunk = gpu.chunk_geometry(spec.level, ir, ii, d, mrd)
        tiles = [chunk.sub(0, y, d, 1, k * d, d) for k, y in enumerate(rows)]
        got = _render(gpu, tiles, len(rows) * d).reshape(len(rows), d)
        for k, y in enumerate(rows):
            want = oracle.render_chunk_counts(spec.level, mrd, ir, ii, d, x0=0, y0=y, w=d, h=1)
            assert np.array_equal(got[k], want[0]), (ir, ii, y)


def test_cfg4_full_frame_properties(gpu, oracle, fullsize):
    """BASELINE config 4 at its FULL size (65536^2, mrd 10000, 16 GiB of counts): the one-launch
    frame equals the same frame rendered chunk by chunk (tiling/launch invariance), the per-chunk
    useful-iteration sums add up to the frame total (a checksum of checksums), every count is in
    [0, mrd-1], and sampled rows agree with the oracle."""
    import torch
    from distributedmandelbrot_b200.frame import CONFIGS, FrameRenderer
    spec = CONFIGS["cfg4"]
    d, mrd, n = spec.definition, spec.mrd, spec.definition ** 2
    fr = FrameRenderer(spec, device=0)
    fr.render_all()
    torch.cuda.synchronize()
    total = fr.useful_iterations()
    # the frame total is the sum of the CPU oracle's per-chunk digests (tests/golden/fullsize_cfg4.json,
    # oracle/gen_fullsize_golden.py) -- no constant here originates from a CUDA build
    doc = fullsize("cfg4")
    assert doc["complete"] and total == doc["useful_total"] == sum(c["useful"] for c in doc["chunks"].values())
    assert int(fr.out.min()) >= 0 and int(fr.out.max()) <= mrd - 1
    per_chunk = torch.empty(n, dtype=torch.int32, device="cuda")
    acc = torch.zeros(1, dtype=torch.int64, device="cuda")
    lib = gpu.lib()
    for c, (ir, ii) in enumerate(spec.chunk_indices()):
        t = gpu.chunk_geometry(spec.level, ir, ii, d, mrd)
        gpu.check(lib.mb_render_counts(gpu.tile_array([t]), 1, per_chunk.data_ptr(), 0, torch.cuda.current_stream().cuda_stream))
        assert torch.equal(per_chunk, fr.out[c * n:(c + 1) * n]), (ir, ii)
        gpu.check(lib.mb_useful_iterations(per_chunk.data_ptr(), n, mrd, acc.data_ptr(), torch.cuda.current_stream().cuda_stream))
    assert int(acc.item()) == total
    frame = fr.out.view(spec.n_chunks, d, d)
    for c in range(5, spec.n_chunks, 37):
        ir, ii = spec.chunk_indices()[c]
        y = (c * 613) % d
        want = oracle.render_chunk_counts(spec.level, mrd, ir, ii, d, x0=0, y0=y, w=d, h=1)
        assert np.array_equal(frame[c, y].cpu().numpy(), want[0]), (ir, ii, y)


def test_peer_frame_buffer_alloc_and_render(gpu, oracle):
    """The exporter side of the peer-memory frame buffer (mb_peer_alloc + zero-copy torch view);
    the importer side needs a second process/GPU and is exercised by `bench.py --gather p2p`."""
    import torch
    from distributedmandelbrot_b200.frame import _RawCudaArray
    d = 128
    ptr, handle = C.c_void_p(), (C.c_ubyte * 64)()
    gpu.check(gpu.lib().mb_peer_alloc(d * d * 4, C.byref(ptr), handle))
    try:
        assert any(handle)
        t = torch.as_tensor(_RawCudaArray(ptr.value, d * d), device="cuda")
        assert t.data_ptr() == ptr.value and t.dtype == torch.int32
        chunk = gpu.chunk_geometry(10, 3, 5, d, 500)
        gpu.check(gpu.lib().mb_render_counts(gpu.tile_array([chunk]), 1, ptr, 0, torch.cuda.current_stream().cuda_stream))
        torch.cuda.synchronize()
        assert np.array_equal(t.cpu().numpy().reshape(d, d), oracle.render_chunk_counts(10, 500, 3, 5, d))
        del t
    finally:
        gpu.check(gpu.lib().mb_peer_free(ptr))


def test_probe_chunk_costs_ranks_chunks(gpu):
    """The low-resolution cost probe orders chunks like their true cost (used for LPT dispatch)."""
    from distributedmandelbrot_b200.frame import CONFIGS, FrameRenderer, probe_chunk_costs
    spec = CONFIGS["cfg2"]
    costs = probe_chunk_costs(spec, samples=32)
    fr = FrameRenderer(spec, device=0)
    fr.render_all()
    full = fr.out.view(spec.n_chunks, -1)
    import torch
    true = torch.where(full == 0, torch.full_like(full, spec.mrd - 1), full).sum(dim=1, dtype=torch.int64).cpu().numpy()
    est = np.array(costs, dtype=np.float64) * (spec.definition / 32) ** 2
    heavy = np.argsort(-true)[:20]
    assert set(np.argsort(-est)[:30]) >= set(heavy[:10])          # the heaviest chunks are found
    assert np.corrcoef(est, true)[0, 1] > 0.95
