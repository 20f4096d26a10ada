"""The hot path's kernel text (csrc/escape_kernels.cuh, csrc/escape_refill.cuh -- unmodified, same template
arguments as the product launches) executed on the CPU (tests/cuda_emul: one fiber per CUDA thread, every *_rn
intrinsic one IEEE operation, built with -ffp-contract=off) against the golden vectors and the CPU oracle, bit for bit.

It checks the LOGIC of the kernels -- coordinates from the 88-byte descriptor, the tested head, deferred blocks with
the exact replay, the |c| ~ 2 annulus, the literal path for exotic inputs, the warp votes, the periodicity shortcut,
the persistent-warp lane refill with its patch counter, the shuffle-packed 128-/64-bit stores and their scalar
fallbacks, the fused uint8 epilogue -- in the container that has no GPU.  `-m gpu` runs the same cases and the
full-size frames on the device (tests/test_gpu_parity.py, tests/test_gpu_fullsize.py); what only the device can show
(that nvcc emits non-contracted arithmetic) stays there.  The emulator is test infrastructure, never the product."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
EMUL = os.path.join(HERE, "cuda_emul")
CSRC = os.path.join(os.path.dirname(HERE), "distributedmandelbrot_b200", "csrc")
SO = os.path.join(EMUL, "libkernels_emul.so")
SOURCES = ["cuda_emul.cpp", "codec_emul.cpp", "escape_emul.cpp"]
VARIANTS = [(False, False), (True, False), (False, True), (True, True)]        # (periodicity shortcut, lane refill)


def build_emulator() -> str:
    """g++ the kernels + the fiber runtime into one shared object (shared with tests/test_codec_emulated.py).
    MB_EMUL_SO overrides it with a prebuilt one (scripts/mutation_check.py runs the suites against mutated kernels)."""
    if os.environ.get("MB_EMUL_SO"):
        return os.environ["MB_EMUL_SO"]
    deps = [os.path.join(EMUL, f) for f in SOURCES + ["cuda_emul.h"]] + \
           [os.path.join(CSRC, f) for f in ("chunk_codec.cuh", "escape_kernels.cuh", "escape_refill.cuh")]
    if not os.path.exists(SO) or os.path.getmtime(SO) < max(os.path.getmtime(d) for d in deps):
        env = {k: v for k, v in os.environ.items() if k not in ("CC", "CXX")}
        # -fsanitize=alignment: a misaligned 32- / 64- / 128-bit access is a fault on the device; x86 would let it pass
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off", "-fsanitize=alignment",
                               "-fno-sanitize-recover=alignment", "-Wall", "-Wno-unknown-pragmas", "-Wno-unused-but-set-variable",
                               "-I", EMUL, "-I", CSRC, "-o", SO] + [os.path.join(EMUL, f) for f in SOURCES], env=env)
    return SO


@pytest.fixture(scope="module")
def emul(native):
    lib = C.CDLL(build_emulator())
    for fn in (lib.emul_render_counts, lib.emul_render_u8):
        fn.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p, C.c_int, C.c_void_p]
        fn.restype = C.c_int
    lib.emul_calc_mb_value.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_void_p, C.c_int]
    lib.emul_calc_mb_value.restype = C.c_int
    return lib


def _render(emul, native, tiles, n_out, dtype="f64", u8=False, executed=None):
    arr = native.tile_array(tiles)
    out = np.full(n_out, 77 if u8 else -7, np.uint8 if u8 else np.int32)
    fn = emul.emul_render_u8 if u8 else emul.emul_render_counts
    ex = C.c_ulonglong(0)
    assert fn(arr, len(tiles), out.ctypes.data, native.DTYPES[dtype], C.byref(ex)) == 0
    if executed is not None:
        executed.append(ex.value)
    return out


def _chunk(emul, native, level, mrd, ir, ii, d, dtype="f64", u8=False, shortcut=False, refill=False):
    t = native.chunk_geometry(level, ir, ii, d, mrd, shortcut=shortcut, refill=refill)
    return _render(emul, native, [t], d * d, dtype=dtype, u8=u8).reshape(d, d)


def _calc(emul, native, r, i, mrd, dtype="f64"):
    r, i = np.ascontiguousarray(r, np.float64), np.ascontiguousarray(i, np.float64)
    out = np.full(r.size, -7, np.int32)
    assert emul.emul_calc_mb_value(r.ctypes.data, i.ctypes.data, r.size, mrd, out.ctypes.data, native.DTYPES[dtype]) == 0
    return out


def test_emulated_kernels_reproduce_the_reference_text_goldens(emul, native, golden_meta, golden_arrays):
    """tests/golden/*: outputs of the reference's own function text (oracle/gen_golden.py), all four kernel variants."""
    for rec in golden_meta["small"]:
        d, args = rec["d"], (rec["level"], rec["mrd"], rec["ir"], rec["ii"])
        for sc, rf in VARIANTS:
            assert np.array_equal(_chunk(emul, native, *args, d, shortcut=sc, refill=rf), golden_arrays[rec["key"] + "_counts"]), (rec, sc, rf)
            assert np.array_equal(_chunk(emul, native, *args, d, u8=True, shortcut=sc, refill=rf), golden_arrays[rec["key"] + "_u8"]), (rec, sc, rf)
    for rec in golden_meta["f32"]:
        d = rec["d"]
        got = _chunk(emul, native, rec["level"], rec["mrd"], rec["ir"], rec["ii"], d, dtype="f32")
        assert np.array_equal(got, golden_arrays[rec["key"] + "_counts"]), rec


@pytest.mark.parametrize("case", [
    # level, ir, ii, d, mrd      (the cases of tests/test_gpu_parity.py::test_counts_vs_oracle)
    (1, 0, 0, 200, 500), (2, 1, 0, 129, 1000), (6, 2, 3, 256, 2000), (16, 4, 8, 256, 1024),
    (10, 3, 5, 300, 1024), (800, 250, 420, 160, 10000), (4000000, 1256353, 2131821, 96, 20000),
    (7, 6, 0, 2, 5), (5, 2, 2, 1, 100), (3, 0, 2, 31, 1), (3, 0, 2, 33, 2), (12, 5, 6, 100, 9),
])
def test_emulated_counts_vs_oracle(emul, native, oracle, case):
    level, ir, ii, d, mrd = case
    want = oracle.render_chunk_counts(level, mrd, ir, ii, d)
    for sc, rf in VARIANTS:
        assert np.array_equal(_chunk(emul, native, level, mrd, ir, ii, d, shortcut=sc, refill=rf), want), (case, sc, rf)
    want32 = oracle.render_chunk_counts(level, mrd, ir, ii, d, dtype="f32")
    for sc, rf in ((False, False), (True, True)):
        assert np.array_equal(_chunk(emul, native, level, mrd, ir, ii, d, dtype="f32", shortcut=sc, refill=rf), want32), (case, sc, rf)


@pytest.mark.parametrize("seed", range(6))
def test_emulated_random_sub_rectangles_and_strides(emul, native, oracle, seed):
    """Ragged tiles: odd origins/sizes, padded pitch, unaligned offsets (scalar-store fallback), several tiles of
    different shape in one launch, int32 and fused-uint8 outputs; nothing outside the tiles is written."""
    rng = np.random.default_rng(100 + seed)
    level = int(rng.integers(1, 30))
    ir, ii = int(rng.integers(0, level)), int(rng.integers(0, level))
    d = int(rng.integers(40, 200))
    mrd = int(rng.choice([2, 64, 300, 1500]))
    sc, rf = VARIANTS[seed % 4]
    chunk = native.chunk_geometry(level, ir, ii, d, mrd, shortcut=sc, refill=rf)
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
        out = _render(emul, native, tiles, off + 8, u8=u8)
        touched = np.zeros(off + 8, bool)
        for (x0, y0, w, h, o, pitch) in places:
            want = whole[y0:y0 + h, x0:x0 + w]
            if u8:
                want = oracle.counts_to_u8(want, mrd)
            idx = (o + np.arange(h)[:, None] * pitch + np.arange(w)[None, :])
            assert np.array_equal(out[idx], want), (seed, u8, x0, y0, w, h, o, pitch)
            touched[idx] = True
        assert (out[~touched] == (77 if u8 else -7)).all(), "kernel wrote outside its tiles"


def test_emulated_every_mrd_around_the_block_boundaries(emul, native, oracle):
    """Tested head of 32 updates, untested blocks of 8 (f64) / 16 (f32) with exact replay, tested tail: sweep mrd
    across all of those boundaries on a tile that straddles the set boundary, patch and refill kernels."""
    level, ir, ii, d = 10, 3, 5, 48
    for dtype in ("f64", "f32"):
        for mrd in list(range(0, 84)) + [95, 96, 97, 127, 128, 129, 130, 255, 1000, 1001, 1003]:
            want = oracle.render_chunk_counts(level, mrd, ir, ii, d, dtype=dtype)
            for sc, rf in ((False, False), (True, True)) if mrd % 3 else VARIANTS:
                assert np.array_equal(_chunk(emul, native, level, mrd, ir, ii, d, dtype=dtype, shortcut=sc, refill=rf), want), (dtype, mrd, sc, rf)


def test_emulated_points_on_the_radius_2_circle_and_exotic_inputs(emul, native, oracle):
    """calc_mb_value (escape_points_kernel): |c| ~ 2 lanes take the test-every-update path; inf / NaN / huge / tiny
    coordinates take the literal (2x)*y path with an IEEE compare."""
    rng = np.random.default_rng(11)
    th = np.concatenate([rng.uniform(0, 2 * np.pi, 1500), np.linspace(np.pi - 1e-6, np.pi + 1e-6, 501)])
    rad = 2.0 + rng.choice([0.0, 1e-16, -1e-16, 4e-16, -4e-16, 1e-13, -1e-13, 1e-11, -1e-11, 2e-11, -2e-11, 1e-9, -1e-9], th.size)
    r = np.concatenate([[-2.0, 2.0, 0.0, 0.0, 1.2, 1.6, -1.2, -1.6, 1.2, -2.0, -2.0, -1.9999999999999998, 0.25, -0.75, -0.75, 3.0, 1e-300, 0.0],
                        rad * np.cos(th), rng.uniform(-2, 2, 1500)])
    i = np.concatenate([[0.0, 0.0, 2.0, -2.0, 1.6, 1.2, 1.6, -1.2, -1.6, 5e-324, 1e-17, 0.0, 0.0, 0.0, 1e-9, 0.0, 1e-300, 5e-324],
                        rad * np.sin(th), rng.uniform(-2, 2, 1500)])
    for mrd in (1, 2, 33, 41, 500, 3000):
        want = np.array([oracle.escape_f64(a, b, mrd) for a, b in zip(r, i)], np.int32)
        assert np.array_equal(_calc(emul, native, r, i, mrd), want), mrd
    vals = [0.0, 1.0, -3.0, 4.0, 4.000001, 1e10, -1e154, 1e200, 1.7e308, np.inf, -np.inf, np.nan, 5e-324, 2.5,
            1.5 * 2.0 ** -44, 2.0 ** -1030, -3 * 2.0 ** -1040, 2.0 ** -501, 2.0 ** -499, 1.75 * 2.0 ** -520, -0.7, 2.0 ** -1022]
    r = np.array([a for a in vals for _ in vals])
    i = np.array([b for _ in vals for b in vals])
    with np.errstate(all="ignore"):
        for mrd in (2, 3, 40, 300):
            want = np.array([oracle.escape_f64(a, b, mrd) for a, b in zip(r, i)], np.int32)
            got = _calc(emul, native, r, i, mrd)
            assert np.array_equal(got, want), (mrd, np.flatnonzero(got != want)[:5])
    # chunks whose corner / edge lies on the circle: level 1 (corners at +-2), level 5 (1.2 / 1.6 grid lines)
    for (level, ir, ii, d, mrd) in [(1, 0, 0, 129, 300), (2, 0, 1, 65, 300), (5, 4, 4, 101, 200), (5, 0, 4, 201, 200)]:
        want = oracle.render_chunk_counts(level, mrd, ir, ii, d)
        for sc, rf in VARIANTS:
            assert np.array_equal(_chunk(emul, native, level, mrd, ir, ii, d, shortcut=sc, refill=rf), want), (level, ir, ii, sc, rf)


def test_emulated_cfg2_frame_against_the_oracle_digests(emul, native, oracle, fullsize):
    """BASELINE cfg2 at its full size (4096 x 4096 whole plane, mrd 1024, 256 chunks of 256 x 256): every chunk's
    (useful iterations, FNV, zero count) equals the digest the CPU oracle alone produced (tests/golden/fullsize_cfg2.json);
    bands of 64 rows with the lane-refill kernel + shortcut for a quarter of the chunks, whole chunks otherwise."""
    from distributedmandelbrot_b200.frame import CONFIGS
    spec, doc = CONFIGS["cfg2"], fullsize("cfg2")
    d, mrd = spec.definition, spec.mrd
    total = 0
    for c, (ir, ii) in enumerate(spec.chunk_indices()):
        fancy = c % 4 == 1
        chunk = native.chunk_geometry(spec.level, ir, ii, d, mrd, shortcut=fancy, refill=fancy)
        tiles = [chunk.sub(0, y, d, 64, y * d, d) for y in range(0, d, 64)] if fancy else [chunk]
        counts = _render(emul, native, tiles, d * d)
        rec = doc["chunks"][str(c)]
        assert (rec["ir"], rec["ii"]) == (ir, ii)
        assert oracle.useful_iterations(counts, mrd) == rec["useful"], (ir, ii)
        assert f"{oracle.fnv1a_words(counts):016x}" == rec["fnv"], (ir, ii)
        assert int(np.count_nonzero(counts == 0)) == rec["zeros"], (ir, ii)
        total += rec["useful"]
    assert total == doc["useful_total"]


def test_emulated_deep_zoom_rows_and_the_shortcut_s_accounting(emul, native, oracle):
    """Rows of the native-size cfg4 boundary chunk (4096 wide, mrd 10000: replays, long cycles) against the oracle,
    all variants; the shortcut executes fewer updates than it is credited with, and never more."""
    from distributedmandelbrot_b200.frame import CONFIGS
    spec = CONFIGS["cfg4"]
    d, mrd, ir, ii = spec.definition, 10000, 250, 420
    y0, h = 768, 8                          # rows that cross the set: ~1300 of their 32768 pixels never escape
    want = oracle.render_chunk_counts(spec.level, mrd, ir, ii, d, x0=0, y0=y0, w=d, h=h)
    useful = oracle.useful_iterations(want, mrd)
    for sc, rf in VARIANTS:
        chunk = native.chunk_geometry(spec.level, ir, ii, d, mrd, shortcut=sc, refill=rf)
        ex = []
        got = _render(emul, native, [chunk.sub(0, y0, d, h, 0, d)], d * h, executed=ex).reshape(h, d)
        assert np.array_equal(got, want), (sc, rf)
        if sc:
            assert 0 < ex[0] < useful, (ex, useful)     # in-set pixels stop at their cycle, the others run as before
        u8 = _render(emul, native, [chunk.sub(0, y0, d, h, 0, d)], d * h, u8=True).reshape(h, d)
        assert np.array_equal(u8, oracle.counts_to_u8(want, mrd)), (sc, rf)


def test_emulated_escape_decided_by_the_last_unit_of_the_high_word(emul, native, oracle):
    """tests/golden/threshold_points.json: the first escaping update is a block end (n = 8k) and |z|^2 there lies in
    [4, 4 + 2^-18) -- `hi32(m) >= 0x40100000` holds with equality.  calc_mb_value (tested head of 32, then blocks) and
    1 x 1 tiles through the patch and the refill kernels (blocks from the first update on), plain and with the shortcut."""
    import json
    pts = json.load(open(os.path.join(HERE, "golden", "threshold_points.json")))["points"]
    r = np.array([float.fromhex(p["r"]) for p in pts])
    i = np.array([float.fromhex(p["i"]) for p in pts])
    n = np.array([p["n"] for p in pts], np.int32)
    assert np.array_equal(_calc(emul, native, r, i, 301), n)
    for mrd in (9, 25, 41, 49, 137):                                     # budgets that end right behind some of the escapes
        assert np.array_equal(_calc(emul, native, r, i, mrd), np.where(n < mrd, n, 0)), mrd
    for sc, rf in VARIANTS:
        tiles = []
        for k in range(len(pts)):
            t = native.Tile()
            t.start_r = t.stop_r = r[k]                                  # definition 1: numpy.linspace(start, stop, 1) = [start]
            t.start_i = t.stop_i = i[k]
            t.step_r = t.step_i = 0.0
            t.definition, t.mrd, t.x0, t.y0, t.w, t.h = 1, 301, 0, 0, 1, 1
            t.out_offset, t.out_pitch = k, 1
            t.flags = (native.TILE_PERIOD_SHORTCUT if sc else 0) | (native.TILE_LANE_REFILL if rf else 0)
            tiles.append(t)
        assert np.array_equal(_render(emul, native, tiles, len(pts)), n), (sc, rf)
    # binary32: |z|^2 == 4.0f exactly at a multiple of 16 (blocks of 16, block-end test `!(m < 4)`)
    p32 = json.load(open(os.path.join(HERE, "golden", "threshold_points.json")))["points_f32"]
    r32 = np.array([float.fromhex(p["r"]) for p in p32])
    i32 = np.array([float.fromhex(p["i"]) for p in p32])
    n32 = np.array([p["n"] for p in p32], np.int32)
    assert np.array_equal(_calc(emul, native, r32, i32, 301, dtype="f32"), n32)
    for sc, rf in VARIANTS:
        tiles = []
        for k in range(len(p32)):
            t = native.Tile()
            t.start_r = t.stop_r = r32[k]
            t.start_i = t.stop_i = i32[k]
            t.step_r = t.step_i = 0.0
            t.definition, t.mrd, t.x0, t.y0, t.w, t.h = 1, 301, 0, 0, 1, 1
            t.out_offset, t.out_pitch = k, 1
            t.flags = (native.TILE_PERIOD_SHORTCUT if sc else 0) | (native.TILE_LANE_REFILL if rf else 0)
            tiles.append(t)
        assert np.array_equal(_render(emul, native, tiles, len(p32), dtype="f32"), n32), (sc, rf)


def test_emulated_device_coordinates_equal_gen_arrays(emul, native, golden_meta, golden_arrays):
    """linspace_at (the kernel's in-register replacement of gen_arrays, .py:19-37) for every sample of both axes:
    the reference's own gen_arrays output (golden grids, incl. level 4 000 000) and numpy.linspace, with the forced
    endpoint -- at shallow levels fl(fl((d-1) * step) + start) is often NOT `stop`."""
    emul.emul_tile_axis.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
    for rec in golden_meta["linspace"]:
        t = native.chunk_geometry(rec["level"], rec["ir"], rec["ii"], rec["d"], 100)
        for axis, key in ((0, "_r"), (1, "_i")):
            out = np.empty(rec["d"], np.float64)
            emul.emul_tile_axis(C.byref(t), axis, out.ctypes.data)
            assert np.array_equal(out, golden_arrays[rec["key"] + key]), (rec, axis)
    rng = np.random.default_rng(7)
    for _ in range(200):
        level = int(rng.integers(1, 5_000_000))
        ir = int(rng.integers(0, level))
        d = int(rng.choice([2, 3, 255, 256, 4096]))
        t = native.chunk_geometry(level, ir, 0, d, 1)
        out = np.empty(d, np.float64)
        emul.emul_tile_axis(C.byref(t), 0, out.ctypes.data)
        start = -2 + (4 / level) * ir
        assert np.array_equal(out, np.linspace(start, start + 4 / level, d)), (level, ir, d)
    # shallow levels, where the width of a chunk is comparable to its position: here the endpoint rule is visible
    forced = 0
    for level in range(1, 40):
        for ir in range(level):
            for d in (96, 160, 200, 257, 1000):
                t = native.chunk_geometry(level, ir, 0, d, 1)
                out = np.empty(d, np.float64)
                emul.emul_tile_axis(C.byref(t), 0, out.ctypes.data)
                start = -2 + (4 / level) * ir
                assert np.array_equal(out, np.linspace(start, start + 4 / level, d)), (level, ir, d)
                forced += int((d - 1) * t.step_r + t.start_r != t.stop_r)
    assert forced > 50          # fl(fl((d-1) * step) + start) != stop in that many of these cases


def test_emulated_helper_kernels(emul, native, oracle, golden_meta, golden_arrays):
    """counts_to_u8_kernel (.py:96-98 over a flat array, vector and scalar paths), useful_iterations_kernel,
    classify_u8_kernel (DataChunk.IsNeverChunk / IsImmediateChunk)."""
    emul.emul_counts_to_u8.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.c_uint32]
    emul.emul_useful_iterations.argtypes = [C.c_void_p, C.c_size_t, C.c_uint32]
    emul.emul_useful_iterations.restype = C.c_ulonglong
    emul.emul_classify_u8.argtypes = [C.c_void_p, C.c_size_t, C.c_void_p]
    for rec in golden_meta["u8"]:                                    # every count for 13 values of mrd, reference text
        mrd = rec["mrd"]
        counts = np.arange(mrd, dtype=np.int32)
        out = np.empty(mrd, np.uint8)
        emul.emul_counts_to_u8(counts.ctypes.data, out.ctypes.data, mrd, mrd)
        assert np.array_equal(out, golden_arrays[rec["key"]]), mrd
    counts = np.arange(1002, dtype=np.int32)[1:]                     # unaligned pointers take the scalar path
    out = np.empty(1004, np.uint8)[3:]
    emul.emul_counts_to_u8(counts.ctypes.data, out.ctypes.data, 1001, 1002)
    assert np.array_equal(out, oracle.counts_to_u8(counts, 1002))
    c = oracle.render_chunk_counts(10, 1024, 3, 5, 200).ravel()
    assert emul.emul_useful_iterations(c.ctypes.data, c.size, 1024) == oracle.useful_iterations(c, 1024)
    flags = np.zeros(2, np.uint32)
    for payload, want in [(np.zeros(5000, np.uint8), [1, 0]), (np.ones(5000, np.uint8), [0, 1]),
                          (np.r_[np.zeros(4999, np.uint8), np.uint8(1)], [0, 0]), (np.r_[np.uint8(0), np.ones(4999, np.uint8)], [0, 0])]:
        emul.emul_classify_u8(payload.ctypes.data, payload.size, flags.ctypes.data)
        assert flags.tolist() == want


def test_emulated_kernels_under_random_warp_interleavings(emul, native, oracle, golden_meta, golden_arrays):
    """The emulator's scheduler in random mode: at every barrier / warp collective the next fiber is drawn at random,
    so the warps of a CTA overtake each other in a different order under every seed.  Results must not depend on it:
    shared-memory traffic without its barrier (descriptor staging, the refill kernel's per-warp slots and patch
    counter, the serializer's list / tile copy / look-back hand-over) would show up as a wrong result."""
    import chunk_codec as codec
    emul.emul_set_schedule.argtypes = [C.c_uint64]
    emul.emul_serialize_chunks.argtypes = [C.c_void_p, C.c_size_t, C.c_uint32, C.c_size_t, C.c_void_p, C.c_size_t, C.c_void_p, C.c_int,
                                           C.c_int, C.c_uint]
    rng = np.random.default_rng(21)
    payload = np.ascontiguousarray(np.repeat(rng.integers(0, 4, 30_000).astype(np.uint8), rng.integers(1, 12, 30_000))[:70_001])
    payload[20_000:40_000] = rng.integers(0, 256, 20_000)             # a tile denser than the list: two rounds
    want_blob = codec.serialize(payload)
    level, mrd, ir, ii, d = 10, 700, 3, 5, 96
    want = oracle.render_chunk_counts(level, mrd, ir, ii, d)
    try:
        for seed in (0x9E3779B97F4A7C15, 12345, 777, 2**63 + 5, 31337):
            emul.emul_set_schedule(seed)
            for sc, rf in VARIANTS:
                assert np.array_equal(_chunk(emul, native, level, mrd, ir, ii, d, shortcut=sc, refill=rf), want), (seed, sc, rf)
            assert np.array_equal(_chunk(emul, native, level, mrd, ir, ii, d, u8=True, refill=True), oracle.counts_to_u8(want, mrd)), seed
            n = payload.size
            stride = (n + 1 + 255) // 256 * 256
            out = np.full(stride, 0xEE, np.uint8)
            info = np.zeros(32, np.uint8)
            for spans in (1, 2):
                assert emul.emul_serialize_chunks(payload.ctypes.data, n, 1, n, out.ctypes.data, stride, info.ctypes.data, 1, spans, 65535) == 0
                assert out[:int(info[:8].view("<u8")[0])].tobytes() == want_blob, (seed, spans)
    finally:
        emul.emul_set_schedule(0)


def test_emulated_native_size_chunk_against_the_oracle_digest(emul, native, oracle, fullsize):
    """A whole 4096 x 4096 DataChunk -- what process_workload renders -- of the BASELINE cfg4 frame (f64, mrd 10000) by
    the CPU oracle's own digest (tests/golden/fullsize_cfg4.json): its cheapest chunk, 3.4e8 pixel-iterations over
    131 072 CTAs.  (The boundary chunk (250, 420), 3.1e9 pixel-iterations, matches its digest too -- 31 s with the patch
    kernel, 145 s with refill + shortcut -- and is left out for time.)"""
    from distributedmandelbrot_b200.frame import CONFIGS
    spec, doc, d = CONFIGS["cfg4"], fullsize("cfg4"), 4096
    rec = {(r["ir"], r["ii"]): r for r in doc["chunks"].values()}[(246, 427)]
    counts = _render(emul, native, [native.chunk_geometry(spec.level, 246, 427, d, spec.mrd)], d * d)
    assert oracle.useful_iterations(counts, spec.mrd) == rec["useful"]
    assert f"{oracle.fnv1a_words(counts):016x}" == rec["fnv"]
    assert int(np.count_nonzero(counts == 0)) == rec["zeros"] and int(counts.max()) == rec["max"]
