"""The CPU oracle against the golden vectors produced by executing the reference's own function
text (oracle/gen_golden.py), and against an independent numpy restatement."""
import hashlib

import numpy as np
import pytest


def test_small_cases_bit_exact(oracle, golden_meta, golden_arrays):
    for rec in golden_meta["small"]:
        u8, counts = oracle.process_workload(rec["level"], rec["mrd"], rec["ir"], rec["ii"], rec["d"], want_counts=True)
        d = rec["d"]
        assert np.array_equal(counts.reshape(d, d), golden_arrays[rec["key"] + "_counts"]), rec
        assert np.array_equal(u8.reshape(d, d), golden_arrays[rec["key"] + "_u8"]), rec
        assert oracle.useful_iterations(counts, rec["mrd"]) == rec["useful"]


def test_medium_cases_hashes(oracle, golden_meta):
    for rec in golden_meta["medium"]:
        if rec["d"] > 1024:
            continue   # the 4096^2 native chunk is checked in test_native_chunk_golden below
        u8, counts = oracle.process_workload(rec["level"], rec["mrd"], rec["ir"], rec["ii"], rec["d"], want_counts=True)
        assert oracle.useful_iterations(counts, rec["mrd"]) == rec["useful"], rec
        assert f"{oracle.fnv1a_words(counts):016x}" == rec["fnv"], rec
        assert int((counts == 0).sum()) == rec["zeros"]
        assert hashlib.sha256(u8.tobytes()).hexdigest() == rec["u8_sha256"], rec


def test_native_chunk_golden(oracle, golden_meta):
    rec = [r for r in golden_meta["medium"] if r["d"] == 4096][0]
    u8, counts = oracle.process_workload(rec["level"], rec["mrd"], rec["ir"], rec["ii"], 4096, want_counts=True)
    assert oracle.useful_iterations(counts, rec["mrd"]) == rec["useful"]
    assert f"{oracle.fnv1a_words(counts):016x}" == rec["fnv"]
    assert hashlib.sha256(u8.tobytes()).hexdigest() == rec["u8_sha256"]


def test_survey_appendix_c_known_answers(oracle):
    """SURVEY.md Appendix C (survey-time strict oracle): sum of pixel-iterations and zero counts."""
    for (level, ir, ii, d, mrd, useful, zeros) in [
        (4, 1, 2, 512, 256, 34_620_920, 129_904),
        (10, 3, 5, 512, 1024, 219_799_424, 212_832),
        (20, 6, 10, 256, 1024, 56_839_714, 54_879),
        (3, 1, 1, 257, 300, 13_531_838, 44_189),
    ]:
        c = oracle.render_chunk_counts(level, mrd, ir, ii, d)
        assert oracle.useful_iterations(c, mrd) == useful
        assert int((c == 0).sum()) == zeros
    # window start (-0.8, 0.05), range 0.1
    c = oracle.render_rect(-0.8, 0.05, 0.1, 512, 1024)
    assert oracle.useful_iterations(c, 1024) == 197_695_090


def test_linspace_matches_numpy_and_golden(oracle, golden_meta, golden_arrays):
    for rec in golden_meta["linspace"]:
        rng, sr, si = oracle.chunk_geometry(rec["level"], rec["ir"], rec["ii"])
        r = oracle.linspace(sr, sr + rng, rec["d"])
        i = oracle.linspace(si, si + rng, rec["d"])
        assert np.array_equal(r, golden_arrays[rec["key"] + "_r"])
        assert np.array_equal(i, golden_arrays[rec["key"] + "_i"])
        assert np.array_equal(r, np.linspace(sr, sr + rng, rec["d"]))
        # Python evaluation of .py:75-78
        chunk_range = (2 - (-2)) / rec["level"]
        assert rng == chunk_range and sr == -2 + chunk_range * rec["ir"] and si == -2 + chunk_range * rec["ii"]


def test_u8_epilogue_all_counts(oracle, golden_meta, golden_arrays):
    for rec in golden_meta["u8"]:
        mrd = rec["mrd"]
        n = np.arange(mrd, dtype=np.int32)
        got = oracle.counts_to_u8(n, mrd)
        assert np.array_equal(got, golden_arrays[rec["key"]]), mrd
        # integer identity used on the device, including the 256 -> 0 wrap
        ident = (((n.astype(np.int64) * 256 + mrd - 1) // mrd) & 0xFF).astype(np.uint8)
        assert np.array_equal(got, ident), mrd
    assert oracle.counts_to_u8(np.array([0, 1021, 1022, 1023], np.int32), 1024).tolist() == [0, 0, 0, 0]


def test_f32_oracle_matches_numpy_restatement(oracle, golden_meta, golden_arrays):
    for rec in golden_meta["f32"]:
        c = oracle.render_chunk_counts(rec["level"], rec["mrd"], rec["ir"], rec["ii"], rec["d"], dtype="f32")
        assert np.array_equal(c, golden_arrays[rec["key"] + "_counts"]), rec


@pytest.mark.parametrize("seed", range(4))
def test_c_oracle_matches_numpy_restatement_random(oracle, seed):
    rng = np.random.default_rng(seed)
    level = int(rng.integers(1, 40))
    ir, ii = int(rng.integers(0, level)), int(rng.integers(0, level))
    d = int(rng.integers(2, 70))
    mrd = int(rng.choice([1, 2, 3, 17, 100, 400]))
    a = oracle.render_chunk_counts(level, mrd, ir, ii, d)
    b = oracle.np_process_counts(level, mrd, ir, ii, d)
    assert np.array_equal(a, b), (level, ir, ii, d, mrd)
    assert np.array_equal(oracle.counts_to_u8(a, max(mrd, 1)), oracle.np_counts_to_u8(a, max(mrd, 1)))


def test_sub_rectangle_indexes_chunk_grid(oracle):
    """A tile of a chunk indexes the chunk-level linspace, never its own."""
    whole = oracle.render_chunk_counts(10, 300, 3, 5, 96)
    part = oracle.render_chunk_counts(10, 300, 3, 5, 96, x0=13, y0=40, w=50, h=33)
    assert np.array_equal(part, whole[40:73, 13:63])


def test_edge_semantics(oracle):
    # z0 = c is never tested: |c| >= 2 still returns >= 1
    assert oracle.escape_f64(2.0, 2.0, 10) == 1
    # c = -2: z1 = 2, |z1|^2 = 4 and the test is `>= 4` (.py:65) -> "escapes" at n = 1
    assert oracle.escape_f64(-2.0, 0.0, 50) == 1
    assert oracle.escape_f64(0.0, 0.0, 50) == 0         # never escapes -> 0 (.py:68)
    assert oracle.escape_f64(3.0, 0.0, 1) == 0          # mrd = 1: range(1, 1) is empty
    assert oracle.escape_f64(3.0, 0.0, 2) == 1
    assert oracle.useful_iterations(np.array([0, 1, 5, 0], np.int32), 10) == 9 + 1 + 5 + 9


def test_frame_driver_equals_per_chunk(oracle):
    """The CPU baseline's whole-frame driver is the same per-pixel path, threaded differently."""
    counts, useful = oracle.render_frame(10, 2, 4, 3, 40, 200)
    for k, (ir, ii) in enumerate([(2 + br, 4 + bi) for bi in range(3) for br in range(3)]):
        assert np.array_equal(counts[k], oracle.render_chunk_counts(10, 200, ir, ii, 40))
    assert useful == oracle.useful_iterations(counts, 200)
    c32, _ = oracle.render_frame(4, 1, 1, 2, 33, 300, dtype="f32")
    assert np.array_equal(c32[3], oracle.render_chunk_counts(4, 300, 2, 2, 33, dtype="f32"))


def test_escaped_stays_escaped_property(oracle):
    """The CUDA kernel tests |z|^2 >= 4 only every 8th update (binary64) and replays the block on a
    hit; that is exact iff a computed |z|^2 >= 4 is followed by 7 more that are >= 4 and finite, for
    every c outside the thin annulus | |c|^2 - 4 | <= 2^-36 (those lanes are never deferred).  The
    kernel header argues it from rounding bounds; this hammers it with the oracle's arithmetic."""
    rng = np.random.default_rng(99)
    n = 400_000
    th = rng.uniform(0, 2 * np.pi, n)
    sets = []
    # (a) hugging the annulus from both sides, down to just outside the band
    rad2 = 4.0 + rng.choice([-1, 1], n) * 2.0 ** rng.uniform(-36, -8, n)
    rad = np.sqrt(rad2)
    sets.append((rad * np.cos(th), rad * np.sin(th)))
    # (b) around -2 on the real axis (the orbit hovers at z = 2) and the tips of the antennae
    sets.append((-2.0 + rng.normal(0, 1e-6, n), rng.normal(0, 1e-6, n)))
    sets.append((-2.0 + 2.0 ** rng.uniform(-50, -5, n) * rng.choice([-1, 1], n), 2.0 ** rng.uniform(-60, -5, n) * rng.choice([-1, 0, 1], n)))
    # (c) the whole plane incl. the corners at |c| = 2.83, and the boundary of the set
    sets.append((rng.uniform(-2, 2, n), rng.uniform(-2, 2, n)))
    t2 = rng.uniform(0, 2 * np.pi, n)
    sets.append((0.5 * np.cos(t2) - 0.25 * np.cos(2 * t2) + rng.normal(0, 3e-3, n), 0.5 * np.sin(t2) - 0.25 * np.sin(2 * t2) + rng.normal(0, 3e-3, n)))
    # (d) actual DataChunk grid points on the circle's neighbourhood (levels with 1.2 / 1.6 grid lines etc.)
    for level in (1, 2, 4, 5, 10, 20):
        g = np.linspace(-2, 2, level * 64 + 1)
        rr, ii = np.meshgrid(g, g)
        sets.append((rr.ravel(), ii.ravel()))
    total = 0
    for r, i in sets:
        total += r.size
        assert oracle.sticky_violations(r, i, 300, extra=7) == 0
        # binary32 kernel: blocks of 16, `!(m < 4)` at the block end, band 2^-9
        assert oracle.sticky_violations_f32(r, i, 300, extra=15) == 0
    assert total > 2_000_000
    # and closer to the binary32 band edge
    rad2 = 4.0 + rng.choice([-1, 1], n) * 2.0 ** rng.uniform(-9, -3, n)
    rad = np.sqrt(rad2)
    assert oracle.sticky_violations_f32(rad * np.cos(th), rad * np.sin(th), 300, extra=15, band=2.0 ** -9) == 0


# ---- the data-parallel build of the oracle and the full-size digests it produced ---------------

def test_simd_oracle_equals_scalar_oracle(oracle):
    """mandel_oracle_simd.c (8/16 pixels per vector, lane-wise IEEE ops) is bit-identical to the
    scalar restatement: ragged widths (padding path), every mrd around small values, both dtypes,
    boundary tiles of all BASELINE windows, sub-rectangles, and one thread vs many."""
    cases = [(1, 0, 0, 131, 300, "f64"), (2, 1, 0, 97, 1000, "f64"), (10, 3, 5, 200, 1024, "f64"), (16, 4, 8, 64, 1024, "f64"),
             (800, 250, 420, 100, 10000, "f64"), (4000000, 1256355, 2131824, 48, 50000, "f64"), (5, 4, 4, 101, 200, "f64"),
             (4, 1, 1, 150, 4096, "f32"), (4, 2, 1, 77, 700, "f32"), (1, 0, 0, 129, 333, "f32"), (7, 6, 0, 2, 5, "f64"),
             (5, 2, 2, 1, 100, "f64"), (3, 0, 2, 17, 1, "f32"), (3, 0, 2, 33, 2, "f64")]
    for (level, ir, ii, d, mrd, dt) in cases:
        a = oracle.render_chunk_counts(level, mrd, ir, ii, d, dtype=dt)
        b = oracle.render_chunk_counts(level, mrd, ir, ii, d, dtype=dt, simd=True)
        assert np.array_equal(a, b), (level, ir, ii, d, mrd, dt)
        c = oracle.render_chunk_counts(level, mrd, ir, ii, d, dtype=dt, simd=True, threads=1)
        assert np.array_equal(a, c)
    for mrd in range(0, 40):
        for dt in ("f64", "f32"):
            assert np.array_equal(oracle.render_chunk_counts(10, mrd, 3, 5, 37, dtype=dt),
                                  oracle.render_chunk_counts(10, mrd, 3, 5, 37, dtype=dt, simd=True)), (mrd, dt)
    a = oracle.render_chunk_counts(10, 500, 3, 5, 300, x0=13, y0=250, w=55, h=9)
    assert np.array_equal(a, oracle.render_chunk_counts(10, 500, 3, 5, 300, x0=13, y0=250, w=55, h=9, simd=True))
    assert oracle.simd_width() in (128, 256, 512)
    _, u0 = oracle.render_frame(16, 0, 0, 4, 32, 200, want_counts=False)
    f1, u1 = oracle.render_frame(16, 0, 0, 4, 32, 200, simd=True)
    assert u0 == u1 == sum(oracle.useful_iterations(f1[c], 200) for c in range(16))


def test_simd_oracle_reproduces_the_reference_text_goldens(oracle, golden_meta, golden_arrays):
    """... and therefore the fixtures generated by executing the reference's own function text."""
    for rec in golden_meta["small"]:
        got = oracle.render_chunk_counts(rec["level"], rec["mrd"], rec["ir"], rec["ii"], rec["d"], simd=True)
        assert np.array_equal(got, golden_arrays[rec["key"] + "_counts"]), rec
    for rec in golden_meta["medium"]:
        if rec["d"] > 1024:
            continue
        got = oracle.render_chunk_counts(rec["level"], rec["mrd"], rec["ir"], rec["ii"], rec["d"], simd=True)
        assert oracle.useful_iterations(got, rec["mrd"]) == rec["useful"] and f"{oracle.fnv1a_words(got):016x}" == rec["fnv"], rec


def test_fullsize_fixture_matches_the_frame_table(oracle):
    """tests/golden/fullsize_*.json: the generator's frame parameters are the product's CONFIGS, the
    chunk order is the product's, and the cheap chunks re-digest to the stored values right here."""
    import json
    import os
    from conftest import GOLDEN
    from distributedmandelbrot_b200.frame import CONFIGS
    seen = 0
    for name in ("cfg2", "cfg3", "cfg4", "cfg5"):
        path = os.path.join(GOLDEN, f"fullsize_{name}.json")
        if not os.path.exists(path):
            continue
        doc = json.load(open(path))
        spec, fr = CONFIGS[name], doc["frame"]
        assert (spec.level, spec.blocks, spec.definition, spec.ir0, spec.ii0, spec.mrd, spec.dtype) == \
            (fr["level"], fr["blocks"], fr["definition"], fr["ir0"], fr["ii0"], fr["mrd"], fr["dtype"])
        idx = spec.chunk_indices()
        for c, rec in doc["chunks"].items():
            assert (rec["ir"], rec["ii"]) == idx[int(c)]
        if doc.get("complete"):
            assert len(doc["chunks"]) == spec.n_chunks
            assert doc["useful_total"] == sum(r["useful"] for r in doc["chunks"].values())
        # re-digest the two cheapest chunks (a few 1e7..1e9 pixel-iterations) with the SCALAR oracle
        cheap = sorted(doc["chunks"].values(), key=lambda r: r["useful"])[:2]
        for rec in cheap:
            if rec["useful"] > 4e9:
                continue
            got = oracle.render_chunk_counts(spec.level, spec.mrd, rec["ir"], rec["ii"], spec.definition, dtype=spec.dtype)
            assert oracle.useful_iterations(got, spec.mrd) == rec["useful"]
            assert f"{oracle.fnv1a_words(got):016x}" == rec["fnv"]
            assert int((got == 0).sum()) == rec["zeros"] and int(got.max()) == rec["max"]
            seen += 1
    assert seen > 0


def test_threshold_points_fixture_is_what_the_oracle_says(oracle):
    """tests/golden/threshold_points.json (oracle/threshold_points.c): coordinates whose first escaping update n is a
    multiple of 8 with 4 <= |z|^2 < 4 + 2^-18 there -- the inputs on which the kernels' integer compare of the high
    word is decided by its last unit, at a block end.  Scalar C oracle and the numpy restatement agree on every one."""
    import json
    import os
    doc = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "threshold_points.json")))
    pts = doc["points"]
    assert len(pts) >= 60 and sum(1 for p in pts if p["n"] >= 40) >= 15
    r = np.array([float.fromhex(p["r"]) for p in pts])
    i = np.array([float.fromhex(p["i"]) for p in pts])
    n = np.array([p["n"] for p in pts], np.int32)
    for k, p in enumerate(pts):
        assert p["n"] % 8 == 0
        assert oracle.escape_f64(r[k], i[k], 301) == p["n"]
        assert oracle.escape_f64(r[k], i[k], p["n"]) == 0                 # one update short of it
        m = float(p["m"])
        assert 4.0 <= m < 4.0 + 2.0 ** -18
    assert np.array_equal(oracle.np_counts(r, i, 301), n)
    # binary32: first escape at a multiple of 16 with |z|^2 == 4.0f exactly (the block-end test there is `!(m < 4)`)
    p32 = doc["points_f32"]
    assert len(p32) >= 8
    for p in p32:
        cr, ci = float.fromhex(p["r"]), float.fromhex(p["i"])
        assert np.float32(cr) == cr and np.float32(ci) == ci and p["n"] % 16 == 0
        assert oracle.lib().mbo_escape_f32(cr, ci, 301) == p["n"] and oracle.lib().mbo_escape_f32(cr, ci, p["n"]) == 0
    r32 = np.array([float.fromhex(p["r"]) for p in p32])
    i32 = np.array([float.fromhex(p["i"]) for p in p32])
    assert np.array_equal(oracle.np_counts(r32, i32, 301, dtype=np.float32), np.array([p["n"] for p in p32], np.int32))
