"""BASELINE frames at their FULL size against the CPU oracle's per-chunk digests.

tests/golden/fullsize_cfg{2,3,4,5}.json hold, for EVERY chunk of the frame, (useful, fnv, zeros,
max) of the int32 counts as the oracle computes them (oracle/gen_fullsize_golden.py: whole native
4096^2 chunks, strict IEEE; no number in there comes from a CUDA build).  The device counts are
copied back chunk by chunk and digested on the host with the oracle's own C helpers, so a single
wrong pixel anywhere in the 65536^2 frame changes an FNV hash and fails the test.
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _spec(name, **kw):
    from distributedmandelbrot_b200.frame import CONFIGS
    return CONFIGS[name].scaled(**kw) if kw else CONFIGS[name]


def _check_frame_against_digests(oracle, fr, spec, doc, label):
    n = spec.definition ** 2
    frame = fr.out.view(spec.n_chunks, n)
    total = 0
    for c in range(spec.n_chunks):
        want = doc["chunks"].get(str(c))
        if want is None:                       # a partially generated fixture checks what it has
            continue
        got = frame[c].cpu().numpy()
        ir, ii = spec.chunk_indices()[c]
        assert (want["ir"], want["ii"]) == (ir, ii)
        assert oracle.useful_iterations(got, spec.mrd) == want["useful"], (label, c, ir, ii)
        assert f"{oracle.fnv1a_words(got):016x}" == want["fnv"], (label, c, ir, ii)
        assert int((got == 0).sum()) == want["zeros"] and int(got.max()) == want["max"], (label, c, ir, ii)
        total += want["useful"]
    if doc.get("complete"):
        assert total == doc["useful_total"]
        assert fr.useful_iterations() == doc["useful_total"]      # the bench's numerator, device-reduced


@pytest.mark.parametrize("name", ["cfg2", "cfg3", "cfg5", "cfg4"])
@pytest.mark.parametrize("shortcut,refill", [(False, False), (True, False), (False, True), (True, True)])
def test_full_frame_equals_oracle_digests(gpu, oracle, fullsize, name, shortcut, refill):
    """Every chunk of the full frame, one launch: the plain kernel, the exact periodicity shortcut, the
    persistent-warp kernel with lane refill, and both together."""
    import torch
    from distributedmandelbrot_b200.frame import FrameRenderer
    doc = fullsize(name)
    spec = _spec(name, shortcut=shortcut, refill=refill)
    fr_doc = doc["frame"]
    assert (spec.level, spec.blocks, spec.definition, spec.ir0, spec.ii0, spec.mrd, spec.dtype) == \
        (fr_doc["level"], fr_doc["blocks"], fr_doc["definition"], fr_doc["ir0"], fr_doc["ii0"], fr_doc["mrd"], fr_doc["dtype"])
    fr = FrameRenderer(spec, device=0)
    fr.render_all()
    torch.cuda.synchronize()
    _check_frame_against_digests(oracle, fr, spec, doc, f"{name} shortcut={shortcut} refill={refill}")
    del fr
    torch.cuda.empty_cache()


def test_cfg4_boundary_chunk_full_compare_with_scalar_oracle(gpu, oracle, fullsize):
    """One whole native cfg4 chunk that straddles the set boundary, compared ELEMENT-WISE with the
    scalar oracle (mbo_escape_f64, the literal restatement of .py:39-68) run on this box's host
    cores -- the deferred-test/replay machinery does most of its work exactly here."""
    doc = fullsize("cfg4")
    spec = _spec("cfg4")
    d, mrd = spec.definition, spec.mrd
    # boundary-heavy but affordable on the CPU: between 8 % and 45 % never-escaping pixels, cheapest first
    cands = sorted((c for c in doc["chunks"].values() if 0.08 * d * d < c["zeros"] < 0.45 * d * d), key=lambda c: c["useful"])
    if not cands:
        pytest.skip("no boundary chunk in the (partial) fixture")
    pick = cands[0]
    from distributedmandelbrot_b200 import worker
    got = worker.process_workload_counts(spec.level, mrd, pick["ir"], pick["ii"], definition=d).reshape(d, d)
    want = oracle.render_chunk_counts(spec.level, mrd, pick["ir"], pick["ii"], d)      # scalar build
    assert np.array_equal(got, want), (pick["ir"], pick["ii"], int((got != want).sum()))
    assert oracle.useful_iterations(want, mrd) == pick["useful"] and f"{oracle.fnv1a_words(want):016x}" == pick["fnv"]


def test_tile_flags_never_change_a_count(gpu, oracle):
    """MB_TILE_PERIOD_SHORTCUT / MB_TILE_LANE_REFILL (and both) never change a count: small chunks across
    both dtypes, mrd values around the save points (powers of two), chunks deep inside the set (every lane
    periodic), on the real axis (y == 0 orbits), at the cusp and at the seahorse zoom."""
    from distributedmandelbrot_b200 import _native as nat
    import torch
    cases = [(1, 0, 0, 200, 3000), (8, 3, 4, 128, 5000), (10, 3, 5, 300, 1024), (16, 5, 8, 256, 4097), (4, 1, 1, 256, 2048),
             (800, 250, 420, 160, 10000), (4000000, 1256353 + 3, 2131821 + 4, 96, 50000), (2, 0, 0, 129, 70), (2, 0, 1, 65, 33),
             (64, 20, 32, 64, 20000), (3, 1, 1, 190, 513)]
    for dtype in ("f64", "f32"):
        for (level, ir, ii, d, mrd) in cases:
            outs = []
            for sc, rf in ((False, False), (True, False), (False, True), (True, True)):
                t = nat.chunk_geometry(level, ir, ii, d, mrd, shortcut=sc, refill=rf)
                out = torch.full((d * d,), -7, dtype=torch.int32, device="cuda")
                nat.check(nat.lib().mb_render_counts(nat.tile_array([t]), 1, out.data_ptr(), nat.DTYPES[dtype],
                                                     torch.cuda.current_stream().cuda_stream))
                outs.append(out.cpu().numpy())
            for k in (1, 2, 3):
                assert np.array_equal(outs[0], outs[k]), (dtype, level, ir, ii, d, mrd, k, int((outs[0] != outs[k]).sum()))
            if mrd <= 5000:
                assert np.array_equal(outs[3].reshape(d, d), oracle.render_chunk_counts(level, mrd, ir, ii, d, dtype=dtype, simd=True))
    # the per-thread switch of the one-call entry points
    from distributedmandelbrot_b200 import worker
    plain = worker.process_workload(8, 3000, 3, 4, definition=160)
    assert nat.set_period_shortcut(True) is False
    try:
        assert np.array_equal(worker.process_workload(8, 3000, 3, 4, definition=160), plain)
        assert np.array_equal(worker.process_workload_counts(10, 2000, 3, 5, definition=128),
                              oracle.render_chunk_counts(10, 2000, 3, 5, 128).ravel())
    finally:
        assert nat.set_period_shortcut(False) is True
    for flags in (2, 3):
        assert nat.set_tile_flags(flags) == 0
        try:
            assert np.array_equal(worker.process_workload(8, 3000, 3, 4, definition=160), plain)
        finally:
            assert nat.set_tile_flags(0) == flags
    # mixed flags inside one launch are refused
    a, b = nat.chunk_geometry(4, 1, 1, 16, 10), nat.chunk_geometry(4, 1, 1, 16, 10, shortcut=True)
    b.out_offset = 256
    out = torch.empty(512, dtype=torch.int32, device="cuda")
    with pytest.raises(ValueError):
        nat.check(nat.lib().mb_render_counts(nat.tile_array([a, b]), 2, out.data_ptr(), 0, None))


@pytest.mark.parametrize("flags", [2, 3])
def test_lane_refill_ragged_tiles_small_mrd_and_the_circle(gpu, oracle, flags):
    """The persistent-warp kernel on everything that is not a clean deep-zoom chunk: ragged multi-tile launches
    with different mrd per tile (int32 and fused uint8, padded pitches, unaligned offsets), every mrd from 0 up
    (tested-only pixels, tails shorter than a trip), and chunks with corners / edges on the |c| = 2 circle."""
    import torch
    nat = gpu

    def render(tiles, n_out, dtype="f64", u8=False):
        out = torch.full((n_out,), 77 if u8 else -7, dtype=torch.uint8 if u8 else torch.int32, device="cuda")
        fn = nat.lib().mb_render_u8 if u8 else nat.lib().mb_render_counts
        nat.check(fn(nat.tile_array(tiles), len(tiles), out.data_ptr(), nat.DTYPES[dtype], torch.cuda.current_stream().cuda_stream))
        torch.cuda.synchronize()
        return out.cpu().numpy()

    for seed in range(5):
        rng = np.random.default_rng(300 + seed)
        tiles, places, off = [], [], int(rng.integers(0, 5))
        for _ in range(int(rng.integers(2, 8))):
            level = int(rng.integers(1, 30))
            ir, ii = int(rng.integers(0, level)), int(rng.integers(0, level))
            d = int(rng.integers(40, 160))
            mrd = int(rng.choice([1, 2, 17, 33, 64, 300, 1500]))
            chunk = nat.chunk_geometry(level, ir, ii, d, mrd)
            chunk.flags = flags
            w, h = int(rng.integers(1, d + 1)), int(rng.integers(1, d + 1))
            x0, y0 = int(rng.integers(0, d - w + 1)), int(rng.integers(0, d - h + 1))
            pitch = w + int(rng.integers(0, 6))
            tiles.append(chunk.sub(x0, y0, w, h, off, pitch))
            places.append((level, ir, ii, d, mrd, x0, y0, w, h, off, pitch))
            off += pitch * h + int(rng.integers(0, 9))
        for u8 in (False, True):
            out = render(tiles, off + 8, u8=u8)
            touched = np.zeros(off + 8, bool)
            for (level, ir, ii, d, mrd, x0, y0, w, h, o, pitch) in places:
                want = oracle.render_chunk_counts(level, mrd, ir, ii, d, x0=x0, y0=y0, w=w, h=h)
                if u8:
                    want = oracle.counts_to_u8(want, mrd)
                idx = (o + np.arange(h)[:, None] * pitch + np.arange(w)[None, :])
                assert np.array_equal(out[idx], want), (seed, u8, level, ir, ii, d, mrd, x0, y0, w, h)
                touched[idx] = True
            assert (out[~touched] == (77 if u8 else -7)).all(), "kernel wrote outside its tiles"
    for dtype in ("f64", "f32"):
        for mrd in list(range(0, 70)) + [95, 96, 97, 127, 128, 129, 1000, 1001]:
            t = nat.chunk_geometry(10, 3, 5, 64, mrd)
            t.flags = flags
            got = render([t], 64 * 64, dtype=dtype).reshape(64, 64)
            assert np.array_equal(got, oracle.render_chunk_counts(10, mrd, 3, 5, 64, dtype=dtype)), (dtype, mrd)
    for (level, ir, ii, d, mrd) in [(1, 0, 0, 129, 300), (2, 0, 1, 65, 300), (5, 4, 4, 101, 200), (5, 0, 4, 201, 200)]:
        t = nat.chunk_geometry(level, ir, ii, d, mrd)
        t.flags = flags
        assert np.array_equal(render([t], d * d).reshape(d, d), oracle.render_chunk_counts(level, mrd, ir, ii, d)), (level, ir, ii)
