#!/usr/bin/env python
"""Per-chunk digests of the BASELINE frames at their FULL size, computed by the CPU oracle.

    python oracle/gen_fullsize_golden.py [cfg3 cfg4 cfg5] [--threads N] [--only K]

TEST INFRASTRUCTURE (see mandel_oracle.h).  Runs on CPU only, needs nothing but oracle/.

Every chunk of cfg2 (256 chunks of 256^2), cfg3 (16 native 4096^2 chunks, binary32), cfg4 (256) and cfg5 (64) is
rendered WHOLE by the data-parallel build of the oracle (mandel_oracle_simd.c; the same lane-wise
IEEE operations as the scalar restatement of .py:39-68, pinned bit-for-bit against it by
tests/test_oracle_golden.py) and reduced to
    useful = sum(n ? n : mrd-1)      the metric's numerator
    fnv    = FNV-1a over the int32 counts in row-major order (oracle.fnv1a_words)
    zeros  = pixels that never escaped
    max    = largest count
and, as a spot check of the SIMD build at full size, `check_rows` rows of every chunk are recomputed
with the scalar oracle and compared.  Output: tests/golden/fullsize_<cfg>.json -- the ONLY source of
the full-size constants in tests/test_gpu_parity.py and bench.py (nothing there is recorded from
the CUDA build any more).  The file is rewritten after every chunk, so an interrupted run resumes.

Chunk order = the product's frame layout: chunk = bi * B + br for indices (ir0 + br, ii0 + bi).
The frame parameters below restate BASELINE.md section 5 / SURVEY.md 8(d);
tests/test_oracle_golden.py::test_fullsize_fixture_matches_the_frame_table checks them against
distributedmandelbrot_b200.frame.CONFIGS.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")
sys.path.insert(0, HERE)
import oracle  # noqa: E402

FRAMES = {
    # name: level, blocks, definition, ir0, ii0, mrd, dtype
    "cfg2": dict(level=16, blocks=16, definition=256, ir0=0, ii0=0, mrd=1024, dtype="f64"),
    "cfg3": dict(level=4, blocks=4, definition=4096, ir0=0, ii0=0, mrd=4096, dtype="f32"),
    "cfg4": dict(level=800, blocks=16, definition=4096, ir0=242, ii0=412, mrd=10000, dtype="f64"),
    "cfg5": dict(level=4000000, blocks=8, definition=4096, ir0=1256353, ii0=2131821, mrd=50000, dtype="f64"),
}
CHECK_ROWS = 2


def path(name: str) -> str:
    return os.path.join(OUT, f"fullsize_{name}.json")


def load(name: str) -> dict:
    if os.path.exists(path(name)):
        with open(path(name)) as f:
            doc = json.load(f)
        if doc.get("frame") == FRAMES[name]:
            return doc
    return {"frame": FRAMES[name], "generator": "oracle/gen_fullsize_golden.py (mandel_oracle_simd.c, strict IEEE, "
            "no contraction)", "check_rows_vs_scalar_oracle": CHECK_ROWS, "chunks": {}}


def save(name: str, doc: dict) -> None:
    ch = doc["chunks"]
    n = FRAMES[name]["blocks"] ** 2
    doc["complete"] = len(ch) == n
    if doc["complete"]:
        doc["useful_total"] = sum(c["useful"] for c in ch.values())
        doc["zeros_total"] = sum(c["zeros"] for c in ch.values())
    tmp = path(name) + ".tmp"
    with open(tmp, "w") as f:
        json.dump(doc, f, indent=0, sort_keys=True)
        f.write("\n")
    os.replace(tmp, path(name))


def digest_chunk(fr: dict, c: int, threads: int) -> dict:
    B, d, mrd, dt = fr["blocks"], fr["definition"], fr["mrd"], fr["dtype"]
    ir, ii = fr["ir0"] + c % B, fr["ii0"] + c // B
    t0 = time.time()
    counts = oracle.render_chunk_counts(fr["level"], mrd, ir, ii, d, dtype=dt, threads=threads, simd=True)
    sec = time.time() - t0
    rng = np.random.default_rng(1000 * c + 7)
    for y in sorted(int(v) for v in rng.choice(d, CHECK_ROWS, replace=False)):
        want = oracle.render_chunk_counts(fr["level"], mrd, ir, ii, d, dtype=dt, threads=threads, x0=0, y0=y, w=d, h=1)
        if not np.array_equal(counts[y], want[0]):
            raise SystemExit(f"SIMD oracle differs from the scalar oracle: chunk {c} ({ir},{ii}) row {y}")
    return {"ir": ir, "ii": ii, "useful": oracle.useful_iterations(counts, mrd), "fnv": f"{oracle.fnv1a_words(counts):016x}",
            "zeros": int((counts == 0).sum()), "max": int(counts.max())}, sec


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("frames", nargs="*", default=["cfg3", "cfg5", "cfg4"])
    ap.add_argument("--threads", type=int, default=0)
    ap.add_argument("--only", type=int, default=None, help="stop after this many new chunks (per frame)")
    args = ap.parse_args()
    os.makedirs(OUT, exist_ok=True)
    for name in args.frames:
        fr = FRAMES[name]
        doc = load(name)
        done = 0
        for c in range(fr["blocks"] ** 2):
            if str(c) in doc["chunks"]:
                continue
            rec, sec = digest_chunk(fr, c, args.threads)       # timings are printed, never stored: the files are reproducible
            doc["chunks"][str(c)] = rec
            save(name, doc)
            done += 1
            print(f"{name} chunk {c:3d} ({rec['ir']},{rec['ii']}): useful {rec['useful']:.6g} zeros {rec['zeros']} "
                  f"{sec:.2f} s ({rec['useful'] / max(sec, 1e-3) / 1e9:.1f} Gpx-it/s)", flush=True)
            if args.only is not None and done >= args.only:
                break
        save(name, doc)
        if doc["complete"]:
            print(f"{name}: complete, useful_total = {doc['useful_total']}", flush=True)


if __name__ == "__main__":
    main()
