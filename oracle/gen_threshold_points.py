#!/usr/bin/env python
"""Regenerates tests/golden/threshold_points.json: builds oracle/threshold_points.c and runs its searches
(TEST INFRASTRUCTURE; ~15 min on 8 cores).  Every point found is re-checked against the scalar oracle before it is
written; tests/test_oracle_golden.py::test_threshold_points_fixture_is_what_the_oracle_says repeats that check on the
committed file, so the fixture does not depend on which points a particular search happened to find (the searches
are seeded per OpenMP thread: OMP_NUM_THREADS=8 reproduces the committed binary64 list; the committed binary32 list
is the first nine hits of the f32 search)."""
import json
import os
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import oracle  # noqa: E402

SEARCHES = [  # r0 r1 i0 i1 points-per-thread mrd type
    ("-0.79", "-0.71", "0.06", "0.14", "150000000", "300", "f64"),      # the cfg4 window
    ("-2.0", "0.5", "-1.25", "1.25", "150000000", "300", "f64"),        # the whole set
    ("-0.79", "-0.71", "0.06", "0.14", "200000000", "300", "f32"),
]


def main():
    import argparse
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(os.path.dirname(HERE), "tests", "golden", "threshold_points.json"))
    ap.add_argument("--scale", type=float, default=1.0, help="fraction of the points per thread (a quick try: 0.05)")
    opts = ap.parse_args()
    exe = os.path.join(tempfile.mkdtemp(prefix="thr_"), "threshold_points")
    env = {k: v for k, v in os.environ.items() if k not in ("CC", "CXX")}
    subprocess.check_call(["gcc", "-O2", "-ffp-contract=off", "-fopenmp", "-o", exe, os.path.join(HERE, "threshold_points.c")], env=env)
    env.setdefault("OMP_NUM_THREADS", "8")
    p64, p32, seen = [], [], set()
    for args in SEARCHES:
        args = (*args[:4], str(max(1, int(int(args[4]) * opts.scale))), *args[5:])
        out = subprocess.run([exe, *args], env=env, capture_output=True, text=True, check=True).stdout
        for line in out.splitlines():
            f = line.split()
            if len(f) < 3 or (f[0], f[1], args[-1]) in seen:
                continue
            seen.add((f[0], f[1], args[-1]))
            cr, ci, n = float.fromhex(f[0]), float.fromhex(f[1]), int(f[2])
            if args[-1] == "f64":
                assert oracle.escape_f64(cr, ci, 301) == n and oracle.escape_f64(cr, ci, n) == 0
                p64.append({"r": f[0], "i": f[1], "n": n, "m": repr(float(f[3]))})
            else:
                assert oracle.lib().mbo_escape_f32(cr, ci, 301) == n and oracle.lib().mbo_escape_f32(cr, ci, n) == 0
                p32.append({"r": f[0], "i": f[1], "n": n})
    p64.sort(key=lambda d: (d["n"], d["r"]))
    p32.sort(key=lambda d: (d["n"], d["r"]))
    doc = {"generator": "oracle/threshold_points.c: strict arithmetic of .py:44-68; binary64 points: first escaping update n = 8k with "
                        "hi32(|z|^2) == 0x40100000; binary32 points (coordinates are binary32 values): n = 16k with |z|^2 == 4.0f exactly",
           "points": p64, "points_f32": p32}
    path = opts.out
    json.dump(doc, open(path, "w"), indent=0)
    print(f"{len(p64)} binary64 and {len(p32)} binary32 points -> {path}")


if __name__ == "__main__":
    main()
