#!/usr/bin/env python
"""How sensitive are the emulated parity suites?  Applies small hand-written mutations to COPIES of the kernel
headers, builds the CPU emulator (tests/cuda_emul) from each mutated copy and runs tests/test_escape_emulated.py and
tests/test_codec_emulated.py against it.  A mutant is `killed` when a test fails.  Mutants that cannot change a result
by construction are listed as `equivalent` and not run.  Writes a table to stdout (profiles/r02_mutation_check.txt).

    python scripts/mutation_check.py [--only SUBSTRING] [--jobs N]
"""
import argparse
import concurrent.futures
import glob
import os
import shutil
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "distributedmandelbrot_b200", "csrc")
EMUL = os.path.join(ROOT, "tests", "cuda_emul")

# (name, file, old, new, which test file kills it fastest)
MUTANTS = [
    # ---- serializer (chunk_codec.cuh) ----
    ("codec: no barrier between list append / look-back and emission", "chunk_codec.cuh",
     "        if (!emits) return;                                            // uniform: nobody is left behind at the barrier\n        __syncthreads();",
     "        if (!emits) return;", "codec"),
    ("codec: look-back window uses lanes < stop (drops the inclusive prefix itself)", "chunk_codec.cuh",
     "                if (lane <= stop) {\n                    c += cv[k];", "                if (lane < stop) {\n                    c += cv[k];", "codec"),
    ("codec: look-back stops after the first 32 predecessors even without an inclusive prefix", "chunk_codec.cuh",
     "                found = incl != 0;", "                found = true;", "codec"),
    ("codec: flag-gathering multiplier off by one bit", "chunk_codec.cuh", "h * 0x00204081u", "h * 0x00204080u", "codec"),
    ("codec: first byte of a later span not compared with its predecessor (prev forced to the byte itself)", "chunk_codec.cuh",
     "__byte_perm(prev, v0.x, 0x6540)", "__byte_perm(v0.x, v0.x, 0x6540)", "codec"),
    ("codec: word phase selector wrong", "chunk_codec.cuh", "0x3210u + 0x1111u * phase", "0x3210u + 0x1111u * (phase & 3u)", "codec"),
    ("codec: value byte of the third record taken from the second", "chunk_codec.cuh",
     "const uint32_t c0 = cp[p0], c1 = cp[p1], c2 = cp[p2], c3 = cp[p3];", "const uint32_t c0 = cp[p0], c1 = cp[p1], c2 = cp[p1], c3 = cp[p3];", "codec"),
    ("codec: start 0 of the chunk treated as closing a run", "chunk_codec.cuh",
     "const uint32_t skip = runs_r ? 0u : 1u;", "const uint32_t skip = 0u;", "codec"),
    ("codec: end-of-chunk sentinel one byte early", "chunk_codec.cuh",
     "s_pos[tot_r] = (uint16_t)(n - tile_base);", "s_pos[tot_r] = (uint16_t)(n - tile_base - 1u);", "codec"),
    ("codec: output not cut off at n + 1 bytes", "chunk_codec.cuh",
     "            slen = (uint32_t)((unsigned long long)n + 1ull - out_lo);  // bytes beyond n + 1 can never be part of an RLE result",
     "            slen = slen;", "codec"),
    ("codec: second round forgets the entries of the first", "chunk_codec.cuh", "            done += tot_r;", "            done += 0;", "codec"),
    ("codec: second round's `before` not handed over", "chunk_codec.cuh",
     "            if (tid == 0) s_last_before = tot_r ? (uint32_t)s_pos[tot_r - 1u] : before;", "            if (tid == 0) s_last_before = before;", "codec"),
    ("codec: odd top entry not peeled (pairs misaligned)", "chunk_codec.cuh", "                    if (e & 1u) {                                      // my top entry",
     "                    if (!(e & 1u)) {                                   // my top entry", "codec"),
    ("codec: pair store halves swapped", "chunk_codec.cuh", "*--qw = off2 + ((j1 << 16) | j0);", "*--qw = off2 + ((j0 << 16) | j1);", "codec"),
    ("codec: packed scan drops the second piece's rank base", "chunk_codec.cuh",
     "        rank[h] = tot + ((excl >> (16 * h)) & 0xffffu);", "        rank[h] = ((excl >> (16 * h)) & 0xffffu);", "codec"),
    ("codec: last start of a tile taken from the first piece only", "chunk_codec.cuh",
     "        if (mask[h]) mylast = (uint32_t)(h * kTile) + span + 32u - (uint32_t)__clz(mask[h]);",
     "        if (mask[h] && h == 0) mylast = (uint32_t)(h * kTile) + span + 32u - (uint32_t)__clz(mask[h]);", "codec"),
    ("codec: Raw loses ties (RLE chosen when 5 * runs == n)", "chunk_codec.cuh",
     "const bool rle = (unsigned long long)runs * 5ull < (unsigned long long)n;", "const bool rle = (unsigned long long)runs * 5ull <= (unsigned long long)n;", "codec"),
    ("codec: Immediate classified as Regular", "chunk_codec.cuh", "(runs == 1 && first == 1) ? 2u : 0u", "(runs == 1 && first == 1) ? 0u : 0u", "codec"),
    ("codec: raw copy's first word keeps a stale header byte", "chunk_codec.cuh",
     "const uint32_t before = base ? (uint32_t)p[base - 1] : 0u;         // header byte", "const uint32_t before = base ? (uint32_t)p[base - 1] : 1u;         // header byte", "codec"),
    ("codec: idle tiles never relay (kRelay larger than any chunk)", "chunk_codec.cuh", "constexpr uint32_t kRelay = 16;", "constexpr uint32_t kRelay = 1u << 30;", "equivalent"),
    # ---- escape kernels ----
    ("escape: replay of block B starts from B's start instead of A's end state swap", "escape_kernels.cuh",
     "if (Defer<Real>::tripped(O::add(bxx, byy))) { x = ax; y = ay; tripped = true; break; }",
     "if (Defer<Real>::tripped(O::add(bxx, byy))) { tripped = true; break; }", "escape"),
    ("escape: block-end test misses m == 4.0 exactly (f64)", "escape_kernels.cuh",
     "static __device__ __forceinline__ bool tripped(double m) { return __double2hiint(m) >= 0x40100000; }",
     "static __device__ __forceinline__ bool tripped(double m) { return __double2hiint(m) > 0x40100000; }", "escape"),
    ("escape: per-update test misses m == 4.0 exactly (f64)", "escape_kernels.cuh",
     "static __device__ __forceinline__ bool escaped(double m) { return __double2hiint(m) >= 0x40100000; }",
     "static __device__ __forceinline__ bool escaped(double m) { return __double2hiint(m) > 0x40100000; }", "escape"),
    ("escape: |c| ~ 2 annulus not treated as risky", "escape_kernels.cuh",
     "const bool risky = (q <= Defer<Real>::kBand) && (q >= -Defer<Real>::kBand);", "const bool risky = false;", "escape"),
    ("escape: y' uses a contracted product (fma(x, y, y) shape): t*2+ci -> t*2.0000000000000004+ci", "escape_kernels.cuh",
     "static __device__ __forceinline__ double twice_plus(double t, double c) { return __fma_rn(t, 2.0, c); }",
     "static __device__ __forceinline__ double twice_plus(double t, double c) { return __fma_rn(t, 2.0000000000000004, c); }", "escape"),
    ("escape: x' = a + cr evaluated as a single fma of x*x (contraction of xx - yy)", "escape_kernels.cuh",
     "    const Real a = O::sub(xx, yy);\n    y = O::twice_plus(t, ci);", "    const Real a = (Real)fma((double)x, (double)x, -(double)yy);\n    y = O::twice_plus(t, ci);", "escape"),
    ("escape: tail takes one update too many", "escape_kernels.cuh",
     "                for (; left > 0; --left) {\n                    update<Real>(x, y, xx, yy, cr, ci);\n                    ++n;\n                    if (O::escaped(O::add(xx, yy))) { result = n; break; }",
     "                for (; left >= 0; --left) {\n                    update<Real>(x, y, xx, yy, cr, ci);\n                    ++n;\n                    if (O::escaped(O::add(xx, yy))) { result = n; break; }", "escape"),
    ("escape: replay reports the block start instead of the exact update", "escape_kernels.cuh",
     "if (O::escaped(O::add(xx, yy))) { result = n + j; break; }", "if (O::escaped(O::add(xx, yy))) { result = n + 1; break; }", "escape"),
    ("escape: uint8 scaling rounds with + mrd instead of + mrd - 1", "escape_kernels.cuh",
     "if (mrd <= (1u << 23)) return ((n * 256u + mrd - 1u) / mrd) & 0xffu;", "if (mrd <= (1u << 23)) return ((n * 256u + mrd) / mrd) & 0xffu;", "escape"),
    ("escape: linspace endpoint not forced to `stop`", "escape_kernels.cuh",
     "return (k == d_minus_1) ? stop : __dadd_rn(__dmul_rn((double)k, step), start);", "return __dadd_rn(__dmul_rn((double)k, step), start);", "escape"),
    ("escape: periodicity shortcut compares x only", "escape_kernels.cuh",
     "if (O::same_bits(x, sx) && O::same_bits(y, sy)) { periodic = true; break; }", "if (O::same_bits(x, sx)) { periodic = true; break; }", "escape"),
    ("escape: shortcut samples the state at every trip (still exact, different work)", "escape_kernels.cuh",
     "if (save_point(n)) { sx = x; sy = y; }", "{ sx = x; sy = y; }", "equivalent"),
    ("escape: tiny-|c| inputs take the fused 2xy path", "escape_kernels.cuh",
     "return fabs(cr) <= 4.0 && fabs(ci) <= 4.0 && !tiny(cr) && !tiny(ci); }", "return fabs(cr) <= 4.0 && fabs(ci) <= 4.0; }", "escape"),
    ("escape f32: per-update test misses m == 4.0f exactly", "escape_kernels.cuh",
     "static __device__ __forceinline__ bool escaped(float m) { return __float_as_int(m) >= 0x40800000; }",
     "static __device__ __forceinline__ bool escaped(float m) { return __float_as_int(m) > 0x40800000; }", "escape"),
    ("escape f32: block-end test misses m == 4.0f exactly", "escape_kernels.cuh",
     "static __device__ __forceinline__ bool tripped(float m) { return !(m < 4.0f); }",
     "static __device__ __forceinline__ bool tripped(float m) { return !(m <= 4.0f); }", "escape"),
    ("escape f32: block-end test written as m >= 4 (an overflow to NaN would read as not escaped)", "escape_kernels.cuh",
     "static __device__ __forceinline__ bool tripped(float m) { return !(m < 4.0f); }",
     "static __device__ __forceinline__ bool tripped(float m) { return m >= 4.0f; }", "escape"),
    # ---- lane-refill kernel ----
    ("refill: replay starts from the wrong block", "escape_refill.cuh",
     "Real rx = tA ? sxx_ : mx_, ry = tA ? sy_ : my_;", "Real rx = tA ? mx_ : sxx_, ry = tA ? my_ : sy_;", "escape"),
    ("refill: replayed escape index misses the first block's updates", "escape_refill.cuh",
     "const int base = tA ? n : n + K;", "const int base = n;", "escape"),
    ("refill: lanes near their update budget are not sent to the tested tail in time", "escape_refill.cuh",
     "(lim - n < 4 * K));", "(lim - n < 2 * K));", "escape"),
    ("refill: tested tail one update short", "escape_refill.cuh",
     "                    for (int j = 1; j <= left; ++j) {", "                    for (int j = 1; j < left; ++j) {", "escape"),
    ("refill: pixel handed to the wrong lane slot", "escape_refill.cuh", "const uint32_t q = pos + rank;", "const uint32_t q = pos + (rank ^ 1u);", "escape"),
    ("refill: 128-bit row store repeats the third count", "escape_refill.cuh",
     "*reinterpret_cast<int4 *>(dst) = make_int4(result, v1, v2, v3);", "*reinterpret_cast<int4 *>(dst) = make_int4(result, v1, v2, v2);", "escape"),
    ("refill: descriptor slot read without the warp barriers around its refresh", "escape_refill.cuh",
     "                __syncwarp();\n                if (lane < sizeof(mb_tile) / 8) {", "                if (lane < sizeof(mb_tile) / 8) {", "escape"),
    ("refill: parked (periodic) lanes store a non-zero result", "escape_refill.cuh",
     "        bool fin = pend;                                             // parked since an earlier trip: result 0\n        pend = false;\n        int result = 0;",
     "        bool fin = pend;\n        int result = pend ? 1 : 0;\n        pend = false;", "escape"),
    ("refill: short budgets (lim < 2K) enter deferred blocks", "escape_refill.cuh",
     "if (lim < 2 * K || risky || !O::ordinary(cr, ci)) {", "if (risky || !O::ordinary(cr, ci)) {", "escape"),
    ("escape: fused uint8 row store drops the high half", "escape_kernels.cuh",
     "if ((lane & 7u) == 0) *reinterpret_cast<uint2 *>(gp) = make_uint2(v, hi);", "if ((lane & 7u) == 0) *reinterpret_cast<uint2 *>(gp) = make_uint2(v, v);", "escape"),
]


def run_one(m, idx):
    name, fname, old, new, target = m
    if target == "equivalent":
        return idx, name, "equivalent", 0.0, ""
    d = tempfile.mkdtemp(prefix="mbmut_")
    try:
        os.makedirs(os.path.join(d, "csrc"))
        os.makedirs(os.path.join(d, "include"))
        for f in glob.glob(os.path.join(CSRC, "*.cuh")):
            shutil.copy(f, os.path.join(d, "csrc"))
        shutil.copy(os.path.join(ROOT, "include", "mandelbrot_b200.h"), os.path.join(d, "include"))
        p = os.path.join(d, "csrc", "escape_kernels.cuh")
        text = open(p).read().replace('"../../include/mandelbrot_b200.h"', '"../include/mandelbrot_b200.h"')
        open(p, "w").write(text)
        p = os.path.join(d, "csrc", fname)
        src = open(p).read()
        if src.count(old) != 1:
            return idx, name, f"NOT APPLIED ({src.count(old)} matches)", 0.0, ""
        open(p, "w").write(src.replace(old, new))
        so = os.path.join(d, "libmut.so")
        env = {k: v for k, v in os.environ.items() if k not in ("CC", "CXX")}
        cc = subprocess.run(["g++", "-O1", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off", "-fsanitize=alignment",
                             "-fno-sanitize-recover=alignment", "-w", "-I", EMUL, "-I", os.path.join(d, "csrc"),
                             "-o", so] + [os.path.join(EMUL, f) for f in ("cuda_emul.cpp", "codec_emul.cpp", "escape_emul.cpp")],
                            env=env, capture_output=True, text=True)
        if cc.returncode != 0:
            return idx, name, "DOES NOT COMPILE", 0.0, cc.stderr[-300:]
        env["MB_EMUL_SO"] = so
        first = "tests/test_codec_emulated.py" if target == "codec" else "tests/test_escape_emulated.py"
        other = "tests/test_escape_emulated.py" if target == "codec" else "tests/test_codec_emulated.py"
        t0 = time.time()
        try:
            r = subprocess.run([sys.executable, "-m", "pytest", first, other, "-x", "-q", "-p", "no:cacheprovider"], cwd=ROOT, env=env,
                               capture_output=True, text=True, timeout=900)
            killed = r.returncode != 0
            tail = [l for l in r.stdout.splitlines() if l.startswith("FAILED") or "passed" in l or "failed" in l][-2:]
        except subprocess.TimeoutExpired:
            killed, tail = True, ["timeout (hang)"]
        return idx, name, "killed" if killed else "SURVIVED", time.time() - t0, " | ".join(tail)[:160]
    finally:
        shutil.rmtree(d, ignore_errors=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default=None)
    ap.add_argument("--jobs", type=int, default=3)
    args = ap.parse_args()
    todo = [(m, i) for i, m in enumerate(MUTANTS) if not args.only or args.only in m[0]]
    res = []
    with concurrent.futures.ThreadPoolExecutor(args.jobs) as ex:
        for r in ex.map(lambda a: run_one(*a), todo):
            res.append(r)
            print(f"{r[0]:2d}  {r[2]:10s} {r[3]:6.0f} s  {r[1]}   {r[4]}", flush=True)
    n_run = sum(1 for r in res if r[2] in ("killed", "SURVIVED"))
    print(f"\n{sum(1 for r in res if r[2] == 'killed')} of {n_run} mutants killed, {sum(1 for r in res if r[2] == 'equivalent')} equivalent (not run)")


if __name__ == "__main__":
    main()
