#!/usr/bin/env python
"""Generate tests/golden/* by EXECUTING THE REFERENCE'S OWN FUNCTION TEXT in this container.

Run in the authoring container only (needs /root/reference, numba, numpy):
    python oracle/gen_golden.py

The reference worker (DistributedMandelbrotWorkerCUDA/DistributedMandelbrotWorkerCUDA.py)
cannot be imported here: `@vectorize(..., target="cuda")` (.py:39) compiles for CUDA at import
and this container has no CUDA driver.  So this script reads the file, takes the
source text of
    gen_arrays        .py:19-37   (verbatim)
    calc_mb_value     .py:39-68   (one edit: target = "cuda" -> "cpu")
    process_workload  .py:70-100  (one edit per case: `definition = 4096` -> the case's d;
                                   `cuda` is a host shim whose to_device/device_array/copy_to_host
                                   are identity operations on numpy arrays)
and exec()s them under the installed numba/numpy.  Nothing from the reference is copied into
the repo: only the OUTPUTS (counts, bytes, hashes) are committed, with this script.

The fixtures pin the oracle (oracle/mandel_oracle.c) and, through it, the CUDA kernels.
The FP32 cases have no upstream counterpart (the reference is FP64-only, .py:39); they come
from the numpy restatement in oracle/oracle.py and are marked "unpinned upstream".
"""
from __future__ import annotations

import hashlib
import json
import os
import re
import sys

import numpy as np

REF = "/root/reference/DistributedMandelbrotWorkerCUDA/DistributedMandelbrotWorkerCUDA.py"
HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")
sys.path.insert(0, HERE)


def _lines(a: int, b: int) -> str:
    with open(REF, encoding="utf-8-sig") as f:
        src = f.read().splitlines()
    return "\n".join(src[a - 1:b]) + "\n"


class _HostArray(np.ndarray):
    def copy_to_host(self):
        return np.asarray(self)


class _CudaShim:
    """Host stand-in for numba.cuda as used at .py:87-94."""

    @staticmethod
    def to_device(a):
        return np.ascontiguousarray(a).view(_HostArray)

    @staticmethod
    def device_array(shape, dtype):
        return np.empty(shape, dtype=dtype).view(_HostArray)


def load_reference(definition: int):
    """Return (namespace, captured) for the reference text with `definition` patched in."""
    import numba
    from numba import vectorize

    ns = {"np": np, "vectorize": vectorize, "cuda": _CudaShim, "MIN_AXIS": -2, "MAX_AXIS": 2}
    exec(compile(_lines(19, 37), REF + ":19-37", "exec"), ns)
    kern = _lines(39, 68)
    assert 'target = "cuda"' in kern
    exec(compile(kern.replace('target = "cuda"', 'target = "cpu"'), REF + ":39-68", "exec"), ns)
    pw = _lines(70, 100)
    assert "definition = 4096;" in pw
    # capture the int32 counts too: keep a handle on `out` before the uint8 scaling (.py:94 -> :96)
    pw = pw.replace("definition = 4096;", f"definition = {int(definition)};")
    pw = pw.replace("    out = out_device.copy_to_host();\n",
                    "    out = out_device.copy_to_host();\n    _CAPTURE['counts'] = out.copy();\n")
    assert "_CAPTURE" in pw
    cap = {}
    ns["_CAPTURE"] = cap
    exec(compile(pw, REF + ":70-100", "exec"), ns)
    ns["_numba_version"] = numba.__version__
    return ns, cap


def fnv1a_words(counts: np.ndarray) -> int:
    h = 0xCBF29CE484222325
    for n in np.asarray(counts, dtype=np.int64).ravel().tolist():
        h = ((h ^ (n & 0xFFFFFFFF)) * 0x100000001B3) & 0xFFFFFFFFFFFFFFFF
    return h


def useful(counts: np.ndarray, mrd: int) -> int:
    c = counts.astype(np.int64)
    return int(np.where(c != 0, c, max(mrd - 1, 0)).sum())


# (level, mrd, ir, ii, d): small cases saved as full arrays
SMALL = [
    (1, 64, 0, 0, 64),        # whole plane, corners at |c| = 2*sqrt(2)
    (1, 1, 0, 0, 16),         # mrd = 1: loop never runs -> all 0
    (1, 2, 0, 0, 16),         # mrd = 2: exactly one update
    (1, 3, 0, 0, 33),
    (2, 256, 0, 1, 64),       # touches real axis (imag == 0 at chunk edge)
    (2, 256, 1, 1, 65),
    (3, 300, 1, 1, 97),       # non power-of-two level, odd d
    (4, 256, 1, 2, 128),      # cfg1 window at reduced definition
    (4, 256, 0, 0, 32),       # corner chunk: everything escapes immediately
    (8, 1024, 3, 4, 96),      # cusp / main cardioid edge
    (10, 1024, 3, 5, 128),    # launch profile level (launchSettings.json:5)
    (20, 1024, 6, 10, 128),
    (16, 2000, 5, 8, 64),     # all-never chunk (inside the main cardioid)
    (800, 10000, 250, 420, 48),   # cfg4 centre chunk, deep iteration
    (4000000, 5000, 1256353, 2131821, 40),  # cfg5 geometry (huge level / indices)
]

# hash-only (bigger) cases
MEDIUM = [
    (4, 256, 1, 2, 512),      # cfg1 / Appendix C row 1
    (10, 1024, 3, 5, 512),    # Appendix C row 2
    (20, 1024, 6, 10, 256),   # Appendix C row 3
    (3, 300, 1, 1, 257),      # Appendix C row 4
    (16, 1024, 7, 8, 256),    # cfg2 tile
    (4, 256, 1, 2, 4096),     # the reference's native chunk size, launch profile 4:256
]


def main() -> None:
    os.makedirs(OUT, exist_ok=True)
    arrays = {}
    meta = {"generator": "oracle/gen_golden.py", "reference_file": REF,
            "numpy": np.__version__, "small": [], "medium": [], "u8": [], "linspace": [], "f32": []}

    for (level, mrd, ir, ii, d) in SMALL + MEDIUM:
        ns, cap = load_reference(d)
        meta["numba"] = ns["_numba_version"]
        u8 = ns["process_workload"](level, mrd, ir, ii)
        counts = cap["counts"]
        assert counts.dtype == np.int32 and counts.shape == (d * d,)
        assert u8.dtype == np.uint8 and u8.shape == (d * d,)
        rec = {"level": level, "mrd": mrd, "ir": ir, "ii": ii, "d": d,
               "useful": useful(counts, mrd), "fnv": f"{fnv1a_words(counts):016x}",
               "zeros": int((counts == 0).sum()), "max": int(counts.max()),
               "u8_sha256": hashlib.sha256(u8.tobytes()).hexdigest()}
        if (level, mrd, ir, ii, d) in SMALL:
            key = f"L{level}_m{mrd}_r{ir}_i{ii}_d{d}"
            arrays[key + "_counts"] = counts.reshape(d, d)
            arrays[key + "_u8"] = u8.reshape(d, d)
            rec["key"] = key
            meta["small"].append(rec)
        else:
            meta["medium"].append(rec)
        print(rec, flush=True)

    # coordinates: gen_arrays (.py:19-37) verbatim
    ns, _ = load_reference(4096)
    for (level, ir, ii, d) in [(1, 0, 0, 4096), (3, 1, 2, 257), (4, 1, 2, 512), (10, 3, 5, 4096),
                               (20, 6, 10, 4096), (800, 250, 420, 4096), (4000000, 1256353, 2131821, 4096),
                               (7, 6, 0, 2)]:
        chunk_range = (2 - (-2)) / level
        start_r = -2 + (chunk_range * ir)
        start_i = -2 + (chunk_range * ii)
        r_rep, i_rep = ns["gen_arrays"](start_r=start_r, start_i=start_i, _range=chunk_range, definition=d)
        r = np.asarray(r_rep[:d]).copy()
        i = np.asarray(i_rep[::d]).copy()
        assert np.array_equal(r_rep.reshape(d, d), np.broadcast_to(r, (d, d)))
        assert np.array_equal(i_rep.reshape(d, d), np.broadcast_to(i[:, None], (d, d)))
        key = f"lin_L{level}_r{ir}_i{ii}_d{d}"
        arrays[key + "_r"] = r
        arrays[key + "_i"] = i
        meta["linspace"].append({"level": level, "ir": ir, "ii": ii, "d": d, "key": key})

    # uint8 epilogue (.py:96-98) over every possible count, text executed literally
    epi = _lines(96, 98)
    for mrd in [1, 2, 3, 255, 256, 257, 300, 1000, 1024, 4096, 10000, 50000, 65535]:
        env = {"np": np, "out": np.arange(0, mrd, dtype=np.int32), "mrd": mrd}
        exec(compile("\n".join(l.strip() for l in epi.splitlines()), REF + ":96-98", "exec"), env)
        u = env["out"]
        assert u.dtype == np.uint8
        arrays[f"u8map_m{mrd}"] = u
        meta["u8"].append({"mrd": mrd, "key": f"u8map_m{mrd}"})

    # FP32 variant: no upstream counterpart -> numpy restatement (unpinned upstream)
    import oracle as _o
    for (level, mrd, ir, ii, d) in [(4, 256, 1, 2, 64), (4, 4096, 1, 1, 48), (1, 500, 0, 0, 96)]:
        c = _o.np_process_counts(level, mrd, ir, ii, d, dtype=np.float32)
        key = f"f32_L{level}_m{mrd}_r{ir}_i{ii}_d{d}"
        arrays[key + "_counts"] = c
        meta["f32"].append({"level": level, "mrd": mrd, "ir": ir, "ii": ii, "d": d, "key": key,
                            "useful": useful(c, mrd), "pinned": "unpinned upstream (numpy restatement)"})

    # DataChunk serialisation (f2): the restated serializer (oracle/chunk_codec.py, following
    # DataChunkSerializer.cs) is pinned against the reference VIEWER's own decoder text
    # (DistributedMandelbrotViewer.py:35-60, pure Python), executed here.
    import chunk_codec
    vsrc = open("/root/reference/DistributedMandelbrotViewer/DistributedMandelbrotViewer.py", encoding="utf-8-sig").read().splitlines()
    vns = {"unpack": __import__("struct").unpack}
    exec(compile("\n".join(vsrc[34:60]) + "\n", "Viewer.py:35-60", "exec"), vns)
    rng = np.random.default_rng(2024)
    payloads = {
        "codec_all7_n4096": np.full(4096, 7, np.uint8),
        "codec_alternating_n1000": (np.arange(1000) % 2).astype(np.uint8),
        "codec_single_n1": np.array([200], np.uint8),
        "codec_runs_n10007": np.repeat(rng.integers(0, 256, 700).astype(np.uint8), rng.integers(1, 30, 700))[:10007],
        "codec_tie_n20": np.repeat(np.array([1, 2, 3, 4], np.uint8), 5),           # 4 runs * 5 B == 20 B: Raw wins the tie
        "codec_chunk_L10_m1024_r3_i5_d128": arrays["L10_m1024_r3_i5_d128_u8"].ravel(),
        "codec_chunk_L4_m256_r0_i0_d32": arrays["L4_m256_r0_i0_d32_u8"].ravel(),
    }
    meta["codec"] = []
    for key, payload in payloads.items():
        blob = chunk_codec.serialize(payload)
        decoded = bytes(vns["chunk_data_to_value_array"](bytearray(blob)))
        assert decoded == payload.tobytes(), key            # the reference's own decoder accepts it
        arrays[key + "_payload"] = payload
        arrays[key + "_serialized"] = np.frombuffer(blob, dtype=np.uint8)
        meta["codec"].append({"key": key, "n": int(payload.size), "codec": int(blob[0]), "serialized_len": len(blob)})

    np.savez_compressed(os.path.join(OUT, "golden_arrays.npz"), **arrays)
    with open(os.path.join(OUT, "golden_meta.json"), "w") as f:
        json.dump(meta, f, indent=1)
    print("wrote", OUT)


if __name__ == "__main__":
    main()
