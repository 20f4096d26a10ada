"""Host side of the C-ABI without a GPU: the library loads, exports every symbol the header
declares, computes chunk geometry exactly, and fails loudly (no CPU fallback) on compute calls."""
import ctypes as C
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_functions():
    src = open(os.path.join(ROOT, "include", "mandelbrot_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    names = re.findall(r"\b(mb_[a-z0-9_]+)\s*\(", src)
    return sorted(set(names))


def test_library_exports_every_declared_symbol(native):
    lib = C.CDLL(native.library_path())
    declared = _declared_functions()
    assert len(declared) >= 20
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/mandelbrot_b200.h but not exported"
    # and the Python binding declares a signature for each of them
    assert set(declared) == set(native.SIGNATURES), set(declared) ^ set(native.SIGNATURES)


def test_tile_struct_layout(native):
    assert C.sizeof(native.Tile) == 88
    assert native.Tile.definition.offset == 48 and native.Tile.out_offset.offset == 72


def test_chunk_geometry_matches_reference_formulas(native, golden_meta, golden_arrays):
    for rec in golden_meta["linspace"]:
        level, ir, ii, d = rec["level"], rec["ir"], rec["ii"], rec["d"]
        t = native.chunk_geometry(level, ir, ii, d, 100)
        chunk_range = (2 - (-2)) / level                  # .py:75
        assert t.start_r == -2 + (chunk_range * ir)       # .py:77
        assert t.start_i == -2 + (chunk_range * ii)       # .py:78
        assert t.stop_r == t.start_r + chunk_range and t.stop_i == t.start_i + chunk_range
        r = np.array([native.lib().mb_tile_coord(C.byref(t), 0, k) for k in range(d)])
        i = np.array([native.lib().mb_tile_coord(C.byref(t), 1, k) for k in range(d)])
        assert np.array_equal(r, golden_arrays[rec["key"] + "_r"])      # gen_arrays output (.py:19-37)
        assert np.array_equal(i, golden_arrays[rec["key"] + "_i"])
        assert np.array_equal(r, np.linspace(t.start_r, t.start_r + chunk_range, d))
        assert (t.x0, t.y0, t.w, t.h, t.out_offset, t.out_pitch, t.mrd) == (0, 0, d, d, 0, d, 100)


def test_chunk_geometry_random_levels(native):
    rng = np.random.default_rng(7)
    for _ in range(200):
        level = int(rng.integers(1, 5_000_000))
        ir, ii = int(rng.integers(0, level)), int(rng.integers(0, level))
        d = int(rng.choice([2, 3, 255, 256, 4096]))
        t = native.chunk_geometry(level, ir, ii, d, 1)
        chunk_range = 4 / level
        start = -2 + chunk_range * ir
        ref = np.linspace(start, start + chunk_range, d)
        ks = [0, 1, d // 2, d - 2, d - 1]
        got = [native.lib().mb_tile_coord(C.byref(t), 0, k) for k in ks]
        assert got == [ref[k] for k in ks]


def test_definition_one_is_linspace_num_1(native):
    t = native.chunk_geometry(4, 1, 2, 1, 10)
    assert native.lib().mb_tile_coord(C.byref(t), 0, 0) == np.linspace(-1.0, 0.0, 1)[0]


def test_argument_errors_mirror_datachunk_ctor(native):
    # DataChunk.cs:99-106
    for args in [(0, 0, 0), (4, 4, 0), (4, 0, 4), (4, 7, 9)]:
        with pytest.raises(ValueError):
            native.chunk_geometry(*args, 4096, 10)
    with pytest.raises(ValueError):
        native.chunk_geometry(4, 0, 0, 0, 10)             # definition 0
    with pytest.raises(ValueError):
        native.chunk_geometry(4, 0, 0, 16, 2**31)         # mrd does not fit the kernel's int32 (.py:39)
    with pytest.raises(ValueError):
        native.chunk_geometry(-1, 0, 0, 16, 10)


def test_no_cpu_fallback_without_a_device(native):
    """Where there is no CUDA device, compute entry points must fail, not compute on the CPU."""
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA device present")
    from distributedmandelbrot_b200 import worker
    with pytest.raises(native.NativeError):
        native.init(0)
    with pytest.raises(native.NativeError):
        worker.process_workload(4, 256, 1, 2)
    with pytest.raises(native.NativeError):
        worker.calc_mb_value(np.zeros(4), np.zeros(4), 10)
    out = np.zeros(16, np.uint8)
    rc = native.lib().mb_process_workload_u8(4, 16, 0, 0, 4, out.ctypes.data_as(C.c_void_p))
    assert rc in (native.MB_ERR_STATE, native.MB_ERR_CUDA)
    assert not out.any()


def test_product_package_never_touches_the_oracle():
    pkg = os.path.join(ROOT, "distributedmandelbrot_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                txt = open(os.path.join(dirpath, f), encoding="utf-8").read()
                assert "mandel_oracle" not in txt, f
                assert not re.search(r"^\s*(import|from)\s+oracle\b", txt, flags=re.M), f


def test_product_build_never_uses_the_kernel_emulator(native):
    """tests/cuda_emul runs the kernels' text on the CPU for the parity tests only: the product build never defines
    MB_EMULATE, no product source outside the two guarded includes knows the emulator, and the shipped library
    exports none of its entry points (there is no CPU execution path in the product)."""
    import subprocess
    from distributedmandelbrot_b200 import build
    assert not any("MB_EMULATE" in flag for flag in build.NVCC_FLAGS)
    pkg = os.path.join(ROOT, "distributedmandelbrot_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cpp", ".h")):
                txt = open(os.path.join(dirpath, f), encoding="utf-8").read()
                assert "MB_EMULATE" not in txt and "cuda_emul" not in txt, f
            elif f.endswith(".cuh"):
                txt = open(os.path.join(dirpath, f), encoding="utf-8").read()
                # a kernel header may only name the emulator inside `#ifdef MB_EMULATE` ... `#else` / `#endif`
                depth_emul = 0
                for line in txt.splitlines():
                    t = line.strip()
                    if t.startswith("#ifdef MB_EMULATE"):
                        depth_emul += 1
                    elif depth_emul and (t.startswith("#else") or t.startswith("#endif")):
                        depth_emul -= 1
                    elif not t.startswith("//"):
                        assert depth_emul or ("cuda_emul" not in t and "mbc_emul" not in t and "EMUL_" not in t), (f, line)
    syms = subprocess.run(["nm", "-D", "--defined-only", build.SO], capture_output=True, text=True, check=True).stdout
    assert "emul" not in syms


def test_header_is_plain_c_and_links_from_c(native, tmp_path):
    """The boundary is a C ABI: include/mandelbrot_b200.h must compile as C99 and a C program must
    link against the library and call the host-only entry points (no GPU needed)."""
    import shutil
    import subprocess
    cc = shutil.which("gcc")
    if cc is None:
        pytest.skip("no gcc")
    src = tmp_path / "use_abi.c"
    src.write_text(r'''
#include <stdio.h>
#include <string.h>
#include "mandelbrot_b200.h"
int main(void) {
    mb_tile t;
    if (mb_abi_version() != MB_ABI_VERSION) return 1;
    if (sizeof(mb_tile) != 88 || sizeof(mb_chunk_info) != 32) return 2;
    if (mb_chunk_geometry(4, 1, 2, 512, 256, &t) != MB_OK) return 3;
    if (t.start_r != -1.0 || t.start_i != 0.0 || t.stop_r != 0.0 || t.w != 512) return 4;
    if (mb_tile_coord(&t, 0, 511) != t.stop_r || mb_tile_coord(&t, 0, 0) != t.start_r) return 5;
    if (mb_chunk_geometry(0, 0, 0, 512, 256, &t) != MB_ERR_INVALID) return 6;      /* "Level must be positive" */
    if (strstr(mb_last_error(), "Level must be positive") == NULL) return 7;
    mb_wsq *q = NULL; uint32_t first, got, stolen;
    if (mb_wsq_create("/mbq_c_abi_test", 2, 10, &q) != MB_OK) return 8;
    if (mb_wsq_pop(q, 1, 3, &first, &got, &stolen) != 1 || first != 5 || got != 3) return 9;
    if (mb_wsq_remaining(q) != 7) return 10;
    mb_wsq_close(q); mb_wsq_unlink("/mbq_c_abi_test");
    printf("ok\n");
    return 0;
}
''')
    exe = tmp_path / "use_abi"
    so = native.library_path()
    subprocess.check_call([cc, "-std=c99", "-Wall", "-Werror", "-pedantic", "-I", os.path.join(ROOT, "include"), str(src),
                           "-o", str(exe), so, f"-Wl,-rpath,{os.path.dirname(so)}"])
    env = dict(os.environ)
    env["LD_LIBRARY_PATH"] = "/usr/local/cuda/lib64:" + env.get("LD_LIBRARY_PATH", "")
    out = subprocess.run([str(exe)], capture_output=True, text=True, env=env)
    assert out.returncode == 0 and out.stdout.strip() == "ok", (out.returncode, out.stdout, out.stderr)


def test_pinned_pool_recycles_blocks(native):
    """process_workload hands out page-locked payload buffers from a recycling pool; the pool logic
    (no CUDA needed: malloc/free injected) must recycle a block only after every view of it died."""
    import gc
    libc = C.CDLL(None)
    libc.malloc.restype, libc.malloc.argtypes, libc.free.argtypes = C.c_void_p, [C.c_size_t], [C.c_void_p]
    freed = []
    pool = native.PinnedPool(alloc=lambda n: libc.malloc(n), free=lambda a: (freed.append(a), libc.free(a)), keep=2)
    a = pool.take(1000)
    assert a.dtype == np.uint8 and a.shape == (1000,) and a.flags.writeable and a.flags.c_contiguous
    a[:] = 7
    addr_a = a.ctypes.data
    view = a[10:20]
    del a
    gc.collect()
    b = pool.take(1000)                       # `view` still alive -> the block must NOT have been recycled
    assert b.ctypes.data != addr_a and pool.allocated == 2
    assert (view == 7).all()
    del view
    gc.collect()
    c = pool.take(1000)                       # now it is
    assert c.ctypes.data == addr_a and pool.allocated == 2
    blocks = [pool.take(64) for _ in range(5)]
    del blocks
    gc.collect()
    assert len(freed) == 3                    # keep = 2 idle blocks of a size, the rest is freed
    del b, c
    gc.collect()
    assert pool.outstanding == 0
    # a caller that hoards results is cut off from page-locked memory at max_bytes, not from results
    small = native.PinnedPool(alloc=lambda n: libc.malloc(n), free=lambda a: libc.free(a), keep=1, max_bytes=2500)
    hoard = [small.take(1000) for _ in range(4)]
    assert small.allocated == 2 and small.outstanding == 2000           # the 3rd and 4th are pageable numpy arrays
    assert all(h.shape == (1000,) and h.dtype == np.uint8 for h in hoard)
    del hoard
    gc.collect()
    assert small.outstanding == 0


def test_bench_parity_guard_is_oracle_derived():
    """bench.py's frame-sum guard reads the CPU oracle's per-chunk digests; neither bench.py nor the GPU
    tests hold a frame-level constant recorded from a CUDA build."""
    import importlib.util
    import json
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(root, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    for name in ("cfg2", "cfg3", "cfg4", "cfg5"):
        doc = json.load(open(os.path.join(root, "tests", "golden", f"fullsize_{name}.json")))
        assert doc["complete"] and bench.expected_useful(name) == doc["useful_total"] == sum(c["useful"] for c in doc["chunks"].values())
    assert bench.expected_useful("cfg4_4x4") is None
    for path in ("bench.py", os.path.join("tests", "test_gpu_parity.py"), os.path.join("tests", "test_gpu_fullsize.py")):
        text = open(os.path.join(root, path)).read()
        assert not re.search(r"\b\d{1,3}(_\d{3}){3,}\b", text), f"{path} holds a large literal constant"
