#!/usr/bin/env python
"""Development-only: build a copy of the library with -DMB_TUNING_VARIANTS and time launch-shape /
HEAD variants of the f64 kernel on interior, boundary and low-iteration workloads."""
import ctypes as C
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from distributedmandelbrot_b200 import build as B

so = os.path.join(ROOT, "gpurun_out", "libmb_tuning.so")
os.makedirs(os.path.dirname(so), exist_ok=True)
if not os.path.exists(so):
    cmd = [B.nvcc()] + B.NVCC_FLAGS + ["-DMB_TUNING_VARIANTS", "-o", so] + [os.path.join(B.CSRC, s) for s in B.SOURCES] + ["-lrt"]
    env = dict(os.environ); env.pop("CC", None); env.pop("CXX", None)
    subprocess.check_call(cmd, env=env)
if "--build-only" in sys.argv:
    sys.exit(0)
B.SO = so                                    # make _native load the tuning build
import torch
from distributedmandelbrot_b200 import _native as nat
from distributedmandelbrot_b200.frame import CONFIGS, FrameRenderer

nat.init(0)
setv = nat.lib().mb_internal_set_variant
setv.argtypes = [C.c_int]
NAMES = {0: "H32 2x2 mc12 (default)", 1: "H16 2x2", 2: "H64 2x2", 3: "H32 1x2 (64thr) mc24", 4: "H32 2x4 (256thr) mc6",
         5: "H32 4x1 (32x4 px)", 6: "H32 1x1 (32thr) mc32", 7: "H32 2x2 mc8", 8: "H0 2x2", 9: "H32 4x2 (256thr)"}
WORK = {"interior": CONFIGS["cfg4"].scaled(name="i", blocks=1),
        "boundary4x4": CONFIGS["cfg4"].scaled(name="b", blocks=4, ir0=248, ii0=418),
        "cfg2": CONFIGS["cfg2"],
        "cfg5_2x2": CONFIGS["cfg5"].scaled(name="s", blocks=2, ir0=1256356, ii0=2131824)}
res = {}
for wname, spec in WORK.items():
    fr = FrameRenderer(spec, device=0)
    setv(0)
    fr.render_all(); torch.cuda.synchronize()
    ref = fr.out.clone(); useful = fr.useful_iterations()
    for v, vn in NAMES.items():
        setv(v)
        ts = []
        for _ in range(4):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fr.render_all(); e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        assert torch.equal(fr.out, ref), (wname, vn)
        res[f"{wname} | {vn}"] = useful / min(ts) / 1e6
        print(f"{wname:12s} {vn:26s} {min(ts):10.3f} ms  {useful / min(ts) / 1e6:9.1f} Gpx-it/s", flush=True)
json.dump(res, open(os.path.join(ROOT, "gpurun_out", "tune_variants.json"), "w"), indent=1)
