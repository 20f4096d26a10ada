#!/usr/bin/env python
"""Development-only: time the refill kernel (shortcut on/off) at 8 vs 10 resident CTAs per SM with a
-DMB_TUNING_VARIANTS build (MB_REFILL_CTAS is read once per process, hence the subprocesses)."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from distributedmandelbrot_b200 import build as B

so = os.path.join(ROOT, "gpurun_out", "libmb_tuning.so")
if "--child" not in sys.argv:
    os.makedirs(os.path.dirname(so), exist_ok=True)
    cmd = [B.nvcc()] + B.NVCC_FLAGS + ["-DMB_TUNING_VARIANTS", "-o", so] + [os.path.join(B.CSRC, s) for s in B.SOURCES] + ["-lrt"]
    env = dict(os.environ); env.pop("CC", None); env.pop("CXX", None)
    if not os.path.exists(so) or "--rebuild" in sys.argv:
        subprocess.check_call(cmd, env=env)
    if "--build-only" in sys.argv:
        sys.exit(0)
    for ctas in ("8", "10"):
        env = dict(os.environ, MB_REFILL_CTAS=ctas)
        subprocess.check_call([sys.executable, os.path.abspath(__file__), "--child"], env=env)
    sys.exit(0)

B.SO = so
B.stale = lambda: False
import torch
from distributedmandelbrot_b200 import _native as nat
from distributedmandelbrot_b200.frame import CONFIGS, FrameRenderer

nat.init(0)
WORK = {"interior": CONFIGS["cfg4"].scaled(name="i", blocks=1),
        "boundary4x4": CONFIGS["cfg4"].scaled(name="b", blocks=4, ir0=248, ii0=418),
        "cfg4_8x8": CONFIGS["cfg4"].scaled(name="q", blocks=8, ir0=246, ii0=416)}
out = {"MB_REFILL_CTAS": os.environ.get("MB_REFILL_CTAS")}
for wname, spec in WORK.items():
    for sc in (False, True):
        fr = FrameRenderer(spec.scaled(refill=True, shortcut=sc), device=0)
        fr.render_all(); torch.cuda.synchronize()
        useful = fr.useful_iterations()
        ts = []
        for _ in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fr.render_all(); e1.record(); torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        out[f"{wname} shortcut={sc}"] = round(useful / (min(ts) * 1e-3) / 1e9, 1)
        del fr
print(json.dumps(out), flush=True)
