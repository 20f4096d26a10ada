#!/usr/bin/env python
"""Drop-in worker behind a (fake) Distributer on loopback TCP: chunks/s for the reference's recorded
launch profile levels (launchSettings.json:5: 4:256, 10:1024), sequential (do_workload_single, as the
reference worker loops) vs pipelined (run_pipelined).  The server side is tests/fake_distributer.py,
a single-threaded Python restatement of Distributer.cs -- like the real one, it is the bottleneck."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from fake_distributer import FakeDistributer

from distributedmandelbrot_b200 import _native as nat
from distributedmandelbrot_b200 import worker

nat.init(0)
res = []
for (level, mrd, count) in [(4, 256, 16), (10, 1024, 40)]:
    wl = [(level, mrd, ir, ii) for ir in range(level) for ii in range(level)][:count]
    worker.process_workload(*wl[0])                       # warm up
    t0 = time.perf_counter()
    for w in wl:
        worker.process_workload(*w)
    compute_only = time.perf_counter() - t0
    for mode in ("sequential", "pipelined"):
        fd = FakeDistributer(wl)
        t0 = time.perf_counter()
        if mode == "sequential":
            n = 0
            while worker.do_workload_single("127.0.0.1", fd.port, verbose=False):
                n += 1
        else:
            n = worker.run_pipelined("127.0.0.1", fd.port, depth=4, verbose=False)
        dt = time.perf_counter() - t0
        fd.close()
        assert n == len(wl) and len(fd.completed) == len(wl) and fd.error is None
        res.append({"level": level, "mrd": mrd, "chunks": n, "mode": mode, "seconds": dt, "chunks_per_s": n / dt,
                    "payload_MB_s": n * 16.777216 / dt, "process_workload_only_chunks_per_s": len(wl) / compute_only})
        print(res[-1], flush=True)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(res, open(os.path.join(ROOT, "gpurun_out", "worker_loop.json"), "w"), indent=1)
