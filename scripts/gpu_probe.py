#!/usr/bin/env python
"""First-contact probe on a B200: FMA-pipe rates and escape-kernel throughput on a few workloads.
Writes gpurun_out/probe.json.  Development tool, not the bench."""
import json
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch

from distributedmandelbrot_b200 import _native as nat
from distributedmandelbrot_b200.frame import CONFIGS, FrameRenderer


def smi():
    try:
        q = "clocks.sm,clocks.max.sm,power.draw,temperature.gpu,clocks_event_reasons.active"
        return subprocess.check_output(["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader", "-i", "0"], text=True).strip()
    except Exception as e:  # noqa: BLE001
        return str(e)


def time_frame(spec, reps=3):
    fr = FrameRenderer(spec, device=0)
    fr.render_all()
    torch.cuda.synchronize()
    useful = fr.useful_iterations()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fr.render_all()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    best = min(ts)
    return {"name": spec.name, "pixels": spec.pixels, "useful": useful, "ms": ts,
            "gpixit_s": useful / (best * 1e-3) / 1e9, "mean_iter_px": useful / spec.pixels, "smi": smi()}


def main():
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    nat.init(0)
    res = {"device": torch.cuda.get_device_name(0), "sms": nat.lib().mb_sm_count(), "smi_idle": smi(), "pipe": {}, "frames": []}
    for dtype in ("f64", "f32"):
        for mode in (0, 1):
            for iters in (2048, 16384):
                rate, ms = nat.measure_pipe_rate(dtype, mode, iters)
                res["pipe"][f"{dtype}_mode{mode}_it{iters}"] = {"tera_ops_s": rate, "ms": ms, "smi": smi()}
                print(dtype, mode, iters, rate, ms, flush=True)
    which = sys.argv[1:] or ["cfg2", "cfg4_1chunk", "cfg4_2x2", "cfg5_1chunk", "cfg3_1chunk"]
    for w in which:
        if w == "cfg4_1chunk":
            spec = CONFIGS["cfg4"].scaled(name="cfg4 chunk (250,420)", blocks=1, ir0=250, ii0=420)
        elif w == "cfg4_2x2":
            spec = CONFIGS["cfg4"].scaled(name="cfg4 2x2 chunks at (249,419)", blocks=2, ir0=249, ii0=419)
        elif w == "cfg5_1chunk":
            spec = CONFIGS["cfg5"].scaled(name="cfg5 chunk centre", blocks=1, ir0=1256357, ii0=2131825)
        elif w == "cfg3_1chunk":
            spec = CONFIGS["cfg3"].scaled(name="cfg3 chunk (1,1)", blocks=1, ir0=1, ii0=1)
        else:
            spec = CONFIGS[w]
        t0 = time.time()
        r = time_frame(spec)
        r["wall_s"] = time.time() - t0
        res["frames"].append(r)
        print(json.dumps(r), flush=True)
    if os.environ.get("PROBE_PERCHUNK"):
        base = CONFIGS[os.environ["PROBE_PERCHUNK"]]
        rows = []
        for (ir, ii) in base.chunk_indices():
            r = time_frame(base.scaled(name=f"{ir},{ii}", blocks=1, ir0=ir, ii0=ii), reps=1)
            rows.append({"ir": ir, "ii": ii, "useful": r["useful"], "ms": r["ms"][0]})
        res["perchunk"] = rows
        tot_u = sum(r["useful"] for r in rows); tot_ms = sum(r["ms"] for r in rows)
        print("perchunk total useful", tot_u, "ms", tot_ms, "Gpx-it/s", tot_u / tot_ms / 1e6, flush=True)
    with open(os.path.join(ROOT, "gpurun_out", "probe.json"), "w") as f:
        json.dump(res, f, indent=1)


if __name__ == "__main__":
    main()
