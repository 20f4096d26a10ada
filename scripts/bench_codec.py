#!/usr/bin/env python
"""Device DataChunk serializer throughput (f2): algorithmic bytes = n read + serialised bytes written."""
import argparse
import ctypes as C
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch

from distributedmandelbrot_b200 import _native as nat

ap = argparse.ArgumentParser()
ap.add_argument("--only", default=None, help="substring of the case name (for ncu captures)")
ap.add_argument("--nb", type=int, nargs="*", default=[1, 64])
ap.add_argument("--iters", type=int, default=10)
ap.add_argument("--out", default="codec_bench.json", help="file name under gpurun_out/")
args = ap.parse_args()
nat.init(0)
L = nat.lib()
n = 4096 * 4096
peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
hbm = peaks.get("hbm_gbs", 6650.0)
res = {"hbm_peak_gbs": hbm, "peak_source": "MEASURED_PEAKS.json" if peaks else "fallback 6650", "cases": []}
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
st = torch.cuda.current_stream().cuda_stream
stride = L.mb_serialize_out_stride(n)
p = torch.empty(n, dtype=torch.uint8, device="cuda")
cases = {}
t = nat.chunk_geometry(4, 1, 2, 4096, 256)
nat.check(L.mb_render_u8(C.byref(t), 1, p.data_ptr(), 0, st))
cases["chunk (4,256,1,2) [RLE]"] = p.clone()
t = nat.chunk_geometry(800, 250, 420, 4096, 10000)
nat.check(L.mb_render_u8(C.byref(t), 1, p.data_ptr(), 0, st))
cases["chunk cfg4 (250,420) mrd=10000 [RLE, 3M runs]"] = p.clone()
cases["noise [raw]"] = torch.randint(0, 256, (n,), dtype=torch.uint8, device="cuda")
cases["constant [RLE, 1 run]"] = torch.ones(n, dtype=torch.uint8, device="cuda")
for nb in args.nb:
    ws_bytes = L.mb_serialize_workspace_bytes(n, nb)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device="cuda")
    out = torch.empty(nb * stride, dtype=torch.uint8, device="cuda")
    info = torch.empty(nb * 32, dtype=torch.uint8, device="cuda")
    for name, one in cases.items():
        if args.only and args.only not in name:
            continue
        payload = one.repeat(nb)
        times = []
        for _ in range(args.iters):
            flush.fill_(1)                       # evict the payloads from the 126 MB L2
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            nat.check(L.mb_serialize_chunks(payload.data_ptr(), n, nb, n, out.data_ptr(), stride, ws.data_ptr(), ws_bytes, info.data_ptr(), st))
            e1.record()
            torch.cuda.synchronize()
            times.append(e0.elapsed_time(e1))
        ms = sorted(times)[len(times) // 2]
        olen = int(info.cpu().numpy().view("<u8")[0])
        gbs = nb * (n + olen) / (ms * 1e-3) / 1e9
        res["cases"].append({"case": name, "chunks_per_call": nb, "codec": int(out[0].item()), "serialized_bytes_per_chunk": olen,
                             "ms_median": ms, "us_per_chunk": ms * 1e3 / nb, "algorithmic_gb_s": gbs, "frac_of_hbm_peak": gbs / hbm})
        print(res["cases"][-1], flush=True)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(res, open(os.path.join(ROOT, "gpurun_out", args.out), "w"), indent=1)
