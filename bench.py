#!/usr/bin/env python
"""bench.py -- Gpixel-iterations/s of the escape-time hot path on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload cfg4] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" is one full render of the workload frame (default cfg4: 65536^2 deep zoom, mrd=10000,
FP64 -- the configuration BASELINE.json's target is quoted on; it fits one GPU: 16 GiB int32).
`value`   = useful pixel-iterations / device time, counts resident in HBM (after the NCCL gather
            to rank 0 when N > 1).  useful = sum over pixels of (n ? n : mrd-1), from the counts.
`e2e`     = same metric through the reference-facing worker API (process_workloads -> uint8
            DataChunk payloads in HOST memory), H2D descriptors + D2H payload inside the timing.
`roofline`= FP64-pipe issue roofline: 7 FP64 instructions per pixel-iteration against
            the DFMA issue rate measured live on this GPU (mb_measure_fma_peak).
`cpu_baseline` = the CPU oracle (strict-IEEE port of the reference loop) on all host cores, on a
            bounded sample of the same frame (same window, reduced definition).
--impl reference times that CPU implementation as the reference arm (the reference has no CPU
path of its own and its numba-CUDA worker cannot run here).
"""
from __future__ import annotations

import argparse
import ctypes
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "Gpixel-iterations/s"
OPS_PER_PIXEL_ITERATION = 7          # FP64/FP32-pipe instructions the rounding sequence needs
ISSUED_PER_PIXEL_ITERATION = 6.125   # what the kernel issues in steady state: the magnitude add runs once per block of 8
NOMINAL_FMA_TERA = {"f64": 148 * 64 * 1.965e9 / 1e12, "f32": 148 * 128 * 1.965e9 / 1e12}


def expected_useful(name: str):
    """Sum of useful pixel-iterations of a full BASELINE frame as the CPU ORACLE computes it
    (tests/golden/fullsize_<cfg>.json, written by oracle/gen_fullsize_golden.py: every chunk rendered whole
    with strict IEEE arithmetic; no number in there comes from a CUDA build).  The rendered frame --
    one GPU or work-stolen across eight -- must land on it: a fast frame that differs is wrong."""
    path = os.path.join(ROOT, "tests", "golden", f"fullsize_{name}.json")
    try:
        with open(path) as f:
            doc = json.load(f)
    except OSError:
        return None
    return doc["useful_total"] if doc.get("complete") else None


def get_spec(name: str):
    from distributedmandelbrot_b200.frame import CONFIGS
    if name in CONFIGS:
        return CONFIGS[name]
    if name == "cfg4_4x4":      # 4x4 chunk sub-block of cfg4 around the zoom centre (for ncu --set full captures)
        return CONFIGS["cfg4"].scaled(name="cfg4 4x4 sub-block at (248,418)", blocks=4, ir0=248, ii0=418)
    if name == "cfg3_1":        # one 4096^2 chunk of cfg3 straddling the set (for the FP32 ncu capture)
        return CONFIGS["cfg3"].scaled(name="cfg3 single chunk (1,1)", blocks=1, ir0=1, ii0=1)
    if name == "cfg4_1":
        return CONFIGS["cfg4"].scaled(name="cfg4 single interior chunk (242,412)", blocks=1)
    raise SystemExit(f"unknown workload {name}")


# ---- clocks during the timed region (B200_PROFILING.md) ----------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, device: int):
        self.device, self.proc, self.lines = device, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(self.device)], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=lambda: self.lines.extend(self.proc.stdout), daemon=True)
            self.t.start()
        except OSError:
            self.proc = None

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        self.proc.wait(timeout=5)
        self.t.join(timeout=5)
        sm, mx, pw, reasons = [], [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2])); pw.append(float(f[3]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        busy = [s for s, p in zip(sm, pw) if p > 0.5 * max(pw)] or sm
        return {"sm_mhz": statistics.median(busy), "sm_max_mhz": max(mx), "power_w_max": max(pw),
                "samples": len(sm), "reasons": sorted(reasons)}


# ---- CPU baseline: the oracle on a bounded sample of the same frame -----------------------------
def cpu_sample(spec, d: int, threads: int, simd: bool = False):
    """Render every chunk of the frame at reduced definition d with the oracle; (useful, seconds)."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle
    t0 = time.perf_counter()
    _counts, useful = oracle.render_frame(spec.level, spec.ir0, spec.ii0, spec.blocks, d, spec.mrd,
                                          dtype=spec.dtype, threads=threads, want_counts=False, simd=simd)
    return useful, time.perf_counter() - t0


def cpu_baseline_simd(spec, d: int, budget_s: float = 8.0):
    """Optional second CPU row: the same loop vectorised (oracle/mandel_oracle_simd.c: -O3,
    AVX-512/AVX2 picked at load time, still no contraction, bit-identical counts) -- what a tuned CPU
    port reaches on these cores.  Same sample window, definition scaled to the budget."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle
    threads = oracle.max_threads()
    useful, sec = cpu_sample(spec, d, threads, simd=True)
    if sec < budget_s / 4:
        d = int(min(2048, d * min(4.0, (budget_s / max(sec, 1e-3)) ** 0.5 * 0.8)))
        useful, sec = cpu_sample(spec, d, threads, simd=True)
    return {"value": useful / sec / 1e9, "unit": METRIC, "cores": threads, "kind": "port, vectorised",
            "vector_bits": oracle.simd_width(),
            "sample": f"same frame, all {spec.n_chunks} chunks at d={d} ({useful:.4g} pixel-iterations in {sec:.1f} s); "
                      f"gcc -O3 -ffp-contract=off, {oracle.simd_width()}-bit vectors, {threads} pthreads"}


def cpu_quota() -> str:
    """cgroup CPU quota of this container, if any (a 128-thread box may grant far fewer core-seconds)."""
    try:
        q, per = open("/sys/fs/cgroup/cpu.max").read().split()
        return "unlimited" if q == "max" else f"{int(q) / int(per):.1f} cores"
    except Exception:  # noqa: BLE001
        return "unknown"


def cpu_baseline(spec, budget_s: float = 15.0):
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle
    threads = oracle.max_threads()
    d = 16
    useful, sec = cpu_sample(spec, d, threads)            # calibration pass
    while sec < budget_s / 6 and d < 1024:
        scale = max(1.5, min(4.0, (budget_s / max(sec, 1e-3)) ** 0.5 * 0.8))
        d = int(min(1024, max(d + 8, d * scale)))
        useful, sec = cpu_sample(spec, d, threads)
    return {"value": useful / sec / 1e9, "unit": METRIC, "cores": threads, "cpu_quota": cpu_quota(), "kind": "port",
            "sample": f"same frame ({spec.name}), all {spec.n_chunks} chunks at reduced definition d={d} "
                      f"({useful:.4g} pixel-iterations in {sec:.1f} s); oracle/mandel_oracle.c, gcc -O2 -ffp-contract=off, "
                      f"{threads} pthreads, one dynamic (chunk,row) queue; cgroup cpu quota: {cpu_quota()}"}, d


def serializer_row(nat, torch, nb: int = 64, iters: int = 5) -> dict:
    """Informational row for the other hand-written kernel on the worker's path: the device
    DataChunk serializer on `nb` copies of the run-dense boundary chunk of the cfg4 frame, one set of launches,
    payloads resident in HBM, L2 flushed before every call, CUDA events on the launching stream.  Algorithmic bytes =
    payload read once + serialised bytes written; peak = MEASURED_PEAKS.json's copy bandwidth.  Same measurement as
    scripts/bench_codec.py (profiles/r02_codec_bench.json).  Not part of `value`, `e2e` or `roofline`."""
    L = nat.lib()
    n = 4096 * 4096
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    peaks = json.load(open(peaks_path)) if os.path.exists(peaks_path) else {}
    hbm = float(peaks.get("hbm_gbs", 6650.0))
    st = torch.cuda.current_stream().cuda_stream
    one = torch.empty(n, dtype=torch.uint8, device="cuda")
    t = nat.chunk_geometry(800, 250, 420, 4096, 10000)
    nat.check(L.mb_render_u8(ctypes.byref(t), 1, one.data_ptr(), 0, st))
    payload = one.repeat(nb)
    stride = L.mb_serialize_out_stride(n)
    ws_bytes = L.mb_serialize_workspace_bytes(n, nb)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device="cuda")
    out = torch.empty(nb * stride, dtype=torch.uint8, device="cuda")
    info = torch.empty(nb * 32, dtype=torch.uint8, device="cuda")
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    times = []
    for _ in range(iters):
        flush.fill_(1)                                   # evict the payloads from the 126 MB L2
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        nat.check(L.mb_serialize_chunks(payload.data_ptr(), n, nb, n, out.data_ptr(), stride, ws.data_ptr(), ws_bytes, info.data_ptr(), st))
        e1.record()
        torch.cuda.synchronize()
        times.append(e0.elapsed_time(e1))
    ms = statistics.median(times)
    rec = info[:32].cpu().numpy()
    olen, runs = int(rec[:8].view("<u8")[0]), int(rec[8:12].view("<u4")[0])
    gbs = nb * (n + olen) / (ms * 1e-3) / 1e9
    return {"kernel": "mbc::rle_tiles<2> (+ status memset, raw_fixup)", "workload": "64 x DataChunk (800, 250, 420) mrd=10000, 16 MiB each",
            "chunks_per_call": nb, "runs_per_chunk": runs, "serialized_bytes_per_chunk": olen, "ms_median": ms, "us_per_chunk": ms * 1e3 / nb,
            "roofline": {"bound": "hbm", "achieved": gbs, "peak": hbm, "unit": "GB/s", "frac": gbs / hbm,
                         "peak_source": "MEASURED_PEAKS.json hbm_gbs" if peaks else "fallback 6650"},
            "l2": "256 MiB written between calls", "iters": iters}


def run_reference(args, spec):
    """Reference arm: the CPU implementation of the path (the oracle port) on all host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle
    threads = oracle.max_threads()
    per_step = max(2.0, min(20.0, 150.0 / (args.steps + args.warmup)))
    base, d = cpu_baseline(spec, budget_s=per_step)
    for _ in range(args.warmup):
        cpu_sample(spec, d, threads)
    t0 = time.perf_counter()
    useful = 0
    for _ in range(args.steps):
        u, _s = cpu_sample(spec, d, threads)
        useful += u
    sec = time.perf_counter() - t0
    value = useful / sec / 1e9
    base["value"] = value
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": METRIC, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": sec / args.steps * 1e3,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": spec.dtype,
        "data": "synthetic (pure function of chunk indices; no RNG)",
        "config": {"workload": spec.name, "level": spec.level, "blocks": spec.blocks, "definition": spec.definition,
                   "sample_definition": d, "mrd": spec.mrd},
        "cpu_baseline": base,
        "e2e": {"value": value, "unit": METRIC, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "note": "the reference has no CPU path (Calc.cs is an unused Lerp) and its numba-CUDA worker is "
                "FMA-contracted and not runnable offline; this arm is the strict-IEEE oracle port on all host cores",
    }), flush=True)


def main() -> None:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg4")
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--gather", default="nccl", choices=["nccl", "p2p"],
                    help="N > 1: nccl = one batched send/recv gather to rank 0; p2p = kernels store directly "
                         "into rank 0's frame buffer over NVLink peer memory (no gather)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-shortcut", action="store_true", help="skip the separately reported periodicity-shortcut legs")
    ap.add_argument("--no-serializer-row", action="store_true", help="skip the informational device-serializer row (N = 1)")
    ap.add_argument("--refill", type=int, default=None, choices=[0, 1],
                    help="1 = persistent-warp kernel with lane refill (MB_TILE_LANE_REFILL), 0 = one warp per 8x4 patch; "
                         "default: the workload's own setting (frame.CONFIGS)")
    args = ap.parse_args()
    spec = get_spec(args.workload)
    if args.refill is not None:
        spec = spec.scaled(refill=bool(args.refill))
    if args.impl == "reference":
        run_reference(args, spec)
        return

    import torch
    import torch.distributed as dist
    from distributedmandelbrot_b200 import _native as nat
    from distributedmandelbrot_b200 import worker
    from distributedmandelbrot_b200.frame import FrameRenderer, InboxWorker, queue_name, shared_frame_buffer

    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world != args.gpus:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}: launch with torch.distributed.run --nproc-per-node {args.gpus}")
    torch.cuda.set_device(local)
    nat.init(local)                        # raises if the CUDA library / device is missing: no fallback
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    barrier = dist.barrier if world > 1 else (lambda: None)

    def allmax(x: float) -> float:
        t = torch.tensor([x], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def gather_list(x: float):
        if world == 1:
            return [x]
        out = [None] * world
        dist.all_gather_object(out, x)
        return out

    shared_out, close_shared = (None, None)
    if world > 1 and args.gather == "p2p":
        shared_out, close_shared = shared_frame_buffer(spec, rank, local)
    fr = FrameRenderer(spec, device=local, rank=rank, world=world, bands_per_claim=4, depth=6, out=shared_out)
    queue = None
    if world > 1:
        qn = queue_name()
        if rank == 0:
            queue = nat.TileQueue(qn, world, spec.n_bands, create=True)
        dist.barrier()
        if rank != 0:
            queue = nat.TileQueue(qn)

    diag = {"busy_ms": [], "gather_ms": [], "tail_ms": []}      # this rank, per timed step

    def step(timed: bool = False):
        if world == 1:
            fr.render_all()                # every band in one launch
            return
        e_a = torch.cuda.Event(enable_timing=True)
        e_b = torch.cuda.Event(enable_timing=True)
        e_c = torch.cuda.Event(enable_timing=True)
        e_a.record()
        fr.render_stolen(queue)            # work-stealing dispatch ...
        e_b.record()
        e_b.synchronize()                  # (the ownership exchange that follows blocks the host anyway)
        t_done = time.monotonic()
        if shared_out is None:
            fr.gather_to_root()            # ... and the one collective: bands -> rank 0
        # (p2p: the stores already landed in rank 0's buffer; the barrier + synchronize below close the step)
        e_c.record()
        if timed:
            e_c.synchronize()
            first, last = ctypes.c_double(), ctypes.c_double()
            nat.check(nat.lib().mb_frame_last_stats(ctypes.byref(first), ctypes.byref(last)))
            diag["busy_ms"].append(e_a.elapsed_time(e_b))                 # this GPU rendering its claims
            diag["gather_ms"].append(e_b.elapsed_time(e_c))              # waiting for the others + the NCCL batch
            diag["tail_ms"].append((t_done - last.value) * 1e3)           # last claim taken -> last kernel finished

    def rearm():
        if world > 1:
            dist.barrier()
            if rank == 0:
                queue.reset(spec.n_bands)
            dist.barrier()

    for _ in range(max(1, args.warmup)):   # at least one untimed frame: the metric's numerator is read from it
        rearm()
        step()
    torch.cuda.synchronize()
    useful = fr.useful_iterations() if rank == 0 else 0     # from the counts resident on rank 0
    want = expected_useful(args.workload)
    if rank == 0 and want is not None and useful != want:
        raise SystemExit(f"PARITY GUARD: frame {args.workload} sums to {useful} useful pixel-iterations, the CPU oracle's "
                         f"per-chunk digests (tests/golden/fullsize_{args.workload}.json) sum to {want} -- the rendered frame is wrong")

    sampler = ClockSampler(local)
    times = []
    launches0, steals0 = fr.launches, fr.steals_total
    if rank == 0:
        sampler.start()
    for _ in range(args.steps):
        rearm()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        step(timed=True)
        e1.record()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        times.append(allmax(e0.elapsed_time(e1)))            # max over ranks
    clocks = sampler.stop() if rank == 0 else None
    launches = torch.tensor([fr.launches - launches0], device=dev)
    steals = torch.tensor([fr.steals_total - steals0], device=dev)
    if world > 1:
        dist.all_reduce(launches)
        dist.all_reduce(steals)
    ms_per_step = sum(times) / args.steps
    per_rank = None
    if world > 1:
        mean = lambda v: sum(v) / max(1, len(v))             # noqa: E731
        per_rank = {k: gather_list(round(mean(v), 3)) for k, v in diag.items()}
        per_rank["bands"] = gather_list(sum(c for _, c in fr.claimed))

    # ---- device-resident frame with the exact periodicity shortcut (separately reported, never `value`) ----
    shortcut = None
    if not args.no_shortcut:
        # the shortcut legs run on the persistent-warp kernel unless --refill 0 asks otherwise: lanes whose orbit closed a
        # cycle are refilled at once instead of idling until the slowest lane of their patch is done (+12 % on cfg4)
        sc_refill = True if args.refill is None else bool(args.refill)
        fs = FrameRenderer(spec.scaled(shortcut=True, refill=sc_refill), device=local, rank=rank, world=world, bands_per_claim=4, depth=6, out=fr.out)
        counter = torch.zeros(1, dtype=torch.int64, device=dev)

        def sc_step():
            if world == 1:
                fs.render_all()
            else:
                fs.render_stolen(queue)
                if shared_out is None:
                    fs.gather_to_root()

        rearm()
        nat.check(nat.lib().mb_set_executed_counter(counter.data_ptr()))
        sc_step()                                              # warm-up pass; also counts the executed updates
        torch.cuda.synchronize()
        nat.check(nat.lib().mb_set_executed_counter(None))
        executed = torch.tensor([int(counter.item())], device=dev)
        if world > 1:
            dist.all_reduce(executed)
        sc_useful = fs.useful_iterations() if rank == 0 else 0
        if rank == 0 and sc_useful != useful:
            raise SystemExit(f"PARITY GUARD: the shortcut frame sums to {sc_useful}, the plain frame to {useful}")
        sc_times = []
        for _ in range(max(1, min(args.steps, 3))):
            rearm()
            barrier()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            sc_step()
            e1.record()
            torch.cuda.synchronize()
            barrier()
            sc_times.append(allmax(e0.elapsed_time(e1)))
        shortcut = {"ms_per_step": statistics.median(sc_times), "executed_updates": int(executed.item())}
        del fs

    # ---- e2e: the reference-facing worker API with HOST buffers -------------------------------------
    # N = 1: worker.process_workloads (whole chunks, pinned ring).  N > 1: frame.InboxWorker -- the same payloads,
    # but chunks are cut into row bands, the bands go through the work-stealing queue and every rank copies its
    # bands device->host straight into ONE shared pinned frame (no gather, no second copy).
    chunks = spec.chunk_indices()
    wl_all = [(spec.level, spec.mrd, ir, ii) for (ir, ii) in chunks]
    e2e = e2e_sc = None
    if spec.dtype == "f64" and args.e2e_steps > 0:
        nbytes = [0]

        def consume(_w, payload):
            nbytes[0] += payload.nbytes

        def run_e2e(use_shortcut: bool):
            secs = []
            use_refill = spec.refill if not use_shortcut else (True if args.refill is None else bool(args.refill))
            if world == 1:
                worker.process_workloads(wl_all[:2], definition=spec.definition, consume=lambda *_: None, shortcut=use_shortcut, refill=use_refill)   # warm the pinned ring
                for _ in range(args.e2e_steps):
                    torch.cuda.synchronize()
                    t0 = time.perf_counter()
                    worker.process_workloads(wl_all, definition=spec.definition, consume=consume, shortcut=use_shortcut, refill=use_refill)
                    torch.cuda.synchronize()
                    secs.append(time.perf_counter() - t0)
                return secs, "worker.process_workloads (uint8 DataChunk payloads in pinned host memory, one per (level, mrd, ir, ii))"
            iw = InboxWorker(wl_all, definition=spec.definition, band_rows=spec.band_rows, rank=rank, world=world, device=local,
                             name=queue_name() + ("_s" if use_shortcut else "_e"), barrier=dist.barrier, shortcut=use_shortcut,
                             bands_per_claim=4, depth=6, refill=use_refill)
            iw.rearm()
            iw.render()                                        # warm-up pass (staging buffers, page-locked mapping)
            for _ in range(args.e2e_steps):
                iw.rearm()
                torch.cuda.synchronize()
                dist.barrier()
                t0 = time.perf_counter()
                iw.render(consume=consume)                     # returns when this rank's bands are in host memory
                dist.barrier()                                 # ... and the frame is complete when all ranks are here
                secs.append(allmax(time.perf_counter() - t0))
            holes = iw.missing()
            iw.close()
            if holes:
                raise SystemExit(f"e2e: {len(holes)} bands never landed in the host frame")
            return secs, ("frame.InboxWorker (band-granular work stealing; every rank copies its uint8 bands device->host into one "
                          "shared page-locked frame; payloads byte-identical to process_workload)")

        secs, api = run_e2e(False)
        e2e = {"value": useful / statistics.median(secs) / 1e9 if rank == 0 else 0.0, "unit": METRIC,
               "h2d_bytes_per_step": 88 * (spec.n_chunks if world == 1 else spec.n_bands), "d2h_bytes_per_step": spec.pixels,
               "api": api, "steps": args.e2e_steps, "s_per_step": statistics.median(secs), "s_min": min(secs), "s_max": max(secs),
               "statistic": "median over steps of the max over ranks"}
        if not args.no_shortcut:
            secs, _ = run_e2e(True)
            e2e_sc = {"value": useful / statistics.median(secs) / 1e9 if rank == 0 else 0.0, "unit": METRIC, "steps": args.e2e_steps,
                      "s_per_step": statistics.median(secs), "s_min": min(secs), "s_max": max(secs)}

    if rank == 0:
        value = useful / (ms_per_step * 1e-3) / 1e9
        peak_fma, _ = nat.measure_pipe_rate(spec.dtype, 0, 16384)        # tera FMA-pipe lane-ops/s, live
        ops_per_s = useful * OPS_PER_PIXEL_ITERATION / (ms_per_step * 1e-3)
        # dominant kernel = escape_tiles_kernel; at N = 1 a step IS one launch of it
        achieved_tflops = ops_per_s * 2 / 1e12 * (1.0 if world == 1 else 1.0 / world)
        peak_tflops = peak_fma * 2
        prof = {}
        tpath = os.path.join(ROOT, "profiles", "ncu_traffic.json")
        if os.path.exists(tpath):
            try:
                prof = json.load(open(tpath))
            except Exception:  # noqa: BLE001
                prof = {}
        out = {
            "metric": METRIC, "value": value, "unit": METRIC, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "ms_min": min(times), "ms_max": max(times), "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": spec.dtype,
            "data": "synthetic (pure function of chunk indices; no RNG, no seeds)",
            "config": {"workload": spec.name, "level": spec.level, "blocks": spec.blocks,
                       "definition": spec.definition, "ir0": spec.ir0, "ii0": spec.ii0, "mrd": spec.mrd,
                       "pixels": spec.pixels, "useful_pixel_iterations": useful,
                       "mean_iterations_per_pixel": useful / spec.pixels, "band_rows": spec.band_rows,
                       "kernel": "escape_refill_kernel (persistent warps, lane refill)" if spec.refill else "escape_tiles_kernel (one warp per 8x4 patch)",
                       "parallelism": "1 GPU, one launch over all bands" if world == 1 else
                                      f"{world} GPUs, shm work-stealing bands + " + ("one batched NCCL send/recv gather to rank 0" if shared_out is None
                                                                                      else "kernel stores directly into rank 0's buffer over NVLink peer memory"),
                       "l2": f"kernel reads only 88-byte descriptors; output {spec.pixels * 4 / 2**30:.1f} GiB int32 per step >> 126 MB L2 (no L2 flush needed)",
                       "steals": int(steals.item()),
                       "parity_guard": ("frame sum equals the CPU oracle's per-chunk digests (tests/golden/fullsize_*.json)" if want is not None
                                        else "none for this workload")},
            "clocks": clocks,
            "gpu_launches": int(launches.item()),
            "roofline": {"bound": "fp64_pipe" if spec.dtype == "f64" else "fp32_pipe",
                         "achieved": achieved_tflops, "peak": peak_tflops, "unit": "TFLOP/s",
                         "frac": achieved_tflops / peak_tflops,
                         "frac_issued": achieved_tflops / peak_tflops * ISSUED_PER_PIXEL_ITERATION / OPS_PER_PIXEL_ITERATION,
                         "per_gpu": world > 1,
                         "convention": f"frac counts the {OPS_PER_PIXEL_ITERATION} pipe instructions the reference's rounding sequence defines per "
                                       f"pixel-iteration (each one FMA slot = 2 FLOP) and may exceed 1; frac_issued counts the "
                                       f"{ISSUED_PER_PIXEL_ITERATION} the kernel issues (magnitude add once per block of 8)",
                         "ncu_pipe_busy": prof.get("fp64_pipe_busy_frac" if spec.dtype == "f64" else "fp32_issue_busy_frac"),
                         "ncu_source": prof.get("pipe_busy_source"),
                         "peak_source": "builder-measured, live in this run: mb_measure_fma_peak (independent FMA chains on all SMs of this GPU); "
                                        "MEASURED_PEAKS.json has no FP64/FP32 FMA entry",
                         "nominal_peak": NOMINAL_FMA_TERA[spec.dtype] * 2,
                         "frac_of_nominal": achieved_tflops / (NOMINAL_FMA_TERA[spec.dtype] * 2),
                         "algorithmic_bytes_per_launch": spec.pixels * 4, "traffic": prof.get(args.workload)},
            "e2e": e2e,
        }
        if per_rank is not None:
            # which of these is large names the limiter: busy_ms spread = imbalance the queue did not fix,
            # tail_ms = size of the last claims, gather_ms (on the fastest rank) = waiting + the collective
            out["per_rank"] = per_rank
        if shortcut is not None:
            out["value_shortcut"] = {"value": useful / (shortcut["ms_per_step"] * 1e-3) / 1e9, "unit": METRIC,
                                     "ms_per_step": shortcut["ms_per_step"], "executed_updates": shortcut["executed_updates"],
                                     "executed_fraction": shortcut["executed_updates"] / useful,
                                     "kernel": "escape_refill_kernel" if (args.refill is None or args.refill) else "escape_tiles_kernel",
                                     "note": "MB_TILE_PERIOD_SHORTCUT: bit-identical counts (parity suite runs with it on), credited work "
                                             "unchanged, executed work reduced; never used for `value` or `roofline`"}
        if e2e_sc is not None:
            out["e2e_shortcut"] = e2e_sc
        if not args.no_serializer_row and world == 1:
            try:
                out["serializer"] = serializer_row(nat, torch)
            except Exception as exc:  # noqa: BLE001 -- informational row only, never fails the bench
                out["serializer"] = {"error": repr(exc)}
        if not args.no_cpu_baseline and world == 1:      # timed on rank 0 at N = 1 only
            out["cpu_baseline"], d_used = cpu_baseline(spec)
            try:
                out["cpu_baseline_simd"] = cpu_baseline_simd(spec, d_used)
            except Exception as exc:  # noqa: BLE001 -- informational row only
                out["cpu_baseline_simd"] = {"error": str(exc)}
        print(json.dumps(out), flush=True)

    if queue is not None:
        dist.barrier()
        queue.close()
    if close_shared is not None:
        del fr
        torch.cuda.synchronize()
        dist.barrier()
        close_shared()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
