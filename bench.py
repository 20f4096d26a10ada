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
# Sum of useful pixel-iterations of the full frames as produced by the parity-tested single-GPU path
# (tests/test_gpu_parity.py::test_cfg4_full_frame_properties; cfg5/cfg2/cfg3 recorded from the same build).
# A multi-GPU run (work stealing + gather) must land on the same number: a fast frame that differs is wrong.
EXPECTED_USEFUL = {"cfg4": 27_914_249_068_737, "cfg5": 6_306_807_763_045, "cfg2": 1_666_561_672, "cfg3": 104_413_119_766}
NOMINAL_FMA_TERA = {"f64": 148 * 64 * 1.965e9 / 1e12, "f32": 148 * 128 * 1.965e9 / 1e12}


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
def cpu_sample(spec, d: int, threads: int):
    """Render every chunk of the frame at reduced definition d with the oracle; (useful, seconds)."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import oracle
    t0 = time.perf_counter()
    _counts, useful = oracle.render_frame(spec.level, spec.ir0, spec.ii0, spec.blocks, d, spec.mrd,
                                          dtype=spec.dtype, threads=threads, want_counts=False)
    return useful, time.perf_counter() - t0


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
    ap.add_argument("--e2e-steps", type=int, default=1)
    ap.add_argument("--gather", default="nccl", choices=["nccl", "p2p"],
                    help="N > 1: nccl = one batched send/recv gather to rank 0; p2p = kernels store directly "
                         "into rank 0's frame buffer over NVLink peer memory (no gather)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    spec = get_spec(args.workload)
    if args.impl == "reference":
        run_reference(args, spec)
        return

    import torch
    import torch.distributed as dist
    from distributedmandelbrot_b200 import _native as nat
    from distributedmandelbrot_b200 import worker
    from distributedmandelbrot_b200.frame import FrameRenderer, probe_chunk_costs, queue_name, shared_frame_buffer

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

    def step():
        if world == 1:
            fr.render_all()                # every band in one launch
        else:
            fr.render_stolen(queue)        # work-stealing dispatch ...
            if shared_out is None:
                fr.gather_to_root()        # ... and the one collective: bands -> rank 0
            # (p2p: the stores already landed in rank 0's buffer; the barrier + synchronize below close the step)

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

    sampler = ClockSampler(local)
    times = []
    launches0 = fr.launches
    if rank == 0:
        sampler.start()
    for _ in range(args.steps):
        rearm()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        step()
        e1.record()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)       # max over ranks
        times.append(float(ms.item()))
    clocks = sampler.stop() if rank == 0 else None
    launches = torch.tensor([fr.launches - launches0], device=dev)
    steals = torch.tensor([fr.steals], device=dev)
    if world > 1:
        dist.all_reduce(launches)
        dist.all_reduce(steals)
    total_ms = sum(times)
    ms_per_step = total_ms / args.steps

    # ---- e2e: the reference-facing worker API with HOST buffers (rank-local share of the chunks)
    # chunks are handed out dynamically by a second work-stealing queue, the way a Distributer
    # hands leases to whichever worker asks next
    chunks = spec.chunk_indices()
    wl_all = [(spec.level, spec.mrd, ir, ii) for (ir, ii) in chunks]
    e2e_s = None
    if spec.dtype == "f64" and args.e2e_steps > 0:
        cq = None
        if world > 1:
            cqn = queue_name() + "_c"
            if rank == 0:
                cq = nat.TileQueue(cqn, world, len(wl_all), create=True)
            dist.barrier()
            if rank != 0:
                cq = nat.TileQueue(cqn)

        order = list(range(len(wl_all)))

        def source():
            if cq is None:
                yield from wl_all
                return
            while (g := cq.pop(rank, 1)) is not None:
                for k in range(g[0], g[0] + g[1]):
                    yield wl_all[order[k]]

        worker.process_workloads(wl_all[:2], definition=spec.definition, consume=lambda *_: None)   # warm the pinned ring
        if world > 1:
            probe_chunk_costs(spec, device=local)                                                   # ... and the probe's torch ops
        total_s = 0.0
        nbytes = 0

        def consume(_w, payload):
            nonlocal nbytes
            nbytes += payload.nbytes

        for _ in range(args.e2e_steps):
            if world > 1:
                dist.barrier()
                if rank == 0:
                    cq.reset(len(wl_all))
                dist.barrier()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            if world > 1:
                # expensive chunks first (a 2 ms low-resolution probe, inside the timed region; every rank
                # computes the same deterministic order)
                costs = probe_chunk_costs(spec, device=local)
                lpt = sorted(range(len(wl_all)), key=lambda k: -costs[k])
                # the queue starts from a static contiguous split: deal the LPT order round-robin so that
                # every rank's own range is "its share, expensive first" and steals take cheap chunks
                order, n_all = [], len(wl_all)
                for r in range(world):
                    lo, hi = n_all * r // world, n_all * (r + 1) // world
                    order += [lpt[j * world + r] for j in range(hi - lo) if j * world + r < n_all]
                missing = [k for k in lpt if k not in set(order)]
                order += missing
            worker.process_workloads(source(), definition=spec.definition, consume=consume)
            torch.cuda.synchronize()
            dt = torch.tensor([time.perf_counter() - t0], device=dev, dtype=torch.float64)
            if world > 1:
                dist.all_reduce(dt, op=dist.ReduceOp.MAX)       # max over ranks, per step
            total_s += float(dt.item())
        e2e_s = total_s / args.e2e_steps
        if cq is not None:
            dist.barrier()
            cq.close()

    if rank == 0 and args.workload in EXPECTED_USEFUL and useful != EXPECTED_USEFUL[args.workload]:
        raise SystemExit(f"PARITY GUARD: frame {args.workload} sums to {useful} useful pixel-iterations, "
                         f"expected {EXPECTED_USEFUL[args.workload]} -- the rendered frame is wrong")
    if rank == 0:
        value = useful / (ms_per_step * 1e-3) / 1e9
        peak_fma, _ = nat.measure_pipe_rate(spec.dtype, 0, 16384)        # tera FMA-pipe lane-ops/s, live
        ops_per_s = useful * OPS_PER_PIXEL_ITERATION / (ms_per_step * 1e-3)
        # dominant kernel = escape_tiles_kernel; at N = 1 a step IS one launch of it
        achieved_tflops = ops_per_s * 2 / 1e12 * (1.0 if world == 1 else 1.0 / world)
        peak_tflops = peak_fma * 2
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "ncu_traffic.json")
        if os.path.exists(tpath):
            try:
                traffic = json.load(open(tpath)).get(args.workload)
            except Exception:  # noqa: BLE001
                traffic = None
        out = {
            "metric": METRIC, "value": value, "unit": METRIC, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": spec.dtype,
            "data": "synthetic (pure function of chunk indices; no RNG, no seeds)",
            "config": {"workload": spec.name, "level": spec.level, "blocks": spec.blocks,
                       "definition": spec.definition, "ir0": spec.ir0, "ii0": spec.ii0, "mrd": spec.mrd,
                       "pixels": spec.pixels, "useful_pixel_iterations": useful,
                       "mean_iterations_per_pixel": useful / spec.pixels, "band_rows": spec.band_rows,
                       "parallelism": "1 GPU, one launch over all bands" if world == 1 else
                                      f"{world} GPUs, shm work-stealing bands + " + ("one batched NCCL send/recv gather to rank 0" if shared_out is None
                                                                                      else "kernel stores directly into rank 0's buffer over NVLink peer memory"),
                       "l2": f"kernel reads only 88-byte descriptors; output {spec.pixels * 4 / 2**30:.1f} GiB int32 per step >> 126 MB L2 (no L2 flush needed)",
                       "steals": int(steals.item()),
                       "parity_guard": "frame sum equals the parity-tested single-GPU value" if args.workload in EXPECTED_USEFUL else "none for this workload"},
            "clocks": clocks,
            "gpu_launches": int(launches.item()),
            "roofline": {"bound": "fp64_pipe" if spec.dtype == "f64" else "fp32_pipe",
                         "achieved": achieved_tflops, "peak": peak_tflops, "unit": "TFLOP/s",
                         "frac": achieved_tflops / peak_tflops,
                         "per_gpu": world > 1,
                         "convention": f"{OPS_PER_PIXEL_ITERATION} pipe instructions per pixel-iteration, each counted as one FMA slot = 2 FLOP",
                         "peak_source": "measured live: mb_measure_fma_peak (independent FMA chains on all SMs) on this GPU",
                         "nominal_peak": NOMINAL_FMA_TERA[spec.dtype] * 2,
                         "algorithmic_bytes_per_launch": spec.pixels * 4, "traffic": traffic},
        }
        if spec.dtype == "f64" and args.e2e_steps > 0:
            out["e2e"] = {"value": useful / e2e_s / 1e9, "unit": METRIC,
                          "h2d_bytes_per_step": 88 * spec.n_chunks, "d2h_bytes_per_step": spec.pixels,
                          "api": "worker.process_workloads (uint8 DataChunk payloads in host memory, one per (level, mrd, ir, ii))",
                          "steps": args.e2e_steps, "s_per_step": e2e_s}
        else:
            out["e2e"] = None
        if not args.no_cpu_baseline and world == 1:      # timed on rank 0 at N = 1 only
            out["cpu_baseline"], _d = cpu_baseline(spec)
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
