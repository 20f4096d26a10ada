#!/usr/bin/env python
"""Summarise ncu artefacts from gpurun_out/ into small tracked text files under profiles/.

    python scripts/summarize_ncu.py launches gpurun_out/launches_r01.csv profiles/r01_launches.txt
    python scripts/summarize_ncu.py report   gpurun_out/prof.ncu-rep     profiles/r01_prof.txt
"""
import collections
import csv
import io
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_warps", "sm__cycles_elapsed.avg.per_second",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_elapsed", "sm__pipe_fmaheavy_cycles_active.avg.pct_of_peak_sustained_elapsed",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "smsp__inst_executed_op_branch.sum",
    "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__thread_inst_executed_per_inst_executed.pct",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__warps_eligible.avg.per_cycle_active",
    "smsp__sass_thread_inst_executed_op_dfma_pred_on.avg.per_cycle_elapsed", "smsp__sass_thread_inst_executed_op_dmul_pred_on.avg.per_cycle_elapsed",
    "smsp__sass_thread_inst_executed_op_dadd_pred_on.avg.per_cycle_elapsed",
    "smsp__sass_thread_inst_executed_op_ffma_pred_on.avg.per_cycle_elapsed", "smsp__sass_thread_inst_executed_op_fmul_pred_on.avg.per_cycle_elapsed",
    "smsp__sass_thread_inst_executed_op_fadd_pred_on.avg.per_cycle_elapsed",
    "smsp__sass_thread_inst_executed_op_dfma_pred_on.sum", "smsp__sass_thread_inst_executed_op_dmul_pred_on.sum",
    "smsp__sass_thread_inst_executed_op_dadd_pred_on.sum", "smsp__sass_thread_inst_executed_op_ffma_pred_on.sum",
    "smsp__sass_thread_inst_executed_op_fmul_pred_on.sum", "smsp__sass_thread_inst_executed_op_fadd_pred_on.sum",
    "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_bytes.sum", "l1tex__t_bytes.sum",
]


def launches(src, dst):
    rows = [r for r in csv.reader(open(src)) if len(r) > 10]
    hdr = rows[0]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = collections.OrderedDict()
    for r in rows[1:]:
        name = r[ki].split("(")[0][:80]
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += float(r[vi].replace(",", ""))
    tot = sum(a[1] for a in agg.values())
    with open(dst, "w") as f:
        f.write(f"# ncu --metrics gpu__time_duration.sum --clock-control none  ({src})\n")
        f.write("# per-launch times are cold-cache and serialised: compare SHARES, not absolutes\n")
        f.write(f"{'kernel':82s} {'launches':>8s} {'total ' + rows[1][ui]:>18s} {'share %':>8s} {'avg/launch':>16s}\n")
        for k, (n, t) in agg.items():
            f.write(f"{k:82s} {n:8d} {t:18.0f} {t / tot * 100:8.2f} {t / n:16.0f}\n")
    print(open(dst).read())


def report(src, dst):
    raw = subprocess.run(["ncu", "-i", src, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    out = [f"# ncu --set full --clock-control none --import-source on   ({src})"]
    for vals in rows[2:]:
        d = dict(zip(hdr, zip(units, vals)))
        out.append(f"\n## kernel: {d['Kernel Name'][1][:120]}   grid {d.get('Grid Size', ('', ''))[1]} block {d.get('Block Size', ('', ''))[1]}")
        for k in KEYS:
            if k in d:
                out.append(f"{k:75s} {d[k][1]:>22s} {d[k][0]}")
        stalls = sorted(((float(v[1].replace(',', '')), k) for k, v in d.items()
                         if k.startswith("smsp__average_warps_issue_stalled") and k.endswith("_per_issue_active.ratio") and v[1]), reverse=True)
        out.append("-- warp stall reasons (average warps stalled per issue-active cycle), top 8")
        for v, k in stalls[:8]:
            out.append(f"   {k.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', ''):40s} {v:8.3f}")
    src_csv = subprocess.run(["ncu", "-i", src, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    srows = list(csv.reader(io.StringIO(src_csv)))
    if len(srows) > 2:
        if srows[0] and srows[0][0] == "Kernel Name":      # first line names the kernel, second is the header
            srows = srows[1:]
        h = srows[0]
        try:
            si = h.index("Source")
            ci = next(i for i, c in enumerate(h) if c.startswith("Warp Stall Sampling (All"))
            ei = next(i for i, c in enumerate(h) if c.startswith("# Instructions Executed") or c == "Instructions Executed")
            hot = sorted(((float(r[ci] or 0), r[si].strip(), r[ei]) for r in srows[1:] if len(r) > max(si, ci, ei) and r[ci]), reverse=True)[:25]
            tot = sum(float(r[ci] or 0) for r in srows[1:] if len(r) > ci and r[ci]) or 1.0
            out.append("\n-- hottest SASS lines by warp-stall samples (source page)")
            for smp, txt, ex in hot:
                out.append(f"   {smp / tot * 100:6.2f}%  exec={ex:>14s}  {txt[:90]}")
        except (ValueError, StopIteration):
            out.append("(source page columns not recognised: " + ", ".join(h[:12]) + ")")
    open(dst, "w").write("\n".join(out) + "\n")
    print("\n".join(out))


if __name__ == "__main__":
    {"launches": launches, "report": report}[sys.argv[1]](sys.argv[2], sys.argv[3])
