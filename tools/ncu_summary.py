#!/usr/bin/env python3
"""Summarise an .ncu-rep (raw page + per-source-line stall samples) into text.
usage: tools/ncu_summary.py report.ncu-rep [top_lines]"""
import csv
import io
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
    "launch__grid_size", "launch__block_size", "launch__waves_per_multiprocessor",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "smsp__inst_executed.sum", "sm__cycles_active.avg", "sm__cycles_elapsed.max",
    "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "smsp__thread_inst_executed_per_inst_executed.ratio",
    "smsp__average_warp_latency_per_inst_issued.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_membar_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio",
    "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct",
]


def ncu(rep, *args):
    return subprocess.run(["ncu", "-i", rep, *args], capture_output=True, text=True).stdout


def main():
    rep = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
    rows = list(csv.reader(io.StringIO(ncu(rep, "--page", "raw", "--csv"))))
    rows = [r for r in rows if len(r) > 10]
    hdr, units = rows[0], rows[1]
    for r in rows[2:]:
        print("===", r[hdr.index("Kernel Name")][:70])
        for k in KEYS:
            if k in hdr:
                print(f"  {k:86s} {r[hdr.index(k)]:>16s} {units[hdr.index(k)]}")
    src = list(csv.reader(io.StringIO(ncu(rep, "--page", "source", "--csv", "--print-source", "cuda,sass"))))
    sections = [i for i, r in enumerate(src) if r and r[0] == "Line No"]
    seen = set()
    for si, start in enumerate(sections):
        fpath = src[start - 2][1] if start >= 2 else ""
        if fpath in seen or "abdpymc" not in fpath:
            continue
        seen.add(fpath)
        h = src[start]
        end = sections[si + 1] - 2 if si + 1 < len(sections) else len(src)
        iL, iS, iI = h.index("Line No"), h.index("# Samples"), h.index("Instructions Executed")
        agg, text = {}, {}
        for r in src[start + 1:end]:
            if len(r) <= iI:
                continue
            try:
                ln = int(r[iL])
            except ValueError:
                continue
            a = agg.setdefault(ln, [0, 0])
            a[0] += int(r[iS]) if r[iS].isdigit() else 0
            a[1] += int(r[iI]) if r[iI].isdigit() else 0
            text[ln] = r[1]
        ts, ti = sum(a[0] for a in agg.values()), sum(a[1] for a in agg.values())
        print(f"--- {fpath}: {ts} stall samples, {ti} warp instructions")
        for ln, a in sorted(agg.items(), key=lambda kv: -kv[1][0])[:top]:
            print(f"  {ln:5d} samples {a[0]:6d} {100 * a[0] / max(ts, 1):5.1f}%   inst {a[1]:9d} {100 * a[1] / max(ti, 1):5.1f}%   {text[ln].strip()[:88]}")


if __name__ == "__main__":
    main()
