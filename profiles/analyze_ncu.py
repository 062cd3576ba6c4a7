#!/usr/bin/env python
"""Summarise an Nsight Compute report (read on the CPU box):  python profiles/analyze_ncu.py rep.ncu-rep [N]
Prints launch facts, issue/pipe utilisation, stall reasons and the N hottest source lines."""
import collections
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
KEYS = ["Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "l1tex__t_sectors_pipe_lsu_mem_local_op_ld.sum", "l1tex__t_sectors_pipe_lsu_mem_local_op_st.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__warps_eligible.avg.per_cycle_active", "smsp__warps_active.avg.per_cycle_active", "sm__cycles_elapsed.avg"]
for vals in rows[2:]:
    d = dict(zip(hdr, zip(vals, units)))
    print("=" * 100)
    for k in KEYS:
        if k in d:
            print(f"{k:72s} {d[k][0]:>22s} {d[k][1]}")
    print("--- warp stall reasons (warps per issue-active cycle)")
    for h in hdr:
        if "smsp__average_warps_issue_stalled" in h and h.endswith("per_issue_active.ratio") and "not_issued" not in h:
            print(f"    {h[34:-23]:28s} {d[h][0]}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
his = [i for i, r in enumerate(rows) if r and r[0] == "Line No"]
if his:
    hi = his[0]
    h = rows[hi]
    iEx, iS = h.index("Instructions Executed"), h.index("# Samples")
    per, samp, text = collections.Counter(), collections.Counter(), {}
    for r in rows[hi + 1:]:
        if len(r) <= iEx:
            continue
        if not r[0].strip():  # SASS rows repeat the counts of their source line
            continue
        try:
            ex, sm = int(r[iEx]), int(r[iS])
        except ValueError:
            continue
        per[r[0]] += ex
        samp[r[0]] += sm
        text.setdefault(r[0], r[1].strip()[:105])
    tot, tots = sum(per.values()), max(1, sum(samp.values()))
    print("=" * 100)
    print(f"hottest source lines (of {tot} warp instructions, {tots} stall samples)")
    print(f"{'line':>6s} {'inst%':>7s} {'samp%':>7s}  source")
    for ln, ex in per.most_common(top):
        print(f"{ln:>6s} {100 * ex / tot:6.2f}% {100 * samp[ln] / tots:6.2f}%  {text[ln]}")
    # coarse view: instructions by 10-line buckets of the hottest file
    print("=" * 100)
    print("instructions by 10-line bucket")
    b = collections.Counter()
    for ln, ex in per.items():
        try:
            b[int(ln) // 10 * 10] += ex
        except ValueError:
            pass
    for k in sorted(b):
        if b[k] / tot > 0.004:
            print(f"  lines {k:4d}-{k + 9:4d}: {100 * b[k] / tot:6.2f}%  {'#' * int(200 * b[k] / tot)}")
