#!/usr/bin/env python
"""profiles/clip_kernel_traffic.json from ONE `ncu --set full` capture of the shipped clip kernel:
    python profiles/make_traffic_json.py <rep.ncu-rep> <cells in the captured launch> "<workload text>"
Stores the SHA-256 of the kernel sources next to the counts; bench.py refuses the file when the sources have changed."""
import csv
import hashlib
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SOURCES = ["clip.cu", "clip_thread.cu", "common.cuh", "tess_math.cuh", "cube_tables.cuh", "Makefile"]


def kernel_sources_sha256() -> str:
    h = hashlib.sha256()
    for f in SOURCES:
        h.update(open(os.path.join(ROOT, "the-tessellator_b200", "csrc", f), "rb").read())
    return h.hexdigest()


def main():
    rep, cells, workload = sys.argv[1], int(sys.argv[2]), sys.argv[3]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units, vals = rows[0], rows[1], rows[2]
    d = dict(zip(hdr, vals))
    u = dict(zip(hdr, units))

    def num(k, scale_unit=True):
        v = float(d[k].replace(",", ""))
        if scale_unit:
            v *= {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0, "ms": 1.0, "us": 1e-3, "s": 1e3, "ns": 1e-6}.get(u[k], 1.0)
        return v

    rd, wr = num("dram__bytes_read.sum"), num("dram__bytes_write.sum")
    out = {
        "kernel": d["Kernel Name"],
        "workload": workload,
        "cells_per_launch": cells,
        "dram_bytes_read_per_launch": rd,
        "dram_bytes_write_per_launch": wr,
        "dram_bytes_per_launch": rd + wr,
        "warp_instructions_per_launch": num("smsp__inst_executed.sum", False),
        "issue_active_pct": num("smsp__issue_active.avg.pct_of_peak_sustained_active", False),
        "threads_per_instruction": num("smsp__thread_inst_executed_per_inst_executed.ratio", False),
        "registers_per_thread": num("launch__registers_per_thread", False),
        "local_load_sectors": num("l1tex__t_sectors_pipe_lsu_mem_local_op_ld.sum", False),
        "local_store_sectors": num("l1tex__t_sectors_pipe_lsu_mem_local_op_st.sum", False),
        "duration_ms_under_ncu": num("gpu__time_duration.sum"),
        "source": os.path.basename(rep) + " (ncu --set full --clock-control none --import-source on, one launch)",
        "kernel_sources_sha256": kernel_sources_sha256(),
    }
    json.dump(out, open(os.path.join(ROOT, "profiles", "clip_kernel_traffic.json"), "w"), indent=1)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
