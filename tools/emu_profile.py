#!/usr/bin/env python
"""Profile of the clip kernel on the CPU warp emulator (tests/emu): work counters per cell and how often every warp
collective of clip.cu runs per cell, by source line.  No GPU needed.

    python tools/emu_profile.py [n_points] [uniform|clustered|bcc] [tier: small|medium|large]

This is how the candidate screen, the packed shuffles and the medium configuration were sized before they were
A/B-timed on the GPU (tools/ab_run.sh)."""
import ctypes as C
import importlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
pkg = importlib.import_module("the-tessellator_b200")
import emu_binding as eb  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 20000
kind = sys.argv[2] if len(sys.argv) > 2 else "uniform"
tier = {"small": False, "medium": "medium", "large": True}[sys.argv[3] if len(sys.argv) > 3 else "small"]
gen = pkg.generators
pts = {"uniform": lambda: gen.uniform(n, 1), "clustered": lambda: gen.clustered(n, 4), "bcc": lambda: gen.bcc(max(2, round((n / 2) ** (1 / 3))), 5)}[kind]()
g = eb.EmuGrid(pts, table_radius=8)
L = eb.lib()
L.emu_line_hist.restype = C.POINTER(C.c_uint64)
L.emu_line_hist_enable(1)
e = g.clip(os_threads=1, large=tier)  # the line histogram is kept by a single host thread
h = np.ctypeslib.as_array(L.emu_line_hist(), shape=(65536,)).copy()
L.emu_line_hist_enable(0)
m = len(pts)
ok = (e.status & 0x16) == 0
print(f"{m} {kind} cells, {int(ok.sum())} finished by this configuration; collectives per cell {e.collectives / m:.1f}")
print("per cell:", {k: round(v / m, 2) for k, v in e.counters.items()})
src = open(os.path.join(ROOT, "the-tessellator_b200", "csrc", "clip.cu")).read().split("\n")
print(" line  per cell  collective")
for ln in np.nonzero(h)[0]:
    if h[ln] / m >= 0.05:
        print("%5d %9.2f  %s" % (ln, h[ln] / m, src[ln - 1].strip()[:110]))
