#!/usr/bin/env python
"""Larger bit-for-bit runs of the clip kernel on the CPU warp emulator (tests/emu) against the oracle: 100k uniform,
60k clustered, 39k BCC and 40k points in an oblong box, both instantiations, the unfinished cells redone in the medium
configuration.  A few minutes on 8 cores; no GPU needed.   python tools/emu_stress.py"""
import os
import sys, importlib, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, 'tests')]
import numpy as np
pkg = importlib.import_module("the-tessellator_b200")
import emu_binding as eb, helpers
gen = pkg.generators
cases = [("uniform", gen.uniform(100000, 101)), ("clustered", gen.clustered(60000, 104)), ("bcc", gen.bcc(27, 105)),
         ("oblong", gen.uniform(40000, 106) * np.array([1.0, 5.0, 0.2]))]
for name, pts in cases:
    t = time.time()
    box = (0, 0, 0, 1, 1, 1) if name != "oblong" else (0, 0, 0, 1, 5, 0.2)
    g = eb.EmuGrid(pts, box)
    for count in (False, True):
        e = g.clip(os_threads=8, count=count)
        if count is False:
            r = g.oracle_cells(nthreads=8)
        ok = (e.status & 0x16) == 0
        class S:
            def __init__(s, x):
                fo = np.asarray(x.face_offsets, np.int64); cnt = np.diff(fo)[ok]
                s.volumes = np.asarray(x.volumes)[ok]; s.face_offsets = np.concatenate([[0], np.cumsum(cnt)])
                sel = np.repeat(ok, np.diff(fo)); s.neighbors = np.asarray(x.neighbors)[sel]; s.areas = np.asarray(x.areas)[sel]
        helpers.assert_cells_identical(S(e), S(r), name)
        assert np.array_equal(e.status[ok], r.status[ok])
        if count and ok.all():
            for k in ("visited", "tested", "vertex_classifications", "cuts", "new_vertices", "table_entries", "faces"):
                assert e.counters[k] == r.counters[k], k
    # redo the unfinished cells in the medium configuration
    bad = np.sort(np.nonzero(~ok)[0]).astype(np.uint32)
    if len(bad):
        em = g.clip(work_slots=bad, large="medium", count=False)
        rm = g.oracle_cells(slots=bad, nthreads=8)
        okm = (em.status & 0x16) == 0
        assert np.array_equal(em.volumes[okm], rm.volumes[okm])
    print(name, len(pts), "cells, finished by the small tables:", int(ok.sum()), "medium redo:", len(bad), "%.0fs" % (time.time() - t), flush=True)
