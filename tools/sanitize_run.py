#!/usr/bin/env python
"""Small workload through every kernel of the library, for compute-sanitizer (tools/final_run.sh):
    compute-sanitizer --tool memcheck  python tools/sanitize_run.py all
    compute-sanitizer --tool racecheck python tools/sanitize_run.py warp
`warp` leaves out the thread-per-cell tier: its producer and consumer warps exchange candidates through shared-memory rings
ordered by st.release / ld.acquire, which racecheck (barrier-based) cannot see."""
import importlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
T = importlib.import_module("the-tessellator_b200")
gen = T.generators
which = sys.argv[1] if len(sys.argv) > 1 else "all"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 6000
u = gen.uniform(200, 54)
th, ph = np.arccos(2 * u[:, 0] - 1), 2 * np.pi * u[:, 1]
shell = 0.5 + 0.3 * np.stack([np.sin(th) * np.cos(ph), np.sin(th) * np.sin(ph), np.cos(th)], axis=1)
bg = gen.uniform(n, 64)
pts = np.concatenate([[[0.5, 0.5, 0.5]], shell, bg[np.linalg.norm(bg - 0.5, axis=1) > 0.35], 0.05 + 0.1 * gen.simple_cubic(5)])  # large cell, medium cells, exact lattice
d = T.Diagram(0)
d.add_particles(pts)
d.initialize(T.Polyhedron(0, 0, 0, 1, 1, 1))
ref = None
for tier in (("small", "fast", "thread") if which == "all" else ("small", "fast")):
    T.set_main_tier(tier)
    for outputs in (7, 7 | 16, 7 | 8):
        b = d.compute_all_cells(outputs=outputs)
        cur = (b.volumes.copy(), b.neighbors.copy(), b.areas.copy())
        if ref is None:
            ref = cur
        assert all(np.array_equal(x, y) for x, y in zip(ref, cur)), (tier, outputs)
        assert np.all((b.status & np.uint32(0xFFFFFFFE)) == 0)  # (the exact lattice block is where the reference's D17 defect lives: flagged cells, no closure)
T.set_main_tier("default")
q = d.compute_cells_at(np.array([[0.5, 0.5, 0.5], [0.2, 0.3, 0.4]]), outputs=7)
d.find_neighbors(gen.uniform(8, 1), 0.05, T._lib.QUERY_REAL_RADIUS)
es = d.expanding_search(gen.uniform(4, 2))
es.expand(0.01, 50)
d.find_cells_in_radius(0.5, 0.5, 0.5, 0.1)
print("sanitize_run ok:", which, len(pts), "points, cell 0 has", len(q.cell_neighbors(0)), "faces")
