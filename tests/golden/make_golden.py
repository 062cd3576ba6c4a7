"""Regenerates the golden fixtures of tests/golden/ (run from the repo root, CPU only):

    python tests/golden/make_golden.py

Sources of truth: the CPU oracle (oracle/, security mode AND the literal no_radius mode where it
is affordable) and, independently, Qhull through scipy.spatial.Voronoi on the mirrored point set.
The reference crate itself cannot produce these numbers, hence no fixture
comes from it.
"""
import importlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import helpers  # noqa: E402
import oracle_binding as ob  # noqa: E402

gen = importlib.import_module("the-tessellator_b200.generators")
OUT = os.path.dirname(os.path.abspath(__file__))
BOX = [0, 0, 0, 1, 1, 1]


def pack(name, pts_desc, r, extra=None):
    nb, ar = helpers.sorted_cells(r.face_offsets, r.neighbors, r.areas)
    d = dict(desc=np.array(pts_desc), volumes=r.volumes, face_offsets=r.face_offsets.astype(np.int64), neighbors=nb.astype(np.int32), areas=ar,
             counters=np.array([r.counters[k] for k in ("visited", "tested", "vertex_classifications", "cuts", "new_vertices", "table_entries", "degenerate_skips", "faces")], dtype=np.int64))
    if extra:
        d.update(extra)
    np.savez_compressed(os.path.join(OUT, name), **d)
    print(name, "cells", r.n, "faces", len(nb), "sum vol", repr(float(r.volumes.sum())))


def main():
    # config 1: 10k uniform, seed 1 — literal no_radius (O(N^2)) must equal security mode bitwise
    pts = gen.uniform(10_000, 1)
    d = ob.Diagram(pts, box=BOX)
    sec = d.compute_cells(mode=ob.MODE_SECURITY)
    lit = d.compute_cells(mode=ob.MODE_NO_RADIUS)
    assert np.array_equal(sec.volumes, lit.volumes) and np.array_equal(sec.neighbors, lit.neighbors) and np.array_equal(sec.areas, lit.areas)
    qn, qv = helpers.qhull_cells(pts, BOX)
    for c in range(len(pts)):
        assert set(int(x) for x in sec.cell_neighbors(c)) == qn[c], c
    assert np.max(np.abs(qv - sec.volumes) / sec.volumes) < 1e-9
    pack("config1_uniform_10k_seed1.npz", "uniform(10000, seed=1), box [0,1]^3", sec, dict(qhull_volumes=qv))

    # small uniform set with non-cubic bounds and no explicit container (bbox container)
    pts = gen.uniform(1500, 9) * np.array([1.0, 2.0, 0.5]) + np.array([-3.0, 10.0, 0.25])
    d = ob.Diagram(pts, box=None)
    sec = d.compute_cells(mode=ob.MODE_SECURITY)
    lit = d.compute_cells(mode=ob.MODE_NO_RADIUS)
    assert np.array_equal(sec.volumes, lit.volumes) and np.array_equal(sec.neighbors, lit.neighbors)
    pack("oblong_1500_seed9_bbox.npz", "uniform(1500, seed=9)*[1,2,.5]+[-3,10,.25], container = bbox", sec, dict(box=d.container_box()))

    # clustered (config 4 recipe, small)
    pts = gen.clustered(4000, 4, k=4, sigma=0.02)
    d = ob.Diagram(pts, box=BOX)
    sec = d.compute_cells(mode=ob.MODE_SECURITY)
    lit = d.compute_cells(mode=ob.MODE_NO_RADIUS)
    assert np.array_equal(sec.volumes, lit.volumes) and np.array_equal(sec.neighbors, lit.neighbors)
    pack("clustered_4000_seed4.npz", "clustered(4000, seed=4, k=4, sigma=0.02), box [0,1]^3", sec, dict(points=pts))

    # jittered BCC (config 5 recipe, small): m=8 -> 1024 points
    pts = gen.bcc(8, 5)
    d = ob.Diagram(pts, box=BOX)
    sec = d.compute_cells(mode=ob.MODE_SECURITY)
    lit = d.compute_cells(mode=ob.MODE_NO_RADIUS)
    assert np.array_equal(sec.volumes, lit.volumes) and np.array_equal(sec.neighbors, lit.neighbors)
    pack("bcc_m8_seed5.npz", "bcc(8, seed=5), box [0,1]^3", sec)

    # exact degeneracy: un-jittered simple cubic 6^3 — interior cells are cubes with exactly 6 faces
    pts = gen.simple_cubic(6)
    d = ob.Diagram(pts, box=BOX)
    sec = d.compute_cells(mode=ob.MODE_SECURITY)
    lit = d.compute_cells(mode=ob.MODE_NO_RADIUS)
    assert np.array_equal(sec.volumes, lit.volumes) and np.array_equal(sec.neighbors, lit.neighbors)
    pack("simple_cubic_6.npz", "simple_cubic(6), box [0,1]^3", sec, dict(status=sec.status))


if __name__ == "__main__":
    main()
