#!/usr/bin/env python
"""Runs one named BASELINE config on one GPU and prints closure / status / timing facts:
    python tools/config_check.py clustered 10000000      (config 4)
    python tools/config_check.py bcc 128                  (config 5 recipe, 2*m^3 points)
"""
import importlib
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
T = importlib.import_module("the-tessellator_b200")
import helpers  # noqa: E402


def main():
    kind, size = sys.argv[1], int(sys.argv[2])
    gen = T.generators
    t0 = time.time()
    pts = {"clustered": lambda: gen.clustered(size, 4), "bcc": lambda: gen.bcc(size, 5), "uniform": lambda: gen.uniform(size, 3)}[kind]()
    print(f"{kind}: {len(pts)} points generated in {time.time() - t0:.1f}s", flush=True)
    d = T.Diagram(0)
    d.add_particles(pts)
    d.initialize(T.Polyhedron(0, 0, 0, 1, 1, 1))
    for it in range(2):
        t0 = time.time()
        b = d.compute_all_cells(outputs=1 | 2 | 4 | 16)
        dt = time.time() - t0
        print(f"run {it}: {dt * 1e3:.1f} ms wall, timings {b.timings()}, binning {d.binning_ms():.2f} ms", flush=True)
    st = b.status
    print("cells", b.n_cells, "faces", b.n_faces, "faces/cell", b.n_faces / b.n_cells)
    print("status histogram", {int(k): int(v) for k, v in zip(*np.unique(st, return_counts=True))})
    print("sum of volumes", repr(float(b.volumes.sum())), "device sum", repr(b.volume_sum()), "min vol", float(b.volumes.min()))
    nf = np.diff(b.face_offsets)
    print("faces per cell: max", int(nf.max()), "p99.9", float(np.percentile(nf, 99.9)))
    print("counters per cell", {k: v / b.n_cells for k, v in b.counters().items()})
    if len(pts) <= 20_000_000:
        print("neighbour symmetry violations", helpers.neighbor_symmetry_violations(b.face_offsets, b.neighbors))
    if len(sys.argv) > 3:  # oracle on a sample
        import oracle_binding as ob

        ids = np.unique((gen.u01(5, np.arange(int(sys.argv[3]), dtype=np.uint64)) * len(pts)).astype(np.uint64))
        r = ob.Diagram(pts, box=[0, 0, 0, 1, 1, 1], table_radius=24).compute_cells(ids=ids, mode=ob.MODE_SECURITY)
        ok = r.status == 0
        bad = 0
        fo = b.face_offsets
        for k, i in enumerate(ids.astype(np.int64)):
            if not ok[k]:
                continue
            if sorted(b.neighbors[fo[i]:fo[i + 1]].tolist()) != sorted(r.cell_neighbors(k).tolist()) or abs(b.volumes[i] - r.volumes[k]) > 1e-12 * r.volumes[k]:
                bad += 1
        print(f"oracle sample: {int(ok.sum())} cells compared, mismatches {bad}")


if __name__ == "__main__":
    main()
