"""Shared helpers of the parity tests (test infrastructure)."""
from __future__ import annotations

import numpy as np

# wall ids of include/tess.h: -1 y_min(F) -2 x_max(R) -3 y_max(B) -4 x_min(L) -5 z_max(U) -6 z_min(D)
WALL_ID = {(0, 0): -4, (0, 1): -2, (1, 0): -1, (1, 1): -3, (2, 0): -6, (2, 1): -5}

VOL_RTOL = 1e-12   # north_star: per-cell volumes within 1e-12 relative (f64)
AREA_RTOL = 1e-12  # north_star: face areas within 1e-12 relative (f64)


def qhull_cells(pts: np.ndarray, box=(0, 0, 0, 1, 1, 1)):
    """Independent reference: Voronoi cells clipped to the box via Qhull.  Every point is mirrored
    across the six walls, which makes each wall the bisector between a point and its own image,
    so all original cells are bounded.  Returns (list of neighbour-id sets, volumes)."""
    from scipy.spatial import ConvexHull, Voronoi

    n = len(pts)
    lo, hi = np.array(box[:3], float), np.array(box[3:], float)
    mir = [pts]
    for ax in range(3):
        for k, w in enumerate((lo[ax], hi[ax])):
            q = pts.copy()
            q[:, ax] = 2 * w - q[:, ax]
            mir.append(q)
    vor = Voronoi(np.concatenate(mir))
    nb = [set() for _ in range(n)]
    for a, b in vor.ridge_points:
        for u, v in ((a, b), (b, a)):
            if u >= n:
                continue
            if v < n:
                nb[u].add(int(v))
            else:
                ax, k = divmod(int(v) // n - 1, 2)
                if int(v) % n == u:
                    nb[u].add(WALL_ID[(ax, k)])
                # a ridge with another point's image lies outside the box: Qhull reports it only
                # when it is degenerate (touching the wall), which seeded random input never is
    vol = np.array([ConvexHull(vor.vertices[vor.regions[vor.point_region[c]]]).volume for c in range(n)])
    return nb, vol


def sorted_cells(offsets: np.ndarray, neighbors: np.ndarray, areas: np.ndarray | None = None):
    """Canonical form of a CSR face list: inside every cell, faces sorted by neighbour id
    (the reference leaves the face order unspecified)."""
    offsets = np.asarray(offsets, dtype=np.int64)
    counts = np.diff(offsets)
    cell = np.repeat(np.arange(len(counts), dtype=np.int64), counts)
    order = np.lexsort((neighbors, cell))
    return neighbors[order], (None if areas is None else areas[order])


def assert_cells_identical(got, ref, what=""):
    """Bit-for-bit: same faces in the same (face-slot) order, same volumes, same areas.  The CUDA path
    allocates half-edge and face slots in pool.rs's LIFO order and sums in the reference's order, so it
    reproduces the oracle exactly, not just within the 1e-12 bar."""
    assert np.array_equal(np.asarray(got.face_offsets, np.int64), np.asarray(ref.face_offsets, np.int64)), f"{what}: face counts differ"
    assert np.array_equal(np.asarray(got.neighbors, np.int64), np.asarray(ref.neighbors, np.int64)), f"{what}: neighbour lists differ (order included)"
    gv, rv = np.asarray(got.volumes), np.asarray(ref.volumes)
    assert np.array_equal(gv, rv), f"{what}: {int((gv != rv).sum())} volumes differ bitwise, max rel {np.max(np.abs(gv - rv) / np.abs(rv)):.3e}"
    ga, ra = np.asarray(got.areas), np.asarray(ref.areas)
    assert np.array_equal(ga, ra), f"{what}: {int((ga != ra).sum())} areas differ bitwise"


def assert_cells_match(got, ref, area_rtol=AREA_RTOL, vol_rtol=VOL_RTOL, what=""):
    """got / ref expose volumes, face_offsets, neighbors, areas (CellBatch or oracle CellResults).
    Topology: the sorted neighbour lists must be identical.  Volumes / areas: relative tolerance."""
    assert np.array_equal(np.asarray(got.face_offsets, np.int64), np.asarray(ref.face_offsets, np.int64)), f"{what}: face counts differ"
    gn, ga = sorted_cells(got.face_offsets, got.neighbors, got.areas)
    rn, ra = sorted_cells(ref.face_offsets, ref.neighbors, ref.areas)
    assert np.array_equal(gn, rn), f"{what}: neighbour sets differ in {int(np.sum(gn != rn))} faces"
    rv = np.asarray(ref.volumes)
    dv = np.abs(np.asarray(got.volumes) - rv)
    assert np.all(dv <= vol_rtol * np.abs(rv)), f"{what}: max rel volume error {np.max(dv / np.abs(rv)):.3e}"
    da = np.abs(ga - ra)
    bad = da > area_rtol * np.abs(ra)
    assert not np.any(bad), f"{what}: {int(bad.sum())} face areas off, max rel {np.max(da[bad] / np.abs(ra[bad])):.3e}"
    return float(np.max(dv / np.abs(rv))), float(np.max(da / np.maximum(np.abs(ra), 1e-300)))


def neighbor_symmetry_violations(offsets, neighbors) -> int:
    """i in nbr(j) <=> j in nbr(i) (walls excluded)."""
    offsets = np.asarray(offsets, dtype=np.int64)
    counts = np.diff(offsets)
    cell = np.repeat(np.arange(len(counts), dtype=np.int64), counts)
    m = neighbors >= 0
    a, b = cell[m], neighbors[m]
    n = len(counts)
    fwd = np.sort(a * n + b)  # (cell, neighbour) pairs are unique: one face per neighbour
    bwd = np.sort(b * n + a)
    return int(np.count_nonzero(fwd != bwd))
