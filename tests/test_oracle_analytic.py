"""Analytic pins of the CPU oracle's full-cell outputs ("missing" 2).

The reference's own tests hold no volume / area / neighbour golden (polyhedron.rs:1099-1118 only asserts "does not
panic"), the committed fixtures are oracle output, and Qhull agrees only to ~1e-10.  Lattices give closed forms that are
independent of all three:
  * un-jittered BCC, lattice constant a: every interior cell is a truncated octahedron of edge L = a*sqrt(2)/4:
    V = 8*sqrt(2)*L^3 = a^3/2; 6 square faces of area L^2 = a^2/8 towards the second shell (distance a) and 8 regular
    hexagons of area (3*sqrt(3)/2)*L^2 = 3*sqrt(3)*a^2/16 towards the first shell (distance a*sqrt(3)/2);
  * simple cubic: every interior cell is a cube, V = a^3, six faces of area a^2 — the 12 edge and 8 corner neighbours'
    bisectors pass exactly through the cube's edges and corners (the reference's Incident case) and must leave no face.
  * points on a line: every cell is a slab between the two mid-planes, V = (x[i+1] - x[i-1]) / 2 in the unit cube.
Exact lattices are also where the reference's own defect D17 lives (find_outgoing_edge, polyhedron.rs:413-434,
returns None when every Outside vertex is adjacent only to Incident / Outside vertices, and the plane is silently skipped):
the oracle reproduces it and flags those cells (status bit 0, DEGENERATE_SKIP).  The closed forms are asserted for every
unflagged cell, and the flagged ones are required to be a minority.  The same checks run on the GPU output in
tests/test_gpu_parity.py (test_analytic_lattices_on_gpu).
"""
import numpy as np
import pytest

BOX = [0, 0, 0, 1, 1, 1]
RTOL = 1e-12  # north_star: volumes and areas within 1e-12 relative in f64


def bcc_expected(m):
    """(interior point ids, their 14 neighbour ids each sorted) for generators.bcc(m, seed, jitter=0)."""
    a = 1.0 / m
    pid = lambda i, j, k, sub: 2 * ((i * m + j) * m + k) + sub  # noqa: E731
    ids, nbrs = [], []
    for i in range(2, m - 2):
        for j in range(2, m - 2):
            for k in range(2, m - 2):
                for sub in (0, 1):
                    ids.append(pid(i, j, k, sub))
                    second = [pid(i + d[0], j + d[1], k + d[2], sub) for d in ((1, 0, 0), (-1, 0, 0), (0, 1, 0), (0, -1, 0), (0, 0, 1), (0, 0, -1))]
                    # sub 0 at (i,j,k) touches the sub-1 sites (i-1..i, j-1..j, k-1..k); sub 1 the sub-0 sites (i..i+1, ...)
                    o = -1 if sub == 0 else 0
                    first = [pid(i + o + di, j + o + dj, k + o + dk, 1 - sub) for di in (0, 1) for dj in (0, 1) for dk in (0, 1)]
                    nbrs.append(sorted(first + second))
    return np.array(ids), np.array(nbrs), a


def check_bcc(volumes, face_offsets, neighbors, areas, status, m):
    ids, want_nb, a = bcc_expected(m)
    ok = np.asarray(status)[ids] == 0  # cells the reference's D17 defect did not touch
    assert ok.mean() > 0.7
    ids, want_nb = ids[ok], want_nb[ok]
    fo = np.asarray(face_offsets, np.int64)
    assert np.all(np.diff(fo)[ids] == 14), "interior BCC cells have 14 faces"
    v = np.asarray(volumes)[ids]
    assert np.max(np.abs(v - a ** 3 / 2) / (a ** 3 / 2)) <= RTOL
    sq, hx = a * a / 8, 3 * np.sqrt(3.0) * a * a / 16
    for row, c in enumerate(ids):
        nb = np.asarray(neighbors[fo[c]:fo[c + 1]], np.int64)
        ar = np.asarray(areas[fo[c]:fo[c + 1]])
        order = np.argsort(nb)
        assert np.array_equal(nb[order], want_nb[row]), c
        # second-shell neighbours carry the same sub-lattice parity as the cell: squares; the others hexagons
        same = (nb % 2) == (c % 2)
        assert same.sum() == 6
        assert np.max(np.abs(ar[same] - sq) / sq) <= RTOL, c
        assert np.max(np.abs(ar[~same] - hx) / hx) <= RTOL, c
    return len(ids)


def check_simple_cubic(volumes, face_offsets, neighbors, areas, status, m):
    a = 1.0 / m
    fo = np.asarray(face_offsets, np.int64)
    status = np.asarray(status)
    ok = status == 0
    assert ok.mean() > 0.5
    n = 0
    for i in range(1, m - 1):
        for j in range(1, m - 1):
            for k in range(1, m - 1):
                c = (i * m + j) * m + k
                if not ok[c]:
                    continue
                nb = np.sort(np.asarray(neighbors[fo[c]:fo[c + 1]], np.int64))
                want = np.sort([c + m * m, c - m * m, c + m, c - m, c + 1, c - 1])
                assert np.array_equal(nb, want), c
                assert abs(volumes[c] - a ** 3) / a ** 3 <= RTOL
                assert np.max(np.abs(np.asarray(areas[fo[c]:fo[c + 1]]) - a * a) / (a * a)) <= RTOL
                n += 1
    # cells on the container's faces: a wall takes the place of the missing neighbour, the cell is still the cube
    assert np.max(np.abs(np.asarray(volumes)[ok] - a ** 3) / a ** 3) <= RTOL
    assert np.all(np.diff(fo)[ok] == 6)
    return n


@pytest.mark.parametrize("m", [8, 12])
def test_oracle_bcc_cells_are_truncated_octahedra(ob, gen, m):
    pts = gen.bcc(m, 5, jitter=0.0)
    r = ob.Diagram(pts, box=BOX).compute_cells(mode=ob.MODE_SECURITY)
    assert check_bcc(r.volumes, r.face_offsets, r.neighbors, r.areas, r.status, m) > 1.4 * (m - 4) ** 3


@pytest.mark.parametrize("m", [6, 9])
def test_oracle_simple_cubic_cells_are_cubes(ob, gen, m):
    pts = gen.simple_cubic(m)
    r = ob.Diagram(pts, box=BOX).compute_cells(mode=ob.MODE_SECURITY)
    assert check_simple_cubic(r.volumes, r.face_offsets, r.neighbors, r.areas, r.status, m) > 0.3 * (m - 2) ** 3
    assert r.counters["degenerate_skips"] > 0  # D17 is reproduced, and flagged


def check_line(volumes, face_offsets, neighbors, areas, x):
    """Points (x[i], 0.5, 0.5) in the unit cube, x ascending."""
    n = len(x)
    fo = np.asarray(face_offsets, np.int64)
    mid = np.concatenate([[0.0], (x[1:] + x[:-1]) / 2, [1.0]])
    want = mid[1:] - mid[:-1]
    assert np.max(np.abs(np.asarray(volumes) - want) / want) <= RTOL
    for c in range(n):
        nb = np.asarray(neighbors[fo[c]:fo[c + 1]], np.int64)
        ar = np.asarray(areas[fo[c]:fo[c + 1]])
        assert len(nb) == 6
        assert sorted(nb[nb >= 0].tolist()) == [i for i in (c - 1, c + 1) if 0 <= i < n]
        assert np.max(np.abs(ar[nb >= 0] - 1.0)) <= RTOL  # the mid-planes span the whole cross-section
        side = ar[(nb == -1) | (nb == -3) | (nb == -5) | (nb == -6)]  # y_min, y_max, z_max, z_min walls: width x 1
        assert np.max(np.abs(side - want[c]) / want[c]) <= RTOL


def line_points(gen, n=40):
    x = np.sort(0.02 + 0.96 * gen.uniform(n, 91)[:, 0])
    return np.stack([x, np.full(n, 0.5), np.full(n, 0.5)], axis=1), x


def test_oracle_cells_of_points_on_a_line_are_slabs(ob, gen):
    pts, x = line_points(gen)
    r = ob.Diagram(pts, box=BOX).compute_cells(mode=ob.MODE_SECURITY)
    assert np.all(r.status == 0)
    check_line(r.volumes, r.face_offsets, r.neighbors, r.areas, x)


def test_oracle_jittered_bcc_stays_close_to_the_analytic_cell(ob, gen):
    """Config 5's input: jitter 1e-3*a moves every bisector by O(1e-3*a), so volumes stay within ~1e-2 relative of a^3/2
    and the 14-face topology holds (the squares have area a^2/8 >> jitter^2)."""
    m = 10
    pts = gen.bcc(m, 5)
    r = ob.Diagram(pts, box=BOX).compute_cells(mode=ob.MODE_SECURITY)
    ids, want_nb, a = bcc_expected(m)
    fo = r.face_offsets.astype(np.int64)
    assert np.all(np.diff(fo)[ids] == 14)
    assert np.max(np.abs(r.volumes[ids] - a ** 3 / 2) / (a ** 3 / 2)) < 2e-2
    for row, c in enumerate(ids[::7]):
        assert np.array_equal(np.sort(r.neighbors[fo[c]:fo[c + 1]]), want_nb[::7][row])
