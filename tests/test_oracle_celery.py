"""Pins the oracle's grid (celery.rs) against the reference's own unit tests (celery.rs:1214-1901).

The reference draws unseeded random points and asserts structure only; here the same assertions
run on seeded points of the same shapes.
"""
import numpy as np
import pytest


def _pts(gen, n, seed, lo=(0, 0, 0), hi=(1, 1, 1)):
    lo, hi = np.array(lo, float), np.array(hi, float)
    return lo + gen.uniform(n, seed) * (hi - lo)


def check_delimiters(cells, sorted_indices, delimiters):
    """celery.rs:1131-1172, assertion for assertion."""
    n = len(sorted_indices)
    assert delimiters[0] == 0
    assert delimiters[-1] == n
    assert len(set(sorted_indices.tolist())) == n
    for i in range(1, len(delimiters)):
        assert delimiters[i] == 0 or i > cells[sorted_indices[delimiters[i] - 1]]
        if delimiters[i] == n:
            assert all(delimiters[j] == n for j in range(i + 1, len(delimiters)))
            return
        c = cells[sorted_indices[delimiters[i]]]
        assert (i == c) or (i > c) or (delimiters[i] == delimiters[i + 1])


BOXES = {
    "unit": ((0, 0, 0), (1, 1, 1)),
    "large": ((-1000, -1000, -1000), (500, 500, 500)),
    "oblong": ((-5, 12, -10000), (-4, 1000, 1)),
}


@pytest.mark.parametrize("box", list(BOXES))
def test_insert_one_point(ob, gen, box):  # celery.rs:1226-1262
    lo, hi = BOXES[box]
    d = ob.Diagram(_pts(gen, 1, 21, lo, hi))
    assert d.cpd == 1 and d.search_order_len == 1 and d.num_delimiters == 2
    check_delimiters(d.cells(), d.sorted_indices(), d.delimiters())


@pytest.mark.parametrize("box", list(BOXES))
def test_insert_one_thousand_points(ob, gen, box):  # celery.rs:1265-1301
    lo, hi = BOXES[box]
    d = ob.Diagram(_pts(gen, 1000, 22, lo, hi))
    assert d.cpd == 10 and d.search_order_len == 6859 and d.num_delimiters == 1001
    check_delimiters(d.cells(), d.sorted_indices(), d.delimiters())


def test_one_million_grid_shape(ob):
    """celery.rs:1303-1314 (commented out there for run time): cpd 93, 804,358 delimiters,
    6,331,625 table entries.  Checked arithmetically (the table itself is 152 MB)."""
    cpd = int(ob.lib().orc_cells_per_dimension(1_000_000))
    assert cpd == 93 and cpd ** 3 + 1 == 804_358 and (2 * cpd - 1) ** 3 == 6_331_625


def test_create_celery(ob):  # celery.rs:1215-1223
    d = ob.Diagram(np.array([[1.2, 3.4, 8.3], [4.2, 7.3, 2.7], [0.3, 1.7, 9.0]]))
    assert d.cpd == 2
    check_delimiters(d.cells(), d.sorted_indices(), d.delimiters())


def test_search_order_is_sorted_with_home_first(ob, gen):  # celery.rs:437-442, 676
    d = ob.Diagram(_pts(gen, 1000, 23))
    dist, ijk = d.search_order()
    assert dist[0] == -1.0 and tuple(ijk[0]) == (0, 0, 0)
    assert np.all(np.diff(dist) >= 0)
    # every offset of [-(cpd-1), cpd-1]^3 exactly once (celery.rs:445-673)
    assert len({tuple(r) for r in ijk}) == 19 ** 3
    # distance of an offset = squared length of ((|o|-1)+ * size) (celery.rs:423-427, 450-457)
    sx, sy, sz = d.cell_info()[:3]
    k = 1234
    t = np.maximum(np.abs(ijk[k]) - 1, 0)
    assert dist[k] == (t[0] * sx) * (t[0] * sx) + (t[1] * sy) * (t[1] * sy) + (t[2] * sz) * (t[2] * sz)


@pytest.mark.parametrize("query,home", [((0.5, 0.5, 0.5), (2, 2, 2)), ((0.0, 0.0, 0.0), (0, 0, 0))])
def test_expanding_search_single_steps(ob, gen, query, home):  # celery.rs:1343-1404
    d = ob.Diagram(_pts(gen, 100, 24))
    assert d.cpd == 5 and d.num_delimiters == 126 and d.search_order_len == 729
    es = d.expanding_search(*query)
    assert es.home() == home
    allr = []
    for _ in range(729):
        allr += es.expand(10.0, 1)
    assert len(allr) == 100 and len(set(allr)) == 100
    assert es.expand(10.0, 1) == []
    assert es.expand(10.0, 50) == []


@pytest.mark.parametrize("cells_to_add", [729, 1000])
def test_expanding_search_all_at_once(ob, gen, cells_to_add):  # celery.rs:1407-1444
    d = ob.Diagram(_pts(gen, 100, 25))
    es = d.expanding_search(0.0, 0.0, 0.0)
    assert es.home() == (0, 0, 0)
    assert len(es.expand(10.0, cells_to_add)) == 100


def test_expand_all_no_radius(ob, gen):  # celery.rs:1447-1456
    d = ob.Diagram(_pts(gen, 100, 26))
    assert len(d.expanding_search(0.0, 0.0, 0.0).expand_all_no_radius()) == 100


def _grid4(gen, seed, extra):
    """79 points on [-2,2]^3 -> 4x4x4 grid of unit cells (celery.rs:1460-1471)."""
    n_rand = 79 - len(extra)
    pts = _pts(gen, n_rand, seed, (-2, -2, -2), (2, 2, 2))
    return np.concatenate([pts, np.array(extra, float)])


def test_expand_all_in_radius(ob, gen):  # celery.rs:1459-1487
    d = ob.Diagram(_grid4(gen, 27, [(2, 2, 2), (-2, -2, -2)]))
    assert d.cpd == 4
    r = d.expanding_search(0.0, 0.0, 0.0).expand_all_in_radius(0.5)
    assert 77 in r and 78 not in r and len(r) < 79
    r = d.expanding_search(0.0, 0.0, 0.0).expand_all_in_radius(10.0)
    assert 77 in r and 78 in r and len(r) == 79


def test_check_cell_in_range(ob, gen):  # celery.rs:1490-1662
    d = ob.Diagram(_grid4(gen, 28, [(2, 2, 2), (-2, -2, -2)]))
    for i in range(4):
        for j in range(4):
            for k in range(4):
                for q in (0.1, 0.9):
                    c = lambda r: d.check_cell_in_range(q, q, q, r, i, j, k)  # noqa: E731
                    assert c(0.5) == (i != 0 and j != 0 and k != 0)
                    medium = (i == 0 and j >= 1 and k >= 1) or (i >= 1 and j == 0 and k >= 1) or (i >= 1 and j >= 1 and k == 0) or (i >= 1 and j >= 1 and k >= 1)
                    assert c(1.1) == medium
                    assert c(1.7) == (i + j + k > 0)
                    assert c(1.8)


def test_find_cells_in_radius(ob, gen):  # celery.rs:1665-1753
    d = ob.Diagram(_grid4(gen, 29, [(2, 2, 2), (-2, -2, -2)]))
    cell = lambda i, j, k: 16 * i + 4 * j + k  # noqa: E731
    f = lambda q, r: set(d.find_cells_in_radius(q, q, q, r))  # noqa: E731
    close1, close2 = f(0.1, 0.5), f(0.9, 0.5)
    med1, med2 = f(0.1, 1.2), f(0.9, 1.2)
    far1, far2 = f(0.1, 1.7), f(0.9, 1.7)
    all1, all2 = f(0.1, 1.8), f(0.9, 1.8)
    for i in range(4):
        for j in range(4):
            for k in range(4):
                c = cell(i, j, k)
                assert (c in close1) == (i in (1, 2) and j in (1, 2) and k in (1, 2))
                assert (c in close2) == (i in (2, 3) and j in (2, 3) and k in (2, 3))
                medium = (i == 0 and j >= 1 and k >= 1) or (i >= 1 and j == 0 and k >= 1) or (i >= 1 and j >= 1 and k == 0) or (i >= 1 and j >= 1 and k >= 1)
                assert (c in med1) == medium
                assert (c in med2) == (i >= 1 and j >= 1 and k >= 1)
                assert (c in far1) == (i + j + k > 0)
                assert (c in far2) == (i != 0 and j != 0 and k != 0)
                assert (c in all2) == (i != 0 and j != 0 and k != 0)
                assert c in all1


def test_find_neighbors_in_cell_radius(ob, gen):  # celery.rs:1756-1827
    extra = [(2, 2, 2), (-2, -2, -2), (0.1, 0.1, 0.1), (-0.3, -0.3, -0.3), (-0.7, -0.7, -0.7), (-1.1, -1.1, -1.1), (1.3, 1.3, 1.3), (1.7, 1.7, 1.7)]
    d = ob.Diagram(_grid4(gen, 30, extra))
    nb = d.find_neighbors_in_cell_radius(0.5, 0.5, 0.5, 1.73)
    assert [nb.count(i) for i in (73, 74, 75, 76, 77, 78)] == [1, 1, 1, 0, 1, 1]


def test_find_neighbors_in_real_radius(ob, gen):  # celery.rs:1830-1901
    extra = [(2, 2, 2), (-2, -2, -2), (0.1, 0.1, 0.1), (-0.1, -0.1, -0.1), (-0.7, -0.7, -0.7), (-1.1, -1.1, -1.1), (1.3, 1.3, 1.3), (1.7, 1.7, 1.7)]
    d = ob.Diagram(_grid4(gen, 31, extra))
    nb = d.find_neighbors_in_real_radius(0.5, 0.5, 0.5, 1.73)
    assert [nb.count(i) for i in (73, 74, 75, 76, 77, 78)] == [1, 1, 0, 0, 1, 0]


def test_truncated_table_is_a_prefix_of_the_full_table(ob, gen):
    """Not in the reference: the GPU and the large-N oracle walk a table cut to |offset| <= R.
    It must be an exact prefix of the full (2cpd-1)^3 table."""
    pts = _pts(gen, 5000, 32, (0, 0, 0), (1, 1.1, 0.9))
    full = ob.Diagram(pts)
    dist_f, ijk_f = full.search_order()
    for R in (1, 2, 3, 5, 8):
        t = ob.Diagram(pts, table_radius=R)
        dist_t, ijk_t = t.search_order()
        assert not t.table_is_full
        assert 1 < len(dist_t) <= (2 * R + 1) ** 3
        assert np.array_equal(dist_t, dist_f[: len(dist_t)])
        assert np.array_equal(ijk_t, ijk_f[: len(dist_t)])
    assert ob.Diagram(pts, table_radius=full.cpd - 1).table_is_full
