"""GPU radius / neighbour-cloud queries against the reference's own unit tests
(celery.rs:1459-1487, 1756-1901) and, on random inputs, against the oracle — ordered lists must be equal."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _grid4(gen, seed, extra):
    n_rand = 79 - len(extra)
    pts = -2.0 + gen.uniform(n_rand, seed) * 4.0
    return np.concatenate([pts, np.array(extra, float)])


def _diagram(tess, pts, groups=None):
    d = tess.Diagram(0)
    d.add_particles(pts, groups)
    d.initialize(None)
    return d


def test_find_neighbors_in_cell_radius(tess, gen):  # celery.rs:1756-1827
    extra = [(2, 2, 2), (-2, -2, -2), (0.1, 0.1, 0.1), (-0.3, -0.3, -0.3), (-0.7, -0.7, -0.7), (-1.1, -1.1, -1.1), (1.3, 1.3, 1.3), (1.7, 1.7, 1.7)]
    d = _diagram(tess, _grid4(gen, 30, extra))
    assert d.grid_info()["cells_per_dimension"] == 4
    nb = d.find_neighbors_in_cell_radius(0.5, 0.5, 0.5, 1.73)
    assert [nb.count(i) for i in (73, 74, 75, 76, 77, 78)] == [1, 1, 1, 0, 1, 1]
    d.close()


def test_find_neighbors_in_real_radius(tess, gen):  # celery.rs:1830-1901
    extra = [(2, 2, 2), (-2, -2, -2), (0.1, 0.1, 0.1), (-0.1, -0.1, -0.1), (-0.7, -0.7, -0.7), (-1.1, -1.1, -1.1), (1.3, 1.3, 1.3), (1.7, 1.7, 1.7)]
    d = _diagram(tess, _grid4(gen, 31, extra))
    nb = d.find_neighbors_in_real_radius(0.5, 0.5, 0.5, 1.73)
    assert [nb.count(i) for i in (73, 74, 75, 76, 77, 78)] == [1, 1, 0, 0, 1, 0]
    d.close()


def test_expand_all_in_radius(tess, gen):  # celery.rs:1459-1487
    d = _diagram(tess, _grid4(gen, 27, [(2, 2, 2), (-2, -2, -2)]))
    r = d.find_neighbors(np.array([[0.0, 0.0, 0.0]]), 0.5, tess._lib.QUERY_NEIGHBOR_CLOUD)[0].tolist()
    assert 77 in r and 78 not in r and len(r) < 79
    r = d.find_neighbors(np.array([[0.0, 0.0, 0.0]]), 10.0, tess._lib.QUERY_NEIGHBOR_CLOUD)[0].tolist()
    assert 77 in r and 78 in r and len(r) == 79
    d.close()


@pytest.mark.parametrize("n", [100, 20_000])
def test_queries_match_oracle_in_order(tess, gen, ob, n):
    pts = gen.uniform(n, 91) * np.array([1.0, 1.3, 0.8])
    d = _diagram(tess, pts)
    od = ob.Diagram(pts)
    sx = od.cell_info()[0]
    qs = np.concatenate([gen.uniform(60, 92) * np.array([1.0, 1.3, 0.8]), pts[:20], [[0.0, 0.0, 0.0], [2.0, 2.0, 2.0], [-1.0, 0.5, 0.4]]])
    for radius in (0.0, 0.4 * sx, 1.5 * sx, 3.7 * sx):
        cell = d.find_neighbors(qs, radius, tess._lib.QUERY_CELL_RADIUS)
        real = d.find_neighbors(qs, radius, tess._lib.QUERY_REAL_RADIUS)
        for i, q in enumerate(qs):
            assert cell[i].tolist() == od.find_neighbors_in_cell_radius(*q, radius)
            assert real[i].tolist() == od.find_neighbors_in_real_radius(*q, radius)
    for max_radius in (0.0, (1.2 * sx) ** 2, (3.1 * sx) ** 2):
        cloud = d.find_neighbors(qs, max_radius, tess._lib.QUERY_NEIGHBOR_CLOUD)
        for i, q in enumerate(qs):
            assert cloud[i].tolist() == od.expanding_search(*q).expand_all_in_radius(max_radius)
    d.close()


def test_compute_neighbor_cloud_with_groups(tess, gen, ob):  # interface.rs:348-365
    pts = gen.uniform(5000, 93)
    groups = (np.arange(5000) % 4).astype(np.uint64)
    d = tess.Diagram(0)
    d.add_particles(pts, groups)
    box = tess.Polyhedron(0, 0, 0, 1, 1, 1)
    d.initialize(box)
    od = ob.Diagram(pts, box=[0, 0, 0, 1, 1, 1], groups=groups)
    sx = od.cell_info()[0]
    for i in (0, 17, 4999):
        cell = d.get_cell_at_index(i, box)
        exp = od.expanding_search(*pts[i]).expand_all_in_radius((2 * sx) ** 2)
        assert cell.compute_neighbor_cloud((2 * sx) ** 2) == exp
        assert cell.compute_neighbor_cloud((2 * sx) ** 2, target_group=3) == [k for k in exp if groups[k] == 3]
    d.close()


def _pts(gen, n, seed, lo=(0, 0, 0), hi=(1, 1, 1)):
    lo, hi = np.array(lo, float), np.array(hi, float)
    return lo + gen.uniform(n, seed) * (hi - lo)


@pytest.mark.parametrize("query", [(0.5, 0.5, 0.5), (0.0, 0.0, 0.0)])
def test_expanding_search_single_steps(tess, gen, query):  # celery.rs:1343-1404
    d = _diagram(tess, _pts(gen, 100, 24))
    assert d.grid_info()["cells_per_dimension"] == 5
    es = d.expanding_search(query)
    allr = []
    for step in range(729):
        allr += es.expand(10.0, 1)
        assert es.current_search_index == step + 1
    assert len(allr) == 100 and len(set(allr)) == 100
    assert es.expand(10.0, 1) == [] and es.expand(10.0, 50) == []
    assert es.current_search_index == 729  # the whole (2*5-1)^3 table
    d.close()


@pytest.mark.parametrize("cells_to_add", [729, 1000])
def test_expanding_search_all_at_once(tess, gen, cells_to_add):  # celery.rs:1407-1444
    d = _diagram(tess, _pts(gen, 100, 25))
    assert len(d.expanding_search((0.0, 0.0, 0.0)).expand(10.0, cells_to_add)) == 100
    assert len(d.expanding_search((0.0, 0.0, 0.0)).expand_all_no_radius()) == 100  # celery.rs:1447-1456
    d.close()


def test_find_cells_in_radius(tess, gen):  # celery.rs:1665-1753
    d = _diagram(tess, _grid4(gen, 29, [(2, 2, 2), (-2, -2, -2)]))
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
    d.close()


def test_cursor_queries_match_oracle_in_order(tess, gen, ob):
    """expand in uneven steps (the table is widened on demand under the cursor) and find_cells_in_radius, many positions at
    once, against the oracle: same lists in the same order, same cursors."""
    pts = gen.uniform(20_000, 41)
    d = _diagram(tess, pts)
    od = ob.Diagram(pts)
    q = np.concatenate([gen.uniform(12, 42), [[0.0, 0.0, 0.0], [1.0, 1.0, 1.0]]])
    sx = od.cell_info()[0]
    es = d.expanding_search(q)
    oes = [od.expanding_search(*p) for p in q]
    for radius, cells in (((1.5 * sx) ** 2, 7), ((1.5 * sx) ** 2, 1000), ((6 * sx) ** 2, 300), ((12 * sx) ** 2, 5000), (float("inf"), 20_000)):
        got = es.expand(radius, cells)
        for i in range(len(q)):
            assert got[i].tolist() == oes[i].expand(radius, cells), (radius, cells, i)
    for r in (0.0, 0.7 * sx, 3.3 * sx):
        for p in q[:6]:
            assert d.find_cells_in_radius(*p, r) == od.find_cells_in_radius(*p, r)
    d.close()
