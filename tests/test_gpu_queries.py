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
