"""The binning pass (the-tessellator_b200/csrc/grid.cu, K1-K4) and the radius queries (query.cu), unchanged, run
thread by thread on the CPU warp emulator (tests/emu/) and compared with the oracle's grid arrays and query
results — the CPU-only counterpart of test_gpu_parity.py::test_grid_matches_oracle and test_gpu_queries.py.
Blocks run on several host threads (so the look-back scan really waits on its predecessors, and arrival ranks
inside a grid cell really vary), lanes forward and backward."""
import numpy as np
import pytest

BOX = (0, 0, 0, 1, 1, 1)


@pytest.fixture(scope="module")
def eb():
    import emu_binding

    return emu_binding


@pytest.mark.parametrize("case", ["n1", "n2", "n10", "n1000", "n30k", "oblong", "clustered", "bcc", "flat"])
@pytest.mark.parametrize("reverse", [False, True])
def test_binning_matches_oracle(eb, gen, ob, case, reverse):
    pts = {
        "n1": lambda: gen.uniform(1, 11), "n2": lambda: gen.uniform(2, 12), "n10": lambda: gen.uniform(10, 13),
        "n1000": lambda: gen.uniform(1000, 14), "n30k": lambda: gen.uniform(30_000, 15),
        "oblong": lambda: gen.uniform(1500, 9) * np.array([1.0, 7.0, 0.1]) + np.array([3.0, -2.0, 0.0]),
        "clustered": lambda: gen.clustered(8000, 4, k=4), "bcc": lambda: gen.bcc(9, 5),
        "flat": lambda: gen.uniform(300, 16) * np.array([1.0, 1.0, 0.0]) + np.array([0.0, 0.0, 0.5]),
    }[case]()
    od = ob.Diagram(pts, table_radius=1)
    g = eb.binning(pts, od, reverse=reverse)
    assert g["oob"] == 0
    assert np.array_equal(g["bounds"], od.bounds())                      # K1: CeleryBounds::new
    assert np.array_equal(g["cell_of"].astype(np.uint64), od.cells())     # K2: get_cells
    assert np.array_equal(g["delim"].astype(np.uint64), od.delimiters())  # K3: get_delimiters
    si = od.sorted_indices()
    assert np.array_equal(g["sorted_idx"].astype(np.uint64), si)          # K4: get_sorted_indices (canonical in-cell order)
    assert np.array_equal(g["sorted"][:, :3], pts[si.astype(np.int64)])
    assert np.array_equal(g["sorted"][:, 3].view(np.int64), si.astype(np.int64))
    cpd = od.cpd
    assert np.array_equal(g["plane_counts"], np.bincount((od.cells() // (cpd * cpd)).astype(np.int64), minlength=cpd).astype(np.uint64))


def test_binning_with_explicit_ids_and_groups(eb, gen, ob):
    """Slab diagrams pass user-visible ids: records carry the id, the in-cell order follows it, sorted_idx keeps the insertion index."""
    pts = gen.uniform(5000, 17)
    perm = np.random.default_rng(3).permutation(5000)
    ids = (perm * 3 + 7).astype(np.int64)  # arbitrary distinct ids
    groups = (np.arange(5000) % 5).astype(np.uint64)
    od = ob.Diagram(pts, table_radius=1)
    g = eb.binning(pts, od, ids=ids, groups=groups)
    cells = od.cells().astype(np.int64)
    order = np.lexsort((ids, cells))  # by grid cell, then by id
    assert np.array_equal(g["sorted_idx"].astype(np.int64), order)
    assert np.array_equal(g["sorted"][:, 3].view(np.int64), ids[order])
    assert np.array_equal(g["sorted"][:, :3], pts[order])
    assert np.array_equal(g["groups_sorted"], groups[order])


def test_binning_flags_particles_outside_the_local_planes(eb, gen, ob):
    pts = gen.uniform(3000, 18)
    od = ob.Diagram(pts, table_radius=1)
    assert eb.binning(pts, od, local=(2, od.cpd - 2))["oob"] == 1


def _grid4(gen, seed, extra):
    pts = -2.0 + gen.uniform(79 - len(extra), seed) * 4.0
    return np.concatenate([pts, np.array(extra, float)])


def test_reference_query_tests(eb, gen):
    """celery.rs:1756-1901 and :1459-1487, on the emulated kernel."""
    extra = [(2, 2, 2), (-2, -2, -2), (0.1, 0.1, 0.1), (-0.3, -0.3, -0.3), (-0.7, -0.7, -0.7), (-1.1, -1.1, -1.1), (1.3, 1.3, 1.3), (1.7, 1.7, 1.7)]
    g = eb.EmuGrid(_grid4(gen, 30, extra), (-2, -2, -2, 2, 2, 2), table_radius=-1)
    assert g.cpd == 4
    nb = eb.radius_query(g, [[0.5, 0.5, 0.5]], 1.73, 0)[0][0]
    assert [nb.count(i) for i in (73, 74, 75, 76, 77, 78)] == [1, 1, 1, 0, 1, 1]
    extra = [(2, 2, 2), (-2, -2, -2), (0.1, 0.1, 0.1), (-0.1, -0.1, -0.1), (-0.7, -0.7, -0.7), (-1.1, -1.1, -1.1), (1.3, 1.3, 1.3), (1.7, 1.7, 1.7)]
    g = eb.EmuGrid(_grid4(gen, 31, extra), (-2, -2, -2, 2, 2, 2), table_radius=-1)
    nb = eb.radius_query(g, [[0.5, 0.5, 0.5]], 1.73, 1)[0][0]
    assert [nb.count(i) for i in (73, 74, 75, 76, 77, 78)] == [1, 1, 0, 0, 1, 0]
    g = eb.EmuGrid(_grid4(gen, 27, [(2, 2, 2), (-2, -2, -2)]), (-2, -2, -2, 2, 2, 2), table_radius=-1)
    r = eb.radius_query(g, [[0.0, 0.0, 0.0]], 0.5, 2)[0][0]
    assert 77 in r and 78 not in r and len(r) < 79
    r = eb.radius_query(g, [[0.0, 0.0, 0.0]], 10.0, 2)[0][0]
    assert 77 in r and 78 in r and len(r) == 79


@pytest.mark.parametrize("reverse", [False, True])
def test_queries_match_oracle_in_order(eb, gen, reverse):
    scale = np.array([1.0, 1.3, 0.8])
    pts = gen.uniform(4000, 91) * scale
    groups = (np.arange(4000) % 4).astype(np.uint64)
    g = eb.EmuGrid(pts, (0, 0, 0, 1, 1.3, 0.8), groups=groups, table_radius=-1)
    od = g.oracle
    sx = g.cell_info[0]
    qs = np.concatenate([gen.uniform(30, 92) * scale, pts[:10], [[0.0, 0.0, 0.0], [2.0, 2.0, 2.0], [-1.0, 0.5, 0.4]]])
    for radius in (0.0, 0.4 * sx, 1.5 * sx, 3.7 * sx):
        cell = eb.radius_query(g, qs, radius, 0, reverse=reverse)[0]
        real = eb.radius_query(g, qs, radius, 1, reverse=reverse)[0]
        for i, q in enumerate(qs):
            assert cell[i] == od.find_neighbors_in_cell_radius(*q, radius)
            assert real[i] == od.find_neighbors_in_real_radius(*q, radius)
    for max_radius in (0.0, (1.2 * sx) ** 2, (3.1 * sx) ** 2):
        cloud = eb.radius_query(g, qs, max_radius, 2, reverse=reverse)[0]
        only3 = eb.radius_query(g, qs, max_radius, 2, target_group=3, reverse=reverse)[0]
        for i, q in enumerate(qs):
            exp = od.expanding_search(*q).expand_all_in_radius(max_radius)
            assert cloud[i] == exp
            assert only3[i] == [k for k in exp if groups[k] == 3]  # interface.rs:359-362
