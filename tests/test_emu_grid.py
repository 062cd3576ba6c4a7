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


# ------------------------------------------------------------------ outputs.cu ---------------
ST_LARGE_PATH = 1 << 31


@pytest.mark.parametrize("reverse", [False, True])
def test_csr_packing_of_staged_rows(eb, gen, ob, reverse):
    """compact_faces + compact_redo (interface.rs:342-384 in batch form): fixed-stride staging rows, some of them
    recomputed by a larger configuration into their own staging, become the CSR arrays — here the oracle's cells
    are staged and must come back unchanged."""
    pts = gen.uniform(3000, 33)
    r = ob.Diagram(pts, box=list(BOX)).compute_cells(mode=ob.MODE_SECURITY)
    fo = r.face_offsets
    n, fstride, rstride = r.n, 40, 128
    nfaces = np.diff(fo).astype(np.uint32)
    status = np.zeros(n, np.uint32)
    redo_rows = np.sort(np.random.default_rng(5).choice(n, 200, replace=False)).astype(np.uint32)
    status[redo_rows] |= ST_LARGE_PATH
    st_nbr, st_area, st_flen = np.full(n * fstride, -7, np.int64), np.full(n * fstride, -7.0), np.zeros(n * fstride, np.uint16)
    for c in range(n):
        if not status[c] & ST_LARGE_PATH:
            k = int(nfaces[c])
            st_nbr[c * fstride: c * fstride + k] = r.cell_neighbors(c)
            st_area[c * fstride: c * fstride + k] = r.cell_areas(c)
            st_flen[c * fstride: c * fstride + k] = np.arange(k) + 3
    rd_nbr, rd_area, rd_flen = np.full(len(redo_rows) * rstride, -8, np.int64), np.full(len(redo_rows) * rstride, -8.0), np.zeros(len(redo_rows) * rstride, np.uint16)
    for w, c in enumerate(redo_rows):
        k = int(nfaces[c])
        rd_nbr[w * rstride: w * rstride + k] = r.cell_neighbors(c)
        rd_area[w * rstride: w * rstride + k] = r.cell_areas(c)
        rd_flen[w * rstride: w * rstride + k] = np.arange(k) + 3
    offsets, nbr, area, flen = eb.pack_faces(status, nfaces, st_nbr, st_area, st_flen, fstride, redo_rows, rd_nbr, rd_area, rd_flen, rstride, reverse=reverse)
    assert np.array_equal(offsets, fo)
    assert np.array_equal(nbr, r.neighbors) and np.array_equal(area, r.areas)
    assert np.array_equal(flen, np.concatenate([np.arange(k) + 3 for k in nfaces]))


def test_chunk_work_lists_volume_sum_and_gathers(eb, gen):
    L = eb.outputs_lib()
    rng = np.random.default_rng(7)
    # streaming: rows [lo, hi) as ascending sorted slots
    n = 5000
    row_of_slot = rng.permutation(n).astype(np.uint32)
    work = np.zeros(n, np.uint32)
    for lo, hi in ((0, 700), (700, 4100), (4100, 5000), (10, 10)):
        m = L.emu_chunk_list(row_of_slot.ctypes.data, n, lo, hi, work.ctypes.data, 4)
        exp = np.nonzero((row_of_slot >= lo) & (row_of_slot < hi))[0]
        assert m == len(exp) and np.array_equal(work[:m], exp)
    # volume closure sum: deterministic whatever the schedule, equal to the exact sum within rounding
    vol = rng.random(100_000) / 100_000
    sums = {L.emu_volume_sum(vol.ctypes.data, len(vol), t, rv) for t in (1, 4) for rv in (0, 1)}
    assert len(sums) == 1
    import math
    assert abs(sums.pop() - math.fsum(vol)) < 1e-13
    # geometry gathers: per-row blocks of the bump-allocated pools -> CSR order
    rows = 300
    nv = rng.integers(4, 30, rows).astype(np.uint32)
    nf = rng.integers(4, 12, rows).astype(np.uint32)
    flen = [rng.integers(3, 8, k) for k in nf]
    nl = np.array([int(f.sum()) for f in flen], np.uint32)
    order = rng.permutation(rows)  # cells finished in another order than their rows
    vbase, lbase = np.zeros(rows, np.uint64), np.zeros(rows, np.uint64)
    cv = cl = 0
    for c in order:
        vbase[c], lbase[c] = cv, cl
        cv += int(nv[c]); cl += int(nl[c])
    vpool, lpool = rng.random((cv, 3)), rng.integers(0, 30, cl).astype(np.uint32)
    voff = np.concatenate([[0], np.cumsum(nv)]).astype(np.uint64)
    face_off = np.concatenate([[0], np.cumsum(nf)]).astype(np.uint64)
    fv_off = np.concatenate([[0], np.cumsum(np.concatenate(flen))]).astype(np.uint64)
    vtx, loops = np.zeros((cv, 3)), np.zeros(cl, np.uint32)
    L.emu_gather(nv.ctypes.data, vbase.ctypes.data, voff.ctypes.data, vpool.ctypes.data, nl.ctypes.data, lbase.ctypes.data, face_off.ctypes.data,
                 fv_off.ctypes.data, lpool.ctypes.data, rows, vtx.ctypes.data, loops.ctypes.data, 4)
    for c in range(rows):
        assert np.array_equal(vtx[int(voff[c]): int(voff[c + 1])], vpool[int(vbase[c]): int(vbase[c]) + int(nv[c])])
        o = int(fv_off[int(face_off[c])])
        assert np.array_equal(loops[o: o + int(nl[c])], lpool[int(lbase[c]): int(lbase[c]) + int(nl[c])])
    st = np.full(1000, 0x80000005, np.uint32)
    L.emu_clear_status_bits(st.ctypes.data, 1000, 0x80000000)
    assert np.all(st == 5)


@pytest.mark.parametrize("reverse", [False, True])
def test_expand_cursor_and_cells_in_radius_on_the_emulator(eb, gen, reverse):
    """query.cu modes 3 (Celery::find_cells_in_radius, celery.rs:753-797) and 4 (ExpandingSearch::expand with a cursor,
    celery.rs:907-963) on the emulated kernel: uneven steps from saved cursors, a truncated table that runs out under the
    cursor (flagged, cursor left at the table's end), all against the oracle — lists, order and cursors."""
    pts = gen.uniform(3000, 93)
    g = eb.EmuGrid(pts, (0, 0, 0, 1, 1, 1), table_radius=-1)
    od = g.oracle
    sx = g.cell_info[0]
    qs = np.concatenate([gen.uniform(10, 94), [[0.0, 0.0, 0.0], [1.0, 1.0, 1.0]]])
    for r in (0.0, 0.7 * sx, 2.6 * sx):
        cells = eb.radius_query(g, qs, r, 3, reverse=reverse)[0]
        for i, q in enumerate(qs):
            assert cells[i] == od.find_cells_in_radius(*q, r)
    oes = [od.expanding_search(*q) for q in qs]
    cur = np.zeros(len(qs), np.uint64)
    for radius, cells_to_add in (((1.5 * sx) ** 2, 1), ((1.5 * sx) ** 2, 5), ((1.5 * sx) ** 2, 1000), ((4 * sx) ** 2, 37), (float("inf"), 400)):
        got, flags, cur = eb.radius_query(g, qs, radius, 4, cursors=cur, cells_to_add=cells_to_add, reverse=reverse)
        assert not flags.any()
        for i in range(len(qs)):
            assert got[i] == oes[i].expand(radius, cells_to_add), (radius, cells_to_add, i)
    # the celery.rs:1343-1404 walk, one table entry per call
    g2 = eb.EmuGrid(gen.uniform(100, 24), (0, 0, 0, 1, 1, 1), table_radius=-1)
    assert g2.cpd == 5 and g2.table_key.size == 729
    cur, seen = np.zeros(1, np.uint64), []
    for step in range(729):
        got, _, cur = eb.radius_query(g2, [[0.5, 0.5, 0.5]], 10.0, 4, cursors=cur, cells_to_add=1, reverse=reverse)
        seen += got[0]
        assert int(cur[0]) == step + 1
    assert len(seen) == 100 and len(set(seen)) == 100
    assert eb.radius_query(g2, [[0.5, 0.5, 0.5]], 10.0, 4, cursors=cur, cells_to_add=50)[0][0] == []
    # a truncated table under the cursor
    g3 = eb.EmuGrid(pts, (0, 0, 0, 1, 1, 1), table_radius=2)
    got, flags, cur = eb.radius_query(g3, qs[:3], float("inf"), 4, cursors=np.zeros(3, np.uint64), cells_to_add=10 ** 6, reverse=reverse)
    assert np.all(flags == 2) and np.all(cur == g3.table_key.size)  # TABLE_EXHAUSTED: the host widens the table and asks again
    for i in range(3):
        assert got[i] == od.expanding_search(*qs[i]).expand(float("inf"), g3.table_key.size)
