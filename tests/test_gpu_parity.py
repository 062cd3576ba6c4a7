"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle and the golden fixtures.

Bars (north_star / BASELINE.md §4): topology bit-exact, volumes / areas within 1e-12 relative,
|sum of volumes - container volume| <= 1e-12.  What is asserted here is stricter: against the live
oracle every output is compared BIT FOR BIT (helpers.assert_cells_identical: volumes, areas, and the
neighbour lists in face-slot order), as are grid arrays, status words and work counters; against
the golden fixtures (stored with faces sorted by neighbour id) the sorted lists must be equal and
the values within the 1e-12 bar (helpers.assert_cells_match) — they are in fact equal.
"""
import os

import numpy as np
import pytest

import helpers

pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
BOX = [0, 0, 0, 1, 1, 1]
CNAMES = ("visited", "tested", "vertex_classifications", "cuts", "new_vertices", "table_entries", "degenerate_skips", "faces")
ALL_OUT = 1 | 2 | 4 | 16


def _diagram(tess, pts, box=BOX, groups=None):
    d = tess.Diagram(0)
    d.add_particles(pts, groups)
    d.initialize(None if box is None else tess.Polyhedron(*box))
    return d


# The main clip pass has three kernels (tess_set_main_tier) and each has two instantiations (with / without the work
# counters, outputs & 16): bench.py times "no counters" on the default tier.  Every oracle comparison below runs on all
# of them: (tier, outputs).
VARIANTS = [("default", ALL_OUT), ("default", 7), ("small", ALL_OUT), ("small", 7), ("fast", ALL_OUT), ("fast", 7), ("thread", ALL_OUT), ("thread", 7)]


@pytest.fixture(params=VARIANTS, ids=lambda v: f"{v[0]}-{'count' if v[1] & 16 else 'nocount'}")
def variant(request, tess):
    tier, outputs = request.param
    tess.set_main_tier(tier)
    yield tier, outputs
    tess.set_main_tier("default")


def test_device_and_library(tess):
    assert tess.device_count() >= 1
    assert os.path.exists(tess._lib.LIB_PATH)


# ------------------------------------------------------------------ binning (K1-K4) ---------
@pytest.mark.parametrize("case", ["n1", "n2", "n10", "n1000", "n100k", "oblong", "clustered", "bcc", "flat"])
def test_grid_matches_oracle(tess, gen, ob, case):
    pts = {
        "n1": lambda: gen.uniform(1, 11),
        "n2": lambda: gen.uniform(2, 11),
        "n10": lambda: gen.uniform(10, 12),
        "n1000": lambda: gen.uniform(1000, 13),
        "n100k": lambda: gen.uniform(100_000, 14),
        "oblong": lambda: gen.uniform(5000, 15) * np.array([1.0, 988.0, 10001.0]) + np.array([-5.0, 12.0, -10000.0]),
        "clustered": lambda: gen.clustered(50_000, 4, k=8),
        "bcc": lambda: gen.bcc(12, 5),
        "flat": lambda: gen.uniform(3000, 16) * np.array([1.0, 1.0, 0.0]) + np.array([0.0, 0.0, 0.5]),  # zero z-extent
    }[case]()
    d = _diagram(tess, pts, box=None if case in ("oblong", "flat") else BOX)
    od = ob.Diagram(pts, box=None if case in ("oblong", "flat") else BOX, table_radius=8)
    gi = d.grid_info()
    assert gi["cells_per_dimension"] == od.cpd
    assert np.array_equal(gi["bounds"], od.bounds())
    assert np.array_equal(np.concatenate([gi["cell_sizes"], gi["inverse_cell_sizes"]]), od.cell_info(), equal_nan=True)
    cells, sidx, delim = d.copy_grid()
    assert np.array_equal(cells, od.cells())
    assert np.array_equal(sidx, od.sorted_indices())  # canonical in-cell order: ascending index
    assert np.array_equal(delim, od.delimiters())
    keys, ijk, full = d.search_order(8)
    okeys, oijk = od.search_order()
    assert full == od.table_is_full
    assert np.array_equal(keys, okeys) and np.array_equal(ijk, oijk)
    d.close()


# ------------------------------------------------------------------ cells vs golden ---------
FIXTURES = ["config1_uniform_10k_seed1.npz", "oblong_1500_seed9_bbox.npz", "clustered_4000_seed4.npz", "bcc_m8_seed5.npz", "simple_cubic_6.npz"]


def _fixture_points(name, gen, g):
    if name.startswith("config1"):
        return gen.uniform(10_000, 1), BOX
    if name.startswith("oblong"):
        return gen.uniform(1500, 9) * np.array([1.0, 2.0, 0.5]) + np.array([-3.0, 10.0, 0.25]), None
    if name.startswith("clustered"):
        return g["points"], BOX
    if name.startswith("bcc"):
        return gen.bcc(8, 5), BOX
    return gen.simple_cubic(6), BOX


class _Gold:
    def __init__(self, g):
        self.volumes, self.face_offsets, self.neighbors, self.areas = g["volumes"], g["face_offsets"], g["neighbors"].astype(np.int64), g["areas"]


@pytest.mark.parametrize("name", FIXTURES)
def test_cells_match_golden_fixture(tess, gen, name, variant):
    tier, outputs = variant
    g = np.load(os.path.join(GOLD, name))
    pts, box = _fixture_points(name, gen, g)
    d = _diagram(tess, pts, box)
    # the fixtures come from the oracle's FULL search table: give the GPU the full table too, so that
    # even the work counters must agree (the default R=8 table + redo pass is covered below)
    b = d.compute_all_cells(outputs=outputs, table_radius=1 << 20)
    helpers.assert_cells_match(b, _Gold(g), area_rtol=0.0, vol_rtol=0.0, what=name)  # rtol 0: equal values
    if outputs & 16:
        c = b.counters()
        assert [c[k] for k in CNAMES] == g["counters"].tolist()
    if "status" in g.files:
        assert np.array_equal(b.status, g["status"])  # degenerate skips flagged on the same cells
    else:
        assert np.all(b.status == 0)
    d.close()


# ------------------------------------------------------------------ cells vs live oracle ----
@pytest.mark.parametrize("case", ["uniform200k", "clustered100k", "bcc16k", "tiny"])
def test_cells_match_oracle(tess, gen, ob, case, variant):
    tier, outputs = variant
    pts = {
        "uniform200k": lambda: gen.uniform(200_000, 51),
        "clustered100k": lambda: gen.clustered(100_000, 4, k=8),
        "bcc16k": lambda: gen.bcc(20, 5),
        "tiny": lambda: gen.uniform(5, 52),
    }[case]()
    d = _diagram(tess, pts)
    b = d.compute_all_cells(outputs=outputs)
    if tier != "default":  # (the thread-per-cell kernel keeps no work counters: with counters the default kernel runs)
        assert b.tier_stats()["main_tier"] == ("fast" if tier == "thread" and outputs & 16 else tier)
    r = ob.Diagram(pts, box=BOX, table_radius=8).compute_cells(mode=ob.MODE_SECURITY)
    ok = r.status == 0  # cells the oracle's own truncated table could finish
    assert ok.mean() > 0.99
    helpers.assert_cells_identical(_Subset(b, ok), _Subset(r, ok), what=case)
    assert np.all(b.status == 0)  # the GPU re-runs table-exhausted cells with a larger table
    assert abs(b.volumes.sum() - 1.0) <= 1e-12
    assert abs(b.volume_sum() - 1.0) <= 1e-12
    if ok.all() and outputs & 16:
        c = b.counters()
        for k in ("tested", "vertex_classifications", "cuts", "new_vertices", "faces", "visited", "table_entries"):
            if case == "clustered100k" and k in ("visited", "table_entries", "tested", "vertex_classifications", "cuts", "new_vertices"):
                continue  # redone cells are counted by the first pass only
            assert c[k] == r.counters[k], k
    d.close()


class _Subset:
    """View of the cells selected by a boolean mask, in CSR form."""

    def __init__(self, r, mask):
        fo = np.asarray(r.face_offsets, np.int64)
        cnt = np.diff(fo)[mask]
        self.volumes = np.asarray(r.volumes)[mask]
        self.face_offsets = np.concatenate([[0], np.cumsum(cnt)])
        sel = np.repeat(mask, np.diff(fo))
        self.neighbors = np.asarray(r.neighbors)[sel]
        self.areas = np.asarray(r.areas)[sel]


def test_exhausted_table_is_redone_not_wrong(tess, gen, ob, variant):
    """A deliberately tiny shell table (R=1) cannot terminate any cell; the redo pass with larger
    tables must still deliver the exact cells."""
    pts = gen.uniform(20_000, 53)
    d = _diagram(tess, pts)
    b = d.compute_all_cells(outputs=variant[1], table_radius=1)
    r = ob.Diagram(pts, box=BOX).compute_cells(mode=ob.MODE_SECURITY)
    helpers.assert_cells_identical(b, r, what="R=1")
    assert np.all(b.status == 0)
    d.close()


def test_large_cell_path(tess, gen, ob, variant):
    """A particle surrounded by a dense shell has hundreds of faces: more than the small tables
    hold (64 vertices / 40 faces), so it must come from the large-cell configuration
    (1024 vertices / 512 faces; beyond that the cell is reported with TESS_STATUS_CAPACITY_OVERFLOW)."""
    u = gen.uniform(300, 54)
    th, ph = np.arccos(2 * u[:, 0] - 1), 2 * np.pi * u[:, 1]
    shell = 0.5 + 0.3 * np.stack([np.sin(th) * np.cos(ph), np.sin(th) * np.sin(ph), np.cos(th)], axis=1)
    pts = np.concatenate([[[0.5, 0.5, 0.5]], shell, gen.uniform(2000, 55)[np.linalg.norm(gen.uniform(2000, 55) - 0.5, axis=1) > 0.35]])
    d = _diagram(tess, pts)
    b = d.compute_all_cells(outputs=variant[1])
    r = ob.Diagram(pts, box=BOX).compute_cells(mode=ob.MODE_SECURITY)
    assert len(r.cell_neighbors(0)) > 100
    helpers.assert_cells_identical(b, r, what="shell")
    assert np.all(b.status == 0)
    assert abs(b.volumes.sum() - 1.0) <= 1e-12
    d.close()


def test_results_are_deterministic(tess, gen):
    pts = gen.uniform(50_000, 56)
    d = _diagram(tess, pts)
    a = d.compute_all_cells()
    b = d.compute_all_cells()
    assert np.array_equal(a.volumes, b.volumes) and np.array_equal(a.neighbors, b.neighbors) and np.array_equal(a.areas, b.areas)
    d2 = _diagram(tess, pts)
    c = d2.compute_all_cells()
    assert np.array_equal(a.volumes, c.volumes) and np.array_equal(a.neighbors, c.neighbors) and np.array_equal(a.areas, c.areas)
    d.close()
    d2.close()


def test_parallel_cut_equals_serial_walk(tess, gen, tmp_path):
    """The kernel cuts in lane-parallel form when no vertex lies on the plane and falls back to the
    reference-shaped serial walk otherwise, and the fast path finds the crossed half-edges either
    through the per-vertex adjacency lists or by sweeping the half-edge table.  TESS_FORCE_SERIAL=1
    disables the fast path, TESS_FORCE_SWEEP=1 the adjacency lists: all three must give bit-identical
    volumes, areas and face order."""
    import subprocess
    import sys

    code = (
        "import sys, importlib, numpy as np; sys.path.insert(0, %r);"
        "T = importlib.import_module('the-tessellator_b200');"
        "pts = np.concatenate([T.generators.uniform(150000, 5), T.generators.bcc(20, 5)]);"
        "d = T.Diagram(0); d.add_particles(pts); d.initialize(T.Polyhedron(0, 0, 0, 1, 1, 1));"
        "b = d.compute_all_cells(outputs=7);"
        "np.savez(sys.argv[1], v=b.volumes, n=b.neighbors, a=b.areas, o=b.face_offsets, s=b.status)"
    ) % os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    outs = {}
    for tag, env in (("par", {}), ("ser", {"TESS_FORCE_SERIAL": "1"}), ("sweep", {"TESS_FORCE_SWEEP": "1"})):
        f = str(tmp_path / f"{tag}.npz")
        subprocess.check_call([sys.executable, "-c", code, f], env=dict(os.environ, **env))
        outs[tag] = np.load(f)
    for k in "vnaos":
        assert np.array_equal(outs["par"][k], outs["ser"][k]), k
        assert np.array_equal(outs["par"][k], outs["sweep"][k]), k


def test_main_tiers_are_bit_identical(tess, gen):
    """The three kernels of the main pass (warp per cell with the serial walk / without it / thread per cell) hand what
    they cannot finish to the next tier.  Random, clustered and exact-lattice input (where nearly every cell is handed
    back) must come out bit-identical from all of them, work counters included."""
    G = gen
    pts = np.concatenate([G.uniform(150000, 5), 0.25 + 0.5 * G.simple_cubic(12), G.clustered(60000, 4, k=4)])
    d = _diagram(tess, pts)
    outs = {}
    try:
        for tier in ("small", "fast", "thread"):
            tess.set_main_tier(tier)
            b = d.compute_all_cells(outputs=7 | 16)
            assert b.tier_stats()["main_tier"] == ("fast" if tier == "thread" else tier)  # no work counters in the thread-per-cell kernel
            c = b.counters()
            outs[tier] = (b.volumes.copy(), b.neighbors.copy(), b.areas.copy(), b.face_offsets.copy(), b.status.copy(), [c[k] for k in sorted(c)])
            b2 = d.compute_all_cells(outputs=7)
            assert b2.tier_stats()["main_tier"] == tier
            if tier == "thread":
                assert 0 < b2.tier_stats()["redo_a"] < len(pts) // 4  # it hands back what its tables cannot hold, and only that
            for x, y in zip(outs[tier][:5], (b2.volumes, b2.neighbors, b2.areas, b2.face_offsets, b2.status)):
                assert np.array_equal(x, y), tier
    finally:
        tess.set_main_tier("default")
    for tier in ("fast", "thread"):
        for k, (x, y) in enumerate(zip(outs["small"], outs[tier])):
            assert np.array_equal(x, y), (tier, k)
    d.close()


# ------------------------------------------------------------------ options ------------------
def test_reference_radius_mode(tess, gen, ob):
    pts = gen.uniform(20_000, 57)
    d = _diagram(tess, pts)
    od = ob.Diagram(pts, box=BOX)
    sx = od.cell_info()[0]
    for radius in (0.0, (1.5 * sx) ** 2, (4 * sx) ** 2):
        b = d.compute_all_cells(search_radius=radius, outputs=ALL_OUT)
        r = od.compute_cells(mode=ob.MODE_REFERENCE_RADIUS, search_radius=radius)
        helpers.assert_cells_identical(b, r, what=f"radius {radius}")
        assert b.counters()["tested"] == r.counters["tested"]
    d.close()


def test_reference_radius_beyond_the_default_table(tess, gen, ob):
    """a radius that reaches past the default search table (half-width 8 grid cells).  The table is sized from
    the radius, and a caller-given table that is too small is widened by the redo passes — never silently cut short."""
    pts = gen.uniform(6000, 59)
    d = _diagram(tess, pts)
    od = ob.Diagram(pts, box=BOX)
    sx = od.cell_info()[0]
    radius = (9.5 * sx) ** 2
    r = od.compute_cells(mode=ob.MODE_REFERENCE_RADIUS, search_radius=radius)
    for kw in (dict(), dict(table_radius=4)):
        b = d.compute_all_cells(search_radius=radius, outputs=ALL_OUT, **kw)
        helpers.assert_cells_identical(b, r, what=f"radius {radius} {kw}")
        assert np.all(b.status == 0)
    assert d.compute_all_cells(search_radius=radius, outputs=ALL_OUT).counters()["tested"] == r.counters["tested"]
    d.close()


def test_target_group(tess, gen, ob):
    pts = gen.uniform(20_000, 58)
    groups = (np.arange(len(pts)) % 3).astype(np.uint64)
    d = _diagram(tess, pts, groups=groups)
    od = ob.Diagram(pts, box=BOX, groups=groups)
    for tg in (0, 2):
        b = d.compute_all_cells(target_group=tg, outputs=ALL_OUT)
        r = od.compute_cells(mode=ob.MODE_SECURITY, target_group=tg)
        helpers.assert_cells_identical(b, r, what=f"group {tg}")
    # a group nobody carries: nothing cuts, every cell is the whole container
    b = d.compute_all_cells(target_group=7)
    assert np.all(np.abs(b.volumes - 1.0) <= 1e-15) and np.all(np.diff(b.face_offsets) == 6)
    assert np.all(b.neighbors < 0)  # only the six container walls
    d.close()


def test_cells_at_query_points(tess, gen, ob):
    pts = gen.uniform(20_000, 59)
    d = _diagram(tess, pts)
    od = ob.Diagram(pts, box=BOX)
    q = gen.uniform(64, 60)
    b = d.compute_cells_at(q, outputs=ALL_OUT)
    for i in range(len(q)):
        r = od.compute_cell_at_point(*q[i])
        assert b.cell_neighbors(i).tolist() == r.neighbors.tolist()
        assert b.volumes[i] == r.volumes[0] and np.array_equal(b.cell_areas(i), r.areas)
    d.close()


def test_query_cells_take_the_redo_tiers(tess, gen, ob):
    """cells at query positions (get_cell_at_particle, interface.rs:211-232) that the small tables or the default
    search table cannot finish — a position inside a dense shell (hundreds of faces), a position in a void (the walk runs
    off the radius-8 table) — are redone tier by tier like every other cell, not returned empty or half clipped."""
    u = gen.uniform(300, 54)
    th, ph = np.arccos(2 * u[:, 0] - 1), 2 * np.pi * u[:, 1]
    shell = 0.5 + 0.3 * np.stack([np.sin(th) * np.cos(ph), np.sin(th) * np.sin(ph), np.cos(th)], axis=1)
    bg = gen.uniform(60_000, 55)
    pts = np.concatenate([shell, bg[(np.linalg.norm(bg - 0.5, axis=1) > 0.35)]])  # a shell around an empty ball, a dense background outside
    d = _diagram(tess, pts)
    od = ob.Diagram(pts, box=BOX)
    q = np.concatenate([[[0.5, 0.5, 0.5], [0.52, 0.47, 0.5]], gen.uniform(20, 60)])
    b = d.compute_cells_at(q, outputs=ALL_OUT)
    assert np.all(b.status == 0)
    for i in range(len(q)):
        r = od.compute_cell_at_point(*q[i])
        assert b.cell_neighbors(i).tolist() == r.neighbors.tolist(), i
        assert b.volumes[i] == r.volumes[0] and np.array_equal(b.cell_areas(i), r.areas), i
    assert len(b.cell_neighbors(0)) > 100
    # the per-cell API on the same position
    cell = d.get_cell_at_particle((0.5, 0.5, 0.5), tess.Polyhedron(*BOX))
    assert cell.compute_volume() == od.compute_cell_at_point(0.5, 0.5, 0.5).volumes[0]
    d.close()


def test_strided_host_input_and_incremental_adds(tess, gen, ob):
    pts = gen.uniform(3000, 61)
    rec = np.zeros((3000, 5))
    rec[:, :3] = pts  # 40-byte records: x, y, z + two payload fields (ToCeleryPoint getters)
    d = tess.Diagram(0)
    d.add_particles(rec[:1000])
    for p in pts[1000:1010]:
        d.add_particle_with_group(p, 0)
    d.add_particles(rec[1010:])
    d.initialize(tess.Polyhedron(*BOX))
    b = d.compute_all_cells()
    r = ob.Diagram(pts, box=BOX).compute_cells(mode=ob.MODE_SECURITY)
    helpers.assert_cells_identical(b, r, what="strided")
    d.close()


# ------------------------------------------------------------------ reference-shaped API -----
def test_per_cell_api_reads_like_the_reference(tess, gen, ob):
    """interface.rs usage: diagram.get_cell_at_index(i, Polyhedron::new(box), None, None) ->
    compute_voronoi_cell -> compute_volume / compute_neighbors / compute_faces."""
    pts = gen.uniform(2000, 62)
    diagram = tess.Diagram()
    for p in pts:
        diagram.add_particle_with_group(p, 0)
    diagram.initialize(tess.Polyhedron(*BOX))
    r = ob.Diagram(pts, box=BOX).compute_cells(mode=ob.MODE_SECURITY, want_vertices=True)
    for i in (0, 7, 1999):
        cell = diagram.get_cell_at_index(i, tess.Polyhedron(*BOX), None, None)
        cell.compute_voronoi_cell()
        assert abs(cell.compute_volume() - r.volumes[i]) <= 1e-12 * r.volumes[i]
        assert sorted(cell.compute_neighbors()) == sorted(r.cell_neighbors(i).tolist())
        faces = cell.compute_faces()
        assert len(faces) == len(r.cell_neighbors(i))
        got = sorted((f.compute_neighbor(), f.compute_area()) for f in faces)
        exp = sorted(zip(r.cell_neighbors(i).tolist(), r.cell_areas(i).tolist()))
        for (gn, ga), (en, ea) in zip(got, exp):
            assert gn == en and abs(ga - ea) <= 1e-12 * ea
        assert cell.original_index() == i
        # vertices: same set of cell-local coordinates (bitwise), order unspecified
        gv = {tuple(v) for v in cell.compute_vertices()}
        ev = {tuple(v) for v in r.cell_vertices(i)}
        assert gv == ev
    with pytest.raises(tess.TessError):
        diagram.get_cell_at_index(0, tess.Polyhedron(0, 0, 0, 2, 2, 2))
    with pytest.raises(tess.TessError):
        diagram.add_particle_with_group((0.5, 0.5, 0.5), 0)  # after initialize (interface.rs:50-51)
    diagram.close()


def test_error_behaviour(tess):
    d = tess.Diagram(0)
    with pytest.raises(tess.TessError) as e:
        d.initialize(tess.Polyhedron(*BOX))  # empty: CeleryBounds::new panics (celery.rs:82)
    assert e.value.code == -1
    with pytest.raises(tess.TessError) as e:
        d.compute_all_cells()
    assert e.value.code == -2
    d.close()


# ------------------------------------------------------------------ slab mode on one device --
def test_slab_decomposition_is_bit_identical(tess, gen):
    """Fake multi-GPU: G slabs computed one after the other on one device with host-side halo
    selection; every cell must equal the whole-domain run bit for bit (same kernel, same grid
    parameters, same candidate order)."""
    pts = gen.uniform(60_000, 63)
    whole = _diagram(tess, pts)
    wb = whole.compute_all_cells()
    gi = whole.grid_info()
    cpd, b = gi["cells_per_dimension"], gi["bounds"]
    gx = np.minimum(((pts[:, 0] - b[0]) * gi["inverse_cell_sizes"][0]).astype(np.int64), cpd - 1)
    gx[pts[:, 0] >= b[1]] = cpd - 1
    G, h = 3, 4
    cuts = [0, cpd // 3, 2 * cpd // 3, cpd]
    seen = np.zeros(len(pts), bool)
    for g in range(G):
        own = (cuts[g], cuts[g + 1])
        local = (max(0, own[0] - h), min(cpd, own[1] + h))
        sel = np.nonzero((gx >= local[0]) & (gx < local[1]))[0]
        # shuffle the local arrival order: the canonical in-cell order must not depend on it
        sel = sel[np.argsort(gen.u01(7, sel.astype(np.uint64)))]
        d = tess.Diagram(0)
        import torch

        xyz = torch.from_numpy(pts[sel]).cuda()
        ids = torch.from_numpy(sel.astype(np.int64)).cuda()
        d.add_particles_device(xyz.data_ptr(), len(sel), ids_ptr=ids.data_ptr())
        d.initialize_slab(tess.Polyhedron(*BOX), b, len(pts), own, local)
        sb = d.compute_all_cells()
        assert np.all((sb.status & 8) == 0)  # halo of 4 planes is enough for uniform input
        ids_out = sb.cell_ids
        assert np.all((gx[ids_out] >= own[0]) & (gx[ids_out] < own[1]))
        seen[ids_out] = True
        assert np.array_equal(sb.volumes, wb.volumes[ids_out])
        wfo = wb.face_offsets
        for k in range(0, len(ids_out), 997):
            i = ids_out[k]
            assert np.array_equal(sb.cell_neighbors(k), wb.neighbors[wfo[i]:wfo[i + 1]])
            assert np.array_equal(sb.cell_areas(k), wb.areas[wfo[i]:wfo[i + 1]])
        # the streamed form of the same slab (rows in sorted order: chunks are runs of slots)
        vol, off, nbr, area, stat = _host_arrays(sb.n_cells, sb.n_faces)
        d.compute_all_cells_to_host(vol, off, nbr, area, stat, n_chunks=4, outputs=1 | 2 | 4)
        assert np.array_equal(vol, sb.volumes) and np.array_equal(nbr, sb.neighbors) and np.array_equal(area, sb.areas)
        assert np.array_equal(off.astype(np.int64), np.asarray(sb.face_offsets).astype(np.int64)) and np.array_equal(stat, sb.status)
        d.close()
    assert seen.all()
    # a halo that is too thin is reported per cell, never silently wrong
    own, local = (cuts[1], cuts[2]), (cuts[1] - 1, cuts[2] + 1)
    sel = np.nonzero((gx >= local[0]) & (gx < local[1]))[0]
    d = tess.Diagram(0)
    import torch

    xyz = torch.from_numpy(pts[sel]).cuda()
    ids = torch.from_numpy(sel.astype(np.int64)).cuda()
    d.add_particles_device(xyz.data_ptr(), len(sel), ids_ptr=ids.data_ptr())
    d.initialize_slab(tess.Polyhedron(*BOX), b, len(pts), own, local)
    sb = d.compute_all_cells()
    flagged = (sb.status & 8) != 0
    assert flagged.any()
    okc = ~flagged
    assert np.array_equal(sb.volumes[okc], wb.volumes[sb.cell_ids[okc]])
    d.close()
    whole.close()


def test_record_exchange_path_on_one_device(tess, gen):
    """The multi-GPU step's own kernels on one device: tess_pack_records routes one rank's particles to three fake ranks
    as 32-byte records (with and without planned counts), each segment goes through tess_diagram_add_records_device +
    initialize_slab + compute, and every cell equals the whole-domain run bit for bit.  A plan that does not fit the
    particles is detected through the packed counts and never writes past a planned segment."""
    import importlib

    import torch

    D = importlib.import_module("the-tessellator_b200.distributed")
    pts = gen.uniform(80_000, 66)
    n = len(pts)
    whole = _diagram(tess, pts)
    wb = whole.compute_all_cells()
    be = D.CudaSlabBackend(0)
    xyz = torch.from_numpy(pts).cuda()
    b6 = be.bounds(xyz).cpu().numpy()
    cpd = D.cells_per_dimension(n)
    cuts = D.slab_cuts(be.plane_histogram(xyz, b6, n).cpu().numpy(), 3)
    lo, hi = D.receive_ranges(cuts, 4)
    owns = [(cuts[g], cuts[g + 1]) for g in range(3)]
    # slab cuts finer than a plane: rows (x, y) of the x-major grid; two fake ranks then share a plane
    rcuts = D.row_cuts(be.row_histogram(xyz, b6, n).cpu().numpy(), 3)
    assert rcuts[0] == 0 and rcuts[-1] == cpd * cpd and any(c % cpd for c in rcuts[1:-1])
    rlo, rhi = D.receive_ranges_rows(rcuts, cpd, 4)
    rowns = [(rcuts[g] // cpd, rcuts[g] % cpd, rcuts[g + 1] // cpd, rcuts[g + 1] % cpd) for g in range(3)]
    seen_rows = np.zeros(n, bool)
    rc_counts, rrec, _ = be.pack_records(xyz, 0, b6, n, rlo, rhi)
    rrec = rrec.clone()
    o = 0
    n_owned_rows = []
    for g in range(3):
        batch, n_owned, flag = be.compute_records(rrec[o:o + rc_counts[g]].contiguous(), BOX, b6, n, rowns[g], (rlo[g], rhi[g]), dict(outputs=7))
        assert int(flag.item()) == 0
        ids = batch.cell_ids
        assert not seen_rows[ids].any()
        seen_rows[ids] = True
        n_owned_rows.append(n_owned)
        assert np.array_equal(batch.volumes, wb.volumes[ids])
        wfo = wb.face_offsets
        for k in range(0, len(ids), 499):
            i = ids[k]
            assert np.array_equal(batch.cell_neighbors(k), wb.neighbors[wfo[i]:wfo[i + 1]]) and np.array_equal(batch.cell_areas(k), wb.areas[wfo[i]:wfo[i + 1]])
        o += rc_counts[g]
    assert seen_rows.all() and max(n_owned_rows) - min(n_owned_rows) < 200  # balanced to a row's worth, not a plane's (1.7k here)
    counts, rec, counts_dev = be.pack_records(xyz, 0, b6, n, lo, hi)
    assert counts == [int(v) for v in counts_dev.cpu().tolist()] and sum(counts) == rec.shape[0] > n
    rec = rec.clone()
    # the same with the counts as a plan: no counting pass, same multiset of records per segment
    counts2, rec2, counts_dev2 = be.pack_records(xyz, 0, b6, n, lo, hi, planned_counts=counts)
    assert counts2 == counts and [int(v) for v in counts_dev2.cpu().tolist()] == counts
    o = 0
    seen = np.zeros(n, bool)
    for g in range(3):
        seg, seg2 = rec[o:o + counts[g]], rec2[o:o + counts[g]]
        ids1 = np.sort(seg[:, 3].contiguous().view(torch.int64).cpu().numpy())
        assert np.array_equal(ids1, np.sort(seg2[:, 3].contiguous().view(torch.int64).cpu().numpy()))
        batch, n_owned, flag = be.compute_records(seg.contiguous(), BOX, b6, n, owns[g], (lo[g], hi[g]), dict(outputs=7))
        assert int(flag.item()) == 0 and n_owned == batch.n_cells
        ids = batch.cell_ids
        seen[ids] = True
        assert np.array_equal(batch.volumes, wb.volumes[ids])
        wfo = wb.face_offsets
        for k in range(0, len(ids), 499):
            i = ids[k]
            assert np.array_equal(batch.cell_neighbors(k), wb.neighbors[wfo[i]:wfo[i + 1]]) and np.array_equal(batch.cell_areas(k), wb.areas[wfo[i]:wfo[i + 1]])
        o += counts[g]
    assert seen.all()
    # a stale plan (counts of another particle set, some too small): detected, nothing written past the segments
    other = torch.from_numpy(gen.uniform(n, 67)).cuda()
    small = [c - 50 for c in counts]
    c3, rec3, dev3 = be.pack_records(other, 0, b6, n, lo, hi, planned_counts=small)
    assert c3 == small and [int(v) for v in dev3.cpu().tolist()] != small and rec3.shape[0] == sum(small)
    # and the whole pipeline as one rank (no process group): plan, then the planned step
    res = D.compute_sharded(be, xyz, 0, n, BOX)
    res2 = D.compute_sharded(be, xyz, 0, n, BOX, plan=res.plan)
    assert res2.rounds == 1 and np.array_equal(res2.batch.volumes, wb.volumes[res2.batch.cell_ids])
    whole.close()


# ------------------------------------------------------------------ full-size configs --------
def _full_size_checks(tess, gen, ob, pts, sample_seed, n_sample, box=BOX):
    d = _diagram(tess, pts, box)
    b = d.compute_all_cells(outputs=ALL_OUT)
    n = len(pts)
    # the instantiations bench.py times (no counters), on every main tier, must give the same arrays bit for bit
    try:
        for tier in ("default", "small", "fast", "thread"):
            tess.set_main_tier(tier)
            t = d.compute_all_cells(outputs=7)
            for name in ("volumes", "face_offsets", "neighbors", "areas", "status"):
                assert np.array_equal(getattr(t, name), getattr(b, name)), (tier, name)
            del t
    finally:
        tess.set_main_tier("default")
    assert np.all(b.status == 0)
    assert abs(b.volume_sum() - 1.0) <= 1e-12
    assert abs(float(np.sum(b.volumes)) - 1.0) <= 1e-12
    assert np.all(b.volumes > 0) and np.all(b.areas >= 0)
    assert helpers.neighbor_symmetry_violations(b.face_offsets, b.neighbors) == 0
    # oracle on a sample of cells
    ids = np.unique((gen.u01(sample_seed, np.arange(n_sample, dtype=np.uint64)) * n).astype(np.uint64))
    r = ob.Diagram(pts, box=BOX, table_radius=8).compute_cells(ids=ids, mode=ob.MODE_SECURITY)
    mask = np.zeros(n, bool)
    mask[ids.astype(np.int64)] = True
    helpers.assert_cells_identical(_Subset(b, mask), r, what=f"sample of {len(ids)}")
    d.close()


def test_config2_one_million_uniform(tess, gen, ob):
    _full_size_checks(tess, gen, ob, gen.uniform(1_000_000, 2), 71, 20_000)


def test_config3_ten_million_uniform(tess, gen, ob):
    _full_size_checks(tess, gen, ob, gen.uniform(10_000_000, 3), 72, 5_000)


def test_config4_ten_million_clustered(tess, gen, ob):
    """Config 4 exactly as SURVEY §8d states it: 10M points, seed 4, 20 % uniform background + 32 Gaussian clusters of
    sigma 0.02.  Every tier of the pipeline runs (wider-table, medium and large redo passes); the oracle sample holds
    random cells plus the cells with the most faces (rim cells: among them the ones that outgrew the small tables on the way)."""
    pts = gen.clustered(10_000_000, 4, k=32, sigma=0.02)
    n = len(pts)
    d = _diagram(tess, pts)
    b = d.compute_all_cells(outputs=7)
    st = b.tier_stats()
    assert st["redo_b"] > 0, st  # medium-path cells exist in this input
    assert np.all(b.status == 0)
    assert abs(b.volume_sum() - 1.0) <= 1e-12
    assert np.all(b.volumes > 0) and np.all(b.areas >= 0)
    nf = np.diff(b.face_offsets.astype(np.int64))  # (no finished cell has more faces than the small tables hold: cells outgrow them on the way)
    rnd = np.unique((gen.u01(74, np.arange(3000, dtype=np.uint64)) * n).astype(np.uint64))
    big = np.argsort(nf)[-400:].astype(np.uint64)
    ids = np.unique(np.concatenate([rnd, big]))
    r = ob.Diagram(pts, box=BOX).compute_cells(ids=ids, mode=ob.MODE_SECURITY)
    mask = np.zeros(n, bool)
    mask[ids.astype(np.int64)] = True
    helpers.assert_cells_identical(_Subset(b, mask), r, what=f"clustered 10M, sample of {len(ids)}")
    # the counting instantiation and the other main tiers give the same arrays
    try:
        for tier, outputs in (("default", ALL_OUT), ("small", 7), ("thread", 7)):
            tess.set_main_tier(tier)
            t = d.compute_all_cells(outputs=outputs)
            for name in ("volumes", "face_offsets", "neighbors", "areas", "status"):
                assert np.array_equal(getattr(t, name), getattr(b, name)), (tier, name)
            del t
    finally:
        tess.set_main_tier("default")
    d.close()


def test_config5_jittered_bcc_on_one_gpu(tess, gen, ob):
    """Config 5's input at the size one test can afford (2 * 128^3 = 4.2M points, jitter 1e-3 a): oracle sample bit for
    bit, and the analytic cell (truncated octahedron: 14 faces, V = a^3 / 2) as an independent anchor — the jitter moves
    every bisector by O(1e-3 a), so volumes stay within 2e-2 relative and the neighbour sets are the lattice's."""
    import test_oracle_analytic as an

    m = 128
    pts = gen.bcc(m, 5)
    _full_size_checks(tess, gen, ob, pts, 75, 4_000)
    d = _diagram(tess, pts)
    b = d.compute_all_cells(outputs=7)
    a = 1.0 / m
    fo = b.face_offsets.astype(np.int64)
    p = np.arange(len(pts), dtype=np.int64)
    site = p // 2
    i, j, k = site // (m * m), (site // m) % m, site % m
    inner = (np.minimum(np.minimum(i, j), k) >= 2) & (np.maximum(np.maximum(i, j), k) <= m - 3)
    assert np.all(np.diff(fo)[inner] == 14)
    assert np.max(np.abs(b.volumes[inner] - a ** 3 / 2) / (a ** 3 / 2)) < 2e-2
    sq, hx = a * a / 8, 3 * np.sqrt(3.0) * a * a / 16
    for c in np.flatnonzero(inner)[:: 40_000]:
        nb, ar = b.neighbors[fo[c]:fo[c + 1]], b.areas[fo[c]:fo[c + 1]]
        same = (nb % 2) == (c % 2)
        assert same.sum() == 6 and np.all(np.abs(ar[same] - sq) / sq < 5e-2) and np.all(np.abs(ar[~same] - hx) / hx < 5e-2)
    d.close()


def test_analytic_lattices_on_gpu(tess, gen):
    """The closed forms that pin the oracle (tests/test_oracle_analytic.py), asserted on the GPU's own output: un-jittered
    BCC (truncated octahedra), simple cubic (cubes) and points on a line (slabs), for every main tier."""
    import test_oracle_analytic as an

    try:
        for tier in ("default", "small", "thread"):
            tess.set_main_tier(tier)
            m = 16
            d = _diagram(tess, gen.bcc(m, 5, jitter=0.0))
            b = d.compute_all_cells(outputs=7, table_radius=1 << 20)
            assert an.check_bcc(b.volumes, b.face_offsets, b.neighbors, b.areas, b.status, m) > 1.4 * (m - 4) ** 3
            d.close()
            m = 12
            d = _diagram(tess, gen.simple_cubic(m))
            b = d.compute_all_cells(outputs=7, table_radius=1 << 20)
            assert an.check_simple_cubic(b.volumes, b.face_offsets, b.neighbors, b.areas, b.status, m) > 0.3 * (m - 2) ** 3
            d.close()
            pts, x = an.line_points(gen)
            d = _diagram(tess, pts)
            b = d.compute_all_cells(outputs=7)
            assert np.all(b.status == 0)
            an.check_line(b.volumes, b.face_offsets, b.neighbors, b.areas, x)
            d.close()
    finally:
        tess.set_main_tier("default")


# ------------------------------------------------------------------ geometry ---
def test_vertices_and_face_loops_match_oracle(tess, gen, ob):
    """Cell::compute_vertices (interface.rs:368) and VoronoiFace::compute_vertices (interface.rs:403 ->
    polyhedron.rs:897-919): per-face ordered vertex loops, bit-identical coordinates and order; the cell's
    vertex list is the same set (its order follows the vertex slots, which the kernel numbers differently)."""
    u = gen.uniform(200, 54)
    th, ph = np.arccos(2 * u[:, 0] - 1), 2 * np.pi * u[:, 1]
    shell = 0.5 + 0.3 * np.stack([np.sin(th) * np.cos(ph), np.sin(th) * np.sin(ph), np.cos(th)], axis=1)
    bg = gen.uniform(6000, 64)
    pts = np.concatenate([[[0.5, 0.5, 0.5]], shell, bg[np.linalg.norm(bg - 0.5, axis=1) > 0.35]])  # cell 0 needs the large-cell path
    d = _diagram(tess, pts)
    b = d.compute_all_cells(outputs=1 | 2 | 4 | 8)
    r = ob.Diagram(pts, box=BOX).compute_cells(mode=ob.MODE_SECURITY, want_vertices=True)
    helpers.assert_cells_identical(b, r, what="geometry run")
    assert np.array_equal(b.vertex_offsets, r.vertex_offsets)
    assert np.array_equal(b.face_vertex_offsets, r.loop_offsets)
    assert len(r.cell_neighbors(0)) > 60
    fo = b.face_offsets
    for c in list(range(0, len(pts), 97)) + [0]:
        assert {tuple(v) for v in b.cell_vertices(c)} == {tuple(v) for v in r.cell_vertices(c)}
        for k in range(fo[c], fo[c + 1]):
            assert np.array_equal(b.face_vertices(c, k), r.face_loop(k)), (c, k)
    # the per-cell API
    box = tess.Polyhedron(*BOX)
    cell = d.get_cell_at_index(5, box)
    faces = cell.compute_faces()
    for j, f in enumerate(faces):
        assert np.array_equal(f.compute_vertices(), r.face_loop(fo[5] + j))
    d.close()


# ------------------------------------------------------------------ awkward inputs -----------
@pytest.mark.parametrize("case", ["duplicates", "on_walls", "outside_box", "two_points", "collinear", "coplanar", "lattice_jitter0"])
def test_awkward_inputs_match_oracle(tess, gen, ob, case, variant):
    """Inputs the reference does not guard against (exact duplicates give a NaN plane that cuts nothing,
    SURVEY D16; particles on or outside the container; exact lattice ties).  Whatever the reference's
    arithmetic does with them, the kernel must do the same, bit for bit, and flag the same cells."""
    base = gen.uniform(3000, 81)
    pts = {
        "duplicates": lambda: np.concatenate([base, base[:200], base[:50]]),
        "on_walls": lambda: np.concatenate([base, np.round(gen.uniform(400, 82), 0) * np.array([1, 1, 0]) + gen.uniform(400, 83) * np.array([0, 0, 1])]),
        "outside_box": lambda: np.concatenate([base, 1.0 + 0.05 * gen.uniform(30, 84), -0.05 * gen.uniform(30, 85)]),
        "two_points": lambda: np.array([[0.25, 0.5, 0.5], [0.75, 0.5, 0.5]]),
        "collinear": lambda: np.stack([np.linspace(0.05, 0.95, 40), np.full(40, 0.5), np.full(40, 0.5)], axis=1),
        "coplanar": lambda: np.concatenate([gen.uniform(500, 86) * np.array([1, 1, 0]) + np.array([0, 0, 0.5])]),
        "lattice_jitter0": lambda: gen.bcc(6, 5, jitter=0.0),
    }[case]()
    d = _diagram(tess, pts)
    b = d.compute_all_cells(outputs=variant[1], table_radius=1 << 20)
    r = ob.Diagram(pts, box=BOX).compute_cells(mode=ob.MODE_SECURITY)
    helpers.assert_cells_identical(b, r, what=case)
    assert np.array_equal(b.status, r.status)
    if variant[1] & 16:
        c = b.counters()
        for k in ("tested", "cuts", "new_vertices", "faces", "degenerate_skips"):
            assert c[k] == r.counters[k], k
    d.close()


# ------------------------------------------------------------------ streamed download --------
def _host_arrays(n, cap):
    return (np.full(n, -1.0), np.full(n + 1, 2**63, dtype=np.uint64), np.full(cap, -99, dtype=np.int64), np.full(cap, -1.0), np.full(n, 0xFFFFFFFF, dtype=np.uint32))


@pytest.mark.parametrize("chunks", [0, 3, 16])
def test_streamed_host_results_equal_download(tess, gen, chunks):
    """tess_compute_all_to_host: rows computed chunk by chunk, each chunk copied while the next is clipped —
    the host arrays must equal tess_compute_all + download bit for bit (and the device batch too)."""
    pts = gen.uniform(120_000, 71)
    d = _diagram(tess, pts)
    ref = d.compute_all_cells(outputs=ALL_OUT)
    n, nf = ref.n_cells, ref.n_faces
    vol, off, nbr, area, stat = _host_arrays(n, nf + 100)
    b = d.compute_all_cells_to_host(vol, off, nbr, area, stat, n_chunks=chunks, outputs=ALL_OUT)
    assert b.n_cells == n and b.n_faces == nf
    assert np.array_equal(vol, ref.volumes) and np.array_equal(off.astype(np.int64), np.asarray(ref.face_offsets).astype(np.int64))
    assert np.array_equal(nbr[:nf], ref.neighbors) and np.array_equal(area[:nf], ref.areas) and np.array_equal(stat, ref.status)
    assert np.all(nbr[nf:] == -99) and np.all(area[nf:] == -1.0)  # nothing written past the faces
    assert np.array_equal(b.volumes, ref.volumes) and np.array_equal(b.neighbors, ref.neighbors) and np.array_equal(b.areas, ref.areas)
    d.close()


def test_streamed_host_results_with_redo_cells(tess, gen):
    """Cells that need a redo pass (voids: table exhausted; a 300-neighbour cell: large path) appear after
    chunks were already copied: the call must repack and copy everything again."""
    u = gen.uniform(300, 54)
    th, ph = np.arccos(2 * u[:, 0] - 1), 2 * np.pi * u[:, 1]
    shell = 0.5 + 0.3 * np.stack([np.sin(th) * np.cos(ph), np.sin(th) * np.sin(ph), np.cos(th)], axis=1)
    bg = gen.uniform(60_000, 55)
    pts = np.concatenate([[[0.5, 0.5, 0.5]], shell, bg[np.linalg.norm(bg - 0.5, axis=1) > 0.35]])
    d = _diagram(tess, pts)
    ref = d.compute_all_cells(outputs=ALL_OUT)
    n, nf = ref.n_cells, ref.n_faces
    vol, off, nbr, area, stat = _host_arrays(n, nf)
    d.compute_all_cells_to_host(vol, off, nbr, area, stat, n_chunks=8, outputs=ALL_OUT)
    assert np.array_equal(vol, ref.volumes) and np.array_equal(off.astype(np.int64), np.asarray(ref.face_offsets).astype(np.int64))
    assert np.array_equal(nbr, ref.neighbors) and np.array_equal(area, ref.areas) and np.array_equal(stat, ref.status)
    assert abs(vol.sum() - 1.0) <= 1e-12
    d.close()


def test_streamed_host_capacity_error_and_optional_arrays(tess, gen):
    pts = gen.uniform(40_000, 72)
    d = _diagram(tess, pts)
    ref = d.compute_all_cells(outputs=ALL_OUT)
    vol, off, nbr, area, stat = _host_arrays(ref.n_cells, ref.n_faces // 2)
    with pytest.raises(tess.TessError):
        d.compute_all_cells_to_host(vol, off, nbr, area, stat, n_chunks=4, outputs=ALL_OUT)
    with pytest.raises(tess.TessError):  # per-cell arrays too short
        d.compute_all_cells_to_host(vol[:-1], off, nbr, area, stat, n_chunks=4, outputs=ALL_OUT)
    vol, off, nbr, area, stat = _host_arrays(ref.n_cells, ref.n_faces)
    d.compute_all_cells_to_host(vol, None, nbr, None, None, n_chunks=4, outputs=1 | 2)  # volumes + neighbours only
    assert np.array_equal(vol, ref.volumes) and np.array_equal(nbr, ref.neighbors) and np.all(area == -1.0)
    d.close()
