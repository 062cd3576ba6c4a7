"""The clip kernel SOURCE (the-tessellator_b200/csrc/clip.cu, unchanged) run lane by lane on the CPU warp emulator
(tests/emu/) against the oracle.  This is test infrastructure: it lets the CPU-only suite exercise every path of
the kernel — lane-parallel cut through the adjacency lists, table-sweep variant, serial walk, large-cell
configuration, radius and group modes — bit for bit, and it checks two things a GPU run cannot show:
every warp collective is reached by all 32 lanes from the same source line, and results do not depend on the
order in which the lanes run between two collectives (forward 0..31 and reverse 31..0 schedules)."""
import numpy as np
import pytest

import helpers

BOX = (0, 0, 0, 1, 1, 1)
BAD = 0x2 | 0x4 | 0x10  # TABLE_EXHAUSTED | CAPACITY_OVERFLOW | INCONSISTENT: cells the product re-runs


@pytest.fixture(scope="module")
def eb():
    import emu_binding

    emu_binding.lib()
    return emu_binding


class _Rows:
    def __init__(self, r, mask):
        fo = np.asarray(r.face_offsets, np.int64)
        cnt = np.diff(fo)[mask]
        self.volumes = np.asarray(r.volumes)[mask]
        self.face_offsets = np.concatenate([[0], np.cumsum(cnt)])
        sel = np.repeat(mask, np.diff(fo))
        self.neighbors = np.asarray(r.neighbors)[sel]
        self.areas = np.asarray(r.areas)[sel]


def _check(e, r, what, expect_all=True):
    ok = (e.status & BAD) == 0
    if expect_all:
        assert ok.all(), f"{what}: {int((~ok).sum())} cells not finished"
    helpers.assert_cells_identical(_Rows(e, ok), _Rows(r, ok), what=what)
    assert np.array_equal(e.status[ok], r.status[ok]), what
    return ok


@pytest.mark.parametrize("case", ["uniform", "clustered", "bcc", "tiny", "on_walls", "duplicates"])
@pytest.mark.parametrize("reverse", [False, True])
def test_small_configuration_matches_oracle(eb, gen, case, reverse):
    base = gen.uniform(1500, 81)
    pts = {
        "uniform": lambda: gen.uniform(3000, 51),
        "clustered": lambda: gen.clustered(3000, 4, k=4),
        "bcc": lambda: gen.bcc(10, 5),
        "tiny": lambda: gen.uniform(5, 52),
        "on_walls": lambda: np.concatenate([base, np.round(gen.uniform(200, 82), 0) * np.array([1, 1, 0]) + gen.uniform(200, 83) * np.array([0, 0, 1])]),
        "duplicates": lambda: np.concatenate([base, base[:100], base[:20]]),
    }[case]()
    g = eb.EmuGrid(pts, BOX, table_radius=-1 if case in ("tiny", "clustered") else 8)
    e = g.clip(reverse=reverse)
    r = g.oracle_cells()
    ok = _check(e, r, case, expect_all=case != "clustered")
    assert ok.mean() > 0.95
    if ok.all():
        for k in ("visited", "tested", "vertex_classifications", "cuts", "new_vertices", "table_entries", "degenerate_skips", "faces"):
            assert e.counters[k] == r.counters[k], k
    assert np.array_equal(e.cell_id, g.sorted_indices)


@pytest.mark.parametrize("case", ["uniform", "clustered", "bcc"])
def test_kernel_without_counters(eb, gen, case):
    """The instantiation the product times (COUNT=false): screened-out planes leave the candidate queue at once."""
    pts = {"uniform": lambda: gen.uniform(3000, 61), "clustered": lambda: gen.clustered(4000, 4, k=4), "bcc": lambda: gen.bcc(9, 5)}[case]()
    g = eb.EmuGrid(pts, BOX, table_radius=-1)
    e = g.clip(count=False)
    ok = _check(e, g.oracle_cells(), case, expect_all=False)
    assert ok.mean() > 0.95
    assert all(v == 0 for v in e.counters.values())


@pytest.mark.parametrize("flags", [1, 2])
def test_cut_variants_are_identical(eb, gen, flags):
    """flags bit 0: serial walk only (TESS_FORCE_SERIAL), bit 1: table sweep instead of adjacency lists."""
    pts = np.concatenate([gen.uniform(1500, 5), gen.bcc(6, 5)])
    g = eb.EmuGrid(pts, BOX)
    r = g.oracle_cells()
    _check(g.clip(flags=flags), r, f"flags {flags}")


@pytest.mark.parametrize("reverse", [False, True])
def test_exact_ties_take_the_serial_walk(eb, gen, reverse):
    """Un-jittered lattices: bisectors through vertices and edges (Incident vertices, valence > 3, SURVEY D17 skips)."""
    for pts in (gen.simple_cubic(6), gen.bcc(5, 5, jitter=0.0)):
        g = eb.EmuGrid(pts, BOX, table_radius=-1)
        e = g.clip(reverse=reverse)
        r = g.oracle_cells()
        _check(e, r, "lattice")
        assert e.counters["degenerate_skips"] == r.counters["degenerate_skips"]


@pytest.mark.parametrize("reverse", [False, True])
def test_large_configuration(eb, gen, reverse):
    """A particle inside a dense shell (hundreds of faces) overflows the small tables and is redone by LargeCfg."""
    u = gen.uniform(150, 54)
    th, ph = np.arccos(2 * u[:, 0] - 1), 2 * np.pi * u[:, 1]
    shell = 0.5 + 0.3 * np.stack([np.sin(th) * np.cos(ph), np.sin(th) * np.sin(ph), np.cos(th)], axis=1)
    bg = gen.uniform(600, 55)
    pts = np.concatenate([[[0.5, 0.5, 0.5]], shell, bg[np.linalg.norm(bg - 0.5, axis=1) > 0.35]])
    g = eb.EmuGrid(pts, BOX, table_radius=-1)
    e = g.clip(reverse=reverse)
    assert e.n_failed >= 1 and (e.status & 0x4).any()  # the centre cell does not fit the small tables
    r = g.oracle_cells()
    _check(e, r, "small pass", expect_all=False)
    slots = np.sort(e.failed_slots)
    el = g.clip(work_slots=slots, large=True, reverse=reverse)
    rl = g.oracle_cells(slots=slots)
    assert max(np.diff(rl.face_offsets)) > 64
    _check(el, rl, "large pass")
    # the medium configuration (256 vertices / 128 faces) cannot hold the centre cell either, but holds the rest
    em = g.clip(work_slots=slots, large="medium", reverse=reverse)
    assert (em.status & 0x4).any()
    _check(em, rl, "medium pass", expect_all=False)
    # and both on ordinary cells
    some = np.arange(0, g.n, 7, dtype=np.uint32)
    _check(g.clip(work_slots=some, large=True, reverse=reverse), g.oracle_cells(slots=some), "large on ordinary cells")
    _check(g.clip(work_slots=some, large="medium", reverse=reverse), g.oracle_cells(slots=some), "medium on ordinary cells")


def test_medium_configuration_on_cluster_rims(eb, gen):
    """Cells of a clustered set that overflow the small tables on the way (not in the end) are finished by MediumCfg."""
    pts = gen.clustered(30000, 4)
    g = eb.EmuGrid(pts, BOX)
    e = g.clip()
    slots = np.sort(e.failed_slots)
    assert len(slots) > 20
    em = g.clip(work_slots=slots, large="medium")
    rm = g.oracle_cells(slots=slots)
    ok = _check(em, rm, "medium", expect_all=False)
    assert ok.mean() > 0.9


def test_reference_radius_and_group_modes(eb, gen, ob):
    pts = gen.uniform(2000, 57)
    groups = (np.arange(len(pts)) % 3).astype(np.uint64)
    g = eb.EmuGrid(pts, BOX, groups=groups, table_radius=-1)
    sx = g.cell_info[0]
    for radius in (0.0, (1.5 * sx) ** 2):
        e = g.clip(search_radius=radius)
        r = g.oracle_cells(mode=ob.MODE_REFERENCE_RADIUS, search_radius=radius)
        _check(e, r, f"radius {radius}")
        assert e.counters["tested"] == r.counters["tested"]
    for tg in (0, 2):
        _check(g.clip(target_group=tg), g.oracle_cells(target_group=tg), f"group {tg}")


@pytest.mark.parametrize("large", [False, "medium", True])
def test_vertex_lists_and_face_loops(eb, gen, large):
    """TESS_OUT_VERTICES: per-face ordered vertex loops bit-identical to
    Polyhedron::compute_face_vertices; the cell's vertex list is the same set of points."""
    pts = gen.uniform(1200, 64)
    g = eb.EmuGrid(pts, BOX, table_radius=-1)
    slots = np.arange(0, g.n, 3 if large else 1, dtype=np.uint32)
    e = g.clip(work_slots=slots, large=large, want_vertices=True)
    r = g.oracle_cells(slots=slots, want_vertices=True)
    _check(e, r, "geometry")
    assert np.array_equal(e.geo["nv"], np.diff(r.vertex_offsets))
    fo = r.face_offsets
    for c in range(0, len(slots), 7):
        assert {tuple(v) for v in e.cell_vertices(c)} == {tuple(v) for v in r.cell_vertices(c)}
        for j in range(int(fo[c + 1] - fo[c])):
            assert np.array_equal(e.face_loop(c, j), r.face_loop(int(fo[c]) + j)), (c, j)


def test_cells_at_query_points(eb, gen):
    """get_cell_at_particle (interface.rs:211-232): a cell around an arbitrary position, no self exclusion."""
    pts = gen.uniform(2000, 59)
    g = eb.EmuGrid(pts, BOX, table_radius=-1)
    q = gen.uniform(48, 60)
    e = g.clip(query_xyz=q)
    for i in range(len(q)):
        r = g.oracle.compute_cell_at_point(*q[i])
        nf = int(e.nfaces[i])
        assert e.neighbors[e.face_offsets[i]: e.face_offsets[i] + nf].tolist() == r.neighbors.tolist()
        assert e.volumes[i] == r.volumes[0] and np.array_equal(e.areas[e.face_offsets[i]: e.face_offsets[i] + nf], r.areas)


@pytest.mark.parametrize("scale", [1e-3, 1.0, 1e3])
def test_other_scales_and_oblong_boxes(eb, gen, scale):
    """The tolerance of the reference is absolute (1e-12) while the candidate screen's margin is relative to the
    cell: small, large and strongly anisotropic containers must stay bit-identical, counters included."""
    pts = gen.uniform(2500, 71) * np.array([scale, 3.0 * scale, 0.25 * scale]) + np.array([5.0 * scale, -scale, 0.0])
    box = (5.0 * scale, -scale, 0.0, 6.0 * scale, 2.0 * scale, 0.25 * scale)
    g = eb.EmuGrid(pts, box, table_radius=-1)
    e = g.clip()
    r = g.oracle_cells()
    _check(e, r, f"scale {scale}")
    for k in ("visited", "tested", "vertex_classifications", "cuts", "new_vertices", "faces"):
        assert e.counters[k] == r.counters[k], k
    assert abs(e.volumes.sum() / (0.75 * scale ** 3) - 1.0) < 1e-12


def test_slab_with_a_thin_halo_flags_the_same_cells_in_both_instantiations(eb, gen):
    """A rank that holds too few ghost planes: a search that reaches a plane it does not hold flags the cell
    (TESS_STATUS_HALO_INSUFFICIENT) instead of being silently wrong.  The instantiation without work counters lets
    its termination threshold lag behind the cuts; the flags must still be exactly those of the counting one."""
    pts = gen.uniform(6000, 77)
    g = eb.EmuGrid(pts, BOX)
    whole = g.oracle_cells(slots=None)  # rows in global grid order
    gid_of_row = g.sorted_indices.copy()
    lo, hi = g.cpd // 3, g.cpd // 3 + 5
    ids = g.restrict_to_planes(pts, (lo, hi))
    cpd = g.cpd
    gx = g.oracle.cells().astype(np.int64)[ids] // (cpd * cpd)
    own = np.nonzero((gx >= lo + 1) & (gx < hi - 1))[0].astype(np.uint32)  # one ghost plane per side: too thin
    a = g.clip(work_slots=own, count=True)
    b = g.clip(work_slots=own, count=False)
    flagged = (a.status & 0x8) != 0
    assert flagged.any() and not flagged.all()
    assert np.array_equal(a.status, b.status)
    assert np.array_equal(a.volumes, b.volumes) and np.array_equal(a.neighbors, b.neighbors) and np.array_equal(a.areas, b.areas)
    # unflagged cells are the cells of the whole diagram
    row_of_gid = np.empty(len(pts), np.int64)
    row_of_gid[gid_of_row] = np.arange(len(pts))
    rows = row_of_gid[ids[own]]
    okc = ~flagged & ((a.status & BAD) == 0)
    assert okc.sum() > 100
    assert np.array_equal(a.volumes[okc], whole.volumes[rows[okc]])
    fo = whole.face_offsets
    for k in np.nonzero(okc)[0][::37]:
        assert a.neighbors[a.face_offsets[k]: a.face_offsets[k + 1]].tolist() == whole.neighbors[fo[rows[k]]: fo[rows[k] + 1]].tolist()


def test_small_configuration_without_the_serial_walk(eb, gen):
    """CLIP_SMALL_FAST: the instantiation that has no serial walk finishes every cell whose planes never touch a
    vertex, bit for bit, and hands the others back flagged like cells that ran out of table; CLIP_SMALL finishes those."""
    pts = np.concatenate([gen.uniform(2500, 5), 0.25 + 0.5 * gen.simple_cubic(6)])  # random cells around an exact lattice
    g = eb.EmuGrid(pts, BOX, table_radius=-1)
    r = g.oracle_cells()
    for count in (False, True):
        e = g.clip(large="fast", count=count)
        handed_back = (e.status & 0x2) != 0
        assert handed_back.any() and (~handed_back).sum() > 1500
        assert not (e.status & (0x4 | 0x10)).any()
        ok = _check(e, r, "fast", expect_all=False)
        assert np.array_equal(ok, ~handed_back)
        slots = np.nonzero(handed_back)[0].astype(np.uint32)
        e2 = g.clip(work_slots=slots, count=count)
        _check(e2, g.oracle_cells(slots=slots), "redone with the serial walk")
    # purely random input never needs the walk
    g2 = eb.EmuGrid(gen.uniform(4000, 6), BOX)
    e = g2.clip(large="fast", count=False)
    assert not (e.status & 0x2).any() or g2.table_full == 0
    _check(e, g2.oracle_cells(), "fast on random input", expect_all=False)


@pytest.mark.parametrize("case", ["uniform", "clustered", "bcc", "cubic", "on_walls"])
@pytest.mark.parametrize("reverse", [False, True])
def test_thread_per_cell_kernel(eb, gen, case, reverse):
    """CLIP_THREAD (clip_thread.cu): consumer warps build one cell per thread, producer warps walk the search table and
    feed them through per-lane rings in shared memory.  Every cell it finishes is the oracle's cell bit for bit; what its
    tables cannot hold — and every cell that meets a vertex ON a plane (the exact lattice) — is handed back flagged like a
    cell that ran out of search table, and the warp-per-cell kernel finishes those."""
    base = gen.uniform(2500, 81)
    pts = {
        "uniform": lambda: gen.uniform(3000, 51),
        "clustered": lambda: gen.clustered(4000, 4, k=4),  # long particle runs per grid cell: the producer's partial steps
        "bcc": lambda: gen.bcc(9, 5),
        "cubic": lambda: np.concatenate([gen.uniform(1500, 5), 0.25 + 0.5 * gen.simple_cubic(6)]),
        "on_walls": lambda: np.concatenate([base, np.round(gen.uniform(300, 82), 0) * np.array([1, 1, 0]) + gen.uniform(300, 83) * np.array([0, 0, 1])]),
    }[case]()
    g = eb.EmuGrid(pts, BOX, table_radius=-1)
    r = g.oracle_cells()
    e = g.clip(large="thread", count=False, reverse=reverse)
    handed_back = (e.status & 0x2) != 0
    assert not (e.status & (0x4 | 0x10)).any()
    ok = _check(e, r, "thread " + case, expect_all=False)
    assert np.array_equal(ok, ~handed_back)
    assert ok.mean() > {"uniform": 0.95, "clustered": 0.75, "bcc": 0.99, "cubic": 0.6, "on_walls": 0.85}[case]
    if handed_back.any():
        slots = np.nonzero(handed_back)[0].astype(np.uint32)
        e2 = g.clip(work_slots=slots, count=False)
        _check(e2, g.oracle_cells(slots=slots), "handed back to the warp-per-cell kernel", expect_all=(case != "clustered"))


def test_thread_per_cell_kernel_truncated_table_and_thin_halo(eb, gen):
    """The producers run ahead of the cell builders with a stale (larger) threshold; the end of a truncated table and a
    missing ghost plane must still be flagged for exactly the cells the reference-shaped walk flags."""
    pts = gen.uniform(6000, 77)
    g = eb.EmuGrid(pts, BOX, table_radius=3)  # some cells terminate inside it, the others run off its end
    a = g.clip(count=False)
    t = g.clip(large="thread", count=False)
    exhausted = (a.status & 0x2) != 0
    assert exhausted.any() and not exhausted.all()
    same = ~exhausted & ((t.status & 0x2) == 0)  # (the thread kernel also hands back what its tables cannot hold)
    assert np.array_equal((t.status & 0x2) != 0, exhausted | ((t.status & 0x2) != 0)) and same.sum() > 500
    assert np.all(((t.status & 0x2) != 0)[exhausted])
    assert np.array_equal(t.volumes[same], a.volumes[same])
    # thin halo
    g = eb.EmuGrid(pts, BOX)
    lo, hi = g.cpd // 3, g.cpd // 3 + 5
    ids = g.restrict_to_planes(pts, (lo, hi))
    cpd = g.cpd
    gx = g.oracle.cells().astype(np.int64)[ids] // (cpd * cpd)
    own = np.nonzero((gx >= lo + 1) & (gx < hi - 1))[0].astype(np.uint32)
    a = g.clip(work_slots=own, count=True)
    t = g.clip(work_slots=own, count=False, large="thread")
    done = (t.status & 0x2) == 0
    assert ((a.status & 0x8) != 0).any() and done.mean() > 0.9
    assert np.array_equal(t.status[done], a.status[done])
    assert np.array_equal(t.volumes[done], a.volumes[done])


def test_thread_per_cell_kernel_radius_and_group_modes(eb, gen, ob):
    pts = gen.uniform(3000, 58)
    groups = (np.arange(len(pts)) % 3).astype(np.uint64)
    g = eb.EmuGrid(pts, BOX, groups=groups, table_radius=-1)
    sx = g.cell_info[0]
    for kw in (dict(search_radius=(1.5 * sx) ** 2), dict(target_group=1), dict(target_group=7)):
        a = g.clip(count=False, **kw)
        t = g.clip(large="thread", count=False, **kw)
        done = (t.status & 0x2) == 0
        assert done.mean() > 0.9, kw
        assert np.array_equal(t.status[done], a.status[done]), kw
        fa, ft = a.face_offsets.astype(np.int64), t.face_offsets.astype(np.int64)
        assert np.array_equal(np.diff(ft)[done], np.diff(fa)[done]), kw
        assert np.array_equal(t.volumes[done], a.volumes[done]), kw
