"""Pins the oracle's start cube and cut decision against polyhedron.rs's unit tests (:951-1118)
and checks the repaired clipper on cuts whose exact result is known."""
import numpy as np
import pytest


def test_new_and_reset(ob):  # polyhedron.rs:952-1006
    p = ob.Polyhedron(-3.0, 40.0, -0.2, 0.0, 100.0, -0.1)
    c = p.counts()
    assert (c["edges"], c["vertices"], c["faces"], c["face_data"], c["root_edge"]) == (24, 8, 6, 0, 0)
    for v in range(8):
        x, y, z = p.vertex(v)
        assert x in (0.0, -3.0) and y in (100.0, 40.0) and z in (-0.1, -0.2)
    p.reset(-1, -1, -1, 1, 1, 1)
    c = p.counts()
    assert (c["edges"], c["vertices"], c["faces"], c["face_data"], c["root_edge"]) == (24, 8, 6, 0, 0)
    for v in range(8):
        assert all(abs(t) == 1.0 for t in p.vertex(v))


def test_build_cube(ob):  # polyhedron.rs:1009-1033 + the numbering of :97-199, :288-383
    p = ob.Polyhedron(-1, -1, -1, 1, 1, 1)
    for e in range(24):
        he = p.edge(e)
        assert p.edge(he["flip"])["flip"] == e
        assert he["face"] == e // 4  # four half-edges per face, in F R B L U D order (D5: DR is in D)
    # vertex order FDL FDR FUR FUL BDL BDR BUR BUL (polyhedron.rs:288-295)
    exp = [(-1, -1, -1), (1, -1, -1), (1, -1, 1), (-1, -1, 1), (-1, 1, -1), (1, 1, -1), (1, 1, 1), (-1, 1, 1)]
    for v, e in enumerate(exp):
        assert tuple(p.vertex(v)) == e
    # FU: flip UF(16), target FUL(3), next FL(1) (polyhedron.rs:320)
    assert p.edge(0) == dict(flip=16, next=1, target=3, face=0)
    # every face loop has an outward normal of magnitude 2*area = 8 and the volume is 8
    for f, n in enumerate([(0, -8, 0), (8, 0, 0), (0, 8, 0), (-8, 0, 0), (0, 0, 8), (0, 0, -8)]):
        assert tuple(p.weighted_normal(f)) == n
    assert p.check() == 0
    assert p.volume() == 8.0


def test_find_outgoing_edge_no_cut(ob):  # polyhedron.rs:1036-1053
    p = ob.Polyhedron(-1, -1, -1, 1, 1, 1)
    assert p.find_outgoing_edge(ob.plane_from_normal_point([1, 1, 1], [2, 2, 2])) is None


def test_find_outgoing_edge_with_cut(ob):  # polyhedron.rs:1056-1096
    p = ob.Polyhedron(-1, -1, -1, 1, 1, 1)
    e = p.find_outgoing_edge(ob.plane_from_normal_point([1, 1, 1], [0.5, 0.5, 0.5]))
    he = p.edge(e)
    assert tuple(p.vertex(he["target"])) == (1, 1, 1)
    src = tuple(p.vertex(p.edge(he["flip"])["target"]))
    assert src in ((-1, 1, 1), (1, -1, 1), (1, 1, -1))
    assert e == 17  # first hit in slot order is RU(4); its flip is UR(17)


def test_cut_with_plane_corner(ob):  # polyhedron.rs:1099-1118 (the reference only checks "no panic")
    p = ob.Polyhedron(-1, -1, -1, 1, 1, 1)
    assert p.cut_with_plane(100, ob.plane_halfway([1, 1, 1]))
    assert p.check() == 0
    assert p.live_counts() == dict(edges=30, vertices=10, faces=7)
    assert p.volume() == 8 - 1.5 ** 3 / 6  # exact: the corner tetrahedron has legs 1.5
    assert p.face(6) == dict(neighbor=100, starting_edge=24)  # the new cap face carries the point index
    # cap triangle: area sqrt(3)/4 * (1.5*sqrt2)^2
    wn = p.weighted_normal(6)
    assert abs(0.5 * np.linalg.norm(wn) - np.sqrt(3) / 4 * (1.5 * np.sqrt(2)) ** 2) < 1e-14
    assert np.all(wn > 0)  # outward (+n)


def test_cut_with_plane_slab_and_sequence(ob):
    p = ob.Polyhedron(-1, -1, -1, 1, 1, 1)
    assert p.cut_with_plane(7, np.array([1.0, 0, 0, 0.5]))  # x <= 0.5 (second trace)
    assert p.check() == 0 and p.live_counts() == dict(edges=24, vertices=8, faces=6) and p.volume() == 6.0
    assert not p.cut_with_plane(8, np.array([1.0, 0, 0, 0.75]))  # nothing outside
    assert p.cut_with_plane(9, np.array([0.0, -1.0, 0, 0.25]))  # y >= -0.25
    assert p.check() == 0 and p.volume() == 1.5 * 1.25 * 2
    nb = sorted(p.face(f)["neighbor"] for f in range(p.counts()["faces"]) if p.face(f))
    assert nb == [-6, -5, -4, -3, 7, 9]  # walls z_min, z_max, x_min, y_max + the two cutters


def test_plane_through_edge_and_d17(ob):
    """Exact-degeneracy behaviour the GPU must reproduce."""
    p = ob.Polyhedron(0, 0, 0, 1, 1, 1)
    # plane x + y = 2 touches the cube along an edge: all Incident -> no cut
    assert not p.cut_with_plane(1, ob.plane_from_normal_point([1, 1, 0], [1, 1, 0]))
    assert p.volume() == 1.0
    # plane x + y = 1 passes through 4 vertices: vertices are Outside but no strictly
    # Inside->Outside edge exists -> the reference skips the plane (volume stays 1)
    assert p.find_outgoing_edge(ob.plane_from_normal_point([1, 1, 0], [0.5, 0.5, 0])) is None
    assert not p.cut_with_plane(2, ob.plane_from_normal_point([1, 1, 0], [0.5, 0.5, 0]))
    assert p.volume() == 1.0


def test_plane_through_vertex_copies_incident_vertex(ob):
    """polyhedron.rs:555-565: an Incident vertex reached by the walk is re-created as a copy.
    Plane x + y + z/2 = 2 passes exactly through BDR (1,1,0); only BUR (1,1,1) is Outside."""
    p = ob.Polyhedron(0, 0, 0, 1, 1, 1)
    pl = ob.plane_from_normal_point([1, 1, 0.5], [1, 1, 0])
    assert ob.lib().orc_vector_location(pl, np.array([1.0, 1.0, 0.0]), 1e-12) == 1  # Incident, exactly
    assert p.cut_with_plane(5, pl)
    assert p.check() == 0
    assert abs(p.volume() - (1 - 1 / 24)) < 1e-15
    lc = p.live_counts()
    assert lc["vertices"] == 9 and lc["faces"] == 7
    pos = [tuple(np.round(p.vertex(v), 12)) for v in range(p.counts()["vertices"]) if p.vertex(v) is not None]
    assert (1.0, 1.0, 1.0) not in pos
    assert pos.count((1.0, 1.0, 0.0)) == 1  # destroyed, then re-created once as a copy
    assert (0.5, 1.0, 1.0) in pos and (1.0, 0.5, 1.0) in pos
    cap = [f for f in range(p.counts()["faces"]) if p.face(f) and p.face(f)["neighbor"] == 5][0]
    assert len(p.face_vertices(cap)) == 3


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_random_cut_sequences_keep_the_mesh_closed(ob, gen, seed):
    """Euler V - E/2 + F = 2, flip/next consistency and sum(area*n) = 0 after every cut."""
    p = ob.Polyhedron(-1, -1, -1, 1, 1, 1)
    pts = (gen.uniform(60, 100 + seed) - 0.5) * 3.0
    vol = p.volume()
    for i, q in enumerate(pts):
        if p.cut_with_plane(i, ob.plane_halfway(q)):
            assert p.check() == 0
            v2 = p.volume()
            assert v2 < vol + 1e-15
            vol = v2
            tot = sum((p.weighted_normal(f) for f in range(p.counts()["faces"]) if p.face(f)), np.zeros(3))
            assert np.linalg.norm(tot) < 1e-13
    assert vol > 0
