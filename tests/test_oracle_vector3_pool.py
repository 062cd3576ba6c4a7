"""Pins the oracle's vector/plane math and pool against the reference's own unit tests.

Each test replays one `#[test]` of /root/reference/src/vector3.rs (:315-476) or pool.rs (:215-1295)
with the same inputs and the same expected values (exact f64 equality where the reference uses
assert_eq!).
"""
import numpy as np

V = lambda *a: np.array(a, dtype=np.float64)  # noqa: E731


# ---------------------------------------------------------------- vector3.rs ---------------
def test_dot(ob):  # vector3.rs:316-322
    L = ob.lib()
    assert L.orc_dot(V(1, 2, 3), V(4, 5, 6)) == 32.0
    assert L.orc_dot(V(4, 5, 6), V(1, 2, 3)) == 32.0


def test_scale(ob):  # vector3.rs:325-331
    L = ob.lib()
    for s, exp in ((1.0, V(1, -2, 0)), (2.0, V(2, -4, 0)), (-3.0, V(-3, 6, 0))):
        o = np.zeros(3)
        L.orc_scale(V(1, -2, 0), s, o)
        assert np.array_equal(o, exp)


def test_add_sub(ob):  # vector3.rs:334-349
    L = ob.lib()
    v1, v2, o = V(-4.5, 0.0, 200.1), V(2.0, -3.4, 4.1), np.zeros(3)
    L.orc_add(v1, v2, o)
    assert np.array_equal(o, V(-2.5, -3.4, 204.2))
    L.orc_add(v2, v1, o)
    assert np.array_equal(o, V(-2.5, -3.4, 204.2))
    L.orc_sub(v1, v2, o)
    assert np.array_equal(o, V(-6.5, 3.4, 196.0))
    L.orc_sub(v2, v1, o)
    assert np.array_equal(o, V(6.5, -3.4, -196.0))


def test_cross(ob):  # vector3.rs:44-50 (no reference test; right-handedness + formula)
    o = np.zeros(3)
    ob.lib().orc_cross(V(1, 0, 0), V(0, 1, 0), o)
    assert np.array_equal(o, V(0, 0, 1))
    ob.lib().orc_cross(V(1, 2, 3), V(4, 5, 6), o)
    assert np.array_equal(o, V(2 * 6 - 3 * 5, 3 * 4 - 1 * 6, 1 * 5 - 2 * 4))


OUTSIDE, INCIDENT, INSIDE = 0, 1, 2


def test_location(ob):  # vector3.rs:352-364
    L = ob.lib()
    assert L.orc_location(1.0, 0.01) == OUTSIDE
    assert L.orc_location(-1.0, 0.01) == INSIDE
    assert L.orc_location(0.005, 0.01) == INCIDENT
    # boundary: `> tol` / `< -tol` are strict (vector3.rs:171-173)
    assert L.orc_location(0.01, 0.01) == INCIDENT
    assert L.orc_location(-0.01, 0.01) == INCIDENT


def test_vector_location(ob):  # vector3.rs:367-383
    L = ob.lib()
    p = V(1, 0, 0, 1)  # plane x = 1
    assert L.orc_vector_location(p, V(4, 2, -6), 0.05) == OUTSIDE
    assert L.orc_vector_location(p, V(-3, -3, -3), 0.05) == INSIDE
    assert L.orc_vector_location(p, V(1, -6, 0), 0.05) == INCIDENT


def test_intersection(ob):  # vector3.rs:386-397 — exact
    o = np.zeros(3)
    ob.lib().orc_intersection(V(1, 0, 0, 1), V(20, 0, 0), V(10, 10, 0), o)
    assert np.array_equal(o, V(1, 19, 0))


def test_bbox(ob):  # vector3.rs:400-437
    L = ob.lib()
    lo, hi = np.zeros(3), np.zeros(3)
    L.orc_bbox_adjust(lo, hi, -1.0, 2.0, 0.0)
    assert np.array_equal(lo, V(-1, 0, 0)) and np.array_equal(hi, V(0, 2, 0))
    L.orc_bbox_adjust(lo, hi, -2.0, -3.0, 1.0)
    assert np.array_equal(lo, V(-2, -3, 0)) and np.array_equal(hi, V(0, 2, 1))
    L.orc_bbox_adjust(lo, hi, 0.0, 0.0, 0.0)
    assert np.array_equal(lo, V(-2, -3, 0)) and np.array_equal(hi, V(0, 2, 1))
    lo, hi = np.zeros(3), np.zeros(3)
    L.orc_bbox_pad(lo, hi, 0.5)
    assert np.array_equal(lo, V(-0.5, -0.5, -0.5)) and np.array_equal(hi, V(0.5, 0.5, 0.5))


def test_halfway_from_origin_to(ob):  # vector3.rs:440-476 — bitwise plane equality
    for sgn in ((1, 1, 1), (-1, 1, 1), (-1, -1, -1)):
        pt = V(*sgn)
        a = ob.plane_halfway(pt)
        b = ob.plane_from_normal_point(0.5 * pt, 0.5 * pt)
        assert np.array_equal(a, b), (a, b)


def test_saturating_cast_and_cbrt(ob):  # float.rs:138-142, celery.rs:161-162
    L = ob.lib()
    assert L.orc_to_usize(-3.5) == 0
    assert L.orc_to_usize(float("nan")) == 0
    assert L.orc_to_usize(2.999) == 2
    assert L.orc_to_usize(1e300) == 2 ** 64 - 1
    # cells_per_dimension pinned by celery.rs:1230,1269,1349,1308 and the 79 -> 4x4x4 comment (:1460)
    for n, cpd in ((1, 1), (79, 4), (100, 5), (1000, 10), (1_000_000, 93)):
        assert L.orc_cells_per_dimension(n) == cpd
    # exact cubes: N/1.25 = 20^3 and 200^3; glibc cbrt is exact there
    assert L.orc_cells_per_dimension(10_000) == 21
    assert L.orc_cells_per_dimension(10_000_000) == 201
    assert L.orc_cells_per_dimension(99_672_064) == 431


# ---------------------------------------------------------------- pool.rs -------------------
VALUE, NEXT, END = 0, 1, 2


def _five(ob):
    p = ob.Pool()
    for i in range(5):
        assert p.add(i) == i  # pool.rs:276-325 add_five
    return p


def test_pool_initial_and_add_one(ob):  # pool.rs:215-253
    p = ob.Pool()
    assert len(p) == 0 and p.first is None
    assert p.add(7) == 0
    assert len(p) == 1 and p.first is None and p.chunk(0) == (VALUE, 7)


def test_pool_remove_one_from_one(ob):  # pool.rs:255-273
    p = ob.Pool()
    p.add(1)
    p.remove(0)
    assert len(p) == 1 and p.first == 0 and p.chunk(0)[0] == END


def test_pool_remove_one_from_five(ob):  # pool.rs:327-375
    p = _five(ob)
    p.remove(2)
    assert len(p) == 5 and p.first == 2
    assert [p.chunk(i)[0] for i in range(5)] == [VALUE, VALUE, END, VALUE, VALUE]


def test_pool_remove_two_from_five(ob):  # pool.rs:378-427
    p = _five(ob)
    p.remove(2)
    p.remove(4)
    assert p.first == 4 and p.chunk(2)[0] == END and p.chunk(4) == (NEXT, 2)
    assert [p.chunk(i) for i in (0, 1, 3)] == [(VALUE, 0), (VALUE, 1), (VALUE, 3)]


def test_pool_remove_two_from_five_reverse(ob):  # pool.rs:430-479
    p = _five(ob)
    p.remove(4)
    p.remove(2)
    assert p.first == 2 and p.chunk(2) == (NEXT, 4) and p.chunk(4)[0] == END


def test_pool_remove_all(ob):  # pool.rs:482-589 (both orders): free list is LIFO
    p = _five(ob)
    for i in range(5):
        p.remove(i)
    assert p.first == 4
    assert [p.chunk(i) for i in range(5)] == [(END, 0), (NEXT, 0), (NEXT, 1), (NEXT, 2), (NEXT, 3)]
    p = _five(ob)
    for i in reversed(range(5)):
        p.remove(i)
    assert p.first == 0
    assert [p.chunk(i) for i in range(5)] == [(NEXT, 1), (NEXT, 2), (NEXT, 3), (NEXT, 4), (END, 0)]


def test_pool_replace(ob):  # pool.rs:592-913: a freed slot is reused, most recently freed first
    p = ob.Pool()
    p.add(1)
    p.remove(0)
    assert p.add(2) == 0 and p.first is None and p.chunk(0) == (VALUE, 2)
    p = _five(ob)
    p.remove(2)
    assert p.add(9) == 2 and p.first is None
    p = _five(ob)
    p.remove(2)
    p.remove(4)
    assert p.add(10) == 4 and p.first == 2
    assert p.add(11) == 2 and p.first is None
    assert p.add(12) == 5 and len(p) == 6
    p = _five(ob)
    for i in range(5):
        p.remove(i)
    assert [p.add(100 + k) for k in range(5)] == [4, 3, 2, 1, 0]
    assert p.first is None


def test_pool_iteration_skips_holes(ob):  # pool.rs:916-1295
    p = ob.Pool()
    assert p.values() == []
    p = _five(ob)
    assert p.values() == [0, 1, 2, 3, 4]
    p.remove(1)
    p.remove(3)
    assert p.values() == [0, 2, 4]
    assert p.has(0) and not p.has(1)
    assert p.add(31) == 3 and p.values() == [0, 2, 31, 4]
    for i in (0, 2, 3, 4):
        p.remove(i)
    assert p.values() == []
