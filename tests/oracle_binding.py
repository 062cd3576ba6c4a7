"""ctypes binding to the CPU oracle (oracle/liboracle.so).

TEST INFRASTRUCTURE: only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
`--impl reference` legs import this.  The product package never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
ORACLE_DIR = os.path.join(os.path.dirname(_HERE), "oracle")
_LIB_PATH = os.path.join(ORACLE_DIR, "liboracle.so")

MODE_NO_RADIUS, MODE_REFERENCE_RADIUS, MODE_SECURITY = 0, 1, 2
STATUS_DEGENERATE_SKIP, STATUS_TABLE_EXHAUSTED = 1, 2

_u64p = np.ctypeslib.ndpointer(dtype=np.uint64, flags="C_CONTIGUOUS")
_f64p = np.ctypeslib.ndpointer(dtype=np.float64, flags="C_CONTIGUOUS")
_i32p = np.ctypeslib.ndpointer(dtype=np.int32, flags="C_CONTIGUOUS")
_i64p = np.ctypeslib.ndpointer(dtype=np.int64, flags="C_CONTIGUOUS")


def _host_tag() -> str:
    """The oracle is built with -march=native (the CPU baseline at the host's best): a library built on another
    machine (the prebuilt file travels with the repo) is rebuilt for the cores it runs on."""
    import hashlib

    try:
        with open("/proc/cpuinfo") as f:
            flags = next((ln for ln in f if ln.startswith("flags")), "")
    except OSError:
        flags = ""
    return hashlib.sha256(flags.encode()).hexdigest()[:16]


def build(force: bool = False) -> str:
    srcs = [os.path.join(ORACLE_DIR, f) for f in ("tess_oracle.cpp", "tess_oracle_capi.cpp", "tess_oracle.hpp", "Makefile")]
    tag_path = _LIB_PATH + ".host"
    tag = _host_tag()
    built_for = open(tag_path).read().strip() if os.path.exists(tag_path) else ""
    stale = force or not os.path.exists(_LIB_PATH) or built_for != tag or any(os.path.getmtime(s) > os.path.getmtime(_LIB_PATH) for s in srcs)
    if stale:
        subprocess.check_call(["make", "-B", "-C", ORACLE_DIR, "liboracle.so"], stdout=subprocess.DEVNULL)
        with open(tag_path, "w") as f:
            f.write(tag)
    return _LIB_PATH


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    L = C.CDLL(build())
    vp, u64, f64, i64, ci = C.c_void_p, C.c_uint64, C.c_double, C.c_int64, C.c_int

    def sig(name, res, *args):
        fn = getattr(L, name)
        fn.restype = res
        fn.argtypes = list(args)

    sig("orc_last_error", C.c_char_p)
    sig("orc_diagram_create", vp, vp, u64, vp, vp, ci)
    sig("orc_diagram_destroy", None, vp)
    for n in ("orc_grid_cpd", "orc_grid_num_points", "orc_grid_num_delimiters", "orc_grid_search_order_len"):
        sig(n, u64, vp)
    sig("orc_grid_table_is_full", ci, vp)
    sig("orc_grid_bounds", None, vp, _f64p)
    sig("orc_grid_cell_info", None, vp, _f64p)
    sig("orc_container_box", None, vp, _f64p)
    sig("orc_grid_copy_cells", None, vp, _u64p)
    sig("orc_grid_copy_sorted_indices", None, vp, _u64p)
    sig("orc_grid_copy_delimiters", None, vp, _u64p)
    sig("orc_grid_copy_search_order", None, vp, _f64p, _i32p)
    sig("orc_grid_home_cell", None, vp, f64, f64, f64, _u64p)
    sig("orc_check_cell_in_range", ci, vp, f64, f64, f64, f64, u64, u64, u64)
    for n in ("orc_find_cells_in_radius", "orc_find_neighbors_in_cell_radius", "orc_find_neighbors_in_real_radius"):
        sig(n, u64, vp, f64, f64, f64, f64, _u64p, u64)
    sig("orc_search_create", vp, vp, f64, f64, f64)
    sig("orc_search_destroy", None, vp)
    sig("orc_search_home", None, vp, _u64p)
    sig("orc_search_expand", u64, vp, f64, u64, _u64p, u64)
    sig("orc_search_expand_all_no_radius", u64, vp, _u64p, u64)
    sig("orc_search_expand_all_in_radius", u64, vp, f64, _u64p, u64)
    sig("orc_compute_cells", vp, vp, vp, u64, ci, f64, i64, ci, ci)
    sig("orc_compute_cell_at_point", vp, vp, f64, f64, f64, ci, f64, i64, ci)
    sig("orc_result_free", None, vp)
    sig("orc_result_n_cells", u64, vp)
    for n, t in (
        ("orc_result_volumes", f64), ("orc_result_face_offsets", u64), ("orc_result_neighbors", i64), ("orc_result_areas", f64),
        ("orc_result_status", C.c_uint32), ("orc_result_max_radius_sq", f64), ("orc_result_vertex_offsets", u64),
        ("orc_result_vertices", f64), ("orc_result_counters", u64),
    ):
        sig(n, C.POINTER(t), vp)
    sig("orc_result_n_loops", u64, vp)
    sig("orc_result_loop_offsets", C.POINTER(u64), vp)
    sig("orc_result_loop_vertices", C.POINTER(f64), vp)
    sig("orc_dot", f64, _f64p, _f64p)
    for n in ("orc_cross", "orc_add", "orc_sub"):
        sig(n, None, _f64p, _f64p, _f64p)
    sig("orc_scale", None, _f64p, f64, _f64p)
    sig("orc_location", ci, f64, f64)
    sig("orc_vector_location", ci, _f64p, _f64p, f64)
    sig("orc_intersection", None, _f64p, _f64p, _f64p, _f64p)
    sig("orc_plane_halfway_from_origin_to", None, _f64p, _f64p)
    sig("orc_plane_from_non_unit_normal_and_point", None, _f64p, _f64p, _f64p)
    sig("orc_bbox_adjust", None, _f64p, _f64p, f64, f64, f64)
    sig("orc_bbox_pad", None, _f64p, _f64p, f64)
    sig("orc_to_usize", u64, f64)
    sig("orc_cells_per_dimension", u64, u64)
    sig("orc_pool_create", vp)
    sig("orc_pool_destroy", None, vp)
    sig("orc_pool_add", u64, vp, i64)
    sig("orc_pool_remove", None, vp, u64)
    sig("orc_pool_len", u64, vp)
    sig("orc_pool_first", i64, vp)
    sig("orc_pool_chunk", ci, vp, u64, C.POINTER(i64))
    sig("orc_pool_has", ci, vp, u64)
    sig("orc_pool_iterate", u64, vp, _i64p, u64)
    sig("orc_poly_create", vp, f64, f64, f64, f64, f64, f64)
    sig("orc_poly_destroy", None, vp)
    sig("orc_poly_reset", None, vp, f64, f64, f64, f64, f64, f64)
    sig("orc_poly_counts", None, vp, _u64p)
    sig("orc_poly_live_counts", None, vp, _u64p)
    sig("orc_poly_edge", ci, vp, u64, _u64p)
    sig("orc_poly_vertex", ci, vp, u64, _f64p)
    sig("orc_poly_face", ci, vp, u64, C.POINTER(i64), C.POINTER(u64))
    sig("orc_poly_find_outgoing_edge", i64, vp, _f64p)
    sig("orc_poly_cut_with_plane", ci, vp, u64, _f64p)
    sig("orc_poly_translate", None, vp, _f64p)
    sig("orc_poly_volume", f64, vp)
    sig("orc_poly_weighted_normal", None, vp, u64, _f64p)
    sig("orc_poly_face_vertices", u64, vp, u64, _f64p, u64)
    sig("orc_poly_check", ci, vp)
    _lib = L
    return L


def _err() -> str:
    return lib().orc_last_error().decode()


def vec(*a) -> np.ndarray:
    return np.array(a, dtype=np.float64)


class CellResults:
    """Copied-out results of orc_compute_cells (CSR over cells, faces in face-slot order)."""

    def __init__(self, handle):
        L = lib()
        m = int(L.orc_result_n_cells(handle))

        def arr(ptr, n, dt):
            return np.ctypeslib.as_array(ptr, shape=(n,)).astype(dt, copy=True) if n else np.zeros(0, dt)

        self.n = m
        self.volumes = arr(L.orc_result_volumes(handle), m, np.float64)
        self.face_offsets = arr(L.orc_result_face_offsets(handle), m + 1, np.uint64).astype(np.int64)
        nf = int(self.face_offsets[-1])
        self.neighbors = arr(L.orc_result_neighbors(handle), nf, np.int64)
        self.areas = arr(L.orc_result_areas(handle), nf, np.float64)
        self.status = arr(L.orc_result_status(handle), m, np.uint32)
        self.max_radius_sq = arr(L.orc_result_max_radius_sq(handle), m, np.float64)
        self.vertex_offsets = arr(L.orc_result_vertex_offsets(handle), m + 1, np.uint64).astype(np.int64)
        nv = int(self.vertex_offsets[-1])
        self.vertices = arr(L.orc_result_vertices(handle), 3 * nv, np.float64).reshape(nv, 3)
        nlo = int(L.orc_result_n_loops(handle))
        self.loop_offsets = arr(L.orc_result_loop_offsets(handle), nlo, np.uint64).astype(np.int64)
        nlv = int(self.loop_offsets[-1]) if nlo else 0
        self.loop_vertices = arr(L.orc_result_loop_vertices(handle), 3 * nlv, np.float64).reshape(nlv, 3)
        c = arr(L.orc_result_counters(handle), 8, np.uint64)
        self.counters = dict(
            visited=int(c[0]), tested=int(c[1]), vertex_classifications=int(c[2]), cuts=int(c[3]),
            new_vertices=int(c[4]), table_entries=int(c[5]), degenerate_skips=int(c[6]), faces=int(c[7]),
        )
        L.orc_result_free(handle)

    def cell_neighbors(self, c: int) -> np.ndarray:
        return self.neighbors[self.face_offsets[c]:self.face_offsets[c + 1]]

    def cell_areas(self, c: int) -> np.ndarray:
        return self.areas[self.face_offsets[c]:self.face_offsets[c + 1]]

    def cell_vertices(self, c: int) -> np.ndarray:
        return self.vertices[self.vertex_offsets[c]:self.vertex_offsets[c + 1]]

    def face_loop(self, k: int) -> np.ndarray:
        """Ordered vertices of global face k (needs want_vertices=True)."""
        return self.loop_vertices[self.loop_offsets[k]:self.loop_offsets[k + 1]]


class Diagram:
    """Oracle restatement of interface.rs `Diagram` (batch form)."""

    def __init__(self, points: np.ndarray, box=None, groups=None, table_radius: int = -1):
        L = lib()
        self._pts = np.ascontiguousarray(points, dtype=np.float64).reshape(-1, 3)
        self.n = self._pts.shape[0]
        g = None if groups is None else np.ascontiguousarray(groups, dtype=np.uint64)
        b = None if box is None else np.ascontiguousarray(box, dtype=np.float64)
        self._h = L.orc_diagram_create(
            self._pts.ctypes.data, self.n, None if g is None else g.ctypes.data, None if b is None else b.ctypes.data, table_radius
        )
        if not self._h:
            raise RuntimeError("oracle: " + _err())

    def close(self):
        if self._h:
            lib().orc_diagram_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- grid
    @property
    def cpd(self) -> int:
        return int(lib().orc_grid_cpd(self._h))

    @property
    def num_delimiters(self) -> int:
        return int(lib().orc_grid_num_delimiters(self._h))

    @property
    def search_order_len(self) -> int:
        return int(lib().orc_grid_search_order_len(self._h))

    @property
    def table_is_full(self) -> bool:
        return bool(lib().orc_grid_table_is_full(self._h))

    def bounds(self) -> np.ndarray:
        o = np.zeros(6)
        lib().orc_grid_bounds(self._h, o)
        return o

    def cell_info(self) -> np.ndarray:
        o = np.zeros(6)
        lib().orc_grid_cell_info(self._h, o)
        return o

    def container_box(self) -> np.ndarray:
        o = np.zeros(6)
        lib().orc_container_box(self._h, o)
        return o

    def cells(self) -> np.ndarray:
        o = np.zeros(self.n, np.uint64)
        lib().orc_grid_copy_cells(self._h, o)
        return o

    def sorted_indices(self) -> np.ndarray:
        o = np.zeros(self.n, np.uint64)
        lib().orc_grid_copy_sorted_indices(self._h, o)
        return o

    def delimiters(self) -> np.ndarray:
        o = np.zeros(self.num_delimiters, np.uint64)
        lib().orc_grid_copy_delimiters(self._h, o)
        return o

    def search_order(self):
        n = self.search_order_len
        d, ijk = np.zeros(n), np.zeros(3 * n, np.int32)
        lib().orc_grid_copy_search_order(self._h, d, ijk)
        return d, ijk.reshape(n, 3)

    def home_cell(self, x, y, z):
        o = np.zeros(3, np.uint64)
        lib().orc_grid_home_cell(self._h, x, y, z, o)
        return tuple(int(v) for v in o)

    # ---- queries (celery.rs:708-855)
    def check_cell_in_range(self, x, y, z, r, i, j, k) -> bool:
        return bool(lib().orc_check_cell_in_range(self._h, x, y, z, r, i, j, k))

    def _list(self, fn, *a):
        cap = max(self.n, self.cpd ** 3) + 8
        o = np.zeros(cap, np.uint64)
        n = int(fn(self._h, *a, o, cap))
        return [int(v) for v in o[:n]]

    def find_cells_in_radius(self, x, y, z, r):
        return self._list(lib().orc_find_cells_in_radius, x, y, z, r)

    def find_neighbors_in_cell_radius(self, x, y, z, r):
        return self._list(lib().orc_find_neighbors_in_cell_radius, x, y, z, r)

    def find_neighbors_in_real_radius(self, x, y, z, r):
        return self._list(lib().orc_find_neighbors_in_real_radius, x, y, z, r)

    def expanding_search(self, x, y, z) -> "ExpandingSearch":
        return ExpandingSearch(self, x, y, z)

    # ---- cells (interface.rs:186-344)
    def compute_cells(self, ids=None, mode=MODE_SECURITY, search_radius=float("nan"), target_group=-1, want_vertices=False, nthreads=0) -> CellResults:
        if ids is None:
            m, p = self.n, None
        else:
            ids = np.ascontiguousarray(ids, dtype=np.uint64)
            m, p = ids.size, ids.ctypes.data
        h = lib().orc_compute_cells(self._h, p, m, mode, search_radius, target_group, int(want_vertices), nthreads)
        if not h:
            raise RuntimeError("oracle: " + _err())
        return CellResults(h)

    def compute_cell_at_point(self, x, y, z, mode=MODE_SECURITY, search_radius=float("nan"), target_group=-1, want_vertices=False) -> CellResults:
        h = lib().orc_compute_cell_at_point(self._h, x, y, z, mode, search_radius, target_group, int(want_vertices))
        if not h:
            raise RuntimeError("oracle: " + _err())
        return CellResults(h)


class ExpandingSearch:
    def __init__(self, diagram: Diagram, x, y, z):
        self._d = diagram
        self._h = lib().orc_search_create(diagram._h, x, y, z)

    def __del__(self):
        try:
            lib().orc_search_destroy(self._h)
        except Exception:
            pass

    def home(self):
        o = np.zeros(3, np.uint64)
        lib().orc_search_home(self._h, o)
        return tuple(int(v) for v in o)

    def _list(self, fn, *a):
        cap = self._d.n + 8
        o = np.zeros(cap, np.uint64)
        n = int(fn(self._h, *a, o, cap))
        return [int(v) for v in o[:n]]

    def expand(self, max_radius, cells_to_add):
        return self._list(lib().orc_search_expand, max_radius, cells_to_add)

    def expand_all_no_radius(self):
        return self._list(lib().orc_search_expand_all_no_radius)

    def expand_all_in_radius(self, max_radius):
        return self._list(lib().orc_search_expand_all_in_radius, max_radius)


class Pool:
    KIND_VALUE, KIND_NEXT, KIND_END = 0, 1, 2

    def __init__(self):
        self._h = lib().orc_pool_create()

    def __del__(self):
        try:
            lib().orc_pool_destroy(self._h)
        except Exception:
            pass

    def add(self, v: int) -> int:
        return int(lib().orc_pool_add(self._h, v))

    def remove(self, i: int):
        lib().orc_pool_remove(self._h, i)

    def __len__(self):
        return int(lib().orc_pool_len(self._h))

    @property
    def first(self):
        f = int(lib().orc_pool_first(self._h))
        return None if f < 0 else f

    def chunk(self, i):
        p = C.c_int64(0)
        k = int(lib().orc_pool_chunk(self._h, i, C.byref(p)))
        return k, int(p.value)

    def has(self, i) -> bool:
        return bool(lib().orc_pool_has(self._h, i))

    def values(self):
        o = np.zeros(len(self) + 1, np.int64)
        n = int(lib().orc_pool_iterate(self._h, o, o.size))
        return [int(v) for v in o[:n]]


class Polyhedron:
    def __init__(self, x0, y0, z0, x1, y1, z1):
        self._h = lib().orc_poly_create(x0, y0, z0, x1, y1, z1)

    def __del__(self):
        try:
            lib().orc_poly_destroy(self._h)
        except Exception:
            pass

    def reset(self, *box):
        lib().orc_poly_reset(self._h, *box)

    def counts(self):
        o = np.zeros(5, np.uint64)
        lib().orc_poly_counts(self._h, o)
        return dict(edges=int(o[0]), vertices=int(o[1]), faces=int(o[2]), face_data=int(o[3]), root_edge=None if o[4] == 0 else int(o[4]) - 1)

    def live_counts(self):
        o = np.zeros(3, np.uint64)
        lib().orc_poly_live_counts(self._h, o)
        return dict(edges=int(o[0]), vertices=int(o[1]), faces=int(o[2]))

    def edge(self, e):
        o = np.zeros(4, np.uint64)
        if not lib().orc_poly_edge(self._h, e, o):
            return None
        f = [None if v == 0 else int(v) - 1 for v in o]
        return dict(flip=f[0], next=f[1], target=f[2], face=f[3])

    def vertex(self, v):
        o = np.zeros(3)
        return o if lib().orc_poly_vertex(self._h, v, o) else None

    def face(self, f):
        nb, se = C.c_int64(0), C.c_uint64(0)
        if not lib().orc_poly_face(self._h, f, C.byref(nb), C.byref(se)):
            return None
        return dict(neighbor=int(nb.value), starting_edge=int(se.value))

    def find_outgoing_edge(self, plane4):
        e = int(lib().orc_poly_find_outgoing_edge(self._h, np.ascontiguousarray(plane4, dtype=np.float64)))
        return None if e < 0 else e

    def cut_with_plane(self, point_index, plane4) -> bool:
        r = lib().orc_poly_cut_with_plane(self._h, point_index, np.ascontiguousarray(plane4, dtype=np.float64))
        if r < 0:
            raise RuntimeError("oracle: " + _err())
        return bool(r)

    def translate(self, s):
        lib().orc_poly_translate(self._h, np.ascontiguousarray(s, dtype=np.float64))

    def volume(self) -> float:
        return float(lib().orc_poly_volume(self._h))

    def weighted_normal(self, f):
        o = np.zeros(3)
        lib().orc_poly_weighted_normal(self._h, f, o)
        return o

    def face_vertices(self, f):
        o = np.zeros(3 * 256)
        n = int(lib().orc_poly_face_vertices(self._h, f, o, 256))
        return o[: 3 * n].reshape(n, 3)

    def check(self) -> int:
        return int(lib().orc_poly_check(self._h))


def plane_halfway(pt):
    o = np.zeros(4)
    lib().orc_plane_halfway_from_origin_to(np.ascontiguousarray(pt, dtype=np.float64), o)
    return o


def plane_from_normal_point(n, pt):
    o = np.zeros(4)
    lib().orc_plane_from_non_unit_normal_and_point(np.ascontiguousarray(n, dtype=np.float64), np.ascontiguousarray(pt, dtype=np.float64), o)
    return o
