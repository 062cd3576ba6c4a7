"""Host-side mirror of the reference's public API (src/interface.rs) over the C ABI.

Names, argument meaning and results follow interface.rs: `Diagram` (:25), `Cell` (:237),
`VoronoiFace` (:393).  The reference computes one cell per call; here the first
`Cell.compute_voronoi_cell()` for a given (search_radius, target_group) runs the whole diagram on
the GPU once (tess_compute_all) and every cell reads its row of that batch.  `compute_all_cells`
is the explicit batch entry point.

Deviations (all forced by defects of the reference):
  * add_particle_with_group / initialize are public (private in the reference, D3);
  * the start polyhedron of a cell must be the diagram's container box (`Polyhedron(x_min, ...)`,
    polyhedron.rs:226): the GPU path clips the axis-aligned container only;
  * faces of the container that survive are reported with neighbour ids -1..-6 (D10).
"""
from __future__ import annotations

import ctypes as C
import math
from dataclasses import dataclass
from typing import Optional, Sequence

import numpy as np

from . import _lib
from ._lib import check


@dataclass(frozen=True)
class Polyhedron:
    """Start shape of a cell: the axis-aligned box of Polyhedron::new (polyhedron.rs:226-233)."""

    x_min: float
    y_min: float
    z_min: float
    x_max: float
    y_max: float
    z_max: float

    def as_box(self) -> np.ndarray:
        return np.array([self.x_min, self.y_min, self.z_min, self.x_max, self.y_max, self.z_max], dtype=np.float64)


class CellBatch:
    """All cells of one tess_compute_all call (CSR over cells)."""

    def __init__(self, handle: int, device: int):
        self._h = C.c_void_p(handle)
        self.device = device
        nc, nf = C.c_uint64(0), C.c_uint64(0)
        check(_lib.lib().tess_result_n_cells(self._h, C.byref(nc), C.byref(nf)))
        self.n_cells, self.n_faces = int(nc.value), int(nf.value)
        self._cache = {}
        self._has_vertices = False

    def close(self):
        if self._h:
            _lib.lib().tess_result_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _view(self, fn: str, n: int, dtype) -> np.ndarray:
        if fn not in self._cache:
            p = C.c_void_p(0)
            check(getattr(_lib.lib(), fn)(self._h, C.byref(p)))
            if n == 0:
                self._cache[fn] = np.zeros(0, dtype)
            else:
                buf = (C.c_char * (n * np.dtype(dtype).itemsize)).from_address(p.value)
                self._cache[fn] = np.frombuffer(buf, dtype=dtype, count=n)  # view into the result's host buffer
        return self._cache[fn]

    @property
    def volumes(self) -> np.ndarray:
        return self._view("tess_result_volumes", self.n_cells, np.float64)

    @property
    def face_offsets(self) -> np.ndarray:
        return self._view("tess_result_face_offsets", self.n_cells + 1, np.uint64).astype(np.int64)

    @property
    def neighbors(self) -> np.ndarray:
        return self._view("tess_result_neighbors", self.n_faces, np.int64)

    @property
    def areas(self) -> np.ndarray:
        return self._view("tess_result_areas", self.n_faces, np.float64)

    @property
    def status(self) -> np.ndarray:
        return self._view("tess_result_status", self.n_cells, np.uint32)

    @property
    def cell_ids(self) -> np.ndarray:
        return self._view("tess_result_cell_ids", self.n_cells, np.int64)

    @property
    def vertex_offsets(self) -> np.ndarray:
        return self._view("tess_result_vertex_offsets", self.n_cells + 1, np.uint64).astype(np.int64)

    @property
    def vertices(self) -> np.ndarray:
        nv = int(self.vertex_offsets[-1])
        return self._view("tess_result_vertices", 3 * nv, np.float64).reshape(nv, 3)

    @property
    def face_vertex_offsets(self) -> np.ndarray:
        return self._view("tess_result_face_vertex_offsets", self.n_faces + 1, np.uint64).astype(np.int64)

    @property
    def face_vertex_indices(self) -> np.ndarray:
        return self._view("tess_result_face_vertex_indices", int(self.face_vertex_offsets[-1]), np.uint32)

    def face_vertices(self, row: int, k: int) -> np.ndarray:
        """Ordered vertex loop of global face k of cell `row`, cell-local coordinates (polyhedron.rs:897-919)."""
        fo = self.face_vertex_offsets
        idx = self.face_vertex_indices[fo[k]:fo[k + 1]].astype(np.int64)
        return self.vertices[int(self.vertex_offsets[row]) + idx]

    def counters(self) -> dict:
        arr = (C.c_uint64 * 8)()
        check(_lib.lib().tess_result_counters(self._h, C.byref(arr)))
        return {k: int(v) for k, v in zip(_lib.COUNTER_NAMES, arr)}

    def volume_sum(self) -> float:
        out = C.c_double(0)
        check(_lib.lib().tess_result_volume_sum(self._h, C.byref(out)))
        return float(out.value)

    def timings(self) -> dict:
        """CUDA-event durations in ms: clip kernel, redo passes (medium / large configurations), scans + compaction, whole call."""
        arr = (C.c_double * 4)()
        check(_lib.lib().tess_result_timings(self._h, C.byref(arr)))
        return dict(clip_ms=arr[0], redo_ms=arr[1], outputs_ms=arr[2], total_ms=arr[3])

    def tier_stats(self) -> dict:
        """Which kernel ran the main clip pass and how many cells the redo passes took over."""
        arr = (C.c_uint64 * 4)()
        check(_lib.lib().tess_result_tier_stats(self._h, C.byref(arr)))
        return dict(main_tier=MAIN_TIERS.get(int(arr[0]), str(int(arr[0]))), redo_a=int(arr[1]), redo_b=int(arr[2]), redo_c=int(arr[3]))

    def download(self, volumes=None, face_offsets=None, neighbors=None, areas=None, status=None, stream: int = 0) -> None:
        """Asynchronous copies into caller-owned (ideally pinned) host arrays; the caller synchronises."""
        ptr = lambda a: None if a is None else a.ctypes.data if hasattr(a, "ctypes") else a.data_ptr()  # noqa: E731
        check(_lib.lib().tess_result_download(self._h, ptr(volumes), ptr(face_offsets), ptr(neighbors), ptr(areas), ptr(status), stream or None))

    def device_views(self) -> dict:
        ptrs = [C.c_void_p(0) for _ in range(6)]
        check(_lib.lib().tess_result_device_views(self._h, *[C.byref(p) for p in ptrs]))
        return dict(zip(("volumes", "face_offsets", "neighbors", "areas", "status", "cell_ids"), [p.value for p in ptrs]))

    # per-cell slices
    def cell_neighbors(self, row: int) -> np.ndarray:
        o = self.face_offsets
        return self.neighbors[o[row]:o[row + 1]]

    def cell_areas(self, row: int) -> np.ndarray:
        o = self.face_offsets
        return self.areas[o[row]:o[row + 1]]

    def cell_vertices(self, row: int) -> np.ndarray:
        o = self.vertex_offsets
        return self.vertices[o[row]:o[row + 1]]


class Diagram:
    """interface.rs:25 `Diagram`: particles + groups + spatial grid + container."""

    def __init__(self, device: int = 0):
        self._h = C.c_void_p(0)
        check(_lib.lib().tess_diagram_create(C.byref(self._h), _lib.TESS_F64, device))
        self.device = device
        self.initialized = False
        self._pending: list = []
        self._pending_groups: list = []
        self._n = 0
        self._box: Optional[np.ndarray] = None
        self._batches: dict = {}
        self._host_pts: list = []  # (first index, array) of host-added particles, for Cell.compute_neighbor_cloud

    def close(self):
        if getattr(self, "_h", None):
            for b in self._batches.values():
                b.close()
            self._batches = {}
            _lib.lib().tess_diagram_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __len__(self):
        return self._n + len(self._pending)

    # ---- building (interface.rs:52-84) -------------------------------------------------------
    def add_particle_with_group(self, particle: Sequence[float], group: int = 0) -> None:
        """interface.rs:52-57.  `particle` exposes x, y, z (ToCeleryPoint, celery.rs:56-60)."""
        if self.initialized:
            raise _lib.TessError(-2, "particles must be added before initialize")
        self._pending.append((float(particle[0]), float(particle[1]), float(particle[2])))
        self._pending_groups.append(int(group))

    def add_particles(self, xyz: np.ndarray, groups: Optional[np.ndarray] = None, stream: int = 0) -> None:
        """Batch form of add_particle_with_group: xyz is (n, >=3) float64; row stride may exceed 24 B."""
        self._flush()
        a = np.asarray(xyz, dtype=np.float64)
        if a.ndim != 2 or a.shape[1] < 3:
            raise ValueError("xyz must be (n, >=3)")
        if a.strides[1] != 8:
            a = np.ascontiguousarray(a)
        g = None if groups is None else np.ascontiguousarray(groups, dtype=np.uint64)
        check(_lib.lib().tess_diagram_add_particles(self._h, a.ctypes.data, a.shape[0], a.strides[0], None if g is None else g.ctypes.data, stream))
        self._host_pts.append((self._n, a))
        self._n += a.shape[0]

    def add_particles_device(self, xyz_ptr: int, n: int, groups_ptr: int = 0, ids_ptr: int = 0, stream: int = 0) -> None:
        """Particles already resident in device memory (packed f64 triples)."""
        self._flush()
        check(_lib.lib().tess_diagram_add_particles_device(self._h, xyz_ptr, n, groups_ptr or None, ids_ptr or None, stream))
        self._n += n

    def add_records_device(self, rec_ptr: int, n: int, stream: int = 0) -> None:
        """Particles resident in device memory as 32-byte records {x, y, z, id} (tess_pack_records after the exchange)."""
        self._flush()
        check(_lib.lib().tess_diagram_add_records_device(self._h, rec_ptr, n, stream))
        self._n += n

    def _position_of(self, index: int):
        for lo, a in self._host_pts:
            if lo <= index < lo + a.shape[0]:
                return tuple(float(v) for v in a[index - lo, :3])
        raise _lib.TessError(-5, "position of a particle added from device memory is not kept on the host")

    def _flush(self):
        if self._pending:
            pts = np.array(self._pending, dtype=np.float64)
            self._host_pts.append((self._n, pts))
            grp = np.array(self._pending_groups, dtype=np.uint64)
            self._pending, self._pending_groups = [], []
            check(_lib.lib().tess_diagram_add_particles(self._h, pts.ctypes.data, pts.shape[0], 24, grp.ctypes.data, 0))
            self._n += pts.shape[0]

    def clear(self) -> None:
        for b in self._batches.values():
            b.close()
        self._batches = {}
        check(_lib.lib().tess_diagram_clear(self._h))
        self._pending, self._pending_groups, self._n = [], [], 0
        self._host_pts = []
        self.initialized = False

    def initialize(self, container: Optional[Polyhedron] = None, stream: int = 0) -> None:
        """interface.rs:60-84.  container None -> bounding box of the particles (:69-79)."""
        self._flush()
        box = None if container is None else container.as_box()
        check(_lib.lib().tess_diagram_initialize(self._h, None if box is None else box.ctypes.data, stream))
        self.initialized = True
        if box is None:
            b = self.grid_info()["bounds"]
            box = np.array([b[0], b[2], b[4], b[1], b[3], b[5]])
        self._box = box

    def initialize_slab(self, container: Polyhedron, bounds, n_global: int, own, local, stream: int = 0) -> None:
        self._flush()
        box = container.as_box()
        s = _lib.Slab()
        for i in range(6):
            s.bounds[i] = float(bounds[i])
        s.n_global = int(n_global)
        # own = (lo_plane, hi_plane), or (lo_plane, lo_row, hi_plane, hi_row) for slab cuts finer than a plane
        if len(own) == 4:
            s.own_lo, s.own_lo_row, s.own_hi, s.own_hi_row = (int(v) for v in own)
        else:
            s.own_lo, s.own_hi = int(own[0]), int(own[1])
            s.own_lo_row = s.own_hi_row = 0
        s.local_lo, s.local_hi = int(local[0]), int(local[1])
        check(_lib.lib().tess_diagram_initialize_slab(self._h, box.ctypes.data, C.byref(s), stream))
        self.initialized = True
        self._box = box

    # ---- grid facts (celery.rs) ---------------------------------------------------------------
    def grid_info(self) -> dict:
        n, cpd = C.c_uint64(0), C.c_uint64(0)
        b, s, i = np.zeros(6), np.zeros(3), np.zeros(3)
        check(_lib.lib().tess_diagram_grid_info(self._h, C.byref(n), C.byref(cpd), b.ctypes.data, s.ctypes.data, i.ctypes.data))
        return dict(n_points=int(n.value), cells_per_dimension=int(cpd.value), bounds=b, cell_sizes=s, inverse_cell_sizes=i)

    def binning_ms(self) -> float:
        arr = (C.c_double * 1)()
        check(_lib.lib().tess_diagram_timings(self._h, C.byref(arr)))
        return float(arr[0])

    def copy_grid(self, n_local_cells: Optional[int] = None):
        info = self.grid_info()
        n = info["n_points"]
        ncl = info["cells_per_dimension"] ** 3 if n_local_cells is None else n_local_cells
        cells, sidx, delim = np.zeros(n, np.uint64), np.zeros(n, np.uint64), np.zeros(ncl + 1, np.uint64)
        check(_lib.lib().tess_diagram_copy_grid(self._h, cells.ctypes.data, sidx.ctypes.data, delim.ctypes.data))
        return cells, sidx, delim

    def search_order(self, table_radius: int = 0):
        ln, full = C.c_uint64(0), C.c_int(0)
        check(_lib.lib().tess_diagram_copy_search_order(self._h, table_radius, C.byref(ln), None, None, C.byref(full)))
        keys, ijk = np.zeros(ln.value), np.zeros(3 * ln.value, np.int32)
        check(_lib.lib().tess_diagram_copy_search_order(self._h, table_radius, C.byref(ln), keys.ctypes.data, ijk.ctypes.data, C.byref(full)))
        return keys, ijk.reshape(-1, 3), bool(full.value)

    # ---- cells --------------------------------------------------------------------------------
    def _opts(self, search_radius, target_group, outputs, table_radius, stream) -> _lib.Opts:
        o = _lib.Opts()
        _lib.lib().tess_opts_default(C.byref(o))
        o.search_radius = float("nan") if search_radius is None else float(search_radius)
        o.target_group = -1 if target_group is None else int(target_group)
        if outputs is not None:
            o.outputs = outputs
        o.table_radius = table_radius
        o.stream = stream or None
        return o

    def compute_all_cells(self, search_radius: Optional[float] = None, target_group: Optional[int] = None, outputs: Optional[int] = None,
                          table_radius: int = 0, stream: int = 0) -> CellBatch:
        """Batch fast path: every cell of the diagram in one call (tess_compute_all)."""
        o = self._opts(search_radius, target_group, outputs, table_radius, stream)
        h = C.c_void_p(0)
        check(_lib.lib().tess_compute_all(self._h, C.byref(o), C.byref(h)))
        return CellBatch(h.value, self.device)

    def compute_all_cells_to_host(self, volumes, face_offsets, neighbors, areas, status, n_chunks: int = 0, search_radius: Optional[float] = None,
                                  target_group: Optional[int] = None, outputs: Optional[int] = None, table_radius: int = 0, stream: int = 0) -> CellBatch:
        """compute_all_cells with the results streamed into caller-owned host arrays while later cells are
        still being computed (tess_compute_all_to_host).  Each argument is a C-contiguous host array of the
        right dtype — numpy, or anything with `data_ptr()` / `numel()` such as a pinned torch tensor — or None:
        volumes f64[n], face_offsets u64/i64[n+1], neighbors i64[cap], areas f64[cap], status u32/i32[n]; the
        capacities in cells and faces are taken from the array lengths (too short -> TessError, nothing overrun).  Page-locked arrays overlap the copy
        with the compute.  Returns the device-side batch; the host arrays are complete on return."""
        def ptr_len(a, itemsize):
            if a is None:
                return None, None
            if hasattr(a, "data_ptr"):
                assert a.is_contiguous() and a.element_size() == itemsize
                return a.data_ptr(), a.numel()
            assert a.flags["C_CONTIGUOUS"] and a.itemsize == itemsize
            return a.ctypes.data, a.size
        pv, nv = ptr_len(volumes, 8)
        po, no = ptr_len(face_offsets, 8)
        pn, nn = ptr_len(neighbors, 8)
        pa, na = ptr_len(areas, 8)
        ps, ns = ptr_len(status, 4)
        cell_caps = [c for c in (nv, None if no is None else no - 1, ns) if c is not None]
        caps = [c for c in (nn, na) if c is not None]
        o = self._opts(search_radius, target_group, outputs, table_radius, stream)
        if pa is not None and not (o.outputs & _lib.OUT_AREAS):
            pa = None
        h = C.c_void_p(0)
        check(_lib.lib().tess_compute_all_to_host(self._h, C.byref(o), int(n_chunks), pv, po, pn, pa, ps, min(cell_caps) if cell_caps else 2**62, min(caps) if caps else 0, C.byref(h)))
        return CellBatch(h.value, self.device)

    def compute_cells_at(self, points: np.ndarray, search_radius: Optional[float] = None, target_group: Optional[int] = None,
                         outputs: Optional[int] = None, table_radius: int = 0, stream: int = 0) -> CellBatch:
        """Batch form of get_cell_at_particle (interface.rs:211-232)."""
        p = np.ascontiguousarray(points, dtype=np.float64).reshape(-1, 3)
        o = self._opts(search_radius, target_group, outputs, table_radius, stream)
        h = C.c_void_p(0)
        check(_lib.lib().tess_compute_at_points(self._h, p.ctypes.data, p.shape[0], C.byref(o), C.byref(h)))
        return CellBatch(h.value, self.device)

    # ---- radius queries (celery.rs:753-855) ---------------------------------------------------
    def find_neighbors(self, points: np.ndarray, radius: float, mode: int, target_group: Optional[int] = None, stream: int = 0) -> list:
        """Batch radius query: one list of particle ids per query point, in the reference's order."""
        p = np.ascontiguousarray(points, dtype=np.float64).reshape(-1, 3)
        h = C.c_void_p(0)
        check(_lib.lib().tess_find_neighbors(self._h, p.ctypes.data, p.shape[0], float(radius), mode, -1 if target_group is None else int(target_group), stream or None, C.byref(h)))
        return _query_lists(h, p.shape[0])

    def find_cells_in_radius(self, x: float, y: float, z: float, radius: float, stream: int = 0) -> list:
        """Celery::find_cells_in_radius (celery.rs:753-797): ids of the grid cells within `radius` of the position."""
        p = np.array([[x, y, z]], dtype=np.float64)
        h = C.c_void_p(0)
        check(_lib.lib().tess_find_cells_in_radius(self._h, p.ctypes.data, 1, float(radius), stream or None, C.byref(h)))
        return _query_lists(h, 1)[0].tolist()

    def expanding_search(self, points) -> "ExpandingSearch":
        """ExpandingSearch::new (celery.rs:882-902) around one position (x, y, z) or around m positions at once."""
        return ExpandingSearch(self, points)

    def find_neighbors_in_cell_radius(self, x: float, y: float, z: float, radius: float) -> list:
        """Celery::find_neighbors_in_cell_radius (celery.rs:802-819)."""
        return self.find_neighbors(np.array([[x, y, z]]), radius, _lib.QUERY_CELL_RADIUS)[0].tolist()

    def find_neighbors_in_real_radius(self, x: float, y: float, z: float, radius: float) -> list:
        """Celery::find_neighbors_in_real_radius (celery.rs:825-855)."""
        return self.find_neighbors(np.array([[x, y, z]]), radius, _lib.QUERY_REAL_RADIUS)[0].tolist()

    def _check_polyhedron(self, polyhedron: Polyhedron):
        if not np.array_equal(polyhedron.as_box(), self._box):
            raise _lib.TessError(-5, "the start polyhedron of a cell must be the diagram's container box")

    def _batch(self, search_radius, target_group, vertices: bool = False) -> CellBatch:
        """All cells for one (search_radius, target_group), computed once; the geometry outputs (vertex lists, face loops) only
        when a caller asks for vertices (Cell.compute_vertices / VoronoiFace.compute_vertices)."""
        key = (None if search_radius is None else float(search_radius), target_group)
        have = self._batches.get(key)
        if have is None or (vertices and not have._has_vertices):
            out = _lib.OUT_VOLUME | _lib.OUT_NEIGHBORS | _lib.OUT_AREAS | (_lib.OUT_VERTICES if vertices else 0)
            new = self.compute_all_cells(search_radius, target_group, outputs=out)
            new._has_vertices = vertices
            self._batches[key] = new  # (cells that hold the earlier batch keep it alive; it is freed with them)
        return self._batches[key]

    def get_cell_at_index(self, index: int, polyhedron: Polyhedron, search_radius: Optional[float] = None, target_group: Optional[int] = None) -> "Cell":
        """interface.rs:186-208."""
        self._check_polyhedron(polyhedron)
        if not (0 <= index < self._n):
            raise IndexError(index)
        return Cell(self, index, None, search_radius, target_group)

    def get_cell_at_particle(self, point: Sequence[float], polyhedron: Polyhedron, search_radius: Optional[float] = None, target_group: Optional[int] = None) -> "Cell":
        """interface.rs:211-232."""
        self._check_polyhedron(polyhedron)
        return Cell(self, None, (float(point[0]), float(point[1]), float(point[2])), search_radius, target_group)


class Cell:
    """interface.rs:237 `Cell`."""

    def __init__(self, diagram: Diagram, index: Optional[int], position, search_radius, target_group):
        self.diagram, self.index, self.position = diagram, index, position
        self.search_radius, self.target_group = search_radius, target_group
        self._batch: Optional[CellBatch] = None
        self._row = 0

    def compute_voronoi_cell(self, vertices: bool = False) -> None:
        """interface.rs:257-313."""
        if self.index is not None:
            self._batch, self._row = self.diagram._batch(self.search_radius, self.target_group, vertices), self.index
        else:
            self._batch = self.diagram.compute_cells_at(
                np.array([self.position]), self.search_radius, self.target_group,
                outputs=_lib.OUT_VOLUME | _lib.OUT_NEIGHBORS | _lib.OUT_AREAS | (_lib.OUT_VERTICES if vertices else 0))
            self._batch._has_vertices = vertices
            self._row = 0

    def _need(self, vertices: bool = False) -> CellBatch:
        if self._batch is None or (vertices and not getattr(self._batch, "_has_vertices", False)):
            self.compute_voronoi_cell(vertices)
            # a cell no tier could finish (more than 1024 vertices / 512 faces, an inconsistent mesh, a search table the
            # redo passes could not widen enough) has no geometry to hand out: say so instead of returning volume 0
            st = int(self._batch.status[self._row])
            bad = st & (_lib.STATUS_TABLE_EXHAUSTED | _lib.STATUS_CAPACITY_OVERFLOW | _lib.STATUS_INCONSISTENT)
            if bad:
                raise _lib.TessError(-6, f"cell could not be completed (status 0x{st:x}: " + ", ".join(
                    n for b, n in ((_lib.STATUS_TABLE_EXHAUSTED, "search table exhausted"), (_lib.STATUS_CAPACITY_OVERFLOW, "mesh tables overflowed"),
                                   (_lib.STATUS_INCONSISTENT, "inconsistent mesh")) if st & b) + ")")
        return self._batch

    def compute_volume(self) -> float:
        """interface.rs:337-339."""
        return float(self._need().volumes[self._row])

    def compute_neighbors(self) -> list:
        """interface.rs:342-344 (face order; container walls are -1..-6)."""
        return [int(v) for v in self._need().cell_neighbors(self._row)]

    def compute_vertices(self) -> np.ndarray:
        """interface.rs:368-370: vertices in cell-local coordinates (relative to the particle)."""
        return self._need(vertices=True).cell_vertices(self._row).copy()

    def compute_faces(self) -> list:
        """interface.rs:373-384."""
        b = self._need()
        lo, hi = int(b.face_offsets[self._row]), int(b.face_offsets[self._row + 1])
        return [VoronoiFace(self, k) for k in range(lo, hi)]

    def compute_neighbor_cloud(self, radius: float, target_group: Optional[int] = None) -> list:
        """interface.rs:348-365: expand_all_in_radius(radius) around the cell's particle, filtered by group."""
        pos = self.position if self.position is not None else self.diagram._position_of(self.index)
        return self.diagram.find_neighbors(np.array([pos]), radius, _lib.QUERY_NEIGHBOR_CLOUD, target_group)[0].tolist()

    def original_index(self) -> Optional[int]:
        """interface.rs:387-389."""
        return self.index

    def status(self) -> int:
        return int(self._need().status[self._row])


class VoronoiFace:
    """interface.rs:393 `VoronoiFace`."""

    def __init__(self, cell: Cell, k: int):
        self.cell, self._k = cell, k

    def compute_area(self) -> float:
        """interface.rs:408-410."""
        return float(self.cell._batch.areas[self._k])

    def compute_neighbor(self) -> int:
        """interface.rs:413-416."""
        return int(self.cell._batch.neighbors[self._k])

    def compute_vertices(self) -> np.ndarray:
        """interface.rs:403-405: the face's vertices in loop order, cell-local coordinates."""
        return self.cell._need(vertices=True).face_vertices(self.cell._row, self._k).copy()


def _query_lists(h, m: int) -> list:
    """CSR result of a query call -> one index array per position; frees the handle."""
    try:
        po, pi = C.c_void_p(0), C.c_void_p(0)
        check(_lib.lib().tess_query_offsets(h, C.byref(po)))
        check(_lib.lib().tess_query_indices(h, C.byref(pi)))
        off = np.ctypeslib.as_array(C.cast(po, C.POINTER(C.c_uint64)), shape=(m + 1,)).astype(np.int64) if m else np.zeros(1, np.int64)
        tot = int(off[-1])
        idx = np.ctypeslib.as_array(C.cast(pi, C.POINTER(C.c_int64)), shape=(tot,)).copy() if tot else np.zeros(0, np.int64)
        return [idx[off[i]:off[i + 1]] for i in range(m)]
    finally:
        _lib.lib().tess_query_free(h)


class ExpandingSearch:
    """celery.rs:865 `ExpandingSearch`: an outward walk over the grid cells around a position, resumable.  Built for one
    position it returns plain lists like the reference; built for m positions every call returns one list per position."""

    def __init__(self, diagram: Diagram, points):
        p = np.ascontiguousarray(points, dtype=np.float64)
        self._single = p.ndim == 1
        self._p = p.reshape(-1, 3)
        self._d = diagram
        self._h = C.c_void_p(0)
        check(_lib.lib().tess_search_create(diagram._h, self._p.ctypes.data, self._p.shape[0], C.byref(self._h)))

    def close(self):
        if self._h:
            _lib.lib().tess_search_free(self._h)
            self._h = C.c_void_p(0)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def expand(self, max_radius: float, cells_to_add: int, stream: int = 0):
        """celery.rs:907-963."""
        h = C.c_void_p(0)
        check(_lib.lib().tess_search_expand(self._h, float(max_radius), min(int(cells_to_add), 2 ** 64 - 1), stream or None, C.byref(h)))
        out = _query_lists(h, self._p.shape[0])
        return out[0].tolist() if self._single else out

    def expand_all_in_radius(self, max_radius: float):
        """celery.rs:1023-1075 (on a search that has not moved yet)."""
        return self.expand(max_radius, 2 ** 64 - 1)

    def expand_all_no_radius(self):
        """celery.rs:971-1018."""
        return self.expand(float("inf"), 2 ** 64 - 1)

    @property
    def current_search_index(self):
        pc = C.c_void_p(0)
        check(_lib.lib().tess_search_cursor(self._h, C.byref(pc)))
        cur = np.ctypeslib.as_array(C.cast(pc, C.POINTER(C.c_uint64)), shape=(self._p.shape[0],)).copy()
        return int(cur[0]) if self._single else cur


MAIN_TIERS = {0: "small", 3: "fast", 4: "thread"}


def set_main_tier(tier) -> None:
    """Kernel of the main clip pass: None / "default", "thread" (one thread per cell), "fast" (one warp per cell, no serial
    walk) or "small" (one warp per cell with the reference-shaped serial walk).  Results are bit-identical for every choice."""
    code = {None: -1, "default": -1, "small": 0, "fast": 3, "thread": 4}[tier]
    check(_lib.lib().tess_set_main_tier(code))


def device_count() -> int:
    return int(_lib.lib().tess_device_count())


def isnan(x) -> bool:
    return isinstance(x, float) and math.isnan(x)
