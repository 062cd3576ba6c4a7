"""TEST INFRASTRUCTURE: ctypes binding of tests/emu/libemu_clip.so — the-tessellator_b200/csrc/clip.cu compiled
by g++ for a lane-by-lane CPU warp emulator (tests/emu/warp_emu.hpp).  It lets the CPU suite run the kernel
SOURCE (not a restatement) against the oracle, check that every warp collective is reached convergently, and
that results do not depend on the order in which lanes execute between collectives.  The product never loads it."""
import ctypes as C
import os
import subprocess

import numpy as np

import oracle_binding as ob

HERE = os.path.dirname(os.path.abspath(__file__))
EMU_DIR = os.path.join(HERE, "emu")
LIB_PATH = os.path.join(EMU_DIR, "libemu_clip.so")
_lib = None


class Args(C.Structure):
    _fields_ = [
        ("particles", C.c_void_p), ("n", C.c_uint32), ("delim", C.c_void_p), ("table_key", C.c_void_p), ("table_ijk", C.c_void_p),
        ("table_len", C.c_uint32), ("table_full", C.c_uint32), ("bounds", C.c_double * 6), ("cell_info", C.c_double * 6), ("cpd", C.c_uint32),
        ("box", C.c_double * 6), ("groups_sorted", C.c_void_p), ("work_slots", C.c_void_p), ("n_work", C.c_uint32), ("query_xyz", C.c_void_p),
        ("target_group", C.c_int64), ("search_radius", C.c_double), ("flags", C.c_uint32), ("large", C.c_int32), ("fstride", C.c_uint32),
        ("vol", C.c_void_p), ("nfaces", C.c_void_p), ("status", C.c_void_p), ("cell_id", C.c_void_p), ("st_nbr", C.c_void_p), ("st_area", C.c_void_p),
        ("counters", C.c_void_p), ("failed_slots", C.c_void_p), ("n_failed", C.c_void_p),
        ("gv_xyz", C.c_void_p), ("gl_idx", C.c_void_p), ("gv_cap", C.c_uint64), ("gl_cap", C.c_uint64), ("g_cursor", C.c_void_p),
        ("nverts", C.c_void_p), ("nloops", C.c_void_p), ("vbase", C.c_void_p), ("lbase", C.c_void_p), ("st_flen", C.c_void_p),
        ("os_threads", C.c_uint32), ("blocks", C.c_uint32), ("reverse", C.c_uint32), ("collectives", C.c_uint64),
        ("local_lo", C.c_uint32), ("local_hi", C.c_uint32),
    ]


def lib():
    global _lib
    if _lib is None:
        subprocess.check_call(["make", "-s", "-C", EMU_DIR, "libemu_clip.so"])
        _lib = C.CDLL(LIB_PATH)
        _lib.emu_clip_run.restype = C.c_int
        _lib.emu_clip_run.argtypes = [C.POINTER(Args)]
        _lib.emu_small_fmax.restype = C.c_uint32
        _lib.emu_large_fmax.restype = C.c_uint32
        _lib.emu_medium_fmax.restype = C.c_uint32
    return _lib


class EmuCells:
    """Rows of one emulated clip launch, shaped like the oracle's CellResults (faces in face-slot order)."""

    def __init__(self, vol, nfaces, status, cell_id, st_nbr, st_area, fstride, counters, n_failed, failed_slots, collectives):
        self.volumes, self.nfaces, self.status, self.cell_id = vol, nfaces, status, cell_id
        self.face_offsets = np.concatenate([[0], np.cumsum(nfaces.astype(np.int64))])
        keep = np.arange(fstride)[None, :] < np.minimum(nfaces, fstride)[:, None]
        self.neighbors = st_nbr.reshape(-1, fstride)[keep]
        self.areas = st_area.reshape(-1, fstride)[keep]
        self.counters = dict(zip(["visited", "tested", "vertex_classifications", "cuts", "new_vertices", "table_entries", "degenerate_skips", "faces"], (int(c) for c in counters)))
        self.n_failed = int(n_failed[0])
        self.failed_slots = failed_slots[: self.n_failed].copy()
        self.collectives = int(collectives)

    def cell_vertices(self, c):
        """Cell::compute_vertices of row c (cell-local coordinates, ascending vertex-slot order)."""
        g = self.geo
        return g["gv"][int(g["vb"][c]): int(g["vb"][c]) + int(g["nv"][c])]

    def face_loop(self, c, j):
        """VoronoiFace::compute_vertices of the j-th face (face-slot order) of row c."""
        g = self.geo
        lens = g["fl"].reshape(-1, self.fstride)[c, : int(self.nfaces[c])].astype(np.int64)
        off = int(g["lb"][c]) + int(lens[:j].sum())
        return self.cell_vertices(c)[g["gl"][off: off + int(lens[j])]]


class EmuGrid:
    """The grid arrays the clip kernel reads, taken from the oracle (tests elsewhere pin the CUDA binning
    pass to the same arrays bit for bit)."""

    def __init__(self, points, box=(0, 0, 0, 1, 1, 1), groups=None, table_radius=8):
        pts = np.ascontiguousarray(points, np.float64).reshape(-1, 3)
        self.oracle = ob.Diagram(pts, box=list(box), groups=groups, table_radius=table_radius)
        d = self.oracle
        self.n = pts.shape[0]
        self.sorted_indices = d.sorted_indices().astype(np.int64)
        rec = np.zeros((self.n, 4), np.float64)
        rec[:, :3] = pts[self.sorted_indices]
        rec[:, 3] = self.sorted_indices.view(np.float64)
        self.particles = rec
        self.delim = d.delimiters().astype(np.uint32)
        key, ijk = d.search_order()
        self.table_key, self.table_ijk = np.ascontiguousarray(key), np.ascontiguousarray(ijk, np.int32)
        self.table_full = int(d.table_is_full)
        self.bounds, self.cell_info, self.cpd = d.bounds(), d.cell_info(), d.cpd
        self.box = np.asarray(box, np.float64)
        self.groups_sorted = None if groups is None else np.ascontiguousarray(np.asarray(groups, np.uint64)[self.sorted_indices])
        assert self.delim.size == self.cpd ** 3 + 1
        self.local = (0, 0)

    def restrict_to_planes(self, points, local):
        """Turn this grid into the slab diagram of a rank that holds only the x-planes [local[0], local[1]): the
        particles of those planes are binned — by the emulated grid.cu, with their global ids — into a slab-local
        delimiter array, as tess_diagram_initialize_slab does.  Returns the global ids in slab slot order."""
        pts = np.ascontiguousarray(points, np.float64).reshape(-1, 3)
        cpd = self.cpd
        gx = (self.oracle.cells() // (cpd * cpd)).astype(np.int64)
        sel = np.nonzero((gx >= local[0]) & (gx < local[1]))[0]
        b = binning(pts[sel], self.oracle, ids=sel.astype(np.int64), local=local)
        assert b["oob"] == 0
        self.particles, self.delim, self.n, self.local = np.ascontiguousarray(b["sorted"]), b["delim"], len(sel), tuple(local)
        self.sorted_indices = b["sorted"][:, 3].view(np.int64).copy()
        return self.sorted_indices

    def clip(self, work_slots=None, large=False, flags=0, search_radius=float("nan"), target_group=-1, os_threads=8, reverse=False, fstride=None,
             query_xyz=None, want_vertices=False, count=True):
        """count=False runs the instantiation without work counters (the one the product times)."""
        L = lib()
        tier = {False: 0, True: 2, "medium": 1, "fast": 3, "thread": 4}[large]  # tess::CLIP_SMALL / CLIP_LARGE / CLIP_MEDIUM / CLIP_SMALL_FAST / CLIP_THREAD
        if fstride is None:
            fstride = (40, int(L.emu_medium_fmax()), int(L.emu_large_fmax()), 40, 40)[tier]
        ws = None if work_slots is None else np.ascontiguousarray(work_slots, np.uint32)
        m = self.n if ws is None else ws.size
        q = None
        if query_xyz is not None:  # get_cell_at_particle: one cell per query position, no self exclusion
            q = np.ascontiguousarray(query_xyz, np.float64).reshape(-1, 3)
            m = q.shape[0]
        vol, nfaces, status, cell_id = np.zeros(m), np.zeros(m, np.uint32), np.zeros(m, np.uint32), np.zeros(m, np.int64)
        st_nbr, st_area = np.zeros(m * fstride, np.int64), np.zeros(m * fstride)
        counters, failed, n_failed = np.zeros(8, np.uint64), np.zeros(max(m, 1), np.uint32), np.zeros(8, np.uint32)
        a = Args()
        a.particles, a.n, a.delim = self.particles.ctypes.data, self.n, self.delim.ctypes.data
        a.table_key, a.table_ijk, a.table_len, a.table_full = self.table_key.ctypes.data, self.table_ijk.ctypes.data, self.table_key.size, self.table_full
        a.bounds, a.cell_info, a.cpd = (C.c_double * 6)(*self.bounds), (C.c_double * 6)(*self.cell_info), self.cpd
        a.box = (C.c_double * 6)(*self.box)
        a.groups_sorted = None if self.groups_sorted is None else self.groups_sorted.ctypes.data
        a.work_slots, a.n_work, a.query_xyz = (None if ws is None else ws.ctypes.data), m, (None if q is None else q.ctypes.data)
        geo = None
        if want_vertices:
            vmax = (64, 256, 1024, 64, 64)[tier]
            geo = dict(gv=np.zeros((m * vmax, 3)), gl=np.zeros(m * 3 * vmax, np.uint32), cur=np.zeros(2, np.uint64), nv=np.zeros(m, np.uint32),
                       nl=np.zeros(m, np.uint32), vb=np.zeros(m, np.uint64), lb=np.zeros(m, np.uint64), fl=np.zeros(m * fstride, np.uint16))
            a.gv_xyz, a.gl_idx, a.gv_cap, a.gl_cap = geo["gv"].ctypes.data, geo["gl"].ctypes.data, m * vmax, m * 3 * vmax
            a.g_cursor, a.nverts, a.nloops = geo["cur"].ctypes.data, geo["nv"].ctypes.data, geo["nl"].ctypes.data
            a.vbase, a.lbase, a.st_flen = geo["vb"].ctypes.data, geo["lb"].ctypes.data, geo["fl"].ctypes.data
        a.target_group, a.search_radius, a.flags, a.large, a.fstride = target_group, search_radius, flags, tier, fstride
        a.vol, a.nfaces, a.status, a.cell_id = vol.ctypes.data, nfaces.ctypes.data, status.ctypes.data, cell_id.ctypes.data
        a.st_nbr, a.st_area, a.counters = st_nbr.ctypes.data, st_area.ctypes.data, (counters.ctypes.data if count else None)
        a.failed_slots, a.n_failed = failed.ctypes.data, n_failed.ctypes.data
        a.os_threads, a.blocks, a.reverse = os_threads, os_threads, int(reverse)
        a.local_lo, a.local_hi = self.local
        rc = L.emu_clip_run(C.byref(a))
        assert rc == 0
        e = EmuCells(vol, nfaces, status, cell_id, st_nbr, st_area, fstride, counters, n_failed, failed, a.collectives)
        if geo is not None:
            e.geo = geo
            e.fstride = fstride
        return e

    def oracle_cells(self, slots=None, **kw):
        """The oracle's cells in the kernel's row order (grid order, or the given sorted slots)."""
        ids = self.sorted_indices if slots is None else self.sorted_indices[np.asarray(slots, np.int64)]
        return self.oracle.compute_cells(ids=ids.astype(np.uint64), **kw)


# ------------------------------------------------------------------------------------------------
# grid.cu (binning pass) and query.cu (radius queries) on the emulator
# ------------------------------------------------------------------------------------------------
class GridArgs(C.Structure):
    _fields_ = [
        ("xyz", C.c_void_p), ("n", C.c_uint32), ("ids", C.c_void_p), ("groups", C.c_void_p), ("bounds", C.c_double * 6), ("cell_info", C.c_double * 6),
        ("cpd", C.c_uint32), ("local_lo", C.c_uint32), ("local_hi", C.c_uint32), ("bounds_out", C.c_void_p), ("cell_of", C.c_void_p), ("delim", C.c_void_p),
        ("sorted", C.c_void_p), ("sorted_idx", C.c_void_p), ("groups_sorted", C.c_void_p), ("plane_counts", C.c_void_p), ("oob", C.c_uint32),
        ("os_threads", C.c_uint32), ("reverse", C.c_uint32),
    ]


class QueryArgs(C.Structure):
    _fields_ = [
        ("particles", C.c_void_p), ("n", C.c_uint32), ("delim", C.c_void_p), ("groups_sorted", C.c_void_p), ("table_key", C.c_void_p), ("table_ijk", C.c_void_p),
        ("table_len", C.c_uint32), ("table_full", C.c_uint32), ("bounds", C.c_double * 6), ("cell_info", C.c_double * 6), ("cpd", C.c_uint32),
        ("xyz", C.c_void_p), ("n_query", C.c_uint32), ("radius", C.c_double), ("mode", C.c_int32), ("target_group", C.c_int64),
        ("offsets", C.c_void_p), ("indices", C.c_void_p), ("cap", C.c_uint64), ("flags", C.c_void_p), ("os_threads", C.c_uint32), ("reverse", C.c_uint32),
        ("cursor_in", C.c_void_p), ("cursor_out", C.c_void_p), ("cells_to_add", C.c_uint64),
    ]


_aux = {}


def _aux_lib(name):
    if name not in _aux:
        subprocess.check_call(["make", "-s", "-C", EMU_DIR, f"libemu_{name}.so"])
        _aux[name] = C.CDLL(os.path.join(EMU_DIR, f"libemu_{name}.so"))
    return _aux[name]


def binning(points, oracle: "ob.Diagram", ids=None, groups=None, local=None, os_threads=4, reverse=False):
    """K1-K4 of grid.cu on the emulator, with the grid parameters tess_diagram_initialize would derive
    (taken from the oracle).  Returns a dict of the arrays the pass produces."""
    L = _aux_lib("grid")
    pts = np.ascontiguousarray(points, np.float64).reshape(-1, 3)
    n, cpd = pts.shape[0], oracle.cpd
    lo, hi = (0, cpd) if local is None else local
    ncl = (hi - lo) * cpd * cpd
    out = dict(bounds=np.zeros(6), cell_of=np.zeros(n + 2, np.uint32), delim=np.zeros(ncl + 1, np.uint32), sorted=np.zeros((n, 4)),
               sorted_idx=np.zeros(n, np.uint32), groups_sorted=np.zeros(n, np.uint64), plane_counts=np.zeros(cpd, np.uint64))
    a = GridArgs()
    a.xyz, a.n = pts.ctypes.data, n
    idv = None if ids is None else np.ascontiguousarray(ids, np.int64)
    grv = None if groups is None else np.ascontiguousarray(groups, np.uint64)
    a.ids, a.groups = (None if idv is None else idv.ctypes.data), (None if grv is None else grv.ctypes.data)
    a.bounds, a.cell_info = (C.c_double * 6)(*oracle.bounds()), (C.c_double * 6)(*oracle.cell_info())
    a.cpd, a.local_lo, a.local_hi = cpd, lo, hi
    a.bounds_out, a.cell_of, a.delim, a.sorted = out["bounds"].ctypes.data, out["cell_of"].ctypes.data, out["delim"].ctypes.data, out["sorted"].ctypes.data
    a.sorted_idx, a.groups_sorted, a.plane_counts = out["sorted_idx"].ctypes.data, (out["groups_sorted"].ctypes.data if grv is not None else None), out["plane_counts"].ctypes.data
    a.os_threads, a.reverse = os_threads, int(reverse)
    assert L.emu_grid_run(C.byref(a)) == 0
    out["cell_of"] = out["cell_of"][:n]
    out["oob"] = int(a.oob)
    return out


def radius_query(grid: EmuGrid, xyz, radius, mode, target_group=-1, os_threads=4, reverse=False, cursors=None, cells_to_add=0):
    """query.cu on the emulator: per query the particle ids in the reference's order.  mode 0 cell radius, 1 real radius,
    2 expand_all_in_radius (search table), 3 find_cells_in_radius (grid cell ids), 4 ExpandingSearch::expand from `cursors`
    (returns the new cursors as a third value)."""
    L = _aux_lib("query")
    q = np.ascontiguousarray(xyz, np.float64).reshape(-1, 3)
    m = q.shape[0]
    cap = max(1, m * grid.n)
    offsets, indices, flags = np.zeros(m + 1, np.uint64), np.zeros(cap, np.int64), np.zeros(m, np.uint32)
    a = QueryArgs()
    a.particles, a.n, a.delim = grid.particles.ctypes.data, grid.n, grid.delim.ctypes.data
    a.groups_sorted = None if grid.groups_sorted is None else grid.groups_sorted.ctypes.data
    a.table_key, a.table_ijk, a.table_len, a.table_full = grid.table_key.ctypes.data, grid.table_ijk.ctypes.data, grid.table_key.size, grid.table_full
    a.bounds, a.cell_info, a.cpd = (C.c_double * 6)(*grid.bounds), (C.c_double * 6)(*grid.cell_info), grid.cpd
    a.xyz, a.n_query, a.radius, a.mode, a.target_group = q.ctypes.data, m, radius, mode, target_group
    a.offsets, a.indices, a.cap, a.flags = offsets.ctypes.data, indices.ctypes.data, cap, flags.ctypes.data
    a.os_threads, a.reverse = os_threads, int(reverse)
    cin = cout = None
    if mode == 4:
        cin, cout = np.ascontiguousarray(cursors, np.uint64), np.zeros(m, np.uint64)
        a.cursor_in, a.cursor_out, a.cells_to_add = cin.ctypes.data, cout.ctypes.data, min(int(cells_to_add), 2 ** 64 - 1)
    assert L.emu_query_run(C.byref(a)) == 0
    o = offsets.astype(np.int64)
    lists = [indices[o[i]:o[i + 1]].tolist() for i in range(m)]
    return (lists, flags, cout) if mode == 4 else (lists, flags)


# ------------------------------------------------------------------------------------------------
# outputs.cu on the emulator
# ------------------------------------------------------------------------------------------------
class PackArgs(C.Structure):
    _fields_ = [
        ("n_rows", C.c_uint64), ("status", C.c_void_p), ("nfaces", C.c_void_p), ("st_nbr", C.c_void_p), ("st_area", C.c_void_p), ("st_flen", C.c_void_p),
        ("fstride", C.c_uint32), ("redo_rows", C.c_void_p), ("n_redo", C.c_uint64), ("redo_nbr", C.c_void_p), ("redo_area", C.c_void_p), ("redo_flen", C.c_void_p),
        ("redo_stride", C.c_uint32), ("offsets", C.c_void_p), ("nbr", C.c_void_p), ("area", C.c_void_p), ("flen", C.c_void_p), ("face_cap", C.c_uint64),
        ("os_threads", C.c_uint32), ("reverse", C.c_uint32),
    ]


def pack_faces(status, nfaces, st_nbr, st_area, st_flen, fstride, redo_rows=None, redo_nbr=None, redo_area=None, redo_flen=None, redo_stride=0,
               os_threads=4, reverse=False):
    """Exclusive scan of the face counts + compact_faces (+ compact_redo): staged rows -> CSR, as tess_compute_all does."""
    L = _aux_lib("outputs")
    n = len(status)
    nf1 = np.concatenate([np.asarray(nfaces, np.uint32), [0]]).astype(np.uint32)
    total = int(nf1.sum())
    offsets, nbr, area, flen = np.zeros(n + 1, np.uint64), np.full(total, -99, np.int64), np.full(total, -1.0), np.zeros(total, np.uint32)
    a = PackArgs()
    a.n_rows, a.status, a.nfaces = n, status.ctypes.data, nf1.ctypes.data
    a.st_nbr, a.st_area, a.st_flen, a.fstride = st_nbr.ctypes.data, st_area.ctypes.data, (None if st_flen is None else st_flen.ctypes.data), fstride
    if redo_rows is not None and len(redo_rows):
        a.redo_rows, a.n_redo, a.redo_nbr, a.redo_area = redo_rows.ctypes.data, len(redo_rows), redo_nbr.ctypes.data, redo_area.ctypes.data
        a.redo_flen, a.redo_stride = (None if redo_flen is None else redo_flen.ctypes.data), redo_stride
    a.offsets, a.nbr, a.area, a.flen, a.face_cap = offsets.ctypes.data, nbr.ctypes.data, area.ctypes.data, (None if st_flen is None else flen.ctypes.data), total
    a.os_threads, a.reverse = os_threads, int(reverse)
    assert L.emu_pack_run(C.byref(a)) == 0
    return offsets.astype(np.int64), nbr, area, flen


def outputs_lib():
    L = _aux_lib("outputs")
    L.emu_chunk_list.restype = C.c_uint64
    L.emu_chunk_list.argtypes = [C.c_void_p, C.c_uint64, C.c_uint32, C.c_uint32, C.c_void_p, C.c_uint32]
    L.emu_volume_sum.restype = C.c_double
    L.emu_volume_sum.argtypes = [C.c_void_p, C.c_uint64, C.c_uint32, C.c_uint32]
    L.emu_gather.argtypes = [C.c_void_p] * 9 + [C.c_uint64, C.c_void_p, C.c_void_p, C.c_uint32]
    L.emu_clear_status_bits.argtypes = [C.c_void_p, C.c_uint64, C.c_uint32]
    return L
