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

    def clip(self, work_slots=None, large=False, flags=0, search_radius=float("nan"), target_group=-1, os_threads=8, reverse=False, fstride=None,
             query_xyz=None, want_vertices=False, count=True):
        """count=False runs the instantiation without work counters (the one the product times)."""
        L = lib()
        tier = {False: 0, True: 2, "medium": 1}[large]  # tess::CLIP_SMALL / CLIP_LARGE / CLIP_MEDIUM
        if fstride is None:
            fstride = (40, int(L.emu_medium_fmax()), int(L.emu_large_fmax()))[tier]
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
            vmax = (64, 256, 1024)[tier]
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
