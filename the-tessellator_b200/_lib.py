"""ctypes binding of libtess_b200.so (include/tess.h).

The library is built in-tree by `build()` (nvcc, sm_100a) and must exist: there is no CPU
fallback and no alternative backend.  Importing this module never needs a GPU; calling a compute
entry point without one fails with TessError.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("TESS_LIB_PATH") or os.path.join(_HERE, "libtess_b200.so")  # override: A/B builds only
CSRC = os.path.join(_HERE, "csrc")
HEADER = os.path.join(os.path.dirname(_HERE), "include", "tess.h")

TESS_F64, TESS_F32 = 0, 1
QUERY_CELL_RADIUS, QUERY_REAL_RADIUS, QUERY_NEIGHBOR_CLOUD = 0, 1, 2
OUT_VOLUME, OUT_NEIGHBORS, OUT_AREAS, OUT_VERTICES, OUT_COUNTERS = 1, 2, 4, 8, 16
STATUS_DEGENERATE_SKIP, STATUS_TABLE_EXHAUSTED, STATUS_CAPACITY_OVERFLOW, STATUS_HALO_INSUFFICIENT, STATUS_INCONSISTENT = 1, 2, 4, 8, 16
COUNTER_NAMES = ("visited", "tested", "vertex_classifications", "cuts", "new_vertices", "table_entries", "degenerate_skips", "faces")

# every symbol include/tess.h declares (tests check the .so exports exactly these)
SYMBOLS = (
    "tess_last_error", "tess_version", "tess_device_count", "tess_opts_default",
    "tess_diagram_create", "tess_diagram_destroy", "tess_diagram_add_particles", "tess_diagram_add_particles_device",
    "tess_diagram_clear", "tess_diagram_initialize", "tess_diagram_initialize_slab", "tess_diagram_grid_info",
    "tess_diagram_copy_grid", "tess_diagram_copy_search_order", "tess_compute_all", "tess_compute_all_to_host", "tess_compute_at_points",
    "tess_result_free", "tess_result_n_cells", "tess_result_volumes", "tess_result_face_offsets", "tess_result_neighbors",
    "tess_result_areas", "tess_result_status", "tess_result_cell_ids", "tess_result_vertex_offsets", "tess_result_vertices",
    "tess_result_face_vertex_offsets", "tess_result_face_vertex_indices",
    "tess_result_counters", "tess_result_volume_sum", "tess_result_device_views",
    "tess_plane_histogram", "tess_row_histogram", "tess_bounds", "tess_pack_for_slabs", "tess_pack_records", "tess_diagram_add_records_device",
    "tess_find_neighbors", "tess_find_cells_in_radius", "tess_search_create", "tess_search_expand", "tess_search_cursor", "tess_search_free", "tess_query_free", "tess_query_offsets", "tess_query_indices", "tess_query_status",
    "tess_result_download", "tess_set_main_tier", "tess_result_tier_stats", "tess_kernel_launch_count", "tess_measure_fp64_peak", "tess_result_timings", "tess_diagram_timings",
)


class TessError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"libtess_b200 error {code}: {msg}")
        self.code = code


class Opts(C.Structure):
    _fields_ = [
        ("search_radius", C.c_double),
        ("target_group", C.c_int64),
        ("outputs", C.c_uint32),
        ("table_radius", C.c_int32),
        ("stream", C.c_void_p),
    ]


class Slab(C.Structure):
    _fields_ = [
        ("bounds", C.c_double * 6),
        ("n_global", C.c_uint64),
        ("own_lo", C.c_uint32),
        ("own_hi", C.c_uint32),
        ("local_lo", C.c_uint32),
        ("local_hi", C.c_uint32),
        ("own_lo_row", C.c_uint32),
        ("own_hi_row", C.c_uint32),
    ]


def build(force: bool = False, verbose: bool = False) -> str:
    """Compile libtess_b200.so for sm_100a with nvcc (cross-compiles without a GPU)."""
    srcs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh")) or f == "Makefile"] + [HEADER]
    stale = force or not os.path.exists(LIB_PATH) or any(os.path.getmtime(s) > os.path.getmtime(LIB_PATH) for s in srcs)
    if stale:
        cmd = ["make", "-C", CSRC] + (["-B"] if force else [])
        out = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        if verbose or out.returncode != 0:
            print(out.stdout)
        if out.returncode != 0:
            raise RuntimeError("building libtess_b200.so failed")
    return LIB_PATH


_lib = None


def lib() -> C.CDLL:
    """Load the CUDA library.  Raises if it has not been built: nothing else can do the work."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise FileNotFoundError(
            f"{LIB_PATH} is missing. Build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(nvcc, sm_100a). the-tessellator_b200 has no CPU fallback."
        )
    L = C.CDLL(LIB_PATH)
    vp, u64, f64, i64, ci, sz = C.c_void_p, C.c_uint64, C.c_double, C.c_int64, C.c_int, C.c_size_t
    P = C.POINTER

    def sig(name, res, *args):
        fn = getattr(L, name)
        fn.restype = res
        fn.argtypes = list(args)

    sig("tess_last_error", C.c_char_p)
    sig("tess_version", ci)
    sig("tess_device_count", ci)
    sig("tess_opts_default", None, P(Opts))
    sig("tess_diagram_create", ci, P(vp), ci, ci)
    sig("tess_diagram_destroy", None, vp)
    sig("tess_diagram_add_particles", ci, vp, vp, sz, sz, vp, vp)
    sig("tess_diagram_add_particles_device", ci, vp, vp, sz, vp, vp, vp)
    sig("tess_diagram_clear", ci, vp)
    sig("tess_diagram_initialize", ci, vp, vp, vp)
    sig("tess_diagram_initialize_slab", ci, vp, vp, P(Slab), vp)
    sig("tess_diagram_grid_info", ci, vp, P(u64), P(u64), vp, vp, vp)
    sig("tess_diagram_copy_grid", ci, vp, vp, vp, vp)
    sig("tess_diagram_copy_search_order", ci, vp, C.c_int32, P(u64), vp, vp, P(ci))
    sig("tess_compute_all", ci, vp, P(Opts), P(vp))
    sig("tess_compute_all_to_host", ci, vp, P(Opts), ci, vp, vp, vp, vp, vp, C.c_uint64, C.c_uint64, P(vp))
    sig("tess_compute_at_points", ci, vp, vp, sz, P(Opts), P(vp))
    sig("tess_result_free", None, vp)
    sig("tess_result_n_cells", ci, vp, P(u64), P(u64))
    for n in ("tess_result_volumes", "tess_result_face_offsets", "tess_result_neighbors", "tess_result_areas", "tess_result_status",
              "tess_result_cell_ids", "tess_result_vertex_offsets", "tess_result_vertices", "tess_result_face_vertex_offsets",
              "tess_result_face_vertex_indices"):
        sig(n, ci, vp, P(vp))
    sig("tess_result_counters", ci, vp, P(u64 * 8))
    sig("tess_result_volume_sum", ci, vp, P(f64))
    sig("tess_result_device_views", ci, vp, P(vp), P(vp), P(vp), P(vp), P(vp), P(vp))
    sig("tess_result_download", ci, vp, vp, vp, vp, vp, vp, vp)
    sig("tess_find_neighbors", ci, vp, vp, sz, f64, ci, i64, vp, P(vp))
    sig("tess_find_cells_in_radius", ci, vp, vp, sz, f64, vp, P(vp))
    sig("tess_search_create", ci, vp, vp, sz, P(vp))
    sig("tess_search_expand", ci, vp, f64, u64, vp, P(vp))
    sig("tess_search_cursor", ci, vp, P(vp))
    sig("tess_search_free", None, vp)
    sig("tess_query_free", None, vp)
    for n in ("tess_query_offsets", "tess_query_indices", "tess_query_status"):
        sig(n, ci, vp, P(vp))
    sig("tess_kernel_launch_count", u64)
    sig("tess_set_main_tier", ci, ci)
    sig("tess_result_tier_stats", ci, vp, P(u64 * 4))
    sig("tess_result_timings", ci, vp, P(f64 * 4))
    sig("tess_diagram_timings", ci, vp, P(f64 * 1))
    sig("tess_measure_fp64_peak", ci, ci, P(f64))
    sig("tess_plane_histogram", ci, vp, sz, vp, u64, vp, vp)
    sig("tess_row_histogram", ci, vp, sz, vp, u64, vp, vp)
    sig("tess_bounds", ci, vp, sz, vp, vp)
    sig("tess_pack_for_slabs", ci, vp, vp, i64, sz, vp, u64, ci, vp, vp, vp, vp, vp, sz, vp)
    sig("tess_pack_records", ci, vp, vp, i64, sz, vp, u64, ci, vp, vp, vp, vp, vp, vp, sz, vp)
    sig("tess_diagram_add_records_device", ci, vp, vp, sz, vp)
    _lib = L
    return L


def check(rc: int) -> None:
    if rc != 0:
        raise TessError(rc, lib().tess_last_error().decode(errors="replace"))
