//! 1:1 declarations of include/tess.h.
#![allow(non_camel_case_types)]
use std::os::raw::{c_char, c_int, c_void};

#[repr(C)]
pub struct tess_diagram {
    _private: [u8; 0],
}
#[repr(C)]
pub struct tess_result {
    _private: [u8; 0],
}
#[repr(C)]
pub struct tess_query {
    _private: [u8; 0],
}
#[repr(C)]
pub struct tess_search {
    _private: [u8; 0],
}

pub const TESS_OK: c_int = 0;
pub const TESS_F64: c_int = 0;
pub const TESS_OUT_VOLUME: u32 = 1;
pub const TESS_OUT_NEIGHBORS: u32 = 2;
pub const TESS_OUT_AREAS: u32 = 4;
pub const TESS_OUT_VERTICES: u32 = 8;
pub const TESS_QUERY_CELL_RADIUS: c_int = 0;
pub const TESS_QUERY_REAL_RADIUS: c_int = 1;
pub const TESS_QUERY_NEIGHBOR_CLOUD: c_int = 2;

#[repr(C)]
#[derive(Clone, Copy)]
pub struct tess_opts {
    pub search_radius: f64,
    pub target_group: i64,
    pub outputs: u32,
    pub table_radius: i32,
    pub stream: *mut c_void,
}

/// One rank's part of a slab-sharded diagram (tess.h `tess_slab`).
#[repr(C)]
#[derive(Clone, Copy)]
pub struct tess_slab {
    pub bounds: [f64; 6],
    pub n_global: u64,
    pub own_lo: u32,
    pub own_hi: u32,
    pub local_lo: u32,
    pub local_hi: u32,
    pub own_lo_row: u32,
    pub own_hi_row: u32,
}

extern "C" {
    pub fn tess_last_error() -> *const c_char;
    pub fn tess_opts_default(o: *mut tess_opts);
    pub fn tess_diagram_create(out: *mut *mut tess_diagram, real_type: c_int, device: c_int) -> c_int;
    pub fn tess_diagram_destroy(d: *mut tess_diagram);
    pub fn tess_diagram_add_particles(d: *mut tess_diagram, xyz: *const c_void, n: usize, stride_bytes: usize, groups: *const u64, stream: *mut c_void) -> c_int;
    pub fn tess_diagram_initialize(d: *mut tess_diagram, box6: *const f64, stream: *mut c_void) -> c_int;
    pub fn tess_diagram_grid_info(d: *const tess_diagram, n_points: *mut u64, cpd: *mut u64, bounds: *mut f64, sizes: *mut f64, inv_sizes: *mut f64) -> c_int;
    pub fn tess_compute_all(d: *const tess_diagram, opts: *const tess_opts, out: *mut *mut tess_result) -> c_int;
    pub fn tess_compute_all_to_host(d: *const tess_diagram, opts: *const tess_opts, n_chunks: c_int, volumes: *mut f64, face_offsets: *mut u64,
                                    neighbors: *mut i64, areas: *mut f64, status: *mut u32, cell_capacity: u64, face_capacity: u64, out: *mut *mut tess_result) -> c_int;
    pub fn tess_compute_at_points(d: *const tess_diagram, xyz: *const f64, m: usize, opts: *const tess_opts, out: *mut *mut tess_result) -> c_int;
    pub fn tess_result_free(r: *mut tess_result);
    pub fn tess_result_n_cells(r: *const tess_result, n_cells: *mut u64, n_faces: *mut u64) -> c_int;
    pub fn tess_result_volumes(r: *mut tess_result, out: *mut *const f64) -> c_int;
    pub fn tess_result_face_offsets(r: *mut tess_result, out: *mut *const u64) -> c_int;
    pub fn tess_result_neighbors(r: *mut tess_result, out: *mut *const i64) -> c_int;
    pub fn tess_result_areas(r: *mut tess_result, out: *mut *const f64) -> c_int;
    pub fn tess_result_status(r: *mut tess_result, out: *mut *const u32) -> c_int;
    pub fn tess_result_vertex_offsets(r: *mut tess_result, out: *mut *const u64) -> c_int;
    pub fn tess_result_vertices(r: *mut tess_result, out: *mut *const f64) -> c_int;
    pub fn tess_result_face_vertex_offsets(r: *mut tess_result, out: *mut *const u64) -> c_int;
    pub fn tess_result_face_vertex_indices(r: *mut tess_result, out: *mut *const u32) -> c_int;
    pub fn tess_result_cell_ids(r: *mut tess_result, out: *mut *const i64) -> c_int;
    pub fn tess_result_counters(r: *mut tess_result, counters: *mut u64) -> c_int;
    pub fn tess_result_volume_sum(r: *mut tess_result, out: *mut f64) -> c_int;
    pub fn tess_version() -> c_int;
    pub fn tess_device_count() -> c_int;
    pub fn tess_diagram_clear(d: *mut tess_diagram) -> c_int;
    pub fn tess_kernel_launch_count() -> u64;
    // radius / neighbour-cloud queries (celery.rs:802-855, :1023-1075; interface.rs:348-365)
    pub fn tess_find_neighbors(d: *const tess_diagram, xyz: *const f64, m: usize, radius: f64, mode: c_int, target_group: i64, stream: *mut c_void,
                               out: *mut *mut tess_query) -> c_int;
    pub fn tess_find_cells_in_radius(d: *const tess_diagram, xyz: *const f64, m: usize, radius: f64, stream: *mut c_void, out: *mut *mut tess_query) -> c_int;
    // ExpandingSearch (celery.rs:865-1075)
    pub fn tess_search_create(d: *const tess_diagram, xyz: *const f64, m: usize, out: *mut *mut tess_search) -> c_int;
    pub fn tess_search_expand(s: *mut tess_search, max_radius: f64, cells_to_add: u64, stream: *mut c_void, out: *mut *mut tess_query) -> c_int;
    pub fn tess_search_cursor(s: *const tess_search, current_search_index: *mut *const u64) -> c_int;
    pub fn tess_search_free(s: *mut tess_search);
    // slab partition (one process per GPU; the two collectives — all-reduce and all-to-all — are the host application's)
    pub fn tess_bounds(xyz_dev: *const f64, n: usize, bounds_dev: *mut f64, stream: *mut c_void) -> c_int;
    pub fn tess_plane_histogram(xyz_dev: *const f64, n: usize, bounds: *const f64, n_global: u64, counts_dev: *mut u64, stream: *mut c_void) -> c_int;
    pub fn tess_row_histogram(xyz_dev: *const f64, n: usize, bounds: *const f64, n_global: u64, counts_dev: *mut u64, stream: *mut c_void) -> c_int;
    pub fn tess_pack_records(xyz_dev: *const f64, ids_dev: *const i64, id_base: i64, n: usize, bounds: *const f64, n_global: u64, n_ranks: c_int, plane_lo: *const u32,
                             plane_hi: *const u32, planned_counts: *const u64, counts_host: *mut u64, counts_dev: *mut u64, out_rec_dev: *mut f64, cap: usize,
                             stream: *mut c_void) -> c_int;
    pub fn tess_diagram_add_records_device(d: *mut tess_diagram, rec_dev: *const f64, n: usize, stream: *mut c_void) -> c_int;
    pub fn tess_diagram_add_particles_device(d: *mut tess_diagram, xyz_dev: *const f64, n: usize, groups_dev: *const u64, ids_dev: *const i64, stream: *mut c_void) -> c_int;
    pub fn tess_diagram_initialize_slab(d: *mut tess_diagram, box6: *const f64, slab: *const tess_slab, stream: *mut c_void) -> c_int;
    pub fn tess_result_download(r: *const tess_result, volumes: *mut f64, face_offsets: *mut u64, neighbors: *mut i64, areas: *mut f64, status: *mut u32, stream: *mut c_void) -> c_int;
    pub fn tess_result_device_views(r: *const tess_result, volumes: *mut *const f64, face_offsets: *mut *const u64, neighbors: *mut *const i64, areas: *mut *const f64,
                                    status: *mut *const u32, cell_ids: *mut *const i64) -> c_int;
    pub fn tess_query_free(q: *mut tess_query);
    pub fn tess_query_offsets(q: *mut tess_query, out: *mut *const u64) -> c_int;
    pub fn tess_query_indices(q: *mut tess_query, out: *mut *const i64) -> c_int;
    pub fn tess_query_status(q: *mut tess_query, out: *mut *const u32) -> c_int;
}
