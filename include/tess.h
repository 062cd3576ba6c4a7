/* tess.h — C ABI of the B200-native Voronoi cell builder (libtess_b200.so).
 *
 * Drop-in boundary for ONE hot path of mcomstock/the-tessellator: per-particle Voronoi cell
 * construction (uniform-grid binning -> shell-ordered neighbour iteration -> half-space clipping
 * -> volume / face areas / neighbour list).  The reference has no FFI layer; its boundary is the
 * Rust API of src/interface.rs.  Each entry point below names the reference item it replaces.
 * The reference API is per-cell (`Diagram::get_cell_at_index(i).compute_voronoi_cell()`); this
 * ABI is batch (all cells of a diagram per call) and the per-cell Rust/C++/Python wrappers read
 * rows of the batch result (INTEGRATION.md shows the Rust-side binding).
 *
 * Conventions
 *   - every function returns 0 (TESS_OK) or a negative tess_error; tess_last_error() gives text
 *     (thread-local).  No C++ exception crosses this boundary.
 *   - the library owns all device memory; callers free only through *_destroy / *_free.
 *   - ids are 64-bit to match Rust `usize`; container-wall faces are reported as neighbour ids
 *     -1..-6 (y_min, x_max, y_max, x_min, z_max, z_min = the reference's face slots F,R,B,L,U,D,
 *     polyhedron.rs:74-81) instead of the reference's `unwrap()` panic (polyhedron.rs:877).
 *   - there is NO CPU fallback: every compute entry point fails with TESS_ERR_CUDA when no
 *     sm_100 device is usable.
 *   - `stream` arguments are `cudaStream_t` passed as void* (NULL = the legacy default stream).
 */
#ifndef TESS_H_
#define TESS_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct tess_diagram tess_diagram; /* interface.rs:25  struct Diagram */
typedef struct tess_result tess_result;   /* the Vec<..> returns of interface.rs:337-384, batched */

typedef enum tess_error {
    TESS_OK = 0,
    TESS_ERR_INVALID = -1,     /* bad argument */
    TESS_ERR_STATE = -2,       /* call order (e.g. add after initialize; interface.rs:65 debug_assert) */
    TESS_ERR_CUDA = -3,        /* CUDA runtime failure or no usable device */
    TESS_ERR_NOMEM = -4,
    TESS_ERR_UNSUPPORTED = -5, /* e.g. real_type != TESS_F64 in this build */
    TESS_ERR_CAPACITY = -6     /* a cell exceeded even the large-cell path's tables */
} tess_error;

/* float.rs:78 — the reference only implements Float64. */
enum { TESS_F64 = 0, TESS_F32 = 1 };

/* tess_opts.outputs */
enum {
    TESS_OUT_VOLUME = 1u << 0,    /* Cell::compute_volume            interface.rs:337 */
    TESS_OUT_NEIGHBORS = 1u << 1, /* Cell::compute_neighbors         interface.rs:342 */
    TESS_OUT_AREAS = 1u << 2,     /* VoronoiFace::compute_area       interface.rs:408 */
    TESS_OUT_VERTICES = 1u << 3,  /* Cell::compute_vertices          interface.rs:368 */
    TESS_OUT_COUNTERS = 1u << 4   /* work counters */
};

/* per-cell status word (tess_result_status) */
enum {
    TESS_STATUS_OK = 0,
    TESS_STATUS_DEGENERATE_SKIP = 1u << 0,   /* a plane had vertices outside but no strictly Inside->Outside
                                                edge and was skipped: the reference's behaviour,
                                                polyhedron.rs:413-434 */
    TESS_STATUS_TABLE_EXHAUSTED = 1u << 1,   /* shell table ran out before the termination test fired */
    TESS_STATUS_CAPACITY_OVERFLOW = 1u << 2, /* vertex/edge/face table overflow (cell left unfinished) */
    TESS_STATUS_HALO_INSUFFICIENT = 1u << 3, /* slab mode: the search reached a grid plane this rank does not hold */
    TESS_STATUS_INCONSISTENT = 1u << 4       /* mesh invariant broken by floating-point fuzz */
};

/* Options of one batch computation = the per-cell arguments of Diagram::get_cell_at_index
 * (interface.rs:186-192).  Initialise with tess_opts_default(). */
typedef struct tess_opts {
    double search_radius;  /* NaN = None: expanding search with the results-preserving 2*r_max
                              termination.  Otherwise ExpandingSearch::expand_all_in_radius(r)
                              semantics (celery.rs:1023-1075, compares the squared table key with r). */
    int64_t target_group;  /* -1 = None; else only particles of this group cut (interface.rs:280-297) */
    uint32_t outputs;      /* TESS_OUT_* mask */
    int32_t table_radius;  /* half-width R of the precomputed shell-offset table; 0 = default */
    void* stream;          /* cudaStream_t */
} tess_opts;

/* Global grid description for slab-sharded (multi-GPU) diagrams.  Every rank passes the SAME
 * bounds / n_global so that all ranks bin with identical parameters (celery.rs:81-189). */
typedef struct tess_slab {
    double bounds[6];    /* x_min,x_max,y_min,y_max,z_min,z_max of ALL points (celery.rs:117-124) */
    uint64_t n_global;   /* total number of points: cpd = floor(cbrt(n/1.25))+1 (celery.rs:161-162) */
    uint32_t own_lo;     /* owned grid x-planes [own_lo, own_hi) : cells of points in these planes are computed */
    uint32_t own_hi;
    uint32_t local_lo;   /* x-planes [local_lo, local_hi) held by this rank (owned + halo) */
    uint32_t local_hi;
    uint32_t own_lo_row; /* finer ownership: the owned cells are the grid rows (x, y) with own_lo*cpd + own_lo_row <= x*cpd + y <  */
    uint32_t own_hi_row; /* own_hi*cpd + own_hi_row (cell ids are x-major, celery.rs:323-324: a contiguous run of the sorted order). */
                         /* 0, 0 = whole planes.  A partly owned plane own_hi must be held locally (own_hi < local_hi).                */
} tess_slab;

const char* tess_last_error(void);
int tess_version(void);
/* number of usable sm_100 devices (0 if none); never fails */
int tess_device_count(void);

void tess_opts_default(tess_opts* o);

/* ---- Diagram (interface.rs:25-233) ------------------------------------------------------- */

/* Diagram::default()  interface.rs:24.  device = CUDA ordinal. */
int tess_diagram_create(tess_diagram** out, int real_type, int device);
void tess_diagram_destroy(tess_diagram* d);

/* Diagram::add_particle_with_group x n  (interface.rs:52-57).  Host AoS: point i has its x,y,z
 * (f64) at xyz + i*stride_bytes (ToCeleryPoint getters, celery.rs:56-60).  groups may be NULL
 * (all group 0).  May be called repeatedly before initialize. */
int tess_diagram_add_particles(tess_diagram* d, const void* xyz, size_t n, size_t stride_bytes, const uint64_t* groups, void* stream);
/* Same, from device memory: packed f64 triples.  ids (device, nullable) are the user-visible
 * ("original", interface.rs:36) indices reported as neighbours; default = insertion order. */
int tess_diagram_add_particles_device(tess_diagram* d, const double* xyz_dev, size_t n, const uint64_t* groups_dev, const int64_t* ids_dev, void* stream);
/* Drop all particles but keep device workspaces (lets a handle be reused step after step). */
int tess_diagram_clear(tess_diagram* d);

/* Diagram::initialize (interface.rs:60-84): bounds, cell sizing, binning (Celery::reset,
 * celery.rs:253-266).  box = container x_min,y_min,z_min,x_max,y_max,z_max (the arguments of
 * Polyhedron::new, polyhedron.rs:226-233); NULL = bounding box of the points. */
int tess_diagram_initialize(tess_diagram* d, const double box[6], void* stream);
/* Slab-sharded variant: this rank holds the particles of grid x-planes [local_lo, local_hi) of a
 * global grid and computes the cells of [own_lo, own_hi).  box must be given. */
int tess_diagram_initialize_slab(tess_diagram* d, const double box[6], const tess_slab* slab, void* stream);

/* Grid facts (celery.rs:146-148, 213, 117-124), valid after initialize. */
int tess_diagram_grid_info(const tess_diagram* d, uint64_t* n_points, uint64_t* cells_per_dimension, double bounds[6], double cell_sizes[3], double inverse_cell_sizes[3]);
/* Copies of Celery::{cells, sorted_indices, delimiters} (celery.rs:198-215) for parity tests.
 * Any pointer may be NULL.  cells/sorted_indices: n entries; delimiters: local cells + 1. */
int tess_diagram_copy_grid(const tess_diagram* d, uint64_t* cells, uint64_t* sorted_indices, uint64_t* delimiters);
/* Bit-exact copy of the truncated, canonically ordered search table (celery.rs:418-679) the
 * clip kernel walks: keys[len], ijk[3*len].  Pass NULL pointers to query len only. */
int tess_diagram_copy_search_order(const tess_diagram* d, int32_t table_radius, uint64_t* len, double* keys, int32_t* ijk, int* is_full);

/* ---- Cells (interface.rs:237-417), batched ----------------------------------------------- */

/* For every particle (every owned particle in slab mode):
 *   Diagram::get_cell_at_index(i, Polyhedron::new(box), search_radius, target_group)
 *   .compute_voronoi_cell()  + compute_volume / compute_neighbors / compute_faces->compute_area.
 * Result rows: whole-domain diagrams -> row i is particle i (insertion order); slab diagrams ->
 * rows follow the rank's grid order and tess_result_cell_ids gives the particle id of each row. */
int tess_compute_all(const tess_diagram* d, const tess_opts* opts, tess_result** out);
/* The same call for a caller that wants the cells in HOST memory — what a loop over
 * Diagram::get_cell_at_index(i) + compute_voronoi_cell + compute_volume / compute_neighbors /
 * compute_faces (interface.rs:186-207, :257-384) leaves behind.  The rows are computed in n_chunks
 * groups (0 = default 8); a finished group is packed and copied to the host buffers on a second
 * stream while the next groups are still being computed, so the device->host transfer overlaps the
 * clip kernel (pass page-locked buffers; pageable ones work but do not overlap).
 *   volumes[cell_capacity], face_offsets[cell_capacity+1], status[cell_capacity],
 *   neighbors[face_capacity], areas[face_capacity];
 * any pointer may be NULL (areas must be NULL unless TESS_OUT_AREAS is set); more cells than
 * cell_capacity or more faces than face_capacity -> TESS_ERR_CAPACITY, nothing is written past them.  The contents equal tess_compute_all + tess_result_download bit for
 * bit; *out holds the same device arrays (n = the rank's owned cells for slab diagrams).  Tiny inputs
 * are not chunked (one copy at the end); TESS_OUT_VERTICES is not supported here.  On return the host
 * buffers are complete. */
int tess_compute_all_to_host(const tess_diagram* d, const tess_opts* opts, int n_chunks, double* volumes, uint64_t* face_offsets, int64_t* neighbors, double* areas,
                             uint32_t* status, uint64_t cell_capacity, uint64_t face_capacity, tess_result** out);
/* Diagram::get_cell_at_particle (interface.rs:211-232): cells of m arbitrary positions (host,
 * packed f64 triples) that are not particles of the diagram (no self exclusion). */
int tess_compute_at_points(const tess_diagram* d, const double* xyz, size_t m, const tess_opts* opts, tess_result** out);

void tess_result_free(tess_result* r);
int tess_result_n_cells(const tess_result* r, uint64_t* n_cells, uint64_t* n_faces);
/* Host views (copied from the device on first use, valid until tess_result_free). */
int tess_result_volumes(tess_result* r, const double** out);         /* n_cells */
int tess_result_face_offsets(tess_result* r, const uint64_t** out);  /* n_cells+1, CSR */
int tess_result_neighbors(tess_result* r, const int64_t** out);      /* n_faces; walls -1..-6 */
int tess_result_areas(tess_result* r, const double** out);           /* n_faces */
int tess_result_status(tess_result* r, const uint32_t** out);        /* n_cells, TESS_STATUS_* */
int tess_result_cell_ids(tess_result* r, const int64_t** out);       /* n_cells */
int tess_result_vertex_offsets(tess_result* r, const uint64_t** out); /* n_cells+1 (TESS_OUT_VERTICES) */
int tess_result_vertices(tess_result* r, const double** out);         /* xyz triples, cell-local coordinates */
/* VoronoiFace::compute_vertices (interface.rs:403-405 -> Polyhedron::compute_face_vertices, polyhedron.rs:897-919):
 * for face k (same numbering as neighbours/areas) the loop entries face_vertex_offsets[k] .. [k+1] are indices into
 * the OWNING CELL's vertex list (add vertex_offsets[cell]); order = from the face's starting edge, following next. */
int tess_result_face_vertex_offsets(tess_result* r, const uint64_t** out); /* n_faces+1 (TESS_OUT_VERTICES) */
int tess_result_face_vertex_indices(tess_result* r, const uint32_t** out);
/* counters[8] = candidates visited, candidates tested, vertex classifications, cuts,
 * new vertices, table entries consumed, degenerate skips, faces */
int tess_result_counters(tess_result* r, uint64_t counters[8]);
/* Sum of all volumes computed on the device (closure check: equals the container volume). */
int tess_result_volume_sum(tess_result* r, double* out);
/* Copy results into caller-owned host buffers (pinned memory makes the copies asynchronous and
 * full-speed).  Any pointer may be NULL; sizes are those of the host views above.  The copies are
 * enqueued on `stream`; the caller synchronises. */
int tess_result_download(const tess_result* r, double* volumes, uint64_t* face_offsets, int64_t* neighbors, double* areas, uint32_t* status, void* stream);
/* Device views for callers that keep results on the GPU (any pointer may be NULL). */
int tess_result_device_views(const tess_result* r, const double** volumes, const uint64_t** face_offsets, const int64_t** neighbors, const double** areas, const uint32_t** status, const int64_t** cell_ids);

/* ---- Radius queries on the grid (celery.rs:753-855, 1023-1075; interface.rs:348-365) --------- */

typedef struct tess_query tess_query;
enum {
    TESS_QUERY_CELL_RADIUS = 0,    /* Celery::find_neighbors_in_cell_radius (celery.rs:802): every particle of every grid
                                      cell within `radius` of the query's cell (adjacent cells count as distance 0) */
    TESS_QUERY_REAL_RADIUS = 1,    /* Celery::find_neighbors_in_real_radius (celery.rs:825): additionally |p - q|^2 <= r^2 */
    TESS_QUERY_NEIGHBOR_CLOUD = 2  /* ExpandingSearch::expand_all_in_radius (celery.rs:1023) as used by
                                      Cell::compute_neighbor_cloud (interface.rs:348): search-table order, stops at the first
                                      entry whose squared key exceeds `radius`; target_group filters (interface.rs:359) */
};
/* m query positions (host, packed f64 triples) -> CSR lists of particle ids in the reference's order.
 * Whole-domain diagrams only. */
int tess_find_neighbors(const tess_diagram* d, const double* xyz, size_t m, double radius, int mode, int64_t target_group, void* stream, tess_query** out);
/* Celery::find_cells_in_radius (celery.rs:753-797): per query position the ids (x*cpd^2 + y*cpd + z, celery.rs:317-325) of the
 * grid cells whose cell distance to the position's cell is within `radius`, in the reference's i, j, k loop order.  The
 * result's indices are grid cell ids, not particle ids. */
int tess_find_cells_in_radius(const tess_diagram* d, const double* xyz, size_t m, double radius, void* stream, tess_query** out);

/* ExpandingSearch (celery.rs:865-1075) as an object: one cursor (current_search_index, celery.rs:873) per position.
 * tess_search_create = ExpandingSearch::new (celery.rs:882-902) for m positions (host, packed f64 triples).
 * tess_search_expand = ExpandingSearch::expand(max_radius, cells_to_add) (celery.rs:907-963) for all of them at once: walks
 * at most cells_to_add search-table entries from each cursor on (entries outside the grid count), stops before the first
 * entry whose squared key exceeds max_radius, appends the particles of the visited cells (original indices, sorted order
 * within a cell) and leaves the cursors where the walks stopped.  expand_all_in_radius(r) = expand(r, UINT64_MAX) on a
 * fresh search; expand_all_no_radius = expand(+inf, UINT64_MAX) (celery.rs:971-1018). */
typedef struct tess_search tess_search;
int tess_search_create(const tess_diagram* d, const double* xyz, size_t m, tess_search** out);
int tess_search_expand(tess_search* s, double max_radius, uint64_t cells_to_add, void* stream, tess_query** out);
int tess_search_cursor(const tess_search* s, const uint64_t** current_search_index);
void tess_search_free(tess_search* s);
void tess_query_free(tess_query* q);
int tess_query_offsets(tess_query* q, const uint64_t** out);  /* m+1 */
int tess_query_indices(tess_query* q, const int64_t** out);   /* offsets[m] particle ids */
int tess_query_status(tess_query* q, const uint32_t** out);   /* m; TESS_STATUS_TABLE_EXHAUSTED cannot occur: the table is sized for the radius */

/* ---- Slab partition helpers (multi-GPU; the collectives themselves are the caller's) ------ */

/* Histogram of particles per global grid x-plane: counts_dev[cpd] (u64, device, zeroed by the call). */
int tess_plane_histogram(const double* xyz_dev, size_t n, const double bounds[6], uint64_t n_global, uint64_t* counts_dev, void* stream);
/* The same per grid row (x, y): counts_dev[cpd*cpd] (u64, device, zeroed by the call), index x*cpd + y — for slab cuts finer
 * than a plane (tess_slab.own_lo_row / own_hi_row). */
int tess_row_histogram(const double* xyz_dev, size_t n, const double bounds[6], uint64_t n_global, uint64_t* counts_dev, void* stream);
/* min/max of packed xyz on the device -> bounds_dev[6] (x_min,x_max,y_min,y_max,z_min,z_max). */
int tess_bounds(const double* xyz_dev, size_t n, double* bounds_dev, void* stream);
/* Route particles to slabs.  plane_lo/plane_hi[g] (host, n_ranks entries) give for destination
 * rank g the x-plane range it must RECEIVE (owned + halo).  A particle is sent to every rank whose
 * range contains its plane.  Outputs (device): send_counts_dev[n_ranks] (u64), and, packed rank
 * after rank, out_xyz_dev / out_ids_dev (capacity `cap` particles).  ids_dev = ids of the inputs
 * (NULL -> id_base + i).  Returns TESS_ERR_NOMEM if cap is too small (send_counts_dev still valid). */
int tess_pack_for_slabs(const double* xyz_dev, const int64_t* ids_dev, int64_t id_base, size_t n, const double bounds[6], uint64_t n_global, int n_ranks, const uint32_t* plane_lo, const uint32_t* plane_hi, uint64_t* send_counts_dev, double* out_xyz_dev, int64_t* out_ids_dev, size_t cap, void* stream);

/* ---- Telemetry used by bench.py ----------------------------------------------------------- */

/* The same routing with one 32-byte record {x, y, z, id (bit pattern of the i64)} per particle, so that ONE all-to-all
 * carries positions and ids, and without a host round trip when the counts are known: planned_counts (host, n_ranks, or
 * NULL) are the per-destination counts of an earlier call on the same particle set; the call then runs no counting pass
 * and does not synchronise the stream.  counts_dev[n_ranks] (device) receives the counts actually packed — with a plan,
 * compare them with it (a mismatch means the set changed and the packed records must be discarded);
 * counts_host (host, nullable) receives the counts used for the layout.  Returns TESS_ERR_NOMEM if cap (records) is too small. */
int tess_pack_records(const double* xyz_dev, const int64_t* ids_dev, int64_t id_base, size_t n, const double bounds[6], uint64_t n_global, int n_ranks, const uint32_t* plane_lo, const uint32_t* plane_hi, const uint64_t* planned_counts, uint64_t* counts_host, uint64_t* counts_dev, double* out_rec_dev, size_t cap, void* stream);
/* Diagram::add_particle (interface.rs:44-56) for n device-resident 32-byte records of tess_pack_records (after the exchange). */
int tess_diagram_add_records_device(tess_diagram* d, const double* rec_dev, size_t n, void* stream);

/* CUDA-event durations (ms, on the launching stream) of the last computation:
 * ms[0] clip kernel (small-cell pass), ms[1] redo passes (wider table, medium and large configurations), ms[2] scans + CSR compaction, ms[3] whole call. */
int tess_result_timings(const tess_result* r, double ms[4]);
/* ms[0] = binning pass (histogram + scan + scatter + gather) of the last initialize. */
int tess_diagram_timings(const tess_diagram* d, double ms[1]);
/* Which kernel runs the main clip pass (process-wide; a tuning and test knob, not part of the reference's API):
 * -1 default (thread per cell; warp per cell for geometry output and query cells), 0 warp per cell with the serial walk,
 * 3 warp per cell without it, 4 thread per cell.  Results are bit-identical for every choice.  The environment variable
 * TESS_MAIN_TIER=thread|fast|small sets the initial value. */
int tess_set_main_tier(int tier);
/* stats[0] = tier that ran the main pass of this result, stats[1..3] = cells redone by the wider-table / medium / large passes. */
int tess_result_tier_stats(const tess_result* r, uint64_t stats[4]);
/* Number of CUDA kernels this library has launched in this process so far. */
uint64_t tess_kernel_launch_count(void);
/* Measures the device's FP64 FMA throughput with a register-resident DFMA loop (a denominator for
 * the clip kernel's roofline; MEASURED_PEAKS.json has no FP64 figure).  Result in TFLOP/s. */
int tess_measure_fp64_peak(int device, double* tflops);

#ifdef __cplusplus
}
#endif
#endif /* TESS_H_ */
