// common.cuh — shared host/device structs of libtess_b200.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <stdexcept>
#include <string>

namespace tess {

// Uniform grid of celery.rs restricted to the x-planes this rank holds.
struct GridSpec {
    double xmin, xmax, ymin, ymax, zmin, zmax;  // CeleryBounds (celery.rs:64-77)
    double sx, sy, sz;                          // cell sizes (celery.rs:132-136)
    double ix, iy, iz;                          // inverse cell sizes (celery.rs:140-144)
    uint32_t cpd;                               // cells_per_dimension (celery.rs:148)
    uint32_t local_lo, local_hi;                // x-planes present locally
    uint32_t own_lo, own_hi;                    // x-planes whose cells are computed here
};

// One entry of the (truncated) search-order table: DistanceIndex of celery.rs:27-32.
struct ShellEntry {
    double key;  // squared min cell-to-cell distance; -1 for the home cell
    int16_t di, dj, dk, pad;
};
static_assert(sizeof(ShellEntry) == 16, "ShellEntry is one 16-byte load");

// Sorted particle record: one 32-byte sector per candidate.
struct __align__(32) Particle {
    double x, y, z;
    int64_t id;  // user-visible ("original") index
};

enum : uint32_t {
    ST_DEGENERATE_SKIP = 1u << 0,
    ST_TABLE_EXHAUSTED = 1u << 1,
    ST_CAPACITY_OVERFLOW = 1u << 2,
    ST_HALO_INSUFFICIENT = 1u << 3,
    ST_INCONSISTENT = 1u << 4,
    ST_LARGE_PATH = 1u << 31,  // internal: the row's faces live in the large-cell staging area
};

enum : int { CNT_VISITED = 0, CNT_TESTED, CNT_VC, CNT_CUTS, CNT_NV, CNT_TABLE, CNT_DEGEN, CNT_FACES, CNT_N };

struct ClipParams {
    const Particle* sorted;        // n_local records in grid order
    const uint32_t* delim;         // local grid cells + 1
    const uint64_t* groups_sorted; // nullable
    const ShellEntry* table;
    uint32_t table_len;
    uint32_t table_full;           // 1: the table covers the whole grid
    GridSpec grid;
    double box[6];                 // container x_min,y_min,z_min,x_max,y_max,z_max
    // work list: either a contiguous slot range, or an explicit list of slots (redo of flagged cells),
    // or explicit query positions (get_cell_at_particle)
    uint32_t slot_begin;
    uint32_t n_work;
    const uint32_t* work_slots;    // nullable; work item w -> sorted slot
    const double* query_xyz;       // nullable; work item w -> position (no self exclusion)
    int64_t target_group;          // -1 = None
    double search_radius;          // NaN = None
    // output row of a cell: row_of_slot[slot] if given, else slot - row_base (query mode: work item)
    const uint32_t* row_of_slot;
    uint32_t row_base;
    // per-row outputs
    double* vol;
    uint32_t* nfaces;
    uint32_t* status;
    int64_t* cell_id;              // nullable
    // face staging [rows][fstride]; indexed by row, or by work item when stage_by_work != 0
    int64_t* st_nbr;
    double* st_area;
    uint32_t fstride;
    uint32_t stage_by_work;
    // optional geometry outputs (TESS_OUT_VERTICES): each cell bump-allocates room for its
    // vertices and its face loops in two pools; they are gathered into CSR order afterwards
    double* gv_xyz;                // nullable; vertex pool, xyz triples (cell-local coordinates)
    uint32_t* gl_idx;              // face-loop pool: per face, the ranks of its vertices in the cell's vertex list
    unsigned long long gv_cap, gl_cap;
    unsigned long long* g_cursor;  // [0] vertices handed out, [1] loop entries handed out
    uint32_t* nverts;              // per row
    uint32_t* nloops;
    unsigned long long* vbase;
    unsigned long long* lbase;
    uint16_t* st_flen;             // per staged face: loop length (indexed like st_nbr)
    unsigned long long* counters;  // CNT_N, nullable
    uint32_t* work_counter;        // dynamic work distribution
    // cells this configuration could not finish (capacity / table exhausted): slots appended here
    uint32_t* failed_slots;        // nullable
    uint32_t* n_failed;            // [0] count of failed cells; [4] those that only exhausted the search table
    uint32_t failed_cap;
    uint32_t mark_large;           // OR ST_LARGE_PATH into the status of every row written
    uint32_t flags;                // A/B checks: bit 0 serial walk only (TESS_FORCE_SERIAL=1), bit 1 always sweep the edge table (TESS_FORCE_SWEEP=1)
};

// radius / neighbour-cloud queries (query.cu)
struct QueryParams {
    const Particle* sorted;
    const uint32_t* delim;
    const uint64_t* groups_sorted;  // nullable
    const ShellEntry* table;        // mode 2
    uint32_t table_len, table_full;
    GridSpec grid;
    const double* xyz;              // n_query positions (device)
    size_t n_query;
    double radius;
    int mode;                       // 0 cell radius, 1 real radius, 2 expand_all_in_radius, 3 find_cells_in_radius (grid cell ids), 4 ExpandingSearch::expand
    const uint64_t* cursor_in;      // mode 4: current_search_index per query (celery.rs:873)
    uint64_t* cursor_out;           // mode 4: ... after the call (written by the counting pass)
    uint64_t cells_to_add;          // mode 4
    int64_t target_group;           // -1 = None
    uint32_t* counts;               // pass 1
    uint32_t* flags;
    const uint64_t* offsets;        // pass 2
    int64_t* indices;
};

#define TESS_CUDA_CHECK(expr)                                                                      \
    do {                                                                                           \
        cudaError_t e__ = (expr);                                                                  \
        if (e__ != cudaSuccess)                                                                    \
            throw std::runtime_error(std::string(#expr) + " failed: " + cudaGetErrorString(e__)); \
    } while (0)

// Kernel launch.  Under TESS_WARP_EMU (tests/emu: the kernel sources compiled by g++ and run lane by lane on the CPU —
// test infrastructure, never part of the library) the launch goes to the emulator instead.
#ifdef TESS_WARP_EMU
#define TESS_LAUNCH(kernel, grid, block, smem, stream, ...) \
    emu_launch_generic((unsigned)(grid), (unsigned)(block), (size_t)(smem), [=]() { kernel(__VA_ARGS__); })
#else
#define TESS_LAUNCH(kernel, grid, block, smem, stream, ...) kernel<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__)
#endif

// every kernel launch of the library is counted (tess_kernel_launch_count)
void note_launch(int n = 1);
unsigned long long launch_count();

// kernels' host launchers (grid.cu / clip.cu / outputs.cu)
void launch_bounds(const double* xyz, size_t n, double* bounds6, cudaStream_t s);
void launch_cell_histogram(const double* xyz, size_t n, const GridSpec& g, uint32_t* cell_of, uint32_t* rank_in_cell, uint32_t* counts, uint32_t* oob_flag, cudaStream_t s);
void launch_exclusive_scan_u32(const uint32_t* in, uint32_t* out, size_t n, void* tmp, size_t tmp_bytes, cudaStream_t s);
size_t scan_tmp_bytes(size_t n);
void launch_scatter_records(const double* xyz, const int64_t* ids, const uint32_t* cell_of, const uint32_t* rank_in_cell, const uint32_t* delim, Particle* arrived, uint32_t* arrived_idx, size_t n, cudaStream_t s);
void launch_rank_fix(const Particle* arrived, const uint32_t* arrived_idx, const GridSpec& g, const uint32_t* delim, const uint64_t* groups, Particle* sorted, uint32_t* sorted_idx, uint64_t* groups_sorted, size_t n, cudaStream_t s);
void launch_plane_histogram(const double* xyz, size_t n, const GridSpec& g, unsigned long long* counts, cudaStream_t s);
void launch_row_histogram(const double* xyz, size_t n, const GridSpec& g, unsigned long long* counts, cudaStream_t s);
void launch_pack_count(const double* xyz, size_t n, const GridSpec& g, int n_ranks, const uint32_t* lo_dev, const uint32_t* hi_dev, unsigned long long* counts, cudaStream_t s);
void launch_pack_scatter(const double* xyz, const int64_t* ids, int64_t id_base, size_t n, const GridSpec& g, int n_ranks, const uint32_t* lo_dev, const uint32_t* hi_dev, const unsigned long long* offsets, const unsigned long long* limits, unsigned long long* cursors, double* out_xyz, int64_t* out_ids, double* out_rec, cudaStream_t s);

// table capacities of the clip kernel (clip.cu); CLIP_SMALL_FAST = the small configuration without the serial walk:
// cells that need it come back flagged like cells that ran out of table and are redone by CLIP_SMALL
// CLIP_THREAD = one thread per cell with tables for the common cell (clip_thread.cu); what it cannot finish is handed back
// the same way
enum : int { CLIP_SMALL = 0, CLIP_MEDIUM = 1, CLIP_LARGE = 2, CLIP_SMALL_FAST = 3, CLIP_THREAD = 4 };
void launch_clip(const ClipParams& p, int tier, cudaStream_t s);
void launch_clip_thread(const ClipParams& p, cudaStream_t s);
uint32_t clip_thread_fmax();
uint32_t clip_thread_vmax();
uint32_t clip_medium_fmax();
uint32_t clip_medium_vmax();
uint32_t clip_small_fmax();
uint32_t clip_small_vmax();
uint32_t clip_large_fmax();
uint32_t clip_large_vmax();

void launch_exclusive_scan_u32_to_u64(const uint32_t* in, uint64_t* out, size_t n, void* tmp, size_t tmp_bytes, cudaStream_t s);
void launch_compact_faces(const uint32_t* status, const uint64_t* offsets, const int64_t* st_nbr, const double* st_area, const uint16_t* st_flen, uint32_t fstride, size_t n_rows, int64_t* nbr, double* area, uint32_t* flen, cudaStream_t s, uint64_t face_cap = ~0ull);
void launch_clear_status_bits(uint32_t* status, size_t n, uint32_t bits, cudaStream_t s);
// streaming: work list (ascending sorted slots) of the rows [row_lo, row_hi)
void launch_chunk_flags(const uint32_t* row_of_slot, uint32_t slot_begin, size_t n, uint32_t row_lo, uint32_t row_hi, uint32_t* flags, cudaStream_t s);
void launch_chunk_scatter(const uint32_t* flags, const uint64_t* pos, uint32_t slot_begin, size_t n, uint32_t* work_slots, cudaStream_t s);
void launch_compact_redo(const uint32_t* work_slots, const uint32_t* row_of_slot, uint32_t row_base, const uint32_t* nfaces, const uint64_t* offsets, const int64_t* st_nbr, const double* st_area, const uint16_t* st_flen, uint32_t fstride, size_t n_work, int64_t* nbr, double* area, uint32_t* flen, cudaStream_t s);
void launch_gather_vertices(const uint32_t* nverts, const unsigned long long* vbase, const uint64_t* voffsets, const double* pool, size_t n_rows, double* vtx, cudaStream_t s);
void launch_gather_loops(const uint32_t* nloops, const unsigned long long* lbase, const uint64_t* face_offsets, const uint64_t* fv_offsets, const uint32_t* pool, size_t n_rows, uint32_t* out, cudaStream_t s);
void launch_volume_sum(const double* vol, size_t n, double* out, cudaStream_t s);
double measure_fp64_peak_tflops();
void launch_radius_query(const QueryParams& p, bool fill, cudaStream_t s);

}  // namespace tess
