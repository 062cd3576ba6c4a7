// emu_clip.cpp — TEST INFRASTRUCTURE.  Compiles the-tessellator_b200/csrc/clip.cu, as it is, for the
// CPU warp emulator (warp_emu.hpp) and exposes one C entry point that runs the clip kernel on grid
// arrays supplied by the caller (the tests take them from the CPU oracle).  Used by
// tests/test_emu_clip.py only; the product library never sees this file.
#include <cuda_runtime.h>  // resolves to tests/emu/shim/cuda_runtime.h

#include <vector>

static unsigned g_os_threads = 1, g_blocks = 1;
static bool g_reverse = false;
static unsigned long long g_collectives = 0;

static void emu_launch_kernel(void (*entry)(const void*), const void* arg, unsigned nthreads, size_t smem) {
    const emu::LaunchStats st = emu::launch(entry, arg, g_blocks, nthreads, smem, g_os_threads, g_reverse);
    g_collectives += st.collectives;
}

#include "../../the-tessellator_b200/csrc/clip.cu"
#include "../../the-tessellator_b200/csrc/clip_thread.cu"

namespace tess {
void note_launch(int) {}
unsigned long long launch_count() { return 0; }
}  // namespace tess

extern "C" {

struct emu_clip_args {
    // grid (celery.rs) as built by the oracle
    const double* particles;  // n x {x, y, z, id (int64 bits)} in grid order
    uint32_t n;
    const uint32_t* delim;    // cpd^3 + 1
    const double* table_key;
    const int32_t* table_ijk;
    uint32_t table_len;
    uint32_t table_full;
    double bounds[6];     // x_min, x_max, y_min, y_max, z_min, z_max
    double cell_info[6];  // sizes xyz, inverse sizes xyz
    uint32_t cpd;
    double box[6];
    const uint64_t* groups_sorted;  // nullable
    // work
    const uint32_t* work_slots;  // nullable: all n cells in slot order
    uint32_t n_work;
    const double* query_xyz;     // nullable
    int64_t target_group;
    double search_radius;
    uint32_t flags;
    int32_t large;
    // outputs, one row per work item
    uint32_t fstride;
    double* vol;
    uint32_t* nfaces;
    uint32_t* status;
    int64_t* cell_id;
    int64_t* st_nbr;
    double* st_area;
    uint64_t* counters;     // 8, nullable
    uint32_t* failed_slots; // n_work, nullable
    uint32_t* n_failed;     // 8
    // optional geometry outputs (TESS_OUT_VERTICES): pools + per-row records, all nullable together
    double* gv_xyz;
    uint32_t* gl_idx;
    uint64_t gv_cap, gl_cap;
    uint64_t* g_cursor;     // 2
    uint32_t* nverts;
    uint32_t* nloops;
    uint64_t* vbase;
    uint64_t* lbase;
    uint16_t* st_flen;
    // emulator
    uint32_t os_threads, blocks, reverse;
    uint64_t collectives;   // out
    // slab diagrams: the x-planes [local_lo, local_hi) are held (delim covers only them); 0, 0 = all
    uint32_t local_lo, local_hi;
};

int emu_clip_run(emu_clip_args* a) {
    using namespace tess;
    std::vector<unsigned char> praw((size_t)a->n * sizeof(Particle) + 64);
    Particle* parts = reinterpret_cast<Particle*>((reinterpret_cast<uintptr_t>(praw.data()) + 31) & ~uintptr_t(31));
    memcpy(parts, a->particles, (size_t)a->n * sizeof(Particle));
    std::vector<ShellEntry> table(a->table_len);
    for (uint32_t t = 0; t < a->table_len; ++t) {
        table[t].key = a->table_key[t];
        table[t].di = (int16_t)a->table_ijk[3 * t];
        table[t].dj = (int16_t)a->table_ijk[3 * t + 1];
        table[t].dk = (int16_t)a->table_ijk[3 * t + 2];
        table[t].pad = 0;
    }
    ClipParams P{};
    P.sorted = parts;
    P.delim = a->delim;
    P.groups_sorted = a->groups_sorted;
    P.table = table.data();
    P.table_len = a->table_len;
    P.table_full = a->table_full;
    GridSpec& g = P.grid;
    g.xmin = a->bounds[0]; g.xmax = a->bounds[1]; g.ymin = a->bounds[2]; g.ymax = a->bounds[3]; g.zmin = a->bounds[4]; g.zmax = a->bounds[5];
    g.sx = a->cell_info[0]; g.sy = a->cell_info[1]; g.sz = a->cell_info[2];
    g.ix = a->cell_info[3]; g.iy = a->cell_info[4]; g.iz = a->cell_info[5];
    g.cpd = a->cpd;
    g.local_lo = 0; g.local_hi = a->cpd; g.own_lo = 0; g.own_hi = a->cpd;
    if (a->local_hi > a->local_lo) { g.local_lo = g.own_lo = a->local_lo; g.local_hi = g.own_hi = a->local_hi; }
    for (int i = 0; i < 6; ++i) P.box[i] = a->box[i];
    P.slot_begin = 0;
    P.n_work = a->n_work;
    P.work_slots = a->work_slots;
    P.query_xyz = a->query_xyz;
    P.target_group = a->target_group;
    P.search_radius = a->search_radius;
    P.row_of_slot = nullptr;
    P.row_base = 0;
    P.vol = a->vol; P.nfaces = a->nfaces; P.status = a->status; P.cell_id = a->cell_id;
    P.st_nbr = a->st_nbr; P.st_area = a->st_area; P.fstride = a->fstride;
    P.stage_by_work = 1;  // rows = work items
    std::vector<uint32_t> rows;
    if (a->work_slots && !a->query_xyz) {
        // the kernel writes per-row results at row_of_slot[slot]: make that the work item
        rows.assign(a->n, 0u);
        for (uint32_t w = 0; w < a->n_work; ++w) rows[a->work_slots[w]] = w;
        P.row_of_slot = rows.data();
    }
    unsigned long long counters[CNT_N] = {0};
    P.counters = a->counters ? counters : nullptr;
    uint32_t work_counter = 0;
    P.work_counter = &work_counter;
    P.failed_slots = a->failed_slots;
    P.n_failed = a->n_failed;
    P.failed_cap = a->n_work;
    P.flags = a->flags;
    P.gv_xyz = a->gv_xyz; P.gl_idx = a->gl_idx; P.gv_cap = a->gv_cap; P.gl_cap = a->gl_cap;
    P.g_cursor = reinterpret_cast<unsigned long long*>(a->g_cursor);
    P.nverts = a->nverts; P.nloops = a->nloops;
    P.vbase = reinterpret_cast<unsigned long long*>(a->vbase); P.lbase = reinterpret_cast<unsigned long long*>(a->lbase);
    P.st_flen = a->st_flen;
    g_os_threads = a->os_threads ? a->os_threads : 1;
    g_blocks = a->blocks ? a->blocks : g_os_threads;
    g_reverse = a->reverse != 0;
    g_collectives = 0;
    launch_clip(P, a->large, nullptr);  // 0 small, 1 medium, 2 large, 3 small without the serial walk, 4 thread per cell (tess::CLIP_*)
    if (a->counters)
        for (int i = 0; i < CNT_N; ++i) a->counters[i] = counters[i];
    a->collectives = g_collectives;
    return 0;
}

// profile of the last launches: how often each warp collective (by clip.cu source line) was executed
static unsigned long long g_hist[65536];
void emu_line_hist_enable(int on) {
    memset(g_hist, 0, sizeof(g_hist));
    emu::g_line_hist = on ? g_hist : nullptr;
}
const unsigned long long* emu_line_hist() { return g_hist; }

uint32_t emu_small_fmax() { return tess::clip_small_fmax(); }
uint32_t emu_large_fmax() { return tess::clip_large_fmax(); }
uint32_t emu_medium_fmax() { return tess::clip_medium_fmax(); }
}
