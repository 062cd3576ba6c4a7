// emu_query.cpp — TEST INFRASTRUCTURE.  the-tessellator_b200/csrc/query.cu (radius / neighbour-cloud queries), as
// it is, compiled for the CPU warp emulator.  Used by tests/test_emu_grid.py only.
#include <cuda_runtime.h>  // resolves to tests/emu/shim/cuda_runtime.h

#include <vector>

#include "../../the-tessellator_b200/csrc/query.cu"

namespace tess {
void note_launch(int) {}
unsigned long long launch_count() { return 0; }
}  // namespace tess

extern "C" {

struct emu_query_args {
    const double* particles;  // n x {x, y, z, id bits} in grid order
    uint32_t n;
    const uint32_t* delim;
    const uint64_t* groups_sorted;  // nullable
    const double* table_key;        // mode 2
    const int32_t* table_ijk;
    uint32_t table_len, table_full;
    double bounds[6], cell_info[6];
    uint32_t cpd;
    const double* xyz;  // n_query x 3
    uint32_t n_query;
    double radius;
    int32_t mode;
    int64_t target_group;
    // outputs
    uint64_t* offsets;  // n_query + 1
    int64_t* indices;   // capacity cap
    uint64_t cap;
    uint32_t* flags;    // n_query
    uint32_t os_threads, reverse;
    // mode 4 (ExpandingSearch::expand): cursors in / out (n_query each), cells_to_add
    const uint64_t* cursor_in;
    uint64_t* cursor_out;
    uint64_t cells_to_add;
};

int emu_query_run(emu_query_args* a) {
    using namespace tess;
    emu::g_os_threads = a->os_threads ? a->os_threads : 1;
    emu::g_reverse = a->reverse != 0;
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
    QueryParams Q{};
    Q.sorted = parts;
    Q.delim = a->delim;
    Q.groups_sorted = a->groups_sorted;
    Q.table = table.data();
    Q.table_len = a->table_len;
    Q.table_full = a->table_full;
    GridSpec& g = Q.grid;
    g.xmin = a->bounds[0]; g.xmax = a->bounds[1]; g.ymin = a->bounds[2]; g.ymax = a->bounds[3]; g.zmin = a->bounds[4]; g.zmax = a->bounds[5];
    g.sx = a->cell_info[0]; g.sy = a->cell_info[1]; g.sz = a->cell_info[2];
    g.ix = a->cell_info[3]; g.iy = a->cell_info[4]; g.iz = a->cell_info[5];
    g.cpd = a->cpd;
    g.local_lo = 0; g.local_hi = a->cpd; g.own_lo = 0; g.own_hi = a->cpd;
    Q.xyz = a->xyz;
    Q.n_query = a->n_query;
    Q.radius = a->radius;
    Q.mode = a->mode;
    Q.target_group = a->target_group;
    Q.cursor_in = a->cursor_in;
    Q.cursor_out = a->cursor_out;
    Q.cells_to_add = a->cells_to_add;
    std::vector<uint32_t> counts(a->n_query + 1, 0u);
    Q.counts = counts.data();
    Q.flags = a->flags;
    launch_radius_query(Q, /*fill=*/false, nullptr);
    uint64_t run = 0;
    for (uint32_t q = 0; q < a->n_query; ++q) {
        a->offsets[q] = run;
        run += counts[q];
    }
    a->offsets[a->n_query] = run;
    if (run > a->cap) return -4;
    Q.offsets = a->offsets;
    Q.indices = a->indices;
    launch_radius_query(Q, /*fill=*/true, nullptr);
    return 0;
}
}
