// emu_grid.cpp — TEST INFRASTRUCTURE.  the-tessellator_b200/csrc/grid.cu (the binning pass K1-K4 and the slab
// helpers), as it is, compiled for the CPU warp emulator; one C entry point runs the pass the way
// tess_diagram_initialize does.  Used by tests/test_emu_grid.py only.
#include <cuda_runtime.h>  // resolves to tests/emu/shim/cuda_runtime.h

#include <vector>

#include "../../the-tessellator_b200/csrc/grid.cu"

namespace tess {
void note_launch(int) {}
unsigned long long launch_count() { return 0; }
}  // namespace tess

extern "C" {

struct emu_grid_args {
    const double* xyz;   // n x 3
    uint32_t n;
    const int64_t* ids;       // nullable (explicit ids: slab diagrams)
    const uint64_t* groups;   // nullable
    double bounds[6];         // x_min, x_max, y_min, y_max, z_min, z_max (grid parameters, as initialize derives them)
    double cell_info[6];      // sizes xyz, inverse sizes xyz
    uint32_t cpd, local_lo, local_hi;
    // outputs
    double* bounds_out;       // 6: K1
    uint32_t* cell_of;        // n
    uint32_t* delim;          // local cells + 1
    double* sorted;           // n x 4 {x, y, z, id bits}
    uint32_t* sorted_idx;     // n
    uint64_t* groups_sorted;  // n, nullable
    uint64_t* plane_counts;   // cpd
    uint32_t oob;             // out: some particle outside the local planes
    uint32_t os_threads, reverse;
};

int emu_grid_run(emu_grid_args* a) {
    using namespace tess;
    emu::g_os_threads = a->os_threads ? a->os_threads : 1;
    emu::g_reverse = a->reverse != 0;
    const size_t n = a->n;
    GridSpec g{};
    g.xmin = a->bounds[0]; g.xmax = a->bounds[1]; g.ymin = a->bounds[2]; g.ymax = a->bounds[3]; g.zmin = a->bounds[4]; g.zmax = a->bounds[5];
    g.sx = a->cell_info[0]; g.sy = a->cell_info[1]; g.sz = a->cell_info[2];
    g.ix = a->cell_info[3]; g.iy = a->cell_info[4]; g.iz = a->cell_info[5];
    g.cpd = a->cpd;
    g.local_lo = a->local_lo; g.local_hi = a->local_hi; g.own_lo = a->local_lo; g.own_hi = a->local_hi;
    const size_t ncl = (size_t)(g.local_hi - g.local_lo) * g.cpd * g.cpd;

    launch_bounds(a->xyz, n, a->bounds_out, nullptr);                                            // K1
    std::vector<uint32_t> rank(n + 2), counts(ncl + 1, 0u), tmp_idx(n);
    uint32_t oob = 0;
    launch_cell_histogram(a->xyz, n, g, a->cell_of, rank.data(), counts.data(), &oob, nullptr);  // K2
    a->oob = oob;
    std::vector<unsigned char> scan_tmp(scan_tmp_bytes(ncl + 1) + 16);
    launch_exclusive_scan_u32(counts.data(), a->delim, ncl + 1, scan_tmp.data(), scan_tmp.size(), nullptr);  // K3
    std::vector<unsigned char> raw((n + 2) * sizeof(Particle) * 2 + 128);
    Particle* arrived = reinterpret_cast<Particle*>((reinterpret_cast<uintptr_t>(raw.data()) + 31) & ~uintptr_t(31));
    Particle* sorted = arrived + n + 1;
    if (!oob) {
        launch_scatter_records(a->xyz, a->ids, a->cell_of, rank.data(), a->delim, arrived, a->ids ? tmp_idx.data() : nullptr, n, nullptr);  // K4
        launch_rank_fix(arrived, a->ids ? tmp_idx.data() : nullptr, g, a->delim, a->groups, sorted, a->sorted_idx, a->groups_sorted, n, nullptr);
        memcpy(a->sorted, sorted, n * sizeof(Particle));
    }
    if (a->plane_counts) {
        std::vector<unsigned long long> pc(g.cpd, 0ull);
        launch_plane_histogram(a->xyz, n, g, pc.data(), nullptr);
        for (uint32_t i = 0; i < g.cpd; ++i) a->plane_counts[i] = pc[i];
    }
    return 0;
}
}
