// emu_outputs.cpp — TEST INFRASTRUCTURE.  the-tessellator_b200/csrc/outputs.cu (CSR packing of the staged face
// lists, redo rows, geometry gathers, chunk work lists, the volume sum), as it is, compiled for the CPU warp
// emulator; linked with emu_grid.cpp for the scan kernel.  Used by tests/test_emu_grid.py only.
#include <cuda_runtime.h>  // resolves to tests/emu/shim/cuda_runtime.h

#include <vector>

#include "../../the-tessellator_b200/csrc/outputs.cu"

extern "C" {

struct emu_pack_args {
    uint64_t n_rows;
    const uint32_t* status;   // n_rows
    const uint32_t* nfaces;   // n_rows + 1 (last entry 0)
    const int64_t* st_nbr;    // n_rows x fstride
    const double* st_area;
    const uint16_t* st_flen;  // nullable
    uint32_t fstride;
    // rows recomputed by a larger configuration: staging indexed by work item
    const uint32_t* redo_rows;  // nullable
    uint64_t n_redo;
    const int64_t* redo_nbr;
    const double* redo_area;
    const uint16_t* redo_flen;
    uint32_t redo_stride;
    // outputs
    uint64_t* offsets;  // n_rows + 1
    int64_t* nbr;
    double* area;
    uint32_t* flen;     // nullable
    uint64_t face_cap;
    uint32_t os_threads, reverse;
};

int emu_pack_run(emu_pack_args* a) {
    using namespace tess;
    emu::g_os_threads = a->os_threads ? a->os_threads : 1;
    emu::g_reverse = a->reverse != 0;
    std::vector<unsigned char> scan_tmp(scan_tmp_bytes(a->n_rows + 1) + 16);
    launch_exclusive_scan_u32_to_u64(a->nfaces, a->offsets, a->n_rows + 1, scan_tmp.data(), scan_tmp.size(), nullptr);
    launch_compact_faces(a->status, a->offsets, a->st_nbr, a->st_area, a->st_flen, a->fstride, a->n_rows, a->nbr, a->area, a->flen, nullptr, a->face_cap);
    if (a->n_redo)
        launch_compact_redo(a->redo_rows, nullptr, 0, a->nfaces, a->offsets, a->redo_nbr, a->redo_area, a->redo_flen, a->redo_stride, a->n_redo, a->nbr, a->area,
                            a->flen, nullptr);
    return 0;
}

// work list of one chunk of rows: ascending slots whose row lies in [row_lo, row_hi)
uint64_t emu_chunk_list(const uint32_t* row_of_slot, uint64_t n, uint32_t row_lo, uint32_t row_hi, uint32_t* work_slots, uint32_t os_threads) {
    using namespace tess;
    emu::g_os_threads = os_threads ? os_threads : 1;
    emu::g_reverse = false;
    std::vector<uint32_t> flags(n + 1);
    std::vector<uint64_t> pos(n + 1);
    std::vector<unsigned char> scan_tmp(scan_tmp_bytes(n + 1) + 16);
    launch_chunk_flags(row_of_slot, 0, n, row_lo, row_hi, flags.data(), nullptr);
    launch_exclusive_scan_u32_to_u64(flags.data(), pos.data(), n + 1, scan_tmp.data(), scan_tmp.size(), nullptr);
    launch_chunk_scatter(flags.data(), pos.data(), 0, n, work_slots, nullptr);
    return pos[n];
}

void emu_gather(const uint32_t* nverts, const uint64_t* vbase, const uint64_t* voffsets, const double* vpool, const uint32_t* nloops, const uint64_t* lbase,
                const uint64_t* face_offsets, const uint64_t* fv_offsets, const uint32_t* lpool, uint64_t n_rows, double* vtx, uint32_t* loops, uint32_t os_threads) {
    using namespace tess;
    emu::g_os_threads = os_threads ? os_threads : 1;
    emu::g_reverse = false;
    launch_gather_vertices(nverts, reinterpret_cast<const unsigned long long*>(vbase), voffsets, vpool, n_rows, vtx, nullptr);
    launch_gather_loops(nloops, reinterpret_cast<const unsigned long long*>(lbase), face_offsets, fv_offsets, lpool, n_rows, loops, nullptr);
}

double emu_volume_sum(const double* vol, uint64_t n, uint32_t os_threads, uint32_t reverse) {
    using namespace tess;
    emu::g_os_threads = os_threads ? os_threads : 1;
    emu::g_reverse = reverse != 0;
    double out = 0.0;
    launch_volume_sum(vol, n, &out, nullptr);
    return out;
}

void emu_clear_status_bits(uint32_t* status, uint64_t n, uint32_t bits) {
    tess::launch_clear_status_bits(status, n, bits, nullptr);
}
}
