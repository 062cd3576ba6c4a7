// outputs.cu — K6: CSR compaction of the per-cell face lists, and K8: the volume closure sum.
//
// The clip kernel leaves each cell's neighbour ids / face areas in a fixed-stride staging row
// (the batched form of the Vec returns of interface.rs:342-384).  After an exclusive scan of the
// face counts these kernels pack the rows into CSR arrays with coalesced writes.
#include <algorithm>

#include "common.cuh"

namespace tess {

namespace {

constexpr int kRowsPerBlock = 256;

// One block packs kRowsPerBlock consecutive rows: every thread walks output positions and finds
// the owning row by binary search in the block's slice of the offsets.
__global__ void __launch_bounds__(256) compact_faces_kernel(const uint32_t* __restrict__ status, const uint64_t* __restrict__ offsets,
                                                            const int64_t* __restrict__ st_nbr, const double* __restrict__ st_area,
                                                            const uint16_t* __restrict__ st_flen, uint32_t fstride, size_t n_rows,
                                                            int64_t* __restrict__ nbr, double* __restrict__ area, uint32_t* __restrict__ flen, uint64_t face_cap) {
    __shared__ uint64_t s_off[kRowsPerBlock + 1];
    const size_t row0 = (size_t)blockIdx.x * kRowsPerBlock;
    const int rows = (int)min((size_t)kRowsPerBlock, n_rows - row0);
    for (int i = threadIdx.x; i <= rows; i += blockDim.x) s_off[i] = offsets[row0 + i];
    __syncthreads();
    const uint64_t begin = s_off[0], end = min(s_off[rows], face_cap);  // never past the arrays' capacity (the host reports the overflow)
    for (uint64_t p = begin + threadIdx.x; p < end; p += blockDim.x) {
        int lo = 0, hi = rows;  // largest r with s_off[r] <= p
        while (hi - lo > 1) {
            const int mid = (lo + hi) >> 1;
            if (s_off[mid] <= p) lo = mid; else hi = mid;
        }
        const size_t row = row0 + lo;
        if (status[row] & ST_LARGE_PATH) continue;  // packed by compact_redo_kernel
        const uint32_t k = (uint32_t)(p - s_off[lo]);
        nbr[p] = st_nbr[row * fstride + k];
        if (area) area[p] = st_area[row * fstride + k];
        if (flen) flen[p] = st_flen[row * fstride + k];
    }
}

// Rows recomputed by the large-cell pass: one warp per work item, staging indexed by work item.
__global__ void __launch_bounds__(128) compact_redo_kernel(const uint32_t* __restrict__ work_slots, const uint32_t* __restrict__ row_of_slot, uint32_t row_base,
                                                           const uint32_t* __restrict__ nfaces, const uint64_t* __restrict__ offsets,
                                                           const int64_t* __restrict__ st_nbr, const double* __restrict__ st_area,
                                                           const uint16_t* __restrict__ st_flen, uint32_t fstride, size_t n_work,
                                                           int64_t* __restrict__ nbr, double* __restrict__ area, uint32_t* __restrict__ flen) {
    const size_t w = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (w >= n_work) return;
    const uint32_t slot = work_slots[w];
    const size_t row = row_of_slot ? row_of_slot[slot] : (size_t)(slot - row_base);
    const uint32_t nf = min(nfaces[row], fstride);  // a row redone later by a larger configuration may exceed this stage's stride
    const uint64_t o = offsets[row];
    for (uint32_t k = lane; k < nf; k += 32) {
        nbr[o + k] = st_nbr[w * fstride + k];
        if (area) area[o + k] = st_area[w * fstride + k];
        if (flen) flen[o + k] = st_flen[w * fstride + k];
    }
}

// Geometry outputs: one warp copies one row's block out of the bump-allocated pools.
__global__ void __launch_bounds__(256) gather_vertices_kernel(const uint32_t* __restrict__ nverts, const unsigned long long* __restrict__ vbase,
                                                              const uint64_t* __restrict__ voffsets, const double* __restrict__ pool, size_t n_rows,
                                                              double* __restrict__ vtx) {
    const size_t w = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (w >= n_rows) return;
    const uint32_t nv = nverts[w];
    const unsigned long long src = vbase[w];
    const uint64_t dst = voffsets[w];
    for (uint32_t k = lane; k < 3 * nv; k += 32) vtx[3 * dst + k] = pool[3 * src + k];
}
__global__ void __launch_bounds__(256) gather_loops_kernel(const uint32_t* __restrict__ nloops, const unsigned long long* __restrict__ lbase,
                                                           const uint64_t* __restrict__ face_offsets, const uint64_t* __restrict__ fv_offsets,
                                                           const uint32_t* __restrict__ pool, size_t n_rows, uint32_t* __restrict__ out) {
    const size_t w = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (w >= n_rows) return;
    const uint32_t nl = nloops[w];
    const unsigned long long src = lbase[w];
    const uint64_t dst = fv_offsets[face_offsets[w]];
    for (uint32_t k = lane; k < nl; k += 32) out[dst + k] = pool[src + k];
}

// Deterministic two-level sum (fixed block partition, fixed tree) so that repeated runs agree bitwise.
__global__ void __launch_bounds__(256) volume_partial_kernel(const double* __restrict__ vol, size_t n, double* __restrict__ partial) {
    __shared__ double s[256];
    const size_t per_block = (n + gridDim.x - 1) / gridDim.x;
    const size_t b = (size_t)blockIdx.x * per_block;
    const size_t e = min(n, b + per_block);
    double acc = 0.0;
    for (size_t i = b + threadIdx.x; i < e; i += blockDim.x) acc += vol[i];
    s[threadIdx.x] = acc;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
        if ((int)threadIdx.x < o) s[threadIdx.x] += s[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) partial[blockIdx.x] = s[0];
}
// Streaming (tess_compute_all_to_host): the sorted slots whose output row lies in [row_lo, row_hi), in
// ascending slot order — the work list of one chunk of rows.  flags -> exclusive scan -> scatter.
__global__ void __launch_bounds__(256) chunk_flags_kernel(const uint32_t* __restrict__ row_of_slot, uint32_t slot_begin, size_t n, uint32_t row_lo, uint32_t row_hi,
                                                          uint32_t* __restrict__ flags) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i > n) return;
    uint32_t f = 0;
    if (i < n) {
        const uint32_t row = row_of_slot[slot_begin + i];
        f = (row >= row_lo && row < row_hi) ? 1u : 0u;
    }
    flags[i] = f;  // flags[n] = 0 closes the scan
}
__global__ void __launch_bounds__(256) chunk_scatter_kernel(const uint32_t* __restrict__ flags, const uint64_t* __restrict__ pos, uint32_t slot_begin, size_t n,
                                                            uint32_t* __restrict__ work_slots) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && flags[i]) work_slots[pos[i]] = slot_begin + (uint32_t)i;
}

__global__ void __launch_bounds__(256) clear_status_bits_kernel(uint32_t* __restrict__ status, size_t n, uint32_t bits) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) status[i] &= ~bits;
}

__global__ void volume_final_kernel(const double* __restrict__ partial, int nb, double* __restrict__ out) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        double acc = 0.0;
        for (int i = 0; i < nb; ++i) acc += partial[i];
        *out = acc;
    }
}

}  // namespace

void launch_compact_faces(const uint32_t* status, const uint64_t* offsets, const int64_t* st_nbr, const double* st_area, const uint16_t* st_flen, uint32_t fstride,
                          size_t n_rows, int64_t* nbr, double* area, uint32_t* flen, cudaStream_t s, uint64_t face_cap) {
    if (!n_rows) return;
    const unsigned int nb = (unsigned int)((n_rows + kRowsPerBlock - 1) / kRowsPerBlock);
    TESS_LAUNCH(compact_faces_kernel, nb, 256, 0, s, status, offsets, st_nbr, st_area, st_flen, fstride, n_rows, nbr, area, flen, face_cap);
    note_launch();
    TESS_CUDA_CHECK(cudaGetLastError());
}

void launch_clear_status_bits(uint32_t* status, size_t n, uint32_t bits, cudaStream_t s) {
    if (!n) return;
    TESS_LAUNCH(clear_status_bits_kernel, (unsigned int)((n + 255) / 256), 256, 0, s, status, n, bits);
    note_launch();
    TESS_CUDA_CHECK(cudaGetLastError());
}

void launch_chunk_flags(const uint32_t* row_of_slot, uint32_t slot_begin, size_t n, uint32_t row_lo, uint32_t row_hi, uint32_t* flags, cudaStream_t s) {
    const unsigned int nb = (unsigned int)((n + 1 + 255) / 256);
    TESS_LAUNCH(chunk_flags_kernel, nb, 256, 0, s, row_of_slot, slot_begin, n, row_lo, row_hi, flags);
    note_launch();
    TESS_CUDA_CHECK(cudaGetLastError());
}

void launch_chunk_scatter(const uint32_t* flags, const uint64_t* pos, uint32_t slot_begin, size_t n, uint32_t* work_slots, cudaStream_t s) {
    if (!n) return;
    const unsigned int nb = (unsigned int)((n + 255) / 256);
    TESS_LAUNCH(chunk_scatter_kernel, nb, 256, 0, s, flags, pos, slot_begin, n, work_slots);
    note_launch();
    TESS_CUDA_CHECK(cudaGetLastError());
}

void launch_compact_redo(const uint32_t* work_slots, const uint32_t* row_of_slot, uint32_t row_base, const uint32_t* nfaces, const uint64_t* offsets,
                         const int64_t* st_nbr, const double* st_area, const uint16_t* st_flen, uint32_t fstride, size_t n_work, int64_t* nbr, double* area,
                         uint32_t* flen, cudaStream_t s) {
    if (!n_work) return;
    const unsigned int nb = (unsigned int)((n_work * 32 + 127) / 128);
    TESS_LAUNCH(compact_redo_kernel, nb, 128, 0, s, work_slots, row_of_slot, row_base, nfaces, offsets, st_nbr, st_area, st_flen, fstride, n_work, nbr, area, flen);
    note_launch();
    TESS_CUDA_CHECK(cudaGetLastError());
}

void launch_gather_vertices(const uint32_t* nverts, const unsigned long long* vbase, const uint64_t* voffsets, const double* pool, size_t n_rows, double* vtx, cudaStream_t s) {
    if (!n_rows) return;
    const unsigned int nb = (unsigned int)((n_rows * 32 + 255) / 256);
    TESS_LAUNCH(gather_vertices_kernel, nb, 256, 0, s, nverts, vbase, voffsets, pool, n_rows, vtx);
    note_launch();
    TESS_CUDA_CHECK(cudaGetLastError());
}

void launch_gather_loops(const uint32_t* nloops, const unsigned long long* lbase, const uint64_t* face_offsets, const uint64_t* fv_offsets, const uint32_t* pool, size_t n_rows,
                         uint32_t* out, cudaStream_t s) {
    if (!n_rows) return;
    const unsigned int nb = (unsigned int)((n_rows * 32 + 255) / 256);
    TESS_LAUNCH(gather_loops_kernel, nb, 256, 0, s, nloops, lbase, face_offsets, fv_offsets, pool, n_rows, out);
    note_launch();
    TESS_CUDA_CHECK(cudaGetLastError());
}

void launch_volume_sum(const double* vol, size_t n, double* out, cudaStream_t s) {
    constexpr int nb = 296;
    double* partial = nullptr;
    TESS_CUDA_CHECK(cudaMallocAsync(&partial, sizeof(double) * nb, s));
    TESS_LAUNCH(volume_partial_kernel, nb, 256, 0, s, vol, n, partial);
    note_launch();
    TESS_LAUNCH(volume_final_kernel, 1, 32, 0, s, partial, nb, out);
    note_launch();
    TESS_CUDA_CHECK(cudaGetLastError());
    TESS_CUDA_CHECK(cudaFreeAsync(partial, s));
}

}  // namespace tess

// ---------------------------------------------------------------------------------------------
// telemetry (not part of the emulated build of tests/emu)
// ---------------------------------------------------------------------------------------------
#ifndef TESS_WARP_EMU
#include <atomic>
namespace tess {
namespace {
std::atomic<unsigned long long> g_launches{0};

// 8 independent DFMA chains per thread, everything in registers.
__global__ void __launch_bounds__(256) fp64_peak_kernel(double* out, int iters, double a, double b) {
    double x0 = threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6, x7 = x0 + 7;
    for (int i = 0; i < iters; ++i) {
        x0 = __fma_rn(x0, a, b); x1 = __fma_rn(x1, a, b); x2 = __fma_rn(x2, a, b); x3 = __fma_rn(x3, a, b);
        x4 = __fma_rn(x4, a, b); x5 = __fma_rn(x5, a, b); x6 = __fma_rn(x6, a, b); x7 = __fma_rn(x7, a, b);
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = ((x0 + x1) + (x2 + x3)) + ((x4 + x5) + (x6 + x7));
}
}  // namespace

void note_launch(int n) { g_launches.fetch_add((unsigned long long)n, std::memory_order_relaxed); }
unsigned long long launch_count() { return g_launches.load(std::memory_order_relaxed); }

double measure_fp64_peak_tflops() {
    int dev = 0, sms = 148;
    TESS_CUDA_CHECK(cudaGetDevice(&dev));
    TESS_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const int blocks = sms * 8, threads = 256, iters = 1 << 14;
    double* out = nullptr;
    TESS_CUDA_CHECK(cudaMalloc(&out, sizeof(double) * blocks * threads));
    cudaEvent_t e0, e1;
    TESS_CUDA_CHECK(cudaEventCreate(&e0));
    TESS_CUDA_CHECK(cudaEventCreate(&e1));
    double best = 0.0;
    for (int rep = 0; rep < 5; ++rep) {
        TESS_CUDA_CHECK(cudaEventRecord(e0));
        fp64_peak_kernel<<<blocks, threads>>>(out, iters, 0.999999, 1e-9);
        TESS_CUDA_CHECK(cudaEventRecord(e1));
        TESS_CUDA_CHECK(cudaEventSynchronize(e1));
        float ms = 0;
        TESS_CUDA_CHECK(cudaEventElapsedTime(&ms, e0, e1));
        const double flops = 2.0 * 8.0 * (double)iters * (double)blocks * threads;
        if (rep > 0) best = std::max(best, flops / (ms * 1e-3) / 1e12);
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(out);
    return best;
}
}  // namespace tess
#endif  // TESS_WARP_EMU
