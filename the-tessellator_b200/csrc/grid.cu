// grid.cu — the binning pass: celery.rs's spatial grid as a GPU counting sort.
//
//   K1 bounds_reduce        CeleryBounds::new            celery.rs:81-125
//   K2 cell_histogram       get_cells / get_cell         celery.rs:269-354  (+ per-cell counts)
//   K3 exclusive_scan       get_delimiters               celery.rs:372-414  (CSR offsets)
//   K4 scatter_records + rank_fix   get_sorted_indices   celery.rs:357-369  (counting sort instead
//                           of sort_unstable_by; order inside a grid cell = ascending particle
//                           index, the canonical choice where the reference leaves it unspecified)
//
// All of it is HBM-bound streaming work: positions are read with 16-byte vector loads, the
// histogram uses warp-aggregated atomics, the scan is single-pass (decoupled look-back).
#include <algorithm>
#include <cfloat>

#include "common.cuh"
#include "tess_math.cuh"

namespace tess {

namespace {

constexpr int kThreads = 256;

__device__ __forceinline__ double warp_min(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const double w = __shfl_xor_sync(0xffffffffu, v, o);
        v = w < v ? w : v;
    }
    return v;
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const double w = __shfl_xor_sync(0xffffffffu, v, o);
        v = w > v ? w : v;
    }
    return v;
}

// Read particles 2 at a time: 48 bytes = three aligned double2 loads (x0 y0 | z0 x1 | y1 z1).
struct Pair {
    double x0, y0, z0, x1, y1, z1;
    bool v0, v1;
};
__device__ __forceinline__ Pair load_pair(const double* __restrict__ xyz, size_t pair, size_t n, bool aligned) {
    Pair p;
    const size_t i0 = 2 * pair;
    p.v0 = i0 < n;
    p.v1 = i0 + 1 < n;
    p.x0 = p.y0 = p.z0 = p.x1 = p.y1 = p.z1 = 0.0;
    if (p.v1 && aligned) {
        const double2* q = reinterpret_cast<const double2*>(xyz + 3 * i0);
        const double2 a = __ldg(q), b = __ldg(q + 1), c = __ldg(q + 2);
        p.x0 = a.x; p.y0 = a.y; p.z0 = b.x; p.x1 = b.y; p.y1 = c.x; p.z1 = c.y;
    } else {
        if (p.v0) { p.x0 = xyz[3 * i0]; p.y0 = xyz[3 * i0 + 1]; p.z0 = xyz[3 * i0 + 2]; }
        if (p.v1) { p.x1 = xyz[3 * i0 + 3]; p.y1 = xyz[3 * i0 + 4]; p.z1 = xyz[3 * i0 + 5]; }
    }
    return p;
}

// ---------------------------------------------------------------- K1 -----------------------
__global__ void __launch_bounds__(kThreads) bounds_partial_kernel(const double* __restrict__ xyz, size_t n, double* __restrict__ partial, bool aligned) {
    double mn[3] = {DBL_MAX, DBL_MAX, DBL_MAX}, mx[3] = {-DBL_MAX, -DBL_MAX, -DBL_MAX};
    const size_t npairs = (n + 1) / 2;
    for (size_t pr = blockIdx.x * (size_t)blockDim.x + threadIdx.x; pr < npairs; pr += (size_t)gridDim.x * blockDim.x) {
        const Pair p = load_pair(xyz, pr, n, aligned);
        if (p.v0) {
            mn[0] = p.x0 < mn[0] ? p.x0 : mn[0]; mx[0] = p.x0 > mx[0] ? p.x0 : mx[0];
            mn[1] = p.y0 < mn[1] ? p.y0 : mn[1]; mx[1] = p.y0 > mx[1] ? p.y0 : mx[1];
            mn[2] = p.z0 < mn[2] ? p.z0 : mn[2]; mx[2] = p.z0 > mx[2] ? p.z0 : mx[2];
        }
        if (p.v1) {
            mn[0] = p.x1 < mn[0] ? p.x1 : mn[0]; mx[0] = p.x1 > mx[0] ? p.x1 : mx[0];
            mn[1] = p.y1 < mn[1] ? p.y1 : mn[1]; mx[1] = p.y1 > mx[1] ? p.y1 : mx[1];
            mn[2] = p.z1 < mn[2] ? p.z1 : mn[2]; mx[2] = p.z1 > mx[2] ? p.z1 : mx[2];
        }
    }
    __shared__ double sm[6][kThreads / 32];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const double a = warp_min(mn[k]), b = warp_max(mx[k]);
        if (lane == 0) { sm[2 * k][w] = a; sm[2 * k + 1][w] = b; }
    }
    __syncthreads();
    if (threadIdx.x < 6) {
        double v = sm[threadIdx.x][0];
        for (int i = 1; i < kThreads / 32; ++i) {
            const double u = sm[threadIdx.x][i];
            v = (threadIdx.x & 1) ? (u > v ? u : v) : (u < v ? u : v);
        }
        partial[blockIdx.x * 6 + threadIdx.x] = v;  // order: xmin,xmax,ymin,ymax,zmin,zmax
    }
}

__global__ void bounds_final_kernel(const double* __restrict__ partial, int nblocks, double* __restrict__ out6) {
    const int k = threadIdx.x;  // 6 threads
    if (k >= 6) return;
    double v = partial[k];
    for (int b = 1; b < nblocks; ++b) {
        const double u = partial[b * 6 + k];
        v = (k & 1) ? (u > v ? u : v) : (u < v ? u : v);
    }
    out6[k] = v;
}

// ---------------------------------------------------------------- K2 -----------------------
__device__ __forceinline__ uint32_t local_cell(double x, double y, double z, const GridSpec& g, bool& oob) {
    const uint32_t gx = axis_index(x, g.xmin, g.xmax, g.ix, g.cpd);
    const uint32_t gy = axis_index(y, g.ymin, g.ymax, g.iy, g.cpd);
    const uint32_t gz = axis_index(z, g.zmin, g.zmax, g.iz, g.cpd);
    oob = gx < g.local_lo || gx >= g.local_hi;
    // Celery::get_cell_from_indices (celery.rs:317-325), x-planes re-based to the local slab
    return ((gx - g.local_lo) * g.cpd + gy) * g.cpd + gz;
}

// Warp-aggregated atomicAdd: lanes that hit the same counter elect a leader that adds the
// group's size once; every lane gets a distinct arrival rank.
__device__ __forceinline__ uint32_t aggregated_inc(uint32_t* counts, uint32_t cell, bool valid) {
    const uint32_t key = valid ? cell : 0xFFFFFFFFu;
    const uint32_t peers = __match_any_sync(0xffffffffu, key);
    const int leader = __ffs(peers) - 1;
    const int lane = threadIdx.x & 31;
    uint32_t base = 0;
    if (valid && lane == leader) base = atomicAdd(&counts[cell], (uint32_t)__popc(peers));
    base = __shfl_sync(0xffffffffu, base, leader);
    return base + __popc(peers & ((1u << lane) - 1u));
}

__global__ void __launch_bounds__(kThreads) cell_histogram_kernel(const double* __restrict__ xyz, size_t n, GridSpec g, uint32_t* __restrict__ cell_of,
                                                                  uint32_t* __restrict__ rank_in_cell, uint32_t* __restrict__ counts,
                                                                  uint32_t* __restrict__ oob_flag, bool aligned) {
    const size_t pr = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    const Pair p = load_pair(xyz, pr, n, aligned);
    bool o0 = false, o1 = false;
    const uint32_t c0 = p.v0 ? local_cell(p.x0, p.y0, p.z0, g, o0) : 0u;
    const uint32_t c1 = p.v1 ? local_cell(p.x1, p.y1, p.z1, g, o1) : 0u;
    if ((p.v0 && o0) || (p.v1 && o1)) atomicOr(oob_flag, 1u);
    const uint32_t r0 = aggregated_inc(counts, c0, p.v0 && !o0);
    const uint32_t r1 = aggregated_inc(counts, c1, p.v1 && !o1);
    if (p.v1) {
        *reinterpret_cast<uint2*>(cell_of + 2 * pr) = make_uint2(c0, c1);
        *reinterpret_cast<uint2*>(rank_in_cell + 2 * pr) = make_uint2(r0, r1);
    } else if (p.v0) {
        cell_of[2 * pr] = c0;
        rank_in_cell[2 * pr] = r0;
    }
}

// ---------------------------------------------------------------- K3 -----------------------
// Single-pass exclusive scan with decoupled look-back.  tile_state[t] = flag(2 bits) | value.
constexpr int kScanItems = 16;
constexpr int kScanTile = kThreads * kScanItems;
constexpr unsigned long long kFlagAgg = 1ull << 62, kFlagPrefix = 2ull << 62, kValMask = (1ull << 62) - 1;

template <typename OutT>
__global__ void __launch_bounds__(kThreads) scan_kernel(const uint32_t* __restrict__ in, OutT* __restrict__ out, size_t n, unsigned int* __restrict__ tile_counter,
                                                        unsigned long long* __restrict__ tile_state) {
    __shared__ unsigned int s_tile;
    __shared__ unsigned long long s_warp[kThreads / 32];
    __shared__ unsigned long long s_excl;
    if (threadIdx.x == 0) s_tile = atomicAdd(tile_counter, 1u);
    __syncthreads();
    const unsigned int tile = s_tile;
    const size_t base = (size_t)tile * kScanTile + (size_t)threadIdx.x * kScanItems;

    uint32_t v[kScanItems];
    unsigned long long tsum = 0;
    if (base + kScanItems <= n) {
        const uint4* q = reinterpret_cast<const uint4*>(in + base);
#pragma unroll
        for (int k = 0; k < kScanItems / 4; ++k) {
            const uint4 u = __ldg(q + k);
            v[4 * k] = u.x; v[4 * k + 1] = u.y; v[4 * k + 2] = u.z; v[4 * k + 3] = u.w;
        }
    } else {
#pragma unroll
        for (int k = 0; k < kScanItems; ++k) v[k] = (base + k < n) ? in[base + k] : 0u;
    }
#pragma unroll
    for (int k = 0; k < kScanItems; ++k) tsum += v[k];

    // block-wide exclusive scan of the per-thread sums
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    unsigned long long incl = tsum;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const unsigned long long t = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += t;
    }
    if (lane == 31) s_warp[w] = incl;
    __syncthreads();
    unsigned long long warp_off = 0, block_total = 0;
#pragma unroll
    for (int i = 0; i < kThreads / 32; ++i) {
        const unsigned long long t = s_warp[i];
        if (i < w) warp_off += t;
        block_total += t;
    }

    if (threadIdx.x == 0) {
        unsigned long long excl = 0;
        if (tile == 0) {
            atomicExch(&tile_state[0], kFlagPrefix | block_total);
        } else {
            atomicExch(&tile_state[tile], kFlagAgg | block_total);
            for (int j = (int)tile - 1; j >= 0; --j) {
                unsigned long long s;
                do { s = atomicAdd(&tile_state[j], 0ull); } while ((s >> 62) == 0ull);
                excl += s & kValMask;
                if ((s >> 62) == 2ull) break;
            }
            atomicExch(&tile_state[tile], kFlagPrefix | (excl + block_total));
        }
        s_excl = excl;
    }
    __syncthreads();
    unsigned long long run = s_excl + warp_off + (incl - tsum);
    if (base + kScanItems <= n) {
#pragma unroll
        for (int k = 0; k < kScanItems; ++k) {
            out[base + k] = (OutT)run;
            run += v[k];
        }
    } else {
#pragma unroll
        for (int k = 0; k < kScanItems; ++k) {
            if (base + k < n) out[base + k] = (OutT)run;
            run += v[k];
        }
    }
}

// ---------------------------------------------------------------- K4 -----------------------
// Counting-sort scatter.  One thread per particle, in input order (coalesced reads): the whole 32-byte
// record {x, y, z, id} goes to the particle's ARRIVAL slot of its grid cell — one full DRAM sector per
// particle, where scattering a 4-byte index would dirty the same sector and leave a 24-byte gather of
// the position for later.
__global__ void __launch_bounds__(kThreads) scatter_records_kernel(const double* __restrict__ xyz, const int64_t* __restrict__ ids,
                                                                   const uint32_t* __restrict__ cell_of, const uint32_t* __restrict__ rank_in_cell,
                                                                   const uint32_t* __restrict__ delim, Particle* __restrict__ arrived,
                                                                   uint32_t* __restrict__ arrived_idx, size_t n) {
    const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t dst = __ldg(delim + cell_of[i]) + rank_in_cell[i];
    const double x = xyz[3 * i], y = xyz[3 * i + 1], z = xyz[3 * i + 2];
    const int64_t id = ids ? ids[i] : (int64_t)i;
    double2* o = reinterpret_cast<double2*>(arrived + dst);
    o[0] = make_double2(x, y);
    o[1] = make_double2(z, __longlong_as_double(id));
    if (ids) arrived_idx[dst] = (uint32_t)i;  // insertion index != id only when ids are explicit (slab diagrams)
}

// Arrival order inside a grid cell depends on atomic timing.  One thread per arrival slot re-ranks its
// record among the few members of its cell by user-visible id — the canonical in-cell order (== insertion
// order when ids are implicit), which also makes a slab-sharded run order candidates exactly like the
// single-GPU run — and writes it to its final slot.  The cell is recomputed from the position (same
// function, same operands as K2), so nothing is gathered: reads and writes are sequential in slot order.
__global__ void __launch_bounds__(kThreads) rank_fix_kernel(const Particle* __restrict__ arrived, const uint32_t* __restrict__ arrived_idx, GridSpec g,
                                                            const uint32_t* __restrict__ delim, const uint64_t* __restrict__ groups,
                                                            Particle* __restrict__ sorted, uint32_t* __restrict__ sorted_idx,
                                                            uint64_t* __restrict__ groups_sorted, size_t n) {
    const size_t s = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (s >= n) return;
    const double2* q = reinterpret_cast<const double2*>(arrived + s);
    const double2 a = q[0], b = q[1];
    const int64_t my = __double_as_longlong(b.y);
    bool oob;
    const uint32_t c = local_cell(a.x, a.y, b.x, g, oob);
    const uint32_t d0 = __ldg(delim + c), d1 = __ldg(delim + c + 1);
    uint32_t r = 0;
    // ids are unique unless the caller supplied duplicates (add_particles_device(ids_dev)): ties break on the arrival slot, so
    // that two records never land on the same sorted slot
    for (uint32_t t = d0; t < d1; ++t) {
        const int64_t id = arrived[t].id;
        r += (id < my || (id == my && t < (uint32_t)s)) ? 1u : 0u;
    }
    const uint32_t dst = d0 + r;
    double2* o = reinterpret_cast<double2*>(sorted + dst);
    o[0] = a;
    o[1] = b;
    const uint32_t i = arrived_idx ? arrived_idx[s] : (uint32_t)my;
    sorted_idx[dst] = i;
    if (groups_sorted) groups_sorted[dst] = groups ? groups[i] : 0ull;
}

// ---------------------------------------------------------------- slab helpers -------------
__global__ void __launch_bounds__(kThreads) row_histogram_kernel(const double* __restrict__ xyz, size_t n, GridSpec g, unsigned long long* __restrict__ counts) {
    const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    const bool valid = i < n;
    uint32_t key = 0xFFFFFFFFu;
    if (valid) key = axis_index(xyz[3 * i], g.xmin, g.xmax, g.ix, g.cpd) * g.cpd + axis_index(xyz[3 * i + 1], g.ymin, g.ymax, g.iy, g.cpd);
    const uint32_t peers = __match_any_sync(0xffffffffu, key);
    const int lane = threadIdx.x & 31;
    if (valid && lane == __ffs(peers) - 1) atomicAdd(&counts[key], (unsigned long long)__popc(peers));
}

__global__ void __launch_bounds__(kThreads) plane_histogram_kernel(const double* __restrict__ xyz, size_t n, GridSpec g, unsigned long long* __restrict__ counts) {
    const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    const bool valid = i < n;
    const uint32_t gx = valid ? axis_index(xyz[3 * i], g.xmin, g.xmax, g.ix, g.cpd) : 0xFFFFFFFFu;
    const uint32_t peers = __match_any_sync(0xffffffffu, gx);
    const int lane = threadIdx.x & 31;
    if (valid && lane == __ffs(peers) - 1) atomicAdd(&counts[gx], (unsigned long long)__popc(peers));
}

// A particle of plane gx goes to every rank g with lo[g] <= gx < hi[g].
__global__ void __launch_bounds__(kThreads) pack_count_kernel(const double* __restrict__ xyz, size_t n, GridSpec g, int n_ranks, const uint32_t* __restrict__ lo,
                                                              const uint32_t* __restrict__ hi, unsigned long long* __restrict__ counts) {
    const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    const bool valid = i < n;
    const uint32_t gx = valid ? axis_index(xyz[3 * i], g.xmin, g.xmax, g.ix, g.cpd) : 0u;
    const int lane = threadIdx.x & 31;
    for (int r = 0; r < n_ranks; ++r) {
        const bool hit = valid && gx >= lo[r] && gx < hi[r];
        const uint32_t m = __ballot_sync(0xffffffffu, hit);
        if (m && lane == __ffs(m) - 1) atomicAdd(&counts[r], (unsigned long long)__popc(m));
    }
}

__global__ void __launch_bounds__(kThreads) pack_scatter_kernel(const double* __restrict__ xyz, const int64_t* __restrict__ ids, int64_t id_base, size_t n, GridSpec g,
                                                                int n_ranks, const uint32_t* __restrict__ lo, const uint32_t* __restrict__ hi,
                                                                const unsigned long long* __restrict__ offsets, const unsigned long long* __restrict__ limits,
                                                                unsigned long long* __restrict__ cursors, double* __restrict__ out_xyz, int64_t* __restrict__ out_ids, double* __restrict__ out_rec) {
    const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    const bool valid = i < n;
    double x = 0, y = 0, z = 0;
    if (valid) { x = xyz[3 * i]; y = xyz[3 * i + 1]; z = xyz[3 * i + 2]; }
    const uint32_t gx = valid ? axis_index(x, g.xmin, g.xmax, g.ix, g.cpd) : 0u;
    const int lane = threadIdx.x & 31;
    for (int r = 0; r < n_ranks; ++r) {
        const bool hit = valid && gx >= lo[r] && gx < hi[r];
        const uint32_t m = __ballot_sync(0xffffffffu, hit);
        if (!m) continue;
        const int leader = __ffs(m) - 1;
        unsigned long long base = 0;
        if (lane == leader) base = atomicAdd(&cursors[r], (unsigned long long)__popc(m));
        base = __shfl_sync(0xffffffffu, base, leader);
        const unsigned long long k = base + __popc(m & ((1u << lane) - 1u));
        if (hit && (!limits || k < limits[r])) {  // (a stale plan's counts may be too small: the excess is dropped, the cursor tells)
            const unsigned long long dst = offsets[r] + k;
            const int64_t id = ids ? ids[i] : id_base + (int64_t)i;
            if (out_rec) {  // one 32-byte record {x, y, z, id}: a single all-to-all carries it
                double4* q = reinterpret_cast<double4*>(out_rec) + dst;
                *q = make_double4(x, y, z, __longlong_as_double(id));
            } else {
                out_xyz[3 * dst] = x; out_xyz[3 * dst + 1] = y; out_xyz[3 * dst + 2] = z;
                out_ids[dst] = id;
            }
        }
    }
}

inline unsigned int blocks_for(size_t n, int per_block) { return (unsigned int)((n + per_block - 1) / per_block); }
inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

}  // namespace

// ---------------------------------------------------------------- launchers ----------------
constexpr int kBoundsBlocks = 148 * 4;

void launch_bounds(const double* xyz, size_t n, double* bounds6, cudaStream_t s) {
    double* partial = nullptr;
    TESS_CUDA_CHECK(cudaMallocAsync(&partial, sizeof(double) * 6 * kBoundsBlocks, s));
    const size_t npairs = (n + 1) / 2;
    const int nb = (int)std::min<size_t>(kBoundsBlocks, std::max<size_t>(1, (npairs + kThreads - 1) / kThreads));
    TESS_LAUNCH(bounds_partial_kernel, nb, kThreads, 0, s, xyz, n, partial, aligned16(xyz));
    note_launch();
    TESS_LAUNCH(bounds_final_kernel, 1, 32, 0, s, partial, nb, bounds6);
    note_launch();
    TESS_CUDA_CHECK(cudaGetLastError());
    TESS_CUDA_CHECK(cudaFreeAsync(partial, s));
}

void launch_cell_histogram(const double* xyz, size_t n, const GridSpec& g, uint32_t* cell_of, uint32_t* rank_in_cell, uint32_t* counts, uint32_t* oob_flag, cudaStream_t s) {
    const size_t npairs = (n + 1) / 2;
    if (!npairs) return;
    TESS_LAUNCH(cell_histogram_kernel, blocks_for(npairs, kThreads), kThreads, 0, s, xyz, n, g, cell_of, rank_in_cell, counts, oob_flag, aligned16(xyz));
    note_launch();
    TESS_CUDA_CHECK(cudaGetLastError());
}

size_t scan_tmp_bytes(size_t n) {
    const size_t tiles = (n + kScanTile - 1) / kScanTile;
    return 16 + sizeof(unsigned long long) * (tiles + 1);
}

template <typename OutT>
static void launch_scan_impl(const uint32_t* in, OutT* out, size_t n, void* tmp, size_t tmp_bytes, cudaStream_t s) {
    if (!n) return;
    const size_t need = scan_tmp_bytes(n);
    if (tmp_bytes < need) throw std::runtime_error("scan: temporary buffer too small");
    TESS_CUDA_CHECK(cudaMemsetAsync(tmp, 0, need, s));
    unsigned int* counter = reinterpret_cast<unsigned int*>(tmp);
    unsigned long long* state = reinterpret_cast<unsigned long long*>(reinterpret_cast<char*>(tmp) + 16);
    const size_t tiles = (n + kScanTile - 1) / kScanTile;
    TESS_LAUNCH(scan_kernel<OutT>, (unsigned int)tiles, kThreads, 0, s, in, out, n, counter, state);
    note_launch();
    TESS_CUDA_CHECK(cudaGetLastError());
}
void launch_exclusive_scan_u32(const uint32_t* in, uint32_t* out, size_t n, void* tmp, size_t tmp_bytes, cudaStream_t s) { launch_scan_impl<uint32_t>(in, out, n, tmp, tmp_bytes, s); }
void launch_exclusive_scan_u32_to_u64(const uint32_t* in, uint64_t* out, size_t n, void* tmp, size_t tmp_bytes, cudaStream_t s) {
    launch_scan_impl<unsigned long long>(in, reinterpret_cast<unsigned long long*>(out), n, tmp, tmp_bytes, s);
}

void launch_scatter_records(const double* xyz, const int64_t* ids, const uint32_t* cell_of, const uint32_t* rank_in_cell, const uint32_t* delim, Particle* arrived,
                            uint32_t* arrived_idx, size_t n, cudaStream_t s) {
    if (!n) return;
    TESS_LAUNCH(scatter_records_kernel, blocks_for(n, kThreads), kThreads, 0, s, xyz, ids, cell_of, rank_in_cell, delim, arrived, arrived_idx, n);
    note_launch();
    TESS_CUDA_CHECK(cudaGetLastError());
}

void launch_rank_fix(const Particle* arrived, const uint32_t* arrived_idx, const GridSpec& g, const uint32_t* delim, const uint64_t* groups, Particle* sorted,
                     uint32_t* sorted_idx, uint64_t* groups_sorted, size_t n, cudaStream_t s) {
    if (!n) return;
    TESS_LAUNCH(rank_fix_kernel, blocks_for(n, kThreads), kThreads, 0, s, arrived, arrived_idx, g, delim, groups, sorted, sorted_idx, groups_sorted, n);
    note_launch();
    TESS_CUDA_CHECK(cudaGetLastError());
}

void launch_row_histogram(const double* xyz, size_t n, const GridSpec& g, unsigned long long* counts, cudaStream_t s) {
    if (!n) return;
    TESS_LAUNCH(row_histogram_kernel, blocks_for(n, kThreads), kThreads, 0, s, xyz, n, g, counts);
    note_launch();
    TESS_CUDA_CHECK(cudaGetLastError());
}

void launch_plane_histogram(const double* xyz, size_t n, const GridSpec& g, unsigned long long* counts, cudaStream_t s) {
    if (!n) return;
    TESS_LAUNCH(plane_histogram_kernel, blocks_for(n, kThreads), kThreads, 0, s, xyz, n, g, counts);
    note_launch();
    TESS_CUDA_CHECK(cudaGetLastError());
}

void launch_pack_count(const double* xyz, size_t n, const GridSpec& g, int n_ranks, const uint32_t* lo_dev, const uint32_t* hi_dev, unsigned long long* counts, cudaStream_t s) {
    if (!n) return;
    TESS_LAUNCH(pack_count_kernel, blocks_for(n, kThreads), kThreads, 0, s, xyz, n, g, n_ranks, lo_dev, hi_dev, counts);
    note_launch();
    TESS_CUDA_CHECK(cudaGetLastError());
}

void launch_pack_scatter(const double* xyz, const int64_t* ids, int64_t id_base, size_t n, const GridSpec& g, int n_ranks, const uint32_t* lo_dev, const uint32_t* hi_dev,
                         const unsigned long long* offsets, const unsigned long long* limits, unsigned long long* cursors, double* out_xyz, int64_t* out_ids, double* out_rec,
                         cudaStream_t s) {
    if (!n) return;
    TESS_LAUNCH(pack_scatter_kernel, blocks_for(n, kThreads), kThreads, 0, s, xyz, ids, id_base, n, g, n_ranks, lo_dev, hi_dev, offsets, limits, cursors, out_xyz, out_ids, out_rec);
    note_launch();
    TESS_CUDA_CHECK(cudaGetLastError());
}

}  // namespace tess
