// query.cu — radius / neighbour-cloud queries on the GPU grid.
//
//   mode 0  Celery::find_neighbors_in_cell_radius   celery.rs:802-819
//   mode 1  Celery::find_neighbors_in_real_radius   celery.rs:825-855
//   mode 2  ExpandingSearch::expand_all_in_radius   celery.rs:1023-1075  (what Cell::compute_neighbor_cloud,
//           interface.rs:348-365, calls; a target group filters the result, :359-362)
//   mode 3  Celery::find_cells_in_radius            celery.rs:753-797    (grid cell ids instead of particles)
//   mode 4  ExpandingSearch::expand                 celery.rs:907-963    (incremental: a cursor per query position)
// all built on find_cells_in_radius (celery.rs:753-797) / check_cell_in_range (:708-743) and the
// search-order table.  One warp answers one query; the kernel runs twice (count, then fill) around an
// exclusive scan, and results come out in the reference's order: cells in i, j, k loop order (modes 0/1)
// or in search-table order (mode 2), particles of one cell in their sorted order.
#include "common.cuh"
#include "tess_math.cuh"

namespace tess {
namespace {

constexpr uint32_t FULL = 0xffffffffu;

/// Celery::max_float / min_float (celery.rs:682-697): "when in doubt, return the second one"
__device__ __forceinline__ double max_float(double a, double b) { return a > b ? a : b; }
__device__ __forceinline__ double min_float(double a, double b) { return a < b ? a : b; }

template <bool FILL>
__global__ void __launch_bounds__(128) radius_query_kernel(const QueryParams P) {
    const size_t q = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (q >= P.n_query) return;
    const GridSpec& G = P.grid;
    const double x = P.xyz[3 * q], y = P.xyz[3 * q + 1], z = P.xyz[3 * q + 2];
    const double r = P.radius;
    unsigned long long out = FILL ? P.offsets[q] : 0ull;
    unsigned long long count = 0;

    // append the particles of grid cell c that pass the filters (warp-cooperative, order preserving)
    auto emit_cell = [&](uint32_t c) {
        const uint32_t d0 = __ldg(P.delim + c), d1 = __ldg(P.delim + c + 1);
        for (uint32_t base = d0; base < d1; base += 32) {
            const uint32_t s = base + lane;
            bool keep = s < d1;
            long long id = 0;
            if (keep) {
                const double2* pq = reinterpret_cast<const double2*>(P.sorted + s);
                const double2 a = __ldg(pq), b = __ldg(pq + 1);
                id = __double_as_longlong(b.y);
                if (P.mode == 1) {
                    // Celery::distance_squared (celery.rs:700-702) with (x1,y1,z1) = the query
                    const double dx = subd(x, a.x), dy = subd(y, a.y), dz = subd(z, b.x);
                    const double d2 = addd(addd(mul(dx, dx), mul(dy, dy)), mul(dz, dz));
                    keep = d2 <= mul(r, r);  // celery.rs:848
                }
                if (keep && P.target_group >= 0) keep = P.groups_sorted ? P.groups_sorted[s] == (uint64_t)P.target_group : P.target_group == 0;
            }
            const uint32_t m = __ballot_sync(FULL, keep);
            if (FILL && keep) P.indices[out + __popc(m & ((1u << lane) - 1u))] = id;
            out += __popc(m);
            count += __popc(m);
        }
    };

    // append one value (a grid cell id)
    auto emit_value = [&](long long v) {
        if (FILL && lane == 0) P.indices[out] = v;
        out += 1;
        count += 1;
    };

    if (P.mode == 4) {
        // ExpandingSearch::expand (celery.rs:907-963): at most cells_to_add table entries from the cursor on — entries
        // outside the grid count too (:937-946 `continue` after the cursor moved) — stopping BEFORE the first entry whose
        // key exceeds max_radius (:925-928) and at the end of the table (:918-920)
        const int cpd = (int)G.cpd;
        const int hx = (int)axis_index(x, G.xmin, G.xmax, G.ix, G.cpd);
        const int hy = (int)axis_index(y, G.ymin, G.ymax, G.iy, G.cpd);
        const int hz = (int)axis_index(z, G.zmin, G.zmax, G.iz, G.cpd);
        unsigned long long t = P.cursor_in[q];
        bool ran_off = false;
        for (unsigned long long it = 0; it < P.cells_to_add; ++it) {
            if (t >= P.table_len) {
                ran_off = !P.table_full;  // a truncated table: the caller widens it and asks again
                break;
            }
            const ShellEntry e = P.table[t];
            if (e.key > r) break;
            ++t;
            const int gx = hx + e.di, gy = hy + e.dj, gz = hz + e.dk;
            if (gx < 0 || gx >= cpd || gy < 0 || gy >= cpd || gz < 0 || gz >= cpd) continue;
            emit_cell(((uint32_t)gx * G.cpd + (uint32_t)gy) * G.cpd + (uint32_t)gz);
        }
        if (!FILL && lane == 0) {
            P.cursor_out[q] = t;
            P.flags[q] = ran_off ? ST_TABLE_EXHAUSTED : 0u;
        }
    } else if (P.mode == 2) {
        // expand_all_in_radius: walk the table until an entry's (squared) key exceeds max_radius (D11)
        const int cpd = (int)G.cpd;
        const int hx = (int)axis_index(x, G.xmin, G.xmax, G.ix, G.cpd);
        const int hy = (int)axis_index(y, G.ymin, G.ymax, G.iy, G.cpd);
        const int hz = (int)axis_index(z, G.zmin, G.zmax, G.iz, G.cpd);
        bool stopped = false;
        for (uint32_t t = 0; t < P.table_len; ++t) {
            const ShellEntry e = P.table[t];
            if (e.key > r) {  // celery.rs:1036
                stopped = true;
                break;
            }
            const int gx = hx + e.di, gy = hy + e.dj, gz = hz + e.dk;
            if (gx < 0 || gx >= cpd || gy < 0 || gy >= cpd || gz < 0 || gz >= cpd) continue;  // celery.rs:1048-1056
            emit_cell(((uint32_t)gx * G.cpd + (uint32_t)gy) * G.cpd + (uint32_t)gz);
        }
        if (!FILL && lane == 0) P.flags[q] = (!stopped && !P.table_full) ? ST_TABLE_EXHAUSTED : 0u;
    } else {
        // find_cells_in_radius (celery.rs:753-797)
        const uint32_t x0 = axis_index(max_float(subd(x, r), G.xmin), G.xmin, G.xmax, G.ix, G.cpd);
        const uint32_t y0 = axis_index(max_float(subd(y, r), G.ymin), G.ymin, G.ymax, G.iy, G.cpd);
        const uint32_t z0 = axis_index(max_float(subd(z, r), G.zmin), G.zmin, G.zmax, G.iz, G.cpd);
        const uint32_t x1 = axis_index(min_float(addd(x, r), G.xmax), G.xmin, G.xmax, G.ix, G.cpd);
        const uint32_t y1 = axis_index(min_float(addd(y, r), G.ymax), G.ymin, G.ymax, G.iy, G.cpd);
        const uint32_t z1 = axis_index(min_float(addd(z, r), G.zmax), G.zmin, G.zmax, G.iz, G.cpd);
        // check_cell_in_range (celery.rs:708-743)
        const int xi = (int)axis_index(x, G.xmin, G.xmax, G.ix, G.cpd);
        const int yi = (int)axis_index(y, G.ymin, G.ymax, G.iy, G.cpd);
        const int zi = (int)axis_index(z, G.zmin, G.zmax, G.iz, G.cpd);
        const double rr = mul(r, r);
        for (uint32_t i = x0; i <= x1; ++i)
            for (uint32_t j = y0; j <= y1; ++j)
                for (uint32_t k = z0; k <= z1; ++k) {
                    const int ox = max(0, abs(xi - (int)i) - 1), oy = max(0, abs(yi - (int)j) - 1), oz = max(0, abs(zi - (int)k) - 1);
                    const double tx = mul((double)ox, G.sx), ty = mul((double)oy, G.sy), tz = mul((double)oz, G.sz);
                    const double ds = addd(addd(mul(tx, tx), mul(ty, ty)), mul(tz, tz));
                    if (ds <= rr) {
                        const uint32_t c = (i * G.cpd + j) * G.cpd + k;  // get_cell_from_indices (celery.rs:317-325)
                        if (P.mode == 3) emit_value((long long)c);
                        else emit_cell(c);
                    }
                }
        if (!FILL && lane == 0) P.flags[q] = 0u;
    }
    if (!FILL && lane == 0) P.counts[q] = (uint32_t)count;
}

}  // namespace

void launch_radius_query(const QueryParams& p, bool fill, cudaStream_t s) {
    if (!p.n_query) return;
    const unsigned int nb = (unsigned int)((p.n_query * 32 + 127) / 128);
    if (fill) TESS_LAUNCH(radius_query_kernel<true>, nb, 128, 0, s, p);
    else TESS_LAUNCH(radius_query_kernel<false>, nb, 128, 0, s, p);
    note_launch();
    TESS_CUDA_CHECK(cudaGetLastError());
}

}  // namespace tess
