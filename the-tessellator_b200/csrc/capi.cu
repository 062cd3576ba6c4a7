// capi.cu — implementation of include/tess.h: host-side orchestration of the binning pass
// (grid.cu), the clip kernel (clip.cu) and the CSR outputs (outputs.cu).
//
// Host responsibilities that mirror reference code:
//   * CeleryCellInfo::new (celery.rs:153-189): cpd = floor(cbrt(N/1.25)) + 1 with *glibc* cbrt
//     (N/1.25 is an exact cube for N = 10k and 10M; a 1-ulp error would change cpd), cell sizes
//     and inverse sizes with the reference's two divisions;
//   * Celery::get_search_order (celery.rs:418-679): the offset table, here truncated to
//     |i|,|j|,|k| <= R and to keys strictly below min_axis sq(R*size) so that it is an exact
//     prefix of the reference's full (2cpd-1)^3 table, ordered canonically by (key, i, j, k).
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/tess.h"
#include "common.cuh"

using namespace tess;

namespace {

thread_local std::string g_err;

int fail(int code, const std::string& msg) {
    g_err = msg;
    return code;
}

struct DevBuf {  // grow-only device buffer
    void* p = nullptr;
    size_t bytes = 0;
    void reserve(size_t need) {
        if (need <= bytes) return;
        if (p) TESS_CUDA_CHECK(cudaFree(p));
        p = nullptr;
        bytes = 0;
        const size_t want = need + need / 8;
        TESS_CUDA_CHECK(cudaMalloc(&p, want));
        bytes = want;
    }
    void grow_keep(size_t need, size_t keep_bytes, cudaStream_t s) {  // reserve preserving contents
        if (need <= bytes) return;
        void* q = nullptr;
        const size_t want = std::max(need + need / 2, (size_t)1 << 20);
        TESS_CUDA_CHECK(cudaMalloc(&q, want));
        if (p && keep_bytes) TESS_CUDA_CHECK(cudaMemcpyAsync(q, p, keep_bytes, cudaMemcpyDeviceToDevice, s));
        if (p) {
            TESS_CUDA_CHECK(cudaStreamSynchronize(s));
            TESS_CUDA_CHECK(cudaFree(p));
        }
        p = q;
        bytes = want;
    }
    void release() {
        if (p) cudaFree(p);
        p = nullptr;
        bytes = 0;
    }
    template <class T>
    T* as() const { return static_cast<T*>(p); }
};

struct ShellTable {
    DevBuf dev;
    uint32_t len = 0;
    bool full = false;
    std::vector<ShellEntry> host;
};

/// float.rs:138-142: `as usize` saturates
size_t to_usize(double v) {
    if (!(v > 0.0)) return 0;
    if (v >= 18446744073709551616.0) return std::numeric_limits<size_t>::max();
    return static_cast<size_t>(v);
}

/// CeleryCellInfo::new (celery.rs:153-181)
void fill_cell_info(GridSpec& g, uint64_t n_points) {
    const double num_points = static_cast<double>(n_points);
    const size_t cpd = to_usize(std::cbrt(num_points / 1.25)) + 1;  // celery.rs:161-162, glibc cbrt
    const double c = static_cast<double>(cpd);
    g.cpd = static_cast<uint32_t>(cpd);
    g.sx = (g.xmax - g.xmin) / c;
    g.sy = (g.ymax - g.ymin) / c;
    g.sz = (g.zmax - g.zmin) / c;
    g.ix = c / (g.xmax - g.xmin);
    g.iy = c / (g.ymax - g.ymin);
    g.iz = c / (g.zmax - g.zmin);
}

/// Celery::get_search_order (celery.rs:418-679), truncated to half-width R.
void build_shell_table(const GridSpec& g, int R, std::vector<ShellEntry>& out, bool& full) {
    const int max_index = static_cast<int>(g.cpd) - 1;  // celery.rs:430
    full = (R <= 0 || R >= max_index);
    const int lim = full ? max_index : R;
    auto sq = [](double x) { return x * x; };
    // celery.rs:423-427; an offset o != 0 carries the distance term of |o|-1 (celery.rs:450-457)
    auto term = [](int o) { return o == 0 ? 0 : std::abs(o) - 1; };
    out.clear();
    out.reserve(static_cast<size_t>(2 * lim + 1) * (2 * lim + 1) * (2 * lim + 1));
    for (int i = -lim; i <= lim; ++i)
        for (int j = -lim; j <= lim; ++j)
            for (int k = -lim; k <= lim; ++k) {
                ShellEntry e;
                e.di = static_cast<int16_t>(i);
                e.dj = static_cast<int16_t>(j);
                e.dk = static_cast<int16_t>(k);
                e.pad = 0;
                if (i == 0 && j == 0 && k == 0) {
                    e.key = -1.0;  // celery.rs:437-442
                } else {
                    e.key = sq(static_cast<double>(term(i)) * g.sx) + sq(static_cast<double>(term(j)) * g.sy) + sq(static_cast<double>(term(k)) * g.sz);
                }
                out.push_back(e);
            }
    // celery.rs:676 sorts by distance only (unstable); the canonical tie-break is (i, j, k)
    std::sort(out.begin(), out.end(), [](const ShellEntry& a, const ShellEntry& b) {
        if (a.key != b.key) return a.key < b.key;
        if (a.di != b.di) return a.di < b.di;
        if (a.dj != b.dj) return a.dj < b.dj;
        return a.dk < b.dk;
    });
    if (!full) {
        const double bound = std::min(sq(static_cast<double>(lim) * g.sx), std::min(sq(static_cast<double>(lim) * g.sy), sq(static_cast<double>(lim) * g.sz)));
        size_t keep = 0;
        while (keep < out.size() && out[keep].key < bound) ++keep;
        out.resize(keep);
    }
}

constexpr int kDefaultTableRadius = 8;

// TESS_TRACE=1: host-side phase timings on stderr (wall clock, each phase closed by a stream sync)
struct Trace {
    bool on;
    cudaStream_t s;
    const char* what;
    std::chrono::steady_clock::time_point t0;
    Trace(const char* w, cudaStream_t st) : on(std::getenv("TESS_TRACE") != nullptr), s(st), what(w), t0(std::chrono::steady_clock::now()) {}
    void mark(const char* phase) {
        if (!on) return;
        cudaStreamSynchronize(s);
        const auto t1 = std::chrono::steady_clock::now();
        std::fprintf(stderr, "[tess trace] %s / %-24s %9.3f ms\n", what, phase, std::chrono::duration<double, std::milli>(t1 - t0).count());
        t0 = t1;
    }
};

}  // namespace

struct tess_diagram {
    int device = 0;
    bool initialized = false;
    bool slab = false;
    // accumulated input
    DevBuf xyz, groups, ids;
    size_t n = 0;
    bool has_groups = false, has_ids = false;
    // grid
    GridSpec grid{};
    double box[6] = {0, 0, 0, 0, 0, 0};
    size_t n_cells_local = 0;
    DevBuf counts, delim, cell_of, rank_in_cell, tmp_idx, arrived, sorted, sorted_idx, groups_sorted, scan_tmp, small;
    uint32_t own_slot_begin = 0, own_slot_end = 0;
    cudaEvent_t ev_bin0 = nullptr, ev_bin1 = nullptr;  // around K2-K4 of the last initialize
    mutable std::mutex mu;
    mutable std::map<int, ShellTable> tables;
    uint32_t table_cpd = 0;
    double table_size[3] = {0, 0, 0};

    const ShellTable& table(int R, cudaStream_t s) const {
        std::lock_guard<std::mutex> lk(mu);
        int key = R;
        if (R <= 0 || R >= static_cast<int>(grid.cpd) - 1) key = 0;  // full
        auto it = tables.find(key);
        if (it != tables.end()) return it->second;
        ShellTable& t = tables[key];
        build_shell_table(grid, key, t.host, t.full);
        t.len = static_cast<uint32_t>(t.host.size());
        t.dev.reserve(sizeof(ShellEntry) * t.host.size());
        TESS_CUDA_CHECK(cudaMemcpyAsync(t.dev.p, t.host.data(), sizeof(ShellEntry) * t.host.size(), cudaMemcpyHostToDevice, s));
        TESS_CUDA_CHECK(cudaStreamSynchronize(s));
        return t;
    }
};

struct tess_result {
    int device = 0;
    cudaStream_t stream = nullptr;  // stream the arrays were allocated on (stream-ordered allocator)
    uint64_t n_cells = 0, n_faces = 0, n_vertices = 0, n_loop_entries = 0;
    // device arrays (cudaMallocAsync on `stream`: cached by the device's memory pool between steps)
    double* vol = nullptr;
    uint32_t* nfaces = nullptr;
    uint32_t* status = nullptr;
    int64_t* cell_id = nullptr;
    uint64_t* offsets = nullptr;
    int64_t* nbr = nullptr;
    double* area = nullptr;
    uint32_t* nverts = nullptr;
    uint64_t* voffsets = nullptr;
    uint64_t* fv_offsets = nullptr;  // per face: start of its vertex loop (n_faces+1)
    uint32_t* fv_idx = nullptr;      // loop entries: rank of the vertex in the cell's vertex list
    double* vtx = nullptr;
    unsigned long long* counters = nullptr;
    unsigned long long counters_redo[CNT_N] = {0, 0, 0, 0, 0, 0, 0, 0};  // work done by the large-cell pass
    double ms_clip = 0, ms_redo = 0, ms_outputs = 0, ms_total = 0;  // CUDA-event durations on the launching stream
    uint64_t tier_stats[4] = {0, 0, 0, 0};  // main tier (CLIP_*), cells redone by pass A / B / C
    // host copies
    std::vector<double> h_vol, h_area, h_vtx;
    std::vector<uint64_t> h_offsets, h_voffsets, h_fv_offsets;
    std::vector<uint32_t> h_fv_idx;
    bool have_fvo = false, have_fvi = false;
    std::vector<int64_t> h_nbr, h_cell_id;
    std::vector<uint32_t> h_status;
    bool have_vol = false, have_area = false, have_offsets = false, have_nbr = false, have_ids = false, have_status = false, have_voff = false, have_vtx = false;
    ~tess_result() {
        cudaSetDevice(device);
        for (void* p : {(void*)vol, (void*)nfaces, (void*)status, (void*)cell_id, (void*)offsets, (void*)nbr, (void*)area, (void*)nverts, (void*)voffsets, (void*)vtx, (void*)counters, (void*)fv_offsets, (void*)fv_idx})
            if (p) cudaFreeAsync(p, stream);
    }
};

#define TESS_TRY try {
#define TESS_CATCH                                            \
    }                                                         \
    catch (const std::bad_alloc&) { return fail(TESS_ERR_NOMEM, "out of host memory"); } \
    catch (const std::exception& e) { return fail(TESS_ERR_CUDA, e.what()); }

static GridSpec spec_from_bounds(const double b[6], uint64_t n_global) {
    GridSpec g{};
    g.xmin = b[0]; g.xmax = b[1]; g.ymin = b[2]; g.ymax = b[3]; g.zmin = b[4]; g.zmax = b[5];
    fill_cell_info(g, n_global);
    g.local_lo = 0; g.local_hi = g.cpd; g.own_lo = 0; g.own_hi = g.cpd;
    return g;
}

extern "C" {

const char* tess_last_error(void) { return g_err.c_str(); }
int tess_version(void) { return 100; }

int tess_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    int ok = 0;
    for (int d = 0; d < n; ++d) {
        int major = 0;
        if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, d) == cudaSuccess && major == 10) ++ok;
    }
    return ok;
}

void tess_opts_default(tess_opts* o) {
    if (!o) return;
    o->search_radius = std::numeric_limits<double>::quiet_NaN();
    o->target_group = -1;
    o->outputs = TESS_OUT_VOLUME | TESS_OUT_NEIGHBORS | TESS_OUT_AREAS;
    o->table_radius = 0;
    o->stream = nullptr;
}

int tess_diagram_create(tess_diagram** out, int real_type, int device) {
    if (!out) return fail(TESS_ERR_INVALID, "out is NULL");
    *out = nullptr;
    if (real_type != TESS_F64) return fail(TESS_ERR_UNSUPPORTED, "only TESS_F64 is implemented (the reference implements only Float64, float.rs:78)");
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) {
        cudaGetLastError();
        return fail(TESS_ERR_CUDA, "no CUDA device available: libtess_b200 has no CPU fallback");
    }
    if (device < 0 || device >= n) return fail(TESS_ERR_INVALID, "bad device ordinal");
    int major = 0;
    cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, device);
    if (major != 10) return fail(TESS_ERR_CUDA, "device is not sm_100 (B200); this library contains sm_100a code only");
    TESS_TRY
    TESS_CUDA_CHECK(cudaSetDevice(device));
    // keep stream-ordered allocations cached between steps
    cudaMemPool_t pool;
    TESS_CUDA_CHECK(cudaDeviceGetDefaultMemPool(&pool, device));
    uint64_t thr = ~0ull;
    TESS_CUDA_CHECK(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr));
    auto* d = new tess_diagram();
    d->device = device;
    *out = d;
    return TESS_OK;
    TESS_CATCH
}

void tess_diagram_destroy(tess_diagram* d) {
    if (!d) return;
    cudaSetDevice(d->device);
    for (DevBuf* b : {&d->xyz, &d->groups, &d->ids, &d->counts, &d->delim, &d->cell_of, &d->rank_in_cell, &d->tmp_idx, &d->arrived, &d->sorted, &d->sorted_idx,
                      &d->groups_sorted, &d->scan_tmp, &d->small})
        b->release();
    for (auto& kv : d->tables) kv.second.dev.release();
    if (d->ev_bin0) cudaEventDestroy(d->ev_bin0);
    if (d->ev_bin1) cudaEventDestroy(d->ev_bin1);
    delete d;
}

static int add_common(tess_diagram* d, size_t n, const void* groups_src, const void* ids_src, cudaMemcpyKind kind, cudaStream_t s) {
    const size_t n0 = d->n;
    if (groups_src && !d->has_groups && n0) {  // earlier particles default to group 0
        d->groups.grow_keep(sizeof(uint64_t) * (n0 + n), 0, s);
        TESS_CUDA_CHECK(cudaMemsetAsync(d->groups.p, 0, sizeof(uint64_t) * n0, s));
    }
    if (groups_src || d->has_groups) {
        d->groups.grow_keep(sizeof(uint64_t) * (n0 + n), d->has_groups ? sizeof(uint64_t) * n0 : 0, s);
        if (groups_src) TESS_CUDA_CHECK(cudaMemcpyAsync(d->groups.as<uint64_t>() + n0, groups_src, sizeof(uint64_t) * n, kind, s));
        else TESS_CUDA_CHECK(cudaMemsetAsync(d->groups.as<uint64_t>() + n0, 0, sizeof(uint64_t) * n, s));
        d->has_groups = true;
    }
    if (ids_src) {
        if (!d->has_ids && n0) return fail(TESS_ERR_INVALID, "ids must be given for all particles or for none");
        d->ids.grow_keep(sizeof(int64_t) * (n0 + n), sizeof(int64_t) * n0, s);
        TESS_CUDA_CHECK(cudaMemcpyAsync(d->ids.as<int64_t>() + n0, ids_src, sizeof(int64_t) * n, kind, s));
        d->has_ids = true;
    } else if (d->has_ids) {
        return fail(TESS_ERR_INVALID, "ids must be given for all particles or for none");
    }
    d->n = n0 + n;
    return TESS_OK;
}

int tess_diagram_add_particles(tess_diagram* d, const void* xyz, size_t n, size_t stride_bytes, const uint64_t* groups, void* stream) {
    if (!d || (!xyz && n)) return fail(TESS_ERR_INVALID, "NULL argument");
    if (d->initialized) return fail(TESS_ERR_STATE, "particles must be added before initialize (interface.rs:50-51)");
    if (stride_bytes < 3 * sizeof(double)) return fail(TESS_ERR_INVALID, "stride_bytes must be >= 24");
    if (d->n + n >= 0xFFFFFFF0ull) return fail(TESS_ERR_INVALID, "more than 2^32 particles on one device");
    if (!n) return TESS_OK;
    TESS_TRY
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    TESS_CUDA_CHECK(cudaSetDevice(d->device));
    d->xyz.grow_keep(sizeof(double) * 3 * (d->n + n), sizeof(double) * 3 * d->n, s);
    double* dst = d->xyz.as<double>() + 3 * d->n;
    if (stride_bytes == 3 * sizeof(double))
        TESS_CUDA_CHECK(cudaMemcpyAsync(dst, xyz, sizeof(double) * 3 * n, cudaMemcpyHostToDevice, s));
    else
        TESS_CUDA_CHECK(cudaMemcpy2DAsync(dst, 3 * sizeof(double), xyz, stride_bytes, 3 * sizeof(double), n, cudaMemcpyHostToDevice, s));
    return add_common(d, n, groups, nullptr, cudaMemcpyHostToDevice, s);
    TESS_CATCH
}

int tess_diagram_add_particles_device(tess_diagram* d, const double* xyz_dev, size_t n, const uint64_t* groups_dev, const int64_t* ids_dev, void* stream) {
    if (!d || (!xyz_dev && n)) return fail(TESS_ERR_INVALID, "NULL argument");
    if (d->initialized) return fail(TESS_ERR_STATE, "particles must be added before initialize (interface.rs:50-51)");
    if (d->n + n >= 0xFFFFFFF0ull) return fail(TESS_ERR_INVALID, "more than 2^32 particles on one device");
    if (!n) return TESS_OK;
    TESS_TRY
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    TESS_CUDA_CHECK(cudaSetDevice(d->device));
    d->xyz.grow_keep(sizeof(double) * 3 * (d->n + n), sizeof(double) * 3 * d->n, s);
    TESS_CUDA_CHECK(cudaMemcpyAsync(d->xyz.as<double>() + 3 * d->n, xyz_dev, sizeof(double) * 3 * n, cudaMemcpyDeviceToDevice, s));
    return add_common(d, n, groups_dev, ids_dev, cudaMemcpyDeviceToDevice, s);
    TESS_CATCH
}

int tess_diagram_clear(tess_diagram* d) {
    if (!d) return fail(TESS_ERR_INVALID, "NULL diagram");
    d->n = 0;
    d->has_groups = d->has_ids = false;
    d->initialized = false;
    d->slab = false;
    return TESS_OK;  // shell tables are kept: initialize drops them only if the grid geometry changes
}

static int initialize_impl(tess_diagram* d, const double* box, const tess_slab* slab, cudaStream_t s) {
    if (d->initialized) return fail(TESS_ERR_STATE, "diagram already initialized (interface.rs:65)");
    if (d->n == 0) return fail(TESS_ERR_INVALID, "no particles (CeleryBounds::new panics on an empty set, celery.rs:82)");
    TESS_CUDA_CHECK(cudaSetDevice(d->device));
    Trace tr("initialize", s);
    const size_t n = d->n;
    d->small.reserve(256);
    double* dev_bounds = d->small.as<double>();             // 6 doubles
    uint32_t* dev_flag = reinterpret_cast<uint32_t*>(d->small.as<char>() + 64);

    GridSpec g{};
    if (slab) {
        g = spec_from_bounds(slab->bounds, slab->n_global);
        if (!(slab->local_lo <= slab->own_lo && slab->own_lo <= slab->own_hi && slab->own_hi <= slab->local_hi && slab->local_hi <= g.cpd))
            return fail(TESS_ERR_INVALID, "slab plane ranges must satisfy local_lo <= own_lo <= own_hi <= local_hi <= cpd");
        if (slab->own_lo_row >= g.cpd || slab->own_hi_row >= g.cpd || (slab->own_hi_row > 0 && slab->own_hi >= slab->local_hi) ||
            (slab->own_lo_row > 0 && slab->own_lo >= slab->local_hi) ||
            (uint64_t)slab->own_lo * g.cpd + slab->own_lo_row > (uint64_t)slab->own_hi * g.cpd + slab->own_hi_row)
            return fail(TESS_ERR_INVALID, "slab row offsets must be < cpd, lie in planes held locally, and keep the owned range non-negative");
        g.local_lo = slab->local_lo; g.local_hi = slab->local_hi; g.own_lo = slab->own_lo; g.own_hi = slab->own_hi;
        d->slab = true;
    } else {
        launch_bounds(d->xyz.as<double>(), n, dev_bounds, s);  // K1: CeleryBounds::new
        double hb[6];
        TESS_CUDA_CHECK(cudaMemcpyAsync(hb, dev_bounds, sizeof(hb), cudaMemcpyDeviceToHost, s));
        TESS_CUDA_CHECK(cudaStreamSynchronize(s));
        g = spec_from_bounds(hb, n);
        d->slab = false;
    }
    if (box) {
        std::memcpy(d->box, box, sizeof(d->box));
    } else {  // interface.rs:69-79: no container given -> bounding box of the points
        if (slab) return fail(TESS_ERR_INVALID, "slab diagrams need an explicit container box");
        d->box[0] = g.xmin; d->box[1] = g.ymin; d->box[2] = g.zmin; d->box[3] = g.xmax; d->box[4] = g.ymax; d->box[5] = g.zmax;
    }
    const uint64_t planes = g.local_hi - g.local_lo;
    const uint64_t ncl = planes * g.cpd * g.cpd;
    if (ncl >= 0xFFFFFFF0ull) return fail(TESS_ERR_INVALID, "local grid has more than 2^32 cells");
    d->grid = g;
    d->n_cells_local = static_cast<size_t>(ncl);

    d->counts.reserve(sizeof(uint32_t) * (ncl + 1));
    d->delim.reserve(sizeof(uint32_t) * (ncl + 1));
    d->cell_of.reserve(sizeof(uint32_t) * (n + 2));
    d->rank_in_cell.reserve(sizeof(uint32_t) * (n + 2));
    if (d->has_ids) d->tmp_idx.reserve(sizeof(uint32_t) * n);
    d->arrived.reserve(sizeof(Particle) * n);
    d->sorted.reserve(sizeof(Particle) * n);
    d->sorted_idx.reserve(sizeof(uint32_t) * n);
    if (d->has_groups) d->groups_sorted.reserve(sizeof(uint64_t) * n);
    d->scan_tmp.reserve(scan_tmp_bytes(ncl + 1));
    tr.mark("bounds+reserve");

    if (!d->ev_bin0) {
        TESS_CUDA_CHECK(cudaEventCreate(&d->ev_bin0));
        TESS_CUDA_CHECK(cudaEventCreate(&d->ev_bin1));
    }
    TESS_CUDA_CHECK(cudaEventRecord(d->ev_bin0, s));
    TESS_CUDA_CHECK(cudaMemsetAsync(d->counts.p, 0, sizeof(uint32_t) * (ncl + 1), s));
    TESS_CUDA_CHECK(cudaMemsetAsync(dev_flag, 0, sizeof(uint32_t), s));
    // K2: cell ids + histogram
    launch_cell_histogram(d->xyz.as<double>(), n, g, d->cell_of.as<uint32_t>(), d->rank_in_cell.as<uint32_t>(), d->counts.as<uint32_t>(), dev_flag, s);
    if (slab) {
        uint32_t flag = 0;
        TESS_CUDA_CHECK(cudaMemcpyAsync(&flag, dev_flag, sizeof(flag), cudaMemcpyDeviceToHost, s));
        TESS_CUDA_CHECK(cudaStreamSynchronize(s));
        if (flag) return fail(TESS_ERR_INVALID, "a particle lies outside the slab's local x-plane range");
    }
    // K3: delimiters = exclusive scan of the counts (celery.rs:372-414); delim[ncl] = n
    launch_exclusive_scan_u32(d->counts.as<uint32_t>(), d->delim.as<uint32_t>(), ncl + 1, d->scan_tmp.p, d->scan_tmp.bytes, s);
    // K4: counting-sort scatter, canonical in-cell order, gather of the particle records
    launch_scatter_records(d->xyz.as<double>(), d->has_ids ? d->ids.as<int64_t>() : nullptr, d->cell_of.as<uint32_t>(), d->rank_in_cell.as<uint32_t>(), d->delim.as<uint32_t>(),
                           d->arrived.as<Particle>(), d->has_ids ? d->tmp_idx.as<uint32_t>() : nullptr, n, s);
    launch_rank_fix(d->arrived.as<Particle>(), d->has_ids ? d->tmp_idx.as<uint32_t>() : nullptr, g, d->delim.as<uint32_t>(), d->has_groups ? d->groups.as<uint64_t>() : nullptr,
                    d->sorted.as<Particle>(), d->sorted_idx.as<uint32_t>(), d->has_groups ? d->groups_sorted.as<uint64_t>() : nullptr, n, s);
    TESS_CUDA_CHECK(cudaEventRecord(d->ev_bin1, s));
    tr.mark("binning kernels");
    if (slab) {
        uint32_t b = 0, e = 0;
        // the owned cells: grid rows (x, y) in [own_lo*cpd + own_lo_row, own_hi*cpd + own_hi_row) — a run of the sorted order
        const size_t cb = (static_cast<size_t>(g.own_lo - g.local_lo) * g.cpd + slab->own_lo_row) * g.cpd,
                     ce = (static_cast<size_t>(g.own_hi - g.local_lo) * g.cpd + slab->own_hi_row) * g.cpd;
        TESS_CUDA_CHECK(cudaMemcpyAsync(&b, d->delim.as<uint32_t>() + cb, sizeof(b), cudaMemcpyDeviceToHost, s));
        TESS_CUDA_CHECK(cudaMemcpyAsync(&e, d->delim.as<uint32_t>() + ce, sizeof(e), cudaMemcpyDeviceToHost, s));
        TESS_CUDA_CHECK(cudaStreamSynchronize(s));
        d->own_slot_begin = b;
        d->own_slot_end = e;
    } else {
        d->own_slot_begin = 0;
        d->own_slot_end = static_cast<uint32_t>(n);
    }
    {
        // the search-order tables depend only on cpd and the cell sizes (celery.rs:423-427)
        std::lock_guard<std::mutex> lk(d->mu);
        const bool same = d->table_cpd == g.cpd && d->table_size[0] == g.sx && d->table_size[1] == g.sy && d->table_size[2] == g.sz;
        if (!same) {
            for (auto& kv : d->tables) kv.second.dev.release();
            d->tables.clear();
            d->table_cpd = g.cpd;
            d->table_size[0] = g.sx; d->table_size[1] = g.sy; d->table_size[2] = g.sz;
        }
    }
    d->initialized = true;
    return TESS_OK;
}

int tess_diagram_initialize(tess_diagram* d, const double box[6], void* stream) {
    if (!d) return fail(TESS_ERR_INVALID, "NULL diagram");
    TESS_TRY
    return initialize_impl(d, box, nullptr, static_cast<cudaStream_t>(stream));
    TESS_CATCH
}

int tess_diagram_initialize_slab(tess_diagram* d, const double box[6], const tess_slab* slab, void* stream) {
    if (!d || !slab || !box) return fail(TESS_ERR_INVALID, "NULL argument");
    TESS_TRY
    return initialize_impl(d, box, slab, static_cast<cudaStream_t>(stream));
    TESS_CATCH
}

int tess_diagram_grid_info(const tess_diagram* d, uint64_t* n_points, uint64_t* cpd, double bounds[6], double cell_sizes[3], double inverse_cell_sizes[3]) {
    if (!d) return fail(TESS_ERR_INVALID, "NULL diagram");
    if (!d->initialized) return fail(TESS_ERR_STATE, "diagram not initialized");
    const GridSpec& g = d->grid;
    if (n_points) *n_points = d->n;
    if (cpd) *cpd = g.cpd;
    if (bounds) { bounds[0] = g.xmin; bounds[1] = g.xmax; bounds[2] = g.ymin; bounds[3] = g.ymax; bounds[4] = g.zmin; bounds[5] = g.zmax; }
    if (cell_sizes) { cell_sizes[0] = g.sx; cell_sizes[1] = g.sy; cell_sizes[2] = g.sz; }
    if (inverse_cell_sizes) { inverse_cell_sizes[0] = g.ix; inverse_cell_sizes[1] = g.iy; inverse_cell_sizes[2] = g.iz; }
    return TESS_OK;
}

int tess_diagram_copy_grid(const tess_diagram* d, uint64_t* cells, uint64_t* sorted_indices, uint64_t* delimiters) {
    if (!d) return fail(TESS_ERR_INVALID, "NULL diagram");
    if (!d->initialized) return fail(TESS_ERR_STATE, "diagram not initialized");
    TESS_TRY
    TESS_CUDA_CHECK(cudaSetDevice(d->device));
    TESS_CUDA_CHECK(cudaDeviceSynchronize());
    std::vector<uint32_t> tmp;
    auto fetch = [&](const DevBuf& b, size_t count, uint64_t* out) {
        tmp.resize(count);
        TESS_CUDA_CHECK(cudaMemcpy(tmp.data(), b.p, sizeof(uint32_t) * count, cudaMemcpyDeviceToHost));
        for (size_t i = 0; i < count; ++i) out[i] = tmp[i];
    };
    if (cells) fetch(d->cell_of, d->n, cells);
    if (sorted_indices) fetch(d->sorted_idx, d->n, sorted_indices);
    if (delimiters) fetch(d->delim, d->n_cells_local + 1, delimiters);
    return TESS_OK;
    TESS_CATCH
}

int tess_diagram_copy_search_order(const tess_diagram* d, int32_t table_radius, uint64_t* len, double* keys, int32_t* ijk, int* is_full) {
    if (!d) return fail(TESS_ERR_INVALID, "NULL diagram");
    if (!d->initialized) return fail(TESS_ERR_STATE, "diagram not initialized");
    TESS_TRY
    TESS_CUDA_CHECK(cudaSetDevice(d->device));
    const ShellTable& t = d->table(table_radius == 0 ? kDefaultTableRadius : table_radius, nullptr);
    if (len) *len = t.len;
    if (is_full) *is_full = t.full ? 1 : 0;
    if (keys && ijk) {
        // read the table back from the DEVICE copy the kernel walks
        std::vector<ShellEntry> h(t.len);
        TESS_CUDA_CHECK(cudaMemcpy(h.data(), t.dev.p, sizeof(ShellEntry) * t.len, cudaMemcpyDeviceToHost));
        for (uint32_t i = 0; i < t.len; ++i) {
            keys[i] = h[i].key;
            ijk[3 * i] = h[i].di; ijk[3 * i + 1] = h[i].dj; ijk[3 * i + 2] = h[i].dk;
        }
    }
    return TESS_OK;
    TESS_CATCH
}

// ------------------------------------------------------------------------------------------------
// compute
// ------------------------------------------------------------------------------------------------
}  // extern "C"

namespace {

template <class T>
T* dmalloc(size_t count, cudaStream_t s) {
    void* p = nullptr;
    TESS_CUDA_CHECK(cudaMallocAsync(&p, std::max<size_t>(1, count) * sizeof(T), s));
    return static_cast<T*>(p);
}

struct Scratch {  // stream-ordered temporaries
    cudaStream_t s;
    std::vector<void*> ptrs;
    explicit Scratch(cudaStream_t s_) : s(s_) {}
    template <class T>
    T* get(size_t count) {
        void* p = nullptr;
        TESS_CUDA_CHECK(cudaMallocAsync(&p, std::max<size_t>(1, count) * sizeof(T), s));
        ptrs.push_back(p);
        return static_cast<T*>(p);
    }
    ~Scratch() {
        for (void* p : ptrs) cudaFreeAsync(p, s);
    }
};

// Host buffers tess_compute_all_to_host streams the results into, chunk of rows by chunk of rows.
struct HostSink {
    double* volumes;
    uint64_t* face_offsets;
    int64_t* neighbors;
    double* areas;
    uint32_t* status;
    uint64_t face_capacity;
    int n_chunks;
};

// tess_set_main_tier / TESS_MAIN_TIER (thread | fast | small): which kernel runs the main clip pass; -1 = the default choice
std::atomic<int> g_main_tier{-2};  // -2: not set, read the environment
int forced_main_tier() {
    int t = g_main_tier.load(std::memory_order_relaxed);
    if (t != -2) return t;
    t = -1;
    if (const char* e = std::getenv("TESS_MAIN_TIER")) {
        const std::string v(e);
        if (v == "thread") t = CLIP_THREAD;
        else if (v == "fast") t = CLIP_SMALL_FAST;
        else if (v == "small") t = CLIP_SMALL;
    }
    return t;
}

int compute_impl(const tess_diagram* d, const tess_opts* opts_in, const double* query_host, size_t n_query, tess_result** out, const HostSink* sink = nullptr) {
    tess_opts o;
    if (opts_in) o = *opts_in; else tess_opts_default(&o);
    cudaStream_t s = static_cast<cudaStream_t>(o.stream);
    TESS_CUDA_CHECK(cudaSetDevice(d->device));
    if (o.target_group >= 0 && !d->has_groups && o.target_group != 0) {
        // every particle is in group 0: nothing can cut
    }
    const bool query = query_host != nullptr;
    const size_t n_rows = query ? n_query : static_cast<size_t>(d->own_slot_end - d->own_slot_begin);
    int R0 = o.table_radius > 0 ? o.table_radius : kDefaultTableRadius;
    if (!(o.search_radius != o.search_radius) && o.table_radius <= 0) {
        // reference-radius mode (expand_all_in_radius, celery.rs:1036) walks every table entry with key <= radius: the table
        // must hold them all, (R * size)^2 > radius on every axis — the sizing tess_find_neighbors uses
        const double smin = std::min(d->grid.sx, std::min(d->grid.sy, d->grid.sz));
        int R = static_cast<int>(d->grid.cpd);  // full table
        if (o.search_radius >= 0 && smin > 0) {
            const double need = std::sqrt(o.search_radius) / smin + 2.0;
            if (need < static_cast<double>(d->grid.cpd)) R = std::max(1, static_cast<int>(need));
        }
        R0 = std::max(R0, R);
    }
    const bool want_area = (o.outputs & TESS_OUT_AREAS) != 0;
    const bool want_vtx = (o.outputs & TESS_OUT_VERTICES) != 0;
    const bool want_cnt = (o.outputs & TESS_OUT_COUNTERS) != 0;

    Trace tr("compute", s);
    std::unique_ptr<tess_result> r(new tess_result());
    r->device = d->device;
    r->stream = s;
    r->n_cells = n_rows;
    r->vol = dmalloc<double>(n_rows, s);
    r->nfaces = dmalloc<uint32_t>(n_rows + 1, s);
    r->status = dmalloc<uint32_t>(n_rows, s);
    r->cell_id = dmalloc<int64_t>(n_rows, s);
    r->offsets = dmalloc<uint64_t>(n_rows + 1, s);
    r->counters = dmalloc<unsigned long long>(CNT_N, s);
    TESS_CUDA_CHECK(cudaMemsetAsync(r->counters, 0, sizeof(unsigned long long) * CNT_N, s));
    TESS_CUDA_CHECK(cudaMemsetAsync(r->nfaces, 0, sizeof(uint32_t) * (n_rows + 1), s));
    if (want_vtx) {
        r->nverts = dmalloc<uint32_t>(n_rows + 1, s);
        r->voffsets = dmalloc<uint64_t>(n_rows + 1, s);
        TESS_CUDA_CHECK(cudaMemsetAsync(r->nverts, 0, sizeof(uint32_t) * (n_rows + 1), s));
    }
    if (n_rows == 0) {
        TESS_CUDA_CHECK(cudaMemsetAsync(r->offsets, 0, sizeof(uint64_t), s));
        if (want_vtx) TESS_CUDA_CHECK(cudaMemsetAsync(r->voffsets, 0, sizeof(uint64_t), s));
        TESS_CUDA_CHECK(cudaStreamSynchronize(s));
        *out = r.release();
        return TESS_OK;
    }

    Scratch tmp(s);
    const uint32_t fstride = 40;  // output capacity of the small path; larger cells go to the large path
    // geometry pools (TESS_OUT_VERTICES): room for 48 vertices / 144 loop entries per cell on average
    // (uniform input needs 27 / 81); exceeding it is reported, never silently truncated
    const unsigned long long gv_cap = want_vtx ? (unsigned long long)n_rows * 48ull + 4096ull : 0ull;
    const unsigned long long gl_cap = want_vtx ? (unsigned long long)n_rows * 144ull + 16384ull : 0ull;
    double* gv_xyz = want_vtx ? tmp.get<double>(3 * gv_cap) : nullptr;
    uint32_t* gl_idx = want_vtx ? tmp.get<uint32_t>(gl_cap) : nullptr;
    unsigned long long* g_cursor = tmp.get<unsigned long long>(2);
    uint32_t* nloops = want_vtx ? tmp.get<uint32_t>(n_rows) : nullptr;
    unsigned long long* vbase = want_vtx ? tmp.get<unsigned long long>(n_rows) : nullptr;
    unsigned long long* lbase = want_vtx ? tmp.get<unsigned long long>(n_rows) : nullptr;
    uint16_t* st_flen = want_vtx ? tmp.get<uint16_t>(n_rows * fstride) : nullptr;
    TESS_CUDA_CHECK(cudaMemsetAsync(g_cursor, 0, sizeof(unsigned long long) * 2, s));
    if (want_vtx) TESS_CUDA_CHECK(cudaMemsetAsync(nloops, 0, sizeof(uint32_t) * n_rows, s));
    int64_t* st_nbr = tmp.get<int64_t>(n_rows * fstride);
    double* st_area = want_area ? tmp.get<double>(n_rows * fstride) : nullptr;
    uint32_t* ctrl = tmp.get<uint32_t>(16);  // [0] work counter, [1] n_failed, [2] n_failed of redo pass A, [5] table-only failures, [8]/[9] n_failed of the medium/large passes, [12]/[13] their table-only failures
    uint32_t* failed = tmp.get<uint32_t>(n_rows);
    double* query_dev = nullptr;
    if (query) {
        query_dev = tmp.get<double>(3 * n_query);
        TESS_CUDA_CHECK(cudaMemcpyAsync(query_dev, query_host, sizeof(double) * 3 * n_query, cudaMemcpyHostToDevice, s));
    }
    void* scan_tmp = tmp.get<char>(scan_tmp_bytes(n_rows + 1));
    TESS_CUDA_CHECK(cudaMemsetAsync(ctrl, 0, sizeof(uint32_t) * 16, s));
    tr.mark("allocations");

    const ShellTable& tab = d->table(R0, s);
    tr.mark("shell table");
    ClipParams P{};
    P.sorted = d->sorted.as<Particle>();
    P.delim = d->delim.as<uint32_t>();
    P.groups_sorted = d->has_groups ? d->groups_sorted.as<uint64_t>() : nullptr;
    P.table = tab.dev.as<ShellEntry>();
    P.table_len = tab.len;
    P.table_full = tab.full ? 1u : 0u;
    P.grid = d->grid;
    std::memcpy(P.box, d->box, sizeof(P.box));
    P.slot_begin = query ? 0u : d->own_slot_begin;
    P.n_work = static_cast<uint32_t>(n_rows);
    P.work_slots = nullptr;
    P.query_xyz = query_dev;
    P.target_group = o.target_group;
    if (o.target_group >= 0 && !d->has_groups) P.target_group = (o.target_group == 0) ? -1 : -2;  // -2: nothing matches
    P.search_radius = o.search_radius;
    P.row_of_slot = (d->slab || query) ? nullptr : d->sorted_idx.as<uint32_t>();
    P.row_base = query ? 0u : d->own_slot_begin;  // (query cells: work items are query indices, row = index)
    P.vol = r->vol; P.nfaces = r->nfaces; P.status = r->status; P.cell_id = r->cell_id;
    P.st_nbr = st_nbr; P.st_area = st_area; P.fstride = fstride; P.stage_by_work = 0;
    P.gv_xyz = gv_xyz; P.gl_idx = gl_idx; P.gv_cap = gv_cap; P.gl_cap = gl_cap; P.g_cursor = g_cursor;
    P.nverts = r->nverts; P.nloops = nloops; P.vbase = vbase; P.lbase = lbase; P.st_flen = st_flen;
    P.counters = want_cnt ? r->counters : nullptr;
    P.work_counter = ctrl;
    P.failed_slots = failed;  // query cells too: what the small tables / the default table cannot finish is redone tier by tier
    P.n_failed = ctrl + 1;
    P.failed_cap = static_cast<uint32_t>(n_rows);
    P.mark_large = 0;
    P.flags = (std::getenv("TESS_FORCE_SERIAL") ? 1u : 0u) | (std::getenv("TESS_FORCE_SWEEP") ? 2u : 0u);
    // Main pass: the warp-per-cell kernel without the serial walk and without divergence guards (CLIP_SMALL_FAST); cells
    // that need the reference-shaped serial walk come back flagged like cells that ran out of search table and take redo
    // pass A (CLIP_SMALL, which has the walk).  CLIP_THREAD (one thread per cell, clip_thread.cu) hands back the same way
    // what its tables cannot hold; it writes no geometry output, keeps no work counters and builds no query cells.  Query cells are computed by
    // CLIP_SMALL and, like every main pass's leftovers, redone tier by tier.  tess_set_main_tier / TESS_MAIN_TIER override.
    int main_tier = query ? CLIP_SMALL : CLIP_SMALL_FAST;
    {
        const int forced = forced_main_tier();
        if (forced == CLIP_SMALL || (forced == CLIP_SMALL_FAST && !query) || (forced == CLIP_THREAD && !query && !want_vtx && !want_cnt)) main_tier = forced;
    }
    r->tier_stats[0] = (uint64_t)main_tier;
    // ---- the pipeline ------------------------------------------------------------------------------
    // The rows are computed in C chunks (C = 1 unless the results stream to host buffers,
    // tess_compute_all_to_host).  Per chunk: clip its cells in sorted (spatial) order; redo what the
    // small tables / the default shell table could not finish; extend the CSR offsets; pack the
    // chunk's rows; and, when streaming, copy them to the host on a second stream.  The clip launch
    // of chunk c+1 is enqueued before chunk c is finished off, so the device never waits for the
    // host, and the copy of chunk c overlaps the clip kernel of chunk c+2.
    const bool streaming = sink && !query && !want_vtx && sink->n_chunks > 1 && n_rows >= (size_t)sink->n_chunks * 1024u;
    const bool listed = streaming && P.row_of_slot != nullptr;  // rows in insertion order: a chunk of rows is a scattered set of slots
    const int C = streaming ? sink->n_chunks : 1;
    const uint64_t face_cap = sink ? sink->face_capacity : 0;
    // Chunk sizes: small at both ends.  When the clip kernel is the slower side (one GPU per host link: 184 ms of compute
    // against 48 ms of copies at 10M cells) only the copy of the LAST chunk is exposed, so the chunks halve towards the end;
    // when the copies are the slower side (8 GPUs sharing the host's memory bandwidth: 28 ms of copies against 22 ms of
    // compute) nothing can be copied before the FIRST chunk is computed, so the chunks also grow from a small first one.
    std::vector<size_t> chunk_row(C + 1, 0);
    {
        std::vector<unsigned long long> w(C);
        unsigned long long sum = 0;
        for (int c = 0; c < C; ++c) sum += (w[c] = 1ull << std::min(std::min(C - 1 - c, c + 2), 4));
        unsigned long long acc = 0;
        for (int c = 0; c < C; ++c) {
            acc += w[c];
            chunk_row[c + 1] = (size_t)((unsigned long long)n_rows * acc / sum);
        }
        chunk_row[C] = n_rows;
    }
    auto row_begin = [&](int c) { return chunk_row[c]; };

    struct Events {
        std::vector<cudaEvent_t> ev;
        cudaStream_t copy_stream = nullptr;
        uint64_t* pinned = nullptr;
        ~Events() {
            for (cudaEvent_t e : ev) cudaEventDestroy(e);
            if (copy_stream) cudaStreamDestroy(copy_stream);
        }
    } E;
    E.ev.resize(3 * (size_t)C + 2);  // per chunk: clip begin, clip end, packed; + whole-call begin / end
    for (auto& e : E.ev) TESS_CUDA_CHECK(cudaEventCreate(&e));
    cudaEvent_t ev_begin = E.ev[3 * (size_t)C], ev_end = E.ev[3 * (size_t)C + 1];
    // page-locked words the per-chunk counts land in: allocated once per host thread (cudaFreeHost
    // synchronises the device, so it is kept out of the per-call path)
    struct PinnedWords {
        uint64_t* p = nullptr;
        ~PinnedWords() { if (p) cudaFreeHost(p); }
    };
    static thread_local PinnedWords pinned_words;
    constexpr size_t kPinnedWords = 4 * 256;  // n_chunks <= 256
    if (!pinned_words.p) TESS_CUDA_CHECK(cudaMallocHost(reinterpret_cast<void**>(&pinned_words.p), sizeof(uint64_t) * kPinnedWords));
    E.pinned = pinned_words.p;
    std::memset(E.pinned, 0, sizeof(uint64_t) * 4 * (size_t)C);
    uint32_t *work_list = nullptr, *flags = nullptr;
    uint64_t* pos = nullptr;
    if (streaming) {
        TESS_CUDA_CHECK(cudaStreamCreateWithFlags(&E.copy_stream, cudaStreamNonBlocking));
        if (listed) {
            work_list = tmp.get<uint32_t>(n_rows);
            flags = tmp.get<uint32_t>(n_rows + 1);
            pos = tmp.get<uint64_t>(n_rows + 1);
        }
        r->nbr = dmalloc<int64_t>(face_cap, s);  // the caller's capacity: the total is not known before the last chunk
        if (want_area) r->area = dmalloc<double>(face_cap, s);
    }
    TESS_CUDA_CHECK(cudaEventRecord(ev_begin, s));

    auto enqueue_clip = [&](int c) {
        const size_t r0 = row_begin(c), r1 = row_begin(c + 1);
        ClipParams Q = P;
        if (listed) {  // the chunk's cells in ascending slot order
            launch_chunk_flags(P.row_of_slot, d->own_slot_begin, n_rows, (uint32_t)r0, (uint32_t)r1, flags, s);
            launch_exclusive_scan_u32_to_u64(flags, pos, n_rows + 1, scan_tmp, scan_tmp_bytes(n_rows + 1), s);
            launch_chunk_scatter(flags, pos, d->own_slot_begin, n_rows, work_list + r0, s);
            Q.n_work = (uint32_t)(r1 - r0);
            Q.work_slots = work_list + r0;
        } else if (C > 1) {  // slab diagrams: rows follow the sorted order, a chunk is a run of slots
            Q.slot_begin = P.slot_begin + (uint32_t)r0;
            Q.n_work = (uint32_t)(r1 - r0);
        }
        TESS_CUDA_CHECK(cudaEventRecord(E.ev[3 * c], s));
        launch_clip(Q, main_tier, s);
        TESS_CUDA_CHECK(cudaEventRecord(E.ev[3 * c + 1], s));
        // failures so far: [0] all, [1] those that only ran out of search table
        TESS_CUDA_CHECK(cudaMemcpyAsync(reinterpret_cast<uint32_t*>(E.pinned + 4 * c), ctrl + 1, sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
        TESS_CUDA_CHECK(cudaMemcpyAsync(reinterpret_cast<uint32_t*>(E.pinned + 4 * c + 1), ctrl + 5, sizeof(uint32_t), cudaMemcpyDeviceToHost, s));
        TESS_CUDA_CHECK(cudaEventRecord(E.ev[3 * c + 2], s));
    };

    // ---- redo passes for `count` failed cells (sorted slots at failed_list):
    //   A: the same small-cell kernel with a wider search table (cells in voids only ran out of table);
    //   B: what still fails (more than 64 vertices / 40 faces at some point) goes to the medium configuration
    //      (256 vertices / 128 faces), with tables of doubled radius while cells run out of table;
    //   C: what outgrows that too goes to the large configuration (1024 vertices / 512 faces), likewise.
    struct Redo {
        uint32_t n_a = 0, n_b = 0, n_c = 0;
        const uint32_t *list_a = nullptr, *list_b = nullptr, *list_c = nullptr;
        int64_t *sa_nbr = nullptr, *md_nbr = nullptr, *lg_nbr = nullptr;
        double *sa_area = nullptr, *md_area = nullptr, *lg_area = nullptr;
        uint16_t *sa_flen = nullptr, *md_flen = nullptr, *lg_flen = nullptr;
    };
    const uint32_t mstride = clip_medium_fmax(), lstride = clip_large_fmax();
    uint32_t total_redo_a = 0, total_redo_b = 0, total_redo_c = 0;
    int final_R = 0;
    auto redo = [&](uint32_t* failed_list, uint32_t count, uint32_t table_only, Redo& ro) -> int {
        const int cpd_m1 = std::max(static_cast<int>(d->grid.cpd) - 1, 1);
        int R = std::min(std::max(3 * R0, 24), cpd_m1);
        uint32_t* list_b = nullptr;
        if (table_only == 0) {  // every failure is a table overflow of the small configuration: straight to pass B
            list_b = failed_list;
            ro.n_b = count;
        } else {
            ro.n_a = count;
            ro.list_a = failed_list;
            ro.sa_nbr = tmp.get<int64_t>((size_t)count * fstride);
            ro.sa_area = want_area ? tmp.get<double>((size_t)count * fstride) : nullptr;
            ro.sa_flen = want_vtx ? tmp.get<uint16_t>((size_t)count * fstride) : nullptr;
            list_b = tmp.get<uint32_t>(count);
            const ShellTable& t2 = d->table(R, s);
            ClipParams Q = P;
            Q.table = t2.dev.as<ShellEntry>();
            Q.table_len = t2.len;
            Q.table_full = t2.full ? 1u : 0u;
            Q.n_work = count;
            Q.work_slots = failed_list;
            Q.st_nbr = ro.sa_nbr; Q.st_area = ro.sa_area; Q.st_flen = ro.sa_flen; Q.fstride = fstride; Q.stage_by_work = 1;
            Q.failed_slots = list_b;
            Q.n_failed = ctrl + 2;
            Q.failed_cap = count;
            Q.mark_large = 1;
            TESS_CUDA_CHECK(cudaMemsetAsync(ctrl + 2, 0, sizeof(uint32_t), s));
            launch_clip(Q, CLIP_SMALL, s);
            TESS_CUDA_CHECK(cudaMemcpyAsync(&ro.n_b, ctrl + 2, sizeof(ro.n_b), cudaMemcpyDeviceToHost, s));
            TESS_CUDA_CHECK(cudaStreamSynchronize(s));
        }
        ro.list_b = list_b;
        // one tier over `n` listed cells: widen the table while some cell only ran out of table; the cells that
        // still fail afterwards (capacity) are left in failed_out
        unsigned long long* redo_counters = nullptr;
        auto run_tier = [&](int tier, const uint32_t* list, uint32_t n, uint32_t stride, int64_t* st_n, double* st_a, uint16_t* st_f, uint32_t* failed_out,
                            uint32_t* still_out) {
            uint32_t* nf = ctrl + (tier == CLIP_MEDIUM ? 8 : 9);
            uint32_t still = 0;
            for (int attempt = 0; attempt < 12; ++attempt) {
                const ShellTable& t2 = d->table(R, s);
                ClipParams Q = P;
                Q.table = t2.dev.as<ShellEntry>();
                Q.table_len = t2.len;
                Q.table_full = t2.full ? 1u : 0u;
                Q.n_work = n;
                Q.work_slots = list;
                Q.st_nbr = st_n; Q.st_area = st_a; Q.st_flen = st_f; Q.fstride = stride; Q.stage_by_work = 1;
                Q.counters = want_cnt ? redo_counters : nullptr;  // only the last attempt's counts are kept
                if (want_cnt) TESS_CUDA_CHECK(cudaMemsetAsync(redo_counters, 0, sizeof(unsigned long long) * CNT_N, s));
                Q.failed_slots = failed_out;
                Q.n_failed = nf;
                Q.failed_cap = n;
                Q.mark_large = 1;
                TESS_CUDA_CHECK(cudaMemsetAsync(nf, 0, sizeof(uint32_t), s));
                TESS_CUDA_CHECK(cudaMemsetAsync(nf + 4, 0, sizeof(uint32_t), s));
                launch_clip(Q, tier, s);
                uint32_t h[5] = {0, 0, 0, 0, 0};
                TESS_CUDA_CHECK(cudaMemcpyAsync(h, nf, sizeof(h), cudaMemcpyDeviceToHost, s));
                TESS_CUDA_CHECK(cudaStreamSynchronize(s));
                still = h[0];
                if (h[4] == 0 || t2.full) break;  // nobody ran out of table: what still fails needs larger tables of the mesh
                R = std::min(2 * R, cpd_m1);
            }
            if (want_cnt) {  // cells this tier finished (the kernel counts only those)
                unsigned long long h[CNT_N];
                TESS_CUDA_CHECK(cudaMemcpyAsync(h, redo_counters, sizeof(h), cudaMemcpyDeviceToHost, s));
                TESS_CUDA_CHECK(cudaStreamSynchronize(s));
                for (int i = 0; i < CNT_N; ++i) r->counters_redo[i] += h[i];
            }
            *still_out = still;
        };
        if (ro.n_b > 0) {
            if (ro.n_b > (1u << 22)) return fail(TESS_ERR_CAPACITY, "more than 2^22 cells need the medium-cell path");
            if (want_cnt) redo_counters = tmp.get<unsigned long long>(CNT_N);
            ro.md_nbr = tmp.get<int64_t>((size_t)ro.n_b * mstride);
            ro.md_area = want_area ? tmp.get<double>((size_t)ro.n_b * mstride) : nullptr;
            ro.md_flen = want_vtx ? tmp.get<uint16_t>((size_t)ro.n_b * mstride) : nullptr;
            uint32_t* list_c = tmp.get<uint32_t>(ro.n_b);
            run_tier(CLIP_MEDIUM, list_b, ro.n_b, mstride, ro.md_nbr, ro.md_area, ro.md_flen, list_c, &ro.n_c);
            ro.list_c = list_c;
        }
        if (ro.n_c > 0) {
            if (ro.n_c > (1u << 20)) return fail(TESS_ERR_CAPACITY, "more than 2^20 cells need the large-cell path");
            // (list_c, like every failed-cell list, is in completion order: rows are found through row_of_slot)
            ro.lg_nbr = tmp.get<int64_t>((size_t)ro.n_c * lstride);
            ro.lg_area = want_area ? tmp.get<double>((size_t)ro.n_c * lstride) : nullptr;
            ro.lg_flen = want_vtx ? tmp.get<uint16_t>((size_t)ro.n_c * lstride) : nullptr;
            uint32_t* failed_d = tmp.get<uint32_t>(ro.n_c);
            uint32_t still = 0;
            run_tier(CLIP_LARGE, ro.list_c, ro.n_c, lstride, ro.lg_nbr, ro.lg_area, ro.lg_flen, failed_d, &still);  // what still fails is reported in status
        }
        total_redo_a += ro.n_a;
        total_redo_b += ro.n_b;
        total_redo_c += ro.n_c;
        final_R = std::max(final_R, R);
        return TESS_OK;
    };

    uint32_t fail_prev = 0, table_only_prev = 0;
    uint64_t total = 0, copied = 0;
    uint32_t* flen_csr = nullptr;
    double ms_clip = 0.0;
    auto finish = [&](int c) -> int {
        const size_t r0 = row_begin(c), r1 = row_begin(c + 1);
        TESS_CUDA_CHECK(cudaEventSynchronize(E.ev[3 * c + 2]));
        {
            float t = 0;
            cudaEventElapsedTime(&t, E.ev[3 * c], E.ev[3 * c + 1]);
            ms_clip += t;
        }
        const uint32_t fail_now = *reinterpret_cast<uint32_t*>(E.pinned + 4 * c), table_only_now = *reinterpret_cast<uint32_t*>(E.pinned + 4 * c + 1);
        Redo ro;
        if (fail_now > fail_prev) {
            const int rc = redo(failed + fail_prev, fail_now - fail_prev, table_only_now - table_only_prev, ro);
            if (rc != TESS_OK) return rc;
        }
        fail_prev = fail_now;
        table_only_prev = table_only_now;
        // CSR offsets: an exclusive scan only looks back, so the scan of the whole array is final up to r1
        // (rows of the next chunk may already hold their counts)
        launch_exclusive_scan_u32_to_u64(r->nfaces, r->offsets, n_rows + 1, scan_tmp, scan_tmp_bytes(n_rows + 1), s);
        TESS_CUDA_CHECK(cudaMemcpyAsync(E.pinned + 4 * c + 2, r->offsets + r1, sizeof(uint64_t), cudaMemcpyDeviceToHost, s));
        TESS_CUDA_CHECK(cudaStreamSynchronize(s));
        total = E.pinned[4 * c + 2];
        if (sink && total > face_cap) return fail(TESS_ERR_CAPACITY, "tess_compute_all_to_host: the cells have more faces than face_capacity");
        if (!streaming) {
            r->nbr = dmalloc<int64_t>(total, s);
            if (want_area) r->area = dmalloc<double>(total, s);
            flen_csr = want_vtx ? tmp.get<uint32_t>(total + 1) : nullptr;
            if (want_vtx) TESS_CUDA_CHECK(cudaMemsetAsync(flen_csr + total, 0, sizeof(uint32_t), s));
        }
        launch_compact_faces(r->status + r0, r->offsets + r0, st_nbr + r0 * fstride, st_area ? st_area + r0 * fstride : nullptr, st_flen ? st_flen + r0 * fstride : nullptr,
                             fstride, r1 - r0, r->nbr, r->area, flen_csr, s);
        if (ro.n_a)  // pass A rows (rows redone again by pass B are overwritten right after)
            launch_compact_redo(ro.list_a, P.row_of_slot, P.row_base, r->nfaces, r->offsets, ro.sa_nbr, ro.sa_area, ro.sa_flen, fstride, ro.n_a, r->nbr, r->area, flen_csr, s);
        if (ro.n_b)  // (likewise for rows redone by pass C)
            launch_compact_redo(ro.list_b, P.row_of_slot, P.row_base, r->nfaces, r->offsets, ro.md_nbr, ro.md_area, ro.md_flen, mstride, ro.n_b, r->nbr, r->area, flen_csr, s);
        if (ro.n_c)
            launch_compact_redo(ro.list_c, P.row_of_slot, P.row_base, r->nfaces, r->offsets, ro.lg_nbr, ro.lg_area, ro.lg_flen, lstride, ro.n_c, r->nbr, r->area, flen_csr, s);
        launch_clear_status_bits(r->status + r0, r1 - r0, ST_LARGE_PATH, s);  // internal marker of the redo passes
        if (streaming) {
            // the packed chunk goes to the host while the next chunks are clipped
            cudaEvent_t packed = E.ev[3 * c];  // (its first use, the clip-begin time stamp, has been read)
            TESS_CUDA_CHECK(cudaEventRecord(packed, s));
            TESS_CUDA_CHECK(cudaStreamWaitEvent(E.copy_stream, packed, 0));
            cudaStream_t cs = E.copy_stream;
            if (sink->volumes) TESS_CUDA_CHECK(cudaMemcpyAsync(sink->volumes + r0, r->vol + r0, sizeof(double) * (r1 - r0), cudaMemcpyDeviceToHost, cs));
            if (sink->status) TESS_CUDA_CHECK(cudaMemcpyAsync(sink->status + r0, r->status + r0, sizeof(uint32_t) * (r1 - r0), cudaMemcpyDeviceToHost, cs));
            if (sink->face_offsets)
                TESS_CUDA_CHECK(cudaMemcpyAsync(sink->face_offsets + r0, r->offsets + r0, sizeof(uint64_t) * (r1 - r0 + (c == C - 1 ? 1 : 0)), cudaMemcpyDeviceToHost, cs));
            if (sink->neighbors && total > copied)
                TESS_CUDA_CHECK(cudaMemcpyAsync(sink->neighbors + copied, r->nbr + copied, sizeof(int64_t) * (total - copied), cudaMemcpyDeviceToHost, cs));
            if (sink->areas && want_area && total > copied)
                TESS_CUDA_CHECK(cudaMemcpyAsync(sink->areas + copied, r->area + copied, sizeof(double) * (total - copied), cudaMemcpyDeviceToHost, cs));
            copied = total;
        }
        return TESS_OK;
    };

    for (int c = 0; c < C; ++c) {
        enqueue_clip(c);
        if (c > 0) {
            const int rc = finish(c - 1);
            if (rc != TESS_OK) return rc;
        }
    }
    {
        const int rc = finish(C - 1);
        if (rc != TESS_OK) return rc;
    }
    r->n_faces = total;
    r->tier_stats[1] = total_redo_a; r->tier_stats[2] = total_redo_b; r->tier_stats[3] = total_redo_c;
    if (std::getenv("TESS_TRACE") && (total_redo_a || total_redo_b))
        std::fprintf(stderr, "[tess trace] redo: pass A (wider table) re-ran %u cells, pass B (medium cells) %u, pass C (large cells) %u, final R=%d\n", total_redo_a, total_redo_b,
                     total_redo_c, final_R);
    tr.mark("clip + redo + pack");

    if (want_vtx) {
        unsigned long long used[2] = {0, 0};
        TESS_CUDA_CHECK(cudaMemcpyAsync(used, g_cursor, sizeof(used), cudaMemcpyDeviceToHost, s));
        launch_exclusive_scan_u32_to_u64(r->nverts, r->voffsets, n_rows + 1, scan_tmp, scan_tmp_bytes(n_rows + 1), s);
        uint64_t tv = 0, tl = 0;
        TESS_CUDA_CHECK(cudaMemcpyAsync(&tv, r->voffsets + n_rows, sizeof(tv), cudaMemcpyDeviceToHost, s));
        r->fv_offsets = dmalloc<uint64_t>(total + 1, s);
        void* scan_tmp2 = tmp.get<char>(scan_tmp_bytes(total + 1));
        launch_exclusive_scan_u32_to_u64(flen_csr, r->fv_offsets, total + 1, scan_tmp2, scan_tmp_bytes(total + 1), s);
        TESS_CUDA_CHECK(cudaMemcpyAsync(&tl, r->fv_offsets + total, sizeof(tl), cudaMemcpyDeviceToHost, s));
        TESS_CUDA_CHECK(cudaStreamSynchronize(s));
        if (used[0] > gv_cap || used[1] > gl_cap)
            return fail(TESS_ERR_CAPACITY, "TESS_OUT_VERTICES: the cells have more vertices / face-loop entries than the geometry pools hold (48 / 144 per cell on average)");
        r->n_vertices = tv;
        r->n_loop_entries = tl;
        r->vtx = dmalloc<double>(3 * tv, s);
        r->fv_idx = dmalloc<uint32_t>(tl, s);
        launch_gather_vertices(r->nverts, vbase, r->voffsets, gv_xyz, n_rows, r->vtx, s);
        launch_gather_loops(nloops, lbase, r->offsets, r->fv_offsets, gl_idx, n_rows, r->fv_idx, s);
    }
    if (sink && !streaming) {  // small inputs: one copy of the finished arrays
        const int rc = tess_result_download(r.get(), sink->volumes, sink->face_offsets, sink->neighbors, want_area ? sink->areas : nullptr, sink->status, s);
        if (rc != TESS_OK) return rc;
    }
    TESS_CUDA_CHECK(cudaEventRecord(ev_end, s));
    TESS_CUDA_CHECK(cudaStreamSynchronize(s));
    if (streaming) TESS_CUDA_CHECK(cudaStreamSynchronize(E.copy_stream));
    tr.mark("outputs");
    {
        float t = 0;
        cudaEventElapsedTime(&t, ev_begin, ev_end);
        r->ms_clip = ms_clip;
        r->ms_redo = 0.0;  // folded into the rest: redo passes, scans, packing (and, when streaming, waiting for copies)
        r->ms_outputs = std::max(0.0, (double)t - ms_clip);
        r->ms_total = t;
    }
    *out = r.release();
    return TESS_OK;
}

}  // namespace

extern "C" {

int tess_compute_all(const tess_diagram* d, const tess_opts* opts, tess_result** out) {
    if (!d || !out) return fail(TESS_ERR_INVALID, "NULL argument");
    *out = nullptr;
    if (!d->initialized) return fail(TESS_ERR_STATE, "diagram not initialized");
    TESS_TRY
    return compute_impl(d, opts, nullptr, 0, out);
    TESS_CATCH
}

int tess_compute_all_to_host(const tess_diagram* d, const tess_opts* opts, int n_chunks, double* volumes, uint64_t* face_offsets, int64_t* neighbors, double* areas,
                             uint32_t* status, uint64_t cell_capacity, uint64_t face_capacity, tess_result** out) {
    if (!d || !out) return fail(TESS_ERR_INVALID, "NULL argument");
    *out = nullptr;
    if (!d->initialized) return fail(TESS_ERR_STATE, "diagram not initialized");
    if (opts && (opts->outputs & TESS_OUT_VERTICES)) return fail(TESS_ERR_INVALID, "tess_compute_all_to_host does not stream vertex geometry; use tess_compute_all");
    if (n_chunks < 0 || n_chunks > 256) return fail(TESS_ERR_INVALID, "n_chunks must be in [0, 256]");
    if (static_cast<uint64_t>(d->own_slot_end - d->own_slot_begin) > cell_capacity)
        return fail(TESS_ERR_CAPACITY, "tess_compute_all_to_host: the diagram has more cells than cell_capacity");
    TESS_TRY
    HostSink sink{volumes, face_offsets, neighbors, areas, status, face_capacity, n_chunks == 0 ? 8 : n_chunks};
    return compute_impl(d, opts, nullptr, 0, out, &sink);
    TESS_CATCH
}

int tess_compute_at_points(const tess_diagram* d, const double* xyz, size_t m, const tess_opts* opts, tess_result** out) {
    if (!d || !out || (!xyz && m)) return fail(TESS_ERR_INVALID, "NULL argument");
    *out = nullptr;
    if (!d->initialized) return fail(TESS_ERR_STATE, "diagram not initialized");
    if (d->slab) return fail(TESS_ERR_UNSUPPORTED, "tess_compute_at_points needs a whole-domain diagram");
    static const double dummy[3] = {0, 0, 0};
    TESS_TRY
    return compute_impl(d, opts, m ? xyz : dummy, m, out);
    TESS_CATCH
}

void tess_result_free(tess_result* r) { delete r; }

int tess_result_n_cells(const tess_result* r, uint64_t* n_cells, uint64_t* n_faces) {
    if (!r) return fail(TESS_ERR_INVALID, "NULL result");
    if (n_cells) *n_cells = r->n_cells;
    if (n_faces) *n_faces = r->n_faces;
    return TESS_OK;
}

}  // extern "C"

namespace {
template <class T>
int host_view(tess_result* r, const T* dev, size_t count, std::vector<T>& host, bool& have, const T** out) {
    if (!r || !out) return fail(TESS_ERR_INVALID, "NULL argument");
    try {
        if (!have) {
            cudaSetDevice(r->device);
            host.resize(count);
            if (count) {
                if (!dev) return fail(TESS_ERR_STATE, "this output was not requested in tess_opts.outputs");
                TESS_CUDA_CHECK(cudaMemcpy(host.data(), dev, sizeof(T) * count, cudaMemcpyDeviceToHost));
            }
            have = true;
        }
        *out = host.data();
        return TESS_OK;
    } catch (const std::bad_alloc&) {
        return fail(TESS_ERR_NOMEM, "out of host memory");
    } catch (const std::exception& e) {
        return fail(TESS_ERR_CUDA, e.what());
    }
}
}  // namespace

extern "C" {

int tess_result_volumes(tess_result* r, const double** out) { return host_view(r, r ? r->vol : nullptr, r ? r->n_cells : 0, r->h_vol, r->have_vol, out); }
int tess_result_face_offsets(tess_result* r, const uint64_t** out) { return host_view(r, r ? r->offsets : nullptr, r ? r->n_cells + 1 : 0, r->h_offsets, r->have_offsets, out); }
int tess_result_neighbors(tess_result* r, const int64_t** out) { return host_view(r, r ? r->nbr : nullptr, r ? r->n_faces : 0, r->h_nbr, r->have_nbr, out); }
int tess_result_areas(tess_result* r, const double** out) { return host_view(r, r ? r->area : nullptr, r ? r->n_faces : 0, r->h_area, r->have_area, out); }
int tess_result_cell_ids(tess_result* r, const int64_t** out) { return host_view(r, r ? r->cell_id : nullptr, r ? r->n_cells : 0, r->h_cell_id, r->have_ids, out); }
int tess_result_vertex_offsets(tess_result* r, const uint64_t** out) { return host_view(r, r ? r->voffsets : nullptr, r ? r->n_cells + 1 : 0, r->h_voffsets, r->have_voff, out); }
int tess_result_vertices(tess_result* r, const double** out) { return host_view(r, r ? r->vtx : nullptr, r ? 3 * r->n_vertices : 0, r->h_vtx, r->have_vtx, out); }
int tess_result_face_vertex_offsets(tess_result* r, const uint64_t** out) { return host_view(r, r ? r->fv_offsets : nullptr, r ? r->n_faces + 1 : 0, r->h_fv_offsets, r->have_fvo, out); }
int tess_result_face_vertex_indices(tess_result* r, const uint32_t** out) { return host_view(r, r ? r->fv_idx : nullptr, r ? r->n_loop_entries : 0, r->h_fv_idx, r->have_fvi, out); }

int tess_result_status(tess_result* r, const uint32_t** out) {
    const bool first = r && !r->have_status;
    const int rc = host_view(r, r ? r->status : nullptr, r ? r->n_cells : 0, r->h_status, r->have_status, out);
    if (rc == TESS_OK && first)
        for (auto& v : r->h_status) v &= ~ST_LARGE_PATH;  // internal bit
    return rc;
}

int tess_result_counters(tess_result* r, uint64_t counters[8]) {
    if (!r || !counters) return fail(TESS_ERR_INVALID, "NULL argument");
    TESS_TRY
    cudaSetDevice(r->device);
    unsigned long long h[CNT_N];
    TESS_CUDA_CHECK(cudaMemcpy(h, r->counters, sizeof(h), cudaMemcpyDeviceToHost));
    for (int i = 0; i < 8; ++i) counters[i] = h[i] + r->counters_redo[i];
    return TESS_OK;
    TESS_CATCH
}

int tess_result_volume_sum(tess_result* r, double* out) {
    if (!r || !out) return fail(TESS_ERR_INVALID, "NULL argument");
    TESS_TRY
    cudaSetDevice(r->device);
    cudaStream_t s = r->stream;
    double* dsum = dmalloc<double>(1, s);
    launch_volume_sum(r->vol, r->n_cells, dsum, s);
    TESS_CUDA_CHECK(cudaMemcpyAsync(out, dsum, sizeof(double), cudaMemcpyDeviceToHost, s));
    TESS_CUDA_CHECK(cudaStreamSynchronize(s));
    cudaFreeAsync(dsum, s);
    return TESS_OK;
    TESS_CATCH
}

int tess_result_download(const tess_result* r, double* volumes, uint64_t* face_offsets, int64_t* neighbors, double* areas, uint32_t* status, void* stream) {
    if (!r) return fail(TESS_ERR_INVALID, "NULL result");
    TESS_TRY
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    TESS_CUDA_CHECK(cudaSetDevice(r->device));
    if (volumes && r->n_cells) TESS_CUDA_CHECK(cudaMemcpyAsync(volumes, r->vol, sizeof(double) * r->n_cells, cudaMemcpyDeviceToHost, s));
    if (face_offsets) TESS_CUDA_CHECK(cudaMemcpyAsync(face_offsets, r->offsets, sizeof(uint64_t) * (r->n_cells + 1), cudaMemcpyDeviceToHost, s));
    if (neighbors && r->n_faces) TESS_CUDA_CHECK(cudaMemcpyAsync(neighbors, r->nbr, sizeof(int64_t) * r->n_faces, cudaMemcpyDeviceToHost, s));
    if (areas && r->n_faces) {
        if (!r->area) return fail(TESS_ERR_STATE, "areas were not requested in tess_opts.outputs");
        TESS_CUDA_CHECK(cudaMemcpyAsync(areas, r->area, sizeof(double) * r->n_faces, cudaMemcpyDeviceToHost, s));
    }
    if (status && r->n_cells) TESS_CUDA_CHECK(cudaMemcpyAsync(status, r->status, sizeof(uint32_t) * r->n_cells, cudaMemcpyDeviceToHost, s));
    return TESS_OK;
    TESS_CATCH
}

uint64_t tess_kernel_launch_count(void) { return launch_count(); }

int tess_set_main_tier(int tier) {
    if (tier != -1 && tier != CLIP_SMALL && tier != CLIP_SMALL_FAST && tier != CLIP_THREAD) return fail(TESS_ERR_INVALID, "tess_set_main_tier: -1 (default), 0 (warp per cell), 3 (warp per cell, no serial walk) or 4 (thread per cell)");
    g_main_tier.store(tier == -1 ? -2 : tier, std::memory_order_relaxed);
    return TESS_OK;
}

int tess_result_tier_stats(const tess_result* r, uint64_t stats[4]) {
    if (!r || !stats) return fail(TESS_ERR_INVALID, "NULL argument");
    for (int i = 0; i < 4; ++i) stats[i] = r->tier_stats[i];
    return TESS_OK;
}

int tess_result_timings(const tess_result* r, double ms[4]) {
    if (!r || !ms) return fail(TESS_ERR_INVALID, "NULL argument");
    ms[0] = r->ms_clip; ms[1] = r->ms_redo; ms[2] = r->ms_outputs; ms[3] = r->ms_total;
    return TESS_OK;
}

int tess_diagram_timings(const tess_diagram* d, double ms[1]) {
    if (!d || !ms) return fail(TESS_ERR_INVALID, "NULL argument");
    if (!d->initialized || !d->ev_bin0) return fail(TESS_ERR_STATE, "diagram not initialized");
    float t = 0;
    if (cudaEventSynchronize(d->ev_bin1) != cudaSuccess || cudaEventElapsedTime(&t, d->ev_bin0, d->ev_bin1) != cudaSuccess) return fail(TESS_ERR_CUDA, "event timing failed");
    ms[0] = t;
    return TESS_OK;
}

int tess_measure_fp64_peak(int device, double* tflops) {
    if (!tflops) return fail(TESS_ERR_INVALID, "NULL argument");
    TESS_TRY
    TESS_CUDA_CHECK(cudaSetDevice(device));
    *tflops = measure_fp64_peak_tflops();
    return TESS_OK;
    TESS_CATCH
}

int tess_result_device_views(const tess_result* r, const double** volumes, const uint64_t** face_offsets, const int64_t** neighbors, const double** areas,
                             const uint32_t** status, const int64_t** cell_ids) {
    if (!r) return fail(TESS_ERR_INVALID, "NULL result");
    if (volumes) *volumes = r->vol;
    if (face_offsets) *face_offsets = r->offsets;
    if (neighbors) *neighbors = r->nbr;
    if (areas) *areas = r->area;
    if (status) *status = r->status;
    if (cell_ids) *cell_ids = r->cell_id;
    return TESS_OK;
}

// ------------------------------------------------------------------------------------------------
// radius queries
// ------------------------------------------------------------------------------------------------
}  // extern "C"

struct tess_query {
    std::vector<uint64_t> offsets;
    std::vector<int64_t> indices;
    std::vector<uint32_t> status;
};

// ExpandingSearch (celery.rs:865-880) for m positions: the positions and each one's current_search_index
struct tess_search {
    const tess_diagram* d = nullptr;
    std::vector<double> xyz;
    std::vector<uint64_t> cursor;
};

extern "C" {

// One query pass pair (count, scan, fill) for m host positions.  mode 4 (ExpandingSearch::expand): cursors in / out (host, m
// entries) and cells_to_add; the search table is widened until no walk runs off the end of a truncated one.
static int run_query(const tess_diagram* d, const double* xyz, size_t m, double radius, int mode, int64_t target_group, const uint64_t* cursor_in,
                     uint64_t* cursor_out, uint64_t cells_to_add, cudaStream_t s, tess_query** out) {
    TESS_CUDA_CHECK(cudaSetDevice(d->device));
    std::unique_ptr<tess_query> q(new tess_query());
    q->offsets.assign(m + 1, 0);
    q->status.assign(m, 0);
    if (m == 0) {
        *out = q.release();
        return TESS_OK;
    }
    Scratch tmp(s);
    double* qx = tmp.get<double>(3 * m);
    uint32_t* counts = tmp.get<uint32_t>(m + 1);
    uint32_t* flags = tmp.get<uint32_t>(m);
    uint64_t* offs = tmp.get<uint64_t>(m + 1);
    void* scan_tmp = tmp.get<char>(scan_tmp_bytes(m + 1));
    uint64_t* cur_in = mode == 4 ? tmp.get<uint64_t>(m) : nullptr;
    uint64_t* cur_out = mode == 4 ? tmp.get<uint64_t>(m) : nullptr;
    TESS_CUDA_CHECK(cudaMemcpyAsync(qx, xyz, sizeof(double) * 3 * m, cudaMemcpyHostToDevice, s));
    if (mode == 4) TESS_CUDA_CHECK(cudaMemcpyAsync(cur_in, cursor_in, sizeof(uint64_t) * m, cudaMemcpyHostToDevice, s));
    QueryParams Q{};
    Q.sorted = d->sorted.as<Particle>();
    Q.delim = d->delim.as<uint32_t>();
    Q.groups_sorted = d->has_groups ? d->groups_sorted.as<uint64_t>() : nullptr;
    Q.grid = d->grid;
    Q.xyz = qx;
    Q.n_query = m;
    Q.radius = radius;
    Q.mode = mode;
    Q.target_group = target_group;
    Q.counts = counts;
    Q.flags = flags;
    Q.cursor_in = cur_in;
    Q.cursor_out = cur_out;
    Q.cells_to_add = cells_to_add;
    int R = 0;
    if (mode == TESS_QUERY_NEIGHBOR_CLOUD || mode == 4) {
        // the table must hold every entry with key <= radius: (R*size)^2 > radius on every axis
        const double smin = std::min(d->grid.sx, std::min(d->grid.sy, d->grid.sz));
        R = static_cast<int>(d->grid.cpd);  // full table
        if (radius >= 0 && smin > 0) {
            const double need = std::sqrt(radius) / smin + 2.0;
            if (need < static_cast<double>(d->grid.cpd)) R = std::max(1, static_cast<int>(need));
        }
        if (mode == 4) R = std::min(R, std::max(kDefaultTableRadius, 1));  // a few cells per call is the common use: start small, widen on demand
    }
    for (;;) {
        if (R > 0) {
            const ShellTable& t = d->table(R, s);
            Q.table = t.dev.as<ShellEntry>();
            Q.table_len = t.len;
            Q.table_full = t.full ? 1u : 0u;
        }
        TESS_CUDA_CHECK(cudaMemsetAsync(counts, 0, sizeof(uint32_t) * (m + 1), s));
        launch_radius_query(Q, /*fill=*/false, s);
        launch_exclusive_scan_u32_to_u64(counts, offs, m + 1, scan_tmp, scan_tmp_bytes(m + 1), s);
        TESS_CUDA_CHECK(cudaMemcpyAsync(q->offsets.data(), offs, sizeof(uint64_t) * (m + 1), cudaMemcpyDeviceToHost, s));
        TESS_CUDA_CHECK(cudaMemcpyAsync(q->status.data(), flags, sizeof(uint32_t) * m, cudaMemcpyDeviceToHost, s));
        TESS_CUDA_CHECK(cudaStreamSynchronize(s));
        if (mode != 4 || Q.table_full) break;
        bool ran_off = false;
        for (size_t i = 0; i < m; ++i) ran_off |= (q->status[i] & ST_TABLE_EXHAUSTED) != 0;
        if (!ran_off || R >= static_cast<int>(d->grid.cpd)) break;  // (a table of half-width cpd is the full table)
        R = std::min(2 * R, static_cast<int>(d->grid.cpd));
    }
    const uint64_t total = q->offsets[m];
    q->indices.resize(total);
    if (total) {
        int64_t* idx = tmp.get<int64_t>(total);
        Q.offsets = offs;
        Q.indices = idx;
        launch_radius_query(Q, /*fill=*/true, s);
        TESS_CUDA_CHECK(cudaMemcpyAsync(q->indices.data(), idx, sizeof(int64_t) * total, cudaMemcpyDeviceToHost, s));
    }
    if (mode == 4) TESS_CUDA_CHECK(cudaMemcpyAsync(cursor_out, cur_out, sizeof(uint64_t) * m, cudaMemcpyDeviceToHost, s));
    TESS_CUDA_CHECK(cudaStreamSynchronize(s));
    *out = q.release();
    return TESS_OK;
}

int tess_find_neighbors(const tess_diagram* d, const double* xyz, size_t m, double radius, int mode, int64_t target_group, void* stream, tess_query** out) {
    if (!d || !out || (!xyz && m)) return fail(TESS_ERR_INVALID, "NULL argument");
    *out = nullptr;
    if (!d->initialized) return fail(TESS_ERR_STATE, "diagram not initialized");
    if (d->slab) return fail(TESS_ERR_UNSUPPORTED, "radius queries need a whole-domain diagram");
    if (mode < 0 || mode > 2) return fail(TESS_ERR_INVALID, "bad query mode");
    TESS_TRY
    return run_query(d, xyz, m, radius, mode, target_group, nullptr, nullptr, 0, static_cast<cudaStream_t>(stream), out);
    TESS_CATCH
}

int tess_find_cells_in_radius(const tess_diagram* d, const double* xyz, size_t m, double radius, void* stream, tess_query** out) {
    if (!d || !out || (!xyz && m)) return fail(TESS_ERR_INVALID, "NULL argument");
    *out = nullptr;
    if (!d->initialized) return fail(TESS_ERR_STATE, "diagram not initialized");
    if (d->slab) return fail(TESS_ERR_UNSUPPORTED, "radius queries need a whole-domain diagram");
    TESS_TRY
    return run_query(d, xyz, m, radius, 3, -1, nullptr, nullptr, 0, static_cast<cudaStream_t>(stream), out);
    TESS_CATCH
}

int tess_search_create(const tess_diagram* d, const double* xyz, size_t m, tess_search** out) {
    if (!d || !out || (!xyz && m)) return fail(TESS_ERR_INVALID, "NULL argument");
    *out = nullptr;
    if (!d->initialized) return fail(TESS_ERR_STATE, "diagram not initialized");
    if (d->slab) return fail(TESS_ERR_UNSUPPORTED, "radius queries need a whole-domain diagram");
    TESS_TRY
    std::unique_ptr<tess_search> sr(new tess_search());
    sr->d = d;
    sr->xyz.assign(xyz, xyz + 3 * m);
    sr->cursor.assign(m, 0);  // current_search_index: 0 (celery.rs:895)
    *out = sr.release();
    return TESS_OK;
    TESS_CATCH
}

int tess_search_expand(tess_search* sr, double max_radius, uint64_t cells_to_add, void* stream, tess_query** out) {
    if (!sr || !out) return fail(TESS_ERR_INVALID, "NULL argument");
    *out = nullptr;
    if (!sr->d->initialized) return fail(TESS_ERR_STATE, "diagram not initialized");
    TESS_TRY
    std::vector<uint64_t> next(sr->cursor.size());
    const int rc = run_query(sr->d, sr->xyz.data(), sr->cursor.size(), max_radius, 4, -1, sr->cursor.data(), next.data(), cells_to_add, static_cast<cudaStream_t>(stream), out);
    if (rc == TESS_OK && !sr->cursor.empty()) sr->cursor.swap(next);
    return rc;
    TESS_CATCH
}

int tess_search_cursor(const tess_search* sr, const uint64_t** out) {
    if (!sr || !out) return fail(TESS_ERR_INVALID, "NULL argument");
    *out = sr->cursor.data();
    return TESS_OK;
}

void tess_search_free(tess_search* sr) { delete sr; }

void tess_query_free(tess_query* q) { delete q; }
int tess_query_offsets(tess_query* q, const uint64_t** out) {
    if (!q || !out) return fail(TESS_ERR_INVALID, "NULL argument");
    *out = q->offsets.data();
    return TESS_OK;
}
int tess_query_indices(tess_query* q, const int64_t** out) {
    if (!q || !out) return fail(TESS_ERR_INVALID, "NULL argument");
    *out = q->indices.data();
    return TESS_OK;
}
int tess_query_status(tess_query* q, const uint32_t** out) {
    if (!q || !out) return fail(TESS_ERR_INVALID, "NULL argument");
    *out = q->status.data();
    return TESS_OK;
}

// ------------------------------------------------------------------------------------------------
// slab helpers
// ------------------------------------------------------------------------------------------------
int tess_bounds(const double* xyz_dev, size_t n, double* bounds_dev, void* stream) {
    if (!xyz_dev || !bounds_dev || !n) return fail(TESS_ERR_INVALID, "NULL/empty argument");
    TESS_TRY
    launch_bounds(xyz_dev, n, bounds_dev, static_cast<cudaStream_t>(stream));
    return TESS_OK;
    TESS_CATCH
}

int tess_plane_histogram(const double* xyz_dev, size_t n, const double bounds[6], uint64_t n_global, uint64_t* counts_dev, void* stream) {
    if (!bounds || !counts_dev || (!xyz_dev && n)) return fail(TESS_ERR_INVALID, "NULL argument");
    TESS_TRY
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const GridSpec g = spec_from_bounds(bounds, n_global);
    TESS_CUDA_CHECK(cudaMemsetAsync(counts_dev, 0, sizeof(uint64_t) * g.cpd, s));
    launch_plane_histogram(xyz_dev, n, g, reinterpret_cast<unsigned long long*>(counts_dev), s);
    return TESS_OK;
    TESS_CATCH
}

int tess_row_histogram(const double* xyz_dev, size_t n, const double bounds[6], uint64_t n_global, uint64_t* counts_dev, void* stream) {
    if (!bounds || !counts_dev || (!xyz_dev && n)) return fail(TESS_ERR_INVALID, "NULL argument");
    TESS_TRY
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const GridSpec g = spec_from_bounds(bounds, n_global);
    TESS_CUDA_CHECK(cudaMemsetAsync(counts_dev, 0, sizeof(uint64_t) * g.cpd * g.cpd, s));
    launch_row_histogram(xyz_dev, n, g, reinterpret_cast<unsigned long long*>(counts_dev), s);
    return TESS_OK;
    TESS_CATCH
}

int tess_pack_for_slabs(const double* xyz_dev, const int64_t* ids_dev, int64_t id_base, size_t n, const double bounds[6], uint64_t n_global, int n_ranks,
                        const uint32_t* plane_lo, const uint32_t* plane_hi, uint64_t* send_counts_dev, double* out_xyz_dev, int64_t* out_ids_dev, size_t cap, void* stream) {
    if (!bounds || !plane_lo || !plane_hi || !send_counts_dev || n_ranks <= 0 || n_ranks > 1024) return fail(TESS_ERR_INVALID, "bad argument");
    TESS_TRY
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const GridSpec g = spec_from_bounds(bounds, n_global);
    Scratch tmp(s);
    uint32_t* lo = tmp.get<uint32_t>(n_ranks);
    uint32_t* hi = tmp.get<uint32_t>(n_ranks);
    unsigned long long* offs = tmp.get<unsigned long long>(n_ranks);
    unsigned long long* cursors = tmp.get<unsigned long long>(n_ranks);
    TESS_CUDA_CHECK(cudaMemcpyAsync(lo, plane_lo, sizeof(uint32_t) * n_ranks, cudaMemcpyHostToDevice, s));
    TESS_CUDA_CHECK(cudaMemcpyAsync(hi, plane_hi, sizeof(uint32_t) * n_ranks, cudaMemcpyHostToDevice, s));
    TESS_CUDA_CHECK(cudaMemsetAsync(send_counts_dev, 0, sizeof(uint64_t) * n_ranks, s));
    launch_pack_count(xyz_dev, n, g, n_ranks, lo, hi, reinterpret_cast<unsigned long long*>(send_counts_dev), s);
    std::vector<unsigned long long> hc(n_ranks), ho(n_ranks);
    TESS_CUDA_CHECK(cudaMemcpyAsync(hc.data(), send_counts_dev, sizeof(uint64_t) * n_ranks, cudaMemcpyDeviceToHost, s));
    TESS_CUDA_CHECK(cudaStreamSynchronize(s));
    unsigned long long acc = 0;
    for (int r = 0; r < n_ranks; ++r) { ho[r] = acc; acc += hc[r]; }
    if (acc > cap || !out_xyz_dev || !out_ids_dev) return fail(TESS_ERR_NOMEM, "pack buffer too small (send counts are valid)");
    TESS_CUDA_CHECK(cudaMemcpyAsync(offs, ho.data(), sizeof(unsigned long long) * n_ranks, cudaMemcpyHostToDevice, s));
    TESS_CUDA_CHECK(cudaMemsetAsync(cursors, 0, sizeof(unsigned long long) * n_ranks, s));
    launch_pack_scatter(xyz_dev, ids_dev, id_base, n, g, n_ranks, lo, hi, offs, nullptr, cursors, out_xyz_dev, out_ids_dev, nullptr, s);
    TESS_CUDA_CHECK(cudaStreamSynchronize(s));
    return TESS_OK;
    TESS_CATCH
}

int tess_pack_records(const double* xyz_dev, const int64_t* ids_dev, int64_t id_base, size_t n, const double bounds[6], uint64_t n_global, int n_ranks,
                      const uint32_t* plane_lo, const uint32_t* plane_hi, const uint64_t* planned_counts, uint64_t* counts_host, uint64_t* counts_dev,
                      double* out_rec_dev, size_t cap, void* stream) {
    if (!bounds || !plane_lo || !plane_hi || !counts_dev || n_ranks <= 0 || n_ranks > 1024) return fail(TESS_ERR_INVALID, "bad argument");
    TESS_TRY
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const GridSpec g = spec_from_bounds(bounds, n_global);
    Scratch tmp(s);
    uint32_t* lo = tmp.get<uint32_t>(n_ranks);
    uint32_t* hi = tmp.get<uint32_t>(n_ranks);
    unsigned long long* offs = tmp.get<unsigned long long>(2 * n_ranks);  // offsets, then limits
    TESS_CUDA_CHECK(cudaMemcpyAsync(lo, plane_lo, sizeof(uint32_t) * n_ranks, cudaMemcpyHostToDevice, s));
    TESS_CUDA_CHECK(cudaMemcpyAsync(hi, plane_hi, sizeof(uint32_t) * n_ranks, cudaMemcpyHostToDevice, s));
    TESS_CUDA_CHECK(cudaMemsetAsync(counts_dev, 0, sizeof(uint64_t) * n_ranks, s));
    std::vector<unsigned long long> hc(n_ranks), ho(2 * n_ranks);
    if (planned_counts) {
        // the caller knows the counts (an unchanged particle set): no counting pass, no host round trip.  The scatter's
        // cursors land in counts_dev; the caller compares them with the plan.
        for (int r = 0; r < n_ranks; ++r) hc[r] = planned_counts[r];
    } else {
        launch_pack_count(xyz_dev, n, g, n_ranks, lo, hi, reinterpret_cast<unsigned long long*>(counts_dev), s);
        TESS_CUDA_CHECK(cudaMemcpyAsync(hc.data(), counts_dev, sizeof(uint64_t) * n_ranks, cudaMemcpyDeviceToHost, s));
        TESS_CUDA_CHECK(cudaStreamSynchronize(s));
        TESS_CUDA_CHECK(cudaMemsetAsync(counts_dev, 0, sizeof(uint64_t) * n_ranks, s));
    }
    unsigned long long acc = 0;
    for (int r = 0; r < n_ranks; ++r) { ho[r] = acc; ho[n_ranks + r] = hc[r]; acc += hc[r]; }
    if (counts_host) for (int r = 0; r < n_ranks; ++r) counts_host[r] = hc[r];
    if (acc > cap || !out_rec_dev) return fail(TESS_ERR_NOMEM, "pack buffer too small (counts_host is valid)");
    TESS_CUDA_CHECK(cudaMemcpyAsync(offs, ho.data(), sizeof(unsigned long long) * 2 * n_ranks, cudaMemcpyHostToDevice, s));
    launch_pack_scatter(xyz_dev, ids_dev, id_base, n, g, n_ranks, lo, hi, offs, offs + n_ranks, reinterpret_cast<unsigned long long*>(counts_dev), nullptr, nullptr, out_rec_dev, s);
    return TESS_OK;
    TESS_CATCH
}

int tess_diagram_add_records_device(tess_diagram* d, const double* rec_dev, size_t n, void* stream) {
    if (!d || (!rec_dev && n)) return fail(TESS_ERR_INVALID, "NULL argument");
    if (d->initialized) return fail(TESS_ERR_STATE, "particles must be added before initialize (interface.rs:50-51)");
    if (d->n + n >= 0xFFFFFFF0ull) return fail(TESS_ERR_INVALID, "more than 2^32 particles on one device");
    if (d->n && !d->has_ids) return fail(TESS_ERR_INVALID, "ids must be given for all particles or for none");
    if (d->has_groups) return fail(TESS_ERR_INVALID, "records carry no groups");
    if (!n) return TESS_OK;
    TESS_TRY
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    TESS_CUDA_CHECK(cudaSetDevice(d->device));
    const size_t n0 = d->n;
    d->xyz.grow_keep(sizeof(double) * 3 * (n0 + n), sizeof(double) * 3 * n0, s);
    d->ids.grow_keep(sizeof(int64_t) * (n0 + n), sizeof(int64_t) * n0, s);
    TESS_CUDA_CHECK(cudaMemcpy2DAsync(d->xyz.as<double>() + 3 * n0, 24, rec_dev, 32, 24, n, cudaMemcpyDeviceToDevice, s));
    TESS_CUDA_CHECK(cudaMemcpy2DAsync(d->ids.as<int64_t>() + n0, 8, rec_dev + 3, 32, 8, n, cudaMemcpyDeviceToDevice, s));
    d->has_ids = true;
    d->n = n0 + n;
    return TESS_OK;
    TESS_CATCH
}

}  // extern "C"
