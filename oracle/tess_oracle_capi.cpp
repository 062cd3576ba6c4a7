// tess_oracle_capi.cpp — flat C entry points over the CPU ORACLE so that tests/ and bench.py's
// cpu_baseline leg can drive it through ctypes.  TEST INFRASTRUCTURE ONLY (see tess_oracle.hpp).
#include <algorithm>
#include <atomic>
#include <cstring>
#include <string>
#include <thread>

#include "tess_oracle.hpp"

using namespace orc;

namespace {
thread_local std::string g_err;

struct OrcResult {
    std::vector<double> volumes;
    std::vector<uint64_t> face_offsets;  // m+1
    std::vector<int64_t> neighbors;
    std::vector<double> areas;
    std::vector<uint32_t> status;
    std::vector<double> max_radius_sq;
    std::vector<uint64_t> vertex_offsets;  // m+1 (only when vertices requested)
    std::vector<double> vertices;          // xyz triples, cell-local
    std::vector<uint64_t> loop_offsets;    // per face (global face order), n_faces+1
    std::vector<double> loop_vertices;     // xyz triples of the face loops
    uint64_t counters[8] = {0, 0, 0, 0, 0, 0, 0, 0};
};
}  // namespace

extern "C" {

const char* orc_last_error(void) { return g_err.c_str(); }

// ---------------------------------------------------------------- diagram ------------------
void* orc_diagram_create(const double* xyz, uint64_t n, const uint64_t* groups, const double* box6, int table_radius) {
    try {
        auto* d = new Diagram();
        d->cell_array.points.reserve(n);
        for (uint64_t i = 0; i < n; ++i) d->add_particle_with_group({xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2]}, groups ? groups[i] : 0);
        d->initialize(box6, table_radius);
        return d;
    } catch (const std::exception& e) {
        g_err = e.what();
        return nullptr;
    }
}
void orc_diagram_destroy(void* d) { delete static_cast<Diagram*>(d); }

uint64_t orc_grid_cpd(const void* d) { return static_cast<const Diagram*>(d)->cell_array.cell_info.cells_per_dimension; }
void orc_grid_bounds(const void* d, double* out6) {  // x_min,x_max,y_min,y_max,z_min,z_max
    const auto& b = static_cast<const Diagram*>(d)->cell_array.bounds;
    out6[0] = b.x_min; out6[1] = b.x_max; out6[2] = b.y_min; out6[3] = b.y_max; out6[4] = b.z_min; out6[5] = b.z_max;
}
void orc_grid_cell_info(const void* d, double* out6) {  // sizes xyz, inverse sizes xyz
    const auto& c = static_cast<const Diagram*>(d)->cell_array.cell_info;
    out6[0] = c.x_cell_size; out6[1] = c.y_cell_size; out6[2] = c.z_cell_size;
    out6[3] = c.x_inverse_cell_size; out6[4] = c.y_inverse_cell_size; out6[5] = c.z_inverse_cell_size;
}
void orc_container_box(const void* d, double* out6) { std::memcpy(out6, static_cast<const Diagram*>(d)->box, 6 * sizeof(double)); }
uint64_t orc_grid_num_points(const void* d) { return static_cast<const Diagram*>(d)->cell_array.points.size(); }
uint64_t orc_grid_num_delimiters(const void* d) { return static_cast<const Diagram*>(d)->cell_array.delimiters.size(); }
uint64_t orc_grid_search_order_len(const void* d) { return static_cast<const Diagram*>(d)->cell_array.search_order.size(); }
int orc_grid_table_is_full(const void* d) { return static_cast<const Diagram*>(d)->cell_array.table_is_full ? 1 : 0; }
void orc_grid_copy_cells(const void* d, uint64_t* out) {
    const auto& v = static_cast<const Diagram*>(d)->cell_array.cells;
    for (size_t i = 0; i < v.size(); ++i) out[i] = v[i];
}
void orc_grid_copy_sorted_indices(const void* d, uint64_t* out) {
    const auto& v = static_cast<const Diagram*>(d)->cell_array.sorted_indices;
    for (size_t i = 0; i < v.size(); ++i) out[i] = v[i];
}
void orc_grid_copy_delimiters(const void* d, uint64_t* out) {
    const auto& v = static_cast<const Diagram*>(d)->cell_array.delimiters;
    for (size_t i = 0; i < v.size(); ++i) out[i] = v[i];
}
void orc_grid_copy_search_order(const void* d, double* dist, int32_t* ijk) {
    const auto& v = static_cast<const Diagram*>(d)->cell_array.search_order;
    for (size_t t = 0; t < v.size(); ++t) {
        dist[t] = v[t].distance;
        ijk[3 * t] = v[t].i; ijk[3 * t + 1] = v[t].j; ijk[3 * t + 2] = v[t].k;
    }
}
void orc_grid_home_cell(const void* d, double x, double y, double z, uint64_t* out3) {
    const auto& c = static_cast<const Diagram*>(d)->cell_array;
    out3[0] = c.x_index(x); out3[1] = c.y_index(y); out3[2] = c.z_index(z);
}

// ---------------------------------------------------------------- queries ------------------
static uint64_t copy_out(const std::vector<size_t>& v, uint64_t* out, uint64_t cap) {
    for (size_t i = 0; i < v.size() && i < cap; ++i) out[i] = v[i];
    return v.size();
}
int orc_check_cell_in_range(const void* d, double x, double y, double z, double r, uint64_t i, uint64_t j, uint64_t k) {
    return static_cast<const Diagram*>(d)->cell_array.check_cell_in_range(x, y, z, r, i, j, k) ? 1 : 0;
}
uint64_t orc_find_cells_in_radius(const void* d, double x, double y, double z, double r, uint64_t* out, uint64_t cap) {
    return copy_out(static_cast<const Diagram*>(d)->cell_array.find_cells_in_radius(x, y, z, r), out, cap);
}
uint64_t orc_find_neighbors_in_cell_radius(const void* d, double x, double y, double z, double r, uint64_t* out, uint64_t cap) {
    return copy_out(static_cast<const Diagram*>(d)->cell_array.find_neighbors_in_cell_radius(x, y, z, r), out, cap);
}
uint64_t orc_find_neighbors_in_real_radius(const void* d, double x, double y, double z, double r, uint64_t* out, uint64_t cap) {
    return copy_out(static_cast<const Diagram*>(d)->cell_array.find_neighbors_in_real_radius(x, y, z, r), out, cap);
}

void* orc_search_create(const void* d, double x, double y, double z) { return new ExpandingSearch(static_cast<const Diagram*>(d)->cell_array, x, y, z); }
void orc_search_destroy(void* s) { delete static_cast<ExpandingSearch*>(s); }
void orc_search_home(const void* s, uint64_t* out3) {
    const auto* es = static_cast<const ExpandingSearch*>(s);
    out3[0] = es->x_cell_index; out3[1] = es->y_cell_index; out3[2] = es->z_cell_index;
}
uint64_t orc_search_expand(void* s, double max_radius, uint64_t cells_to_add, uint64_t* out, uint64_t cap) {
    return copy_out(static_cast<ExpandingSearch*>(s)->expand(max_radius, cells_to_add), out, cap);
}
uint64_t orc_search_expand_all_no_radius(void* s, uint64_t* out, uint64_t cap) {
    return copy_out(static_cast<ExpandingSearch*>(s)->expand_all_no_radius(), out, cap);
}
uint64_t orc_search_expand_all_in_radius(void* s, double max_radius, uint64_t* out, uint64_t cap) {
    return copy_out(static_cast<ExpandingSearch*>(s)->expand_all_in_radius(max_radius), out, cap);
}

// ---------------------------------------------------------------- cells --------------------
// ids == NULL -> cells 0..m-1.  mode: 0 no_radius, 1 reference_radius, 2 security.
// target_group < 0 -> None.  nthreads <= 0 -> hardware_concurrency.
void* orc_compute_cells(const void* dv, const uint64_t* ids, uint64_t m, int mode, double search_radius, int64_t target_group, int want_vertices, int nthreads) {
    const auto* d = static_cast<const Diagram*>(dv);
    try {
        std::vector<CellResult> cells(m);
        int nt = nthreads > 0 ? nthreads : static_cast<int>(std::thread::hardware_concurrency());
        nt = std::max(1, std::min<int>(nt, static_cast<int>(std::max<uint64_t>(m, 1))));
        std::atomic<uint64_t> next{0};
        std::atomic<bool> failed{false};
        std::string err;
        auto work = [&]() {
            try {
                for (;;) {
                    const uint64_t b = next.fetch_add(256);
                    if (b >= m) break;
                    const uint64_t e = std::min<uint64_t>(m, b + 256);
                    for (uint64_t c = b; c < e; ++c)
                        cells[c] = d->compute_cell_at_index(ids ? ids[c] : c, static_cast<SearchMode>(mode), search_radius, target_group, want_vertices != 0);
                }
            } catch (const std::exception& ex) {
                if (!failed.exchange(true)) err = ex.what();
            }
        };
        if (nt == 1) {
            work();
        } else {
            std::vector<std::thread> th;
            for (int t = 0; t < nt; ++t) th.emplace_back(work);
            for (auto& t : th) t.join();
        }
        if (failed) {
            g_err = err;
            return nullptr;
        }
        auto* r = new OrcResult();
        r->volumes.resize(m);
        r->status.resize(m);
        r->max_radius_sq.resize(m);
        r->face_offsets.assign(m + 1, 0);
        r->vertex_offsets.assign(m + 1, 0);
        for (uint64_t c = 0; c < m; ++c) {
            r->face_offsets[c + 1] = r->face_offsets[c] + cells[c].neighbors.size();
            r->vertex_offsets[c + 1] = r->vertex_offsets[c] + cells[c].vertices.size();
        }
        r->loop_offsets.assign(1, 0);
        r->neighbors.reserve(r->face_offsets[m]);
        r->areas.reserve(r->face_offsets[m]);
        r->vertices.reserve(3 * r->vertex_offsets[m]);
        for (uint64_t c = 0; c < m; ++c) {
            const CellResult& cr = cells[c];
            r->volumes[c] = cr.volume;
            r->status[c] = cr.status;
            r->max_radius_sq[c] = cr.max_radius_sq;
            r->neighbors.insert(r->neighbors.end(), cr.neighbors.begin(), cr.neighbors.end());
            r->areas.insert(r->areas.end(), cr.areas.begin(), cr.areas.end());
            for (const Vec3& v : cr.vertices) {
                r->vertices.push_back(v.x); r->vertices.push_back(v.y); r->vertices.push_back(v.z);
            }
            for (uint32_t sz : cr.face_loop_sizes) r->loop_offsets.push_back(r->loop_offsets.back() + sz);
            for (const Vec3& v : cr.face_loop_vertices) {
                r->loop_vertices.push_back(v.x); r->loop_vertices.push_back(v.y); r->loop_vertices.push_back(v.z);
            }
            r->counters[0] += cr.counters.visited;
            r->counters[1] += cr.counters.tested;
            r->counters[2] += cr.counters.vertex_classifications;
            r->counters[3] += cr.counters.cuts;
            r->counters[4] += cr.counters.new_vertices;
            r->counters[5] += cr.counters.table_entries;
            r->counters[6] += cr.counters.degenerate_skips;
            r->counters[7] += cr.neighbors.size();
        }
        return r;
    } catch (const std::exception& e) {
        g_err = e.what();
        return nullptr;
    }
}

// One cell around an arbitrary position (interface.rs:211-232).
void* orc_compute_cell_at_point(const void* dv, double x, double y, double z, int mode, double search_radius, int64_t target_group, int want_vertices) {
    const auto* d = static_cast<const Diagram*>(dv);
    try {
        CellResult cr = d->compute_cell_at_point({x, y, z}, static_cast<SearchMode>(mode), search_radius, target_group, want_vertices != 0);
        auto* r = new OrcResult();
        r->volumes = {cr.volume};
        r->status = {cr.status};
        r->max_radius_sq = {cr.max_radius_sq};
        r->face_offsets = {0, cr.neighbors.size()};
        r->vertex_offsets = {0, cr.vertices.size()};
        r->neighbors = cr.neighbors;
        r->areas = cr.areas;
        for (const Vec3& v : cr.vertices) {
            r->vertices.push_back(v.x); r->vertices.push_back(v.y); r->vertices.push_back(v.z);
        }
        return r;
    } catch (const std::exception& e) {
        g_err = e.what();
        return nullptr;
    }
}

void orc_result_free(void* r) { delete static_cast<OrcResult*>(r); }
uint64_t orc_result_n_cells(const void* r) { return static_cast<const OrcResult*>(r)->volumes.size(); }
const double* orc_result_volumes(const void* r) { return static_cast<const OrcResult*>(r)->volumes.data(); }
const uint64_t* orc_result_face_offsets(const void* r) { return static_cast<const OrcResult*>(r)->face_offsets.data(); }
const int64_t* orc_result_neighbors(const void* r) { return static_cast<const OrcResult*>(r)->neighbors.data(); }
const double* orc_result_areas(const void* r) { return static_cast<const OrcResult*>(r)->areas.data(); }
const uint32_t* orc_result_status(const void* r) { return static_cast<const OrcResult*>(r)->status.data(); }
const double* orc_result_max_radius_sq(const void* r) { return static_cast<const OrcResult*>(r)->max_radius_sq.data(); }
const uint64_t* orc_result_vertex_offsets(const void* r) { return static_cast<const OrcResult*>(r)->vertex_offsets.data(); }
const double* orc_result_vertices(const void* r) { return static_cast<const OrcResult*>(r)->vertices.data(); }
const uint64_t* orc_result_counters(const void* r) { return static_cast<const OrcResult*>(r)->counters; }
uint64_t orc_result_n_loops(const void* r) { return static_cast<const OrcResult*>(r)->loop_offsets.size(); }
const uint64_t* orc_result_loop_offsets(const void* r) { return static_cast<const OrcResult*>(r)->loop_offsets.data(); }
const double* orc_result_loop_vertices(const void* r) { return static_cast<const OrcResult*>(r)->loop_vertices.data(); }

// ---------------------------------------------------------------- unit hooks ---------------
// vector3.rs
double orc_dot(const double* a, const double* b) { return dot({a[0], a[1], a[2]}, {b[0], b[1], b[2]}); }
void orc_cross(const double* a, const double* b, double* o) { Vec3 c = cross({a[0], a[1], a[2]}, {b[0], b[1], b[2]}); o[0] = c.x; o[1] = c.y; o[2] = c.z; }
void orc_scale(const double* a, double s, double* o) { Vec3 c = scale({a[0], a[1], a[2]}, s); o[0] = c.x; o[1] = c.y; o[2] = c.z; }
void orc_add(const double* a, const double* b, double* o) { Vec3 c = add({a[0], a[1], a[2]}, {b[0], b[1], b[2]}); o[0] = c.x; o[1] = c.y; o[2] = c.z; }
void orc_sub(const double* a, const double* b, double* o) { Vec3 c = sub({a[0], a[1], a[2]}, {b[0], b[1], b[2]}); o[0] = c.x; o[1] = c.y; o[2] = c.z; }
int orc_location(double sd, double tol) { return static_cast<int>(Plane::location(sd, tol)); }
int orc_vector_location(const double* plane4, const double* v, double tol) {
    Plane p{{plane4[0], plane4[1], plane4[2]}, plane4[3]};
    return static_cast<int>(p.vector_location({v[0], v[1], v[2]}, tol));
}
void orc_intersection(const double* plane4, const double* a, const double* b, double* o) {
    Plane p{{plane4[0], plane4[1], plane4[2]}, plane4[3]};
    Vec3 c = p.intersection({a[0], a[1], a[2]}, {b[0], b[1], b[2]});
    o[0] = c.x; o[1] = c.y; o[2] = c.z;
}
void orc_plane_halfway_from_origin_to(const double* pt, double* plane4) {
    Plane p = Plane::halfway_from_origin_to({pt[0], pt[1], pt[2]});
    plane4[0] = p.unit_normal.x; plane4[1] = p.unit_normal.y; plane4[2] = p.unit_normal.z; plane4[3] = p.plane_offset;
}
void orc_plane_from_non_unit_normal_and_point(const double* n, const double* pt, double* plane4) {
    Plane p = Plane::build_from_non_unit_normal_and_point({n[0], n[1], n[2]}, {pt[0], pt[1], pt[2]});
    plane4[0] = p.unit_normal.x; plane4[1] = p.unit_normal.y; plane4[2] = p.unit_normal.z; plane4[3] = p.plane_offset;
}
void orc_bbox_adjust(double* low3, double* high3, double x, double y, double z) {
    BoundingBox b{{low3[0], low3[1], low3[2]}, {high3[0], high3[1], high3[2]}};
    b.adjust_to_contain(x, y, z);
    low3[0] = b.low.x; low3[1] = b.low.y; low3[2] = b.low.z; high3[0] = b.high.x; high3[1] = b.high.y; high3[2] = b.high.z;
}
void orc_bbox_pad(double* low3, double* high3, double p) {
    BoundingBox b{{low3[0], low3[1], low3[2]}, {high3[0], high3[1], high3[2]}};
    b.pad(p);
    low3[0] = b.low.x; low3[1] = b.low.y; low3[2] = b.low.z; high3[0] = b.high.x; high3[1] = b.high.y; high3[2] = b.high.z;
}
uint64_t orc_to_usize(double v) { return to_usize(v); }
uint64_t orc_cells_per_dimension(uint64_t n) { return to_usize(std::cbrt(static_cast<double>(n) / 1.25)) + 1; }

// pool.rs — a Pool<int64_t> driven by a tiny op interpreter so the reference's pool tests replay.
void* orc_pool_create(void) { return new Pool<int64_t>(); }
void orc_pool_destroy(void* p) { delete static_cast<Pool<int64_t>*>(p); }
uint64_t orc_pool_add(void* p, int64_t v) { return static_cast<Pool<int64_t>*>(p)->add(v); }
void orc_pool_remove(void* p, uint64_t i) { static_cast<Pool<int64_t>*>(p)->remove(i); }
uint64_t orc_pool_len(const void* p) { return static_cast<const Pool<int64_t>*>(p)->len(); }
int64_t orc_pool_first(const void* p) { const auto* q = static_cast<const Pool<int64_t>*>(p); return q->first ? static_cast<int64_t>(*q->first) : -1; }
// kind: 0 Value, 1 NextIndex, 2 End; payload = value or next index
int orc_pool_chunk(const void* p, uint64_t i, int64_t* payload) {
    const auto& c = static_cast<const Pool<int64_t>*>(p)->data[i];
    *payload = c.kind == Pool<int64_t>::Value ? c.value : static_cast<int64_t>(c.next);
    return static_cast<int>(c.kind);
}
int orc_pool_has(const void* p, uint64_t i) { return static_cast<const Pool<int64_t>*>(p)->has(i) ? 1 : 0; }
uint64_t orc_pool_iterate(const void* p, int64_t* out, uint64_t cap) {
    const auto* q = static_cast<const Pool<int64_t>*>(p);
    uint64_t n = 0;
    for (size_t i = 0; i < q->len(); ++i)
        if (const int64_t* v = q->get(i)) { if (n < cap) out[n] = *v; ++n; }
    return n;
}

// polyhedron.rs
void* orc_poly_create(double x0, double y0, double z0, double x1, double y1, double z1) { return new Polyhedron(x0, y0, z0, x1, y1, z1); }
void orc_poly_destroy(void* p) { delete static_cast<Polyhedron*>(p); }
void orc_poly_reset(void* p, double x0, double y0, double z0, double x1, double y1, double z1) { static_cast<Polyhedron*>(p)->reset(x0, y0, z0, x1, y1, z1); }
void orc_poly_counts(const void* pv, uint64_t* out5) {  // edge slots, vertex slots, face slots, face_data len, root_edge(+1, 0 = None)
    const auto* p = static_cast<const Polyhedron*>(pv);
    out5[0] = p->edges.len(); out5[1] = p->vertices.len(); out5[2] = p->faces.len(); out5[3] = p->face_data.size();
    out5[4] = p->root_edge ? *p->root_edge + 1 : 0;
}
void orc_poly_live_counts(const void* pv, uint64_t* out3) {
    const auto* p = static_cast<const Polyhedron*>(pv);
    out3[0] = p->edges.live(); out3[1] = p->vertices.live(); out3[2] = p->faces.live();
}
// edge -> flip,next,target,face (+1; 0 = None); returns 0 if the slot is free
int orc_poly_edge(const void* pv, uint64_t e, uint64_t* out4) {
    const auto* p = static_cast<const Polyhedron*>(pv);
    const HalfEdge* h = p->edges.get(e);
    if (!h) return 0;
    out4[0] = h->flip ? *h->flip + 1 : 0; out4[1] = h->next ? *h->next + 1 : 0;
    out4[2] = h->target ? *h->target + 1 : 0; out4[3] = h->face ? *h->face + 1 : 0;
    return 1;
}
int orc_poly_vertex(const void* pv, uint64_t v, double* out3) {
    const auto* p = static_cast<const Polyhedron*>(pv);
    const Vec3* x = p->vertices.get(v);
    if (!x) return 0;
    out3[0] = x->x; out3[1] = x->y; out3[2] = x->z;
    return 1;
}
int orc_poly_face(const void* pv, uint64_t f, int64_t* neighbor, uint64_t* starting_edge) {
    const auto* p = static_cast<const Polyhedron*>(pv);
    const Face* x = p->faces.get(f);
    if (!x) return 0;
    *neighbor = x->point_index ? static_cast<int64_t>(*x->point_index) : -static_cast<int64_t>(x->wall + 1);
    *starting_edge = x->starting_edge_index;
    return 1;
}
int64_t orc_poly_find_outgoing_edge(void* pv, const double* plane4) {
    Plane pl{{plane4[0], plane4[1], plane4[2]}, plane4[3]};
    OptIdx e = static_cast<Polyhedron*>(pv)->find_outgoing_edge(pl);
    return e ? static_cast<int64_t>(*e) : -1;
}
int orc_poly_cut_with_plane(void* pv, uint64_t point_index, const double* plane4) {
    Plane pl{{plane4[0], plane4[1], plane4[2]}, plane4[3]};
    try {
        return static_cast<Polyhedron*>(pv)->cut_with_plane(point_index, pl) ? 1 : 0;
    } catch (const std::exception& e) {
        g_err = e.what();
        return -1;
    }
}
void orc_poly_translate(void* pv, const double* s) { static_cast<Polyhedron*>(pv)->translate({s[0], s[1], s[2]}); }
double orc_poly_volume(void* pv) {
    auto* p = static_cast<Polyhedron*>(pv);
    p->face_data.clear();  // D13
    return p->compute_volume();
}
void orc_poly_weighted_normal(const void* pv, uint64_t f, double* out3) {
    Vec3 w = static_cast<const Polyhedron*>(pv)->weighted_normal(f);
    out3[0] = w.x; out3[1] = w.y; out3[2] = w.z;
}
uint64_t orc_poly_face_vertices(const void* pv, uint64_t f, double* out, uint64_t cap_vertices) {
    auto v = static_cast<const Polyhedron*>(pv)->compute_face_vertices(f);
    for (size_t i = 0; i < v.size() && i < cap_vertices; ++i) { out[3 * i] = v[i].x; out[3 * i + 1] = v[i].y; out[3 * i + 2] = v[i].z; }
    return v.size();
}
// Structural self-check used by the tests: flip(flip(e)) == e, face/next consistency,
// source(e) == target(flip(e)) continuity, Euler characteristic.  Returns 0 if consistent.
int orc_poly_check(const void* pv) {
    const auto* p = static_cast<const Polyhedron*>(pv);
    size_t E = 0;
    for (size_t e = 0; e < p->edges.len(); ++e) {
        const HalfEdge* h = p->edges.get(e);
        if (!h) continue;
        ++E;
        if (!h->flip || !h->next || !h->target || !h->face) return 1;
        const HalfEdge* f = p->edges.get(*h->flip);
        const HalfEdge* n = p->edges.get(*h->next);
        if (!f || !n) return 2;
        if (!f->flip || *f->flip != e) return 3;
        if (!n->face || *n->face != *h->face) return 4;
        if (!p->vertices.has(*h->target)) return 5;
        if (!p->faces.has(*h->face)) return 6;
        // the next edge starts where this one ends: source(next) = target(flip(next)) == target(e)
        const HalfEdge* nf = p->edges.get(*n->flip);
        if (!nf || !nf->target || *nf->target != *h->target) return 7;
    }
    const long V = static_cast<long>(p->vertices.live()), F = static_cast<long>(p->faces.live());
    if (V - static_cast<long>(E / 2) + F != 2) return 8;
    for (size_t f = 0; f < p->faces.len(); ++f) {
        const Face* x = p->faces.get(f);
        if (!x) continue;
        const HalfEdge* s = p->edges.get(x->starting_edge_index);
        if (!s || !s->face || *s->face != f) return 9;
    }
    return 0;
}

}  // extern "C"
