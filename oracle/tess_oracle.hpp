// tess_oracle.hpp — CPU ORACLE for the per-particle Voronoi cell path.
//
// THIS IS TEST INFRASTRUCTURE, NOT THE PRODUCT.  Only tests/, __graft_entry__.smoke() and
// bench.py's cpu_baseline / `--impl reference` legs may load it.  The product
// (the-tessellator_b200/csrc) never includes, links or calls anything in this directory.
//
// It is a C++17 restatement of the reference crate's algorithm (mcomstock/the-tessellator,
// src/{float,vector3,pool,celery,polyhedron,interface}.rs), operation by operation in f64,
// no FMA contraction (build with -ffp-contract=off), left-associated sums exactly as the
// Rust expressions are written.  Every function cites the reference file:line it follows.
//
// PARITY STATUS
//   * grid (celery.rs), vector/plane math (vector3.rs), pool (pool.rs) and start cube
//     (polyhedron.rs) are PINNED by the reference's own unit tests (tests/test_oracle_*.py
//     replay every one of them).
//   * full-cell outputs (volume / face areas / neighbour list): **parity unpinned** — the
//     reference has no test, golden vector or runnable code path that produces them (its
//     cut_with_plane is unfinished; no Rust toolchain exists in the build image).  They are
//     anchored instead on (a) an independent Qhull cross-check, (b) closure invariants and
//     (c) agreement of this file's literal O(N^2) `no_radius` mode with its terminated mode.
//
// Deliberate deviations from the reference source (numbering = SURVEY.md §2.3):
//   D1/D2  container = caller box (else bbox of points); the start polyhedron must be built.
//   D4     no Morton permutation: internal index == original index (user-visible ids are
//          original ids either way).
//   D5     cube edge DR belongs to face D.
//   D6     the cap face loop is closed (see Polyhedron::cut_with_plane).
//   D7/D8  each vertex/edge/face is freed exactly once; what survives is what is reachable
//          from root_edge through flip/next.
//   D9     debug println!s dropped.
//   D10    surviving container faces report neighbour ids -1..-6 (face slots 0..5 = F,R,B,L,U,D).
//   D11    expand_all_in_radius keeps the literal `distance(squared) > max_radius` compare.
//   D12    unspecified sort orders are made canonical: points inside a grid cell by index,
//          search offsets by (distance, i, j, k).
//   D16    self is excluded by index.
//   D17    "no strictly Inside->Outside edge" => the plane is skipped (reference behaviour),
//          and the cell's status gets ORC_STATUS_DEGENERATE_SKIP.
//   NEW    `security` mode: the shell walk stops at the first search_order entry whose
//          distance exceeds 4*max|v|^2 and candidates with |r|^2 >= 4*max|v|^2 are skipped.
//          Both tests can only remove candidates for which find_outgoing_edge would have
//          returned None (n.v - |r|/2 <= |v| - |r|/2 <= 0 < tol), so results are unchanged.
#pragma once

#include <cmath>
#include <cstddef>
#include <cstdint>
#include <limits>
#include <optional>
#include <stdexcept>
#include <vector>

namespace orc {

// ------------------------------------------------------------------------------------------
// float.rs
// ------------------------------------------------------------------------------------------

/// float.rs:138-142 — `self.0 as usize`: Rust float->int casts saturate (NaN -> 0).
inline size_t to_usize(double v) {
    if (!(v > 0.0)) return 0;  // NaN, negatives, zero
    if (v >= 18446744073709551616.0) return std::numeric_limits<size_t>::max();
    return static_cast<size_t>(v);
}

/// float.rs:158-170 — total order used by sort(): NaN sorts as Less.
inline int float_cmp(double a, double b) {
    if (std::isnan(a)) return -1;
    if (std::isnan(b)) return 1;
    return (a < b) ? -1 : (a > b) ? 1 : 0;
}

// ------------------------------------------------------------------------------------------
// vector3.rs
// ------------------------------------------------------------------------------------------

struct Vec3 {
    double x = 0, y = 0, z = 0;
};

/// vector3.rs:38-40
inline double dot(const Vec3& a, const Vec3& b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
/// vector3.rs:44-50
inline Vec3 cross(const Vec3& a, const Vec3& b) {
    return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
}
/// vector3.rs:53-59
inline Vec3 scale(const Vec3& a, double s) { return {a.x * s, a.y * s, a.z * s}; }
/// vector3.rs:63-65
inline double mag_sq(const Vec3& a) { return dot(a, a); }
/// vector3.rs:69-71
inline double mag(const Vec3& a) { return std::sqrt(mag_sq(a)); }
/// vector3.rs:75-77 — scale by the reciprocal (three multiplies, not three divides)
inline Vec3 unit(const Vec3& a) { return scale(a, 1.0 / mag(a)); }
/// vector3.rs:94-104
inline Vec3 add(const Vec3& a, const Vec3& b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
/// vector3.rs:106-116
inline Vec3 sub(const Vec3& a, const Vec3& b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
/// vector3.rs:118-128
inline Vec3 neg(const Vec3& a) { return {-a.x, -a.y, -a.z}; }
/// vector3.rs:81-83
inline Vec3 midpoint(const Vec3& a, const Vec3& b) { return scale(add(a, b), 0.5); }

/// vector3.rs:146-153
enum class Loc : int { Outside = 0, Incident = 1, Inside = 2 };

/// vector3.rs:158-250
struct Plane {
    Vec3 unit_normal;
    double plane_offset = 0;

    /// vector3.rs:170-178
    static Loc location(double sd, double tol) {
        if (sd > tol) return Loc::Outside;
        if (sd < -tol) return Loc::Inside;
        return Loc::Incident;
    }
    /// vector3.rs:208-210
    double offset_inverse(const Vec3& v) const { return dot(unit_normal, v); }
    /// vector3.rs:191-193
    double signed_distance(const Vec3& v) const { return offset_inverse(v) - plane_offset; }
    /// vector3.rs:183-185
    Loc vector_location(const Vec3& v, double tol) const { return location(signed_distance(v), tol); }
    /// vector3.rs:213-219
    Vec3 intersection(const Vec3& a, const Vec3& b) const {
        const double a_offset = offset_inverse(a);
        const double b_offset = offset_inverse(b);
        return add(a, scale(sub(b, a), (plane_offset - a_offset) / (b_offset - a_offset)));
    }
    /// vector3.rs:229-239
    static Plane build_from_normal_and_point(const Vec3& n, const Vec3& p) { return {n, dot(n, p)}; }
    /// vector3.rs:244-249
    static Plane build_from_non_unit_normal_and_point(const Vec3& n, const Vec3& p) {
        return build_from_normal_and_point(unit(n), p);
    }
    /// vector3.rs:223-225
    static Plane halfway_from_origin_to(const Vec3& p) {
        return build_from_normal_and_point(unit(p), scale(p, 0.5));
    }
};

/// vector3.rs:254-299
struct BoundingBox {
    Vec3 low, high;
    void adjust_to_contain(double x, double y, double z) {
        if (x < low.x) low.x = x;
        if (y < low.y) low.y = y;
        if (z < low.z) low.z = z;
        if (x > high.x) high.x = x;
        if (y > high.y) high.y = y;
        if (z > high.z) high.z = z;
    }
    void pad(double p) {
        low = {low.x - p, low.y - p, low.z - p};
        high = {high.x + p, high.y + p, high.z + p};
    }
};

// ------------------------------------------------------------------------------------------
// pool.rs — slab with a LIFO free list
// ------------------------------------------------------------------------------------------

template <class T>
struct Pool {
    enum Kind : int { Value = 0, NextIndex = 1, End = 2 };  // pool.rs:24-28
    struct Chunk {
        Kind kind;
        size_t next;
        T value;
    };
    std::vector<Chunk> data;
    std::optional<size_t> first;  // pool.rs:39

    /// pool.rs:85-110
    size_t add(const T& value) {
        if (first) {
            const size_t i = *first;
            if (data[i].kind == NextIndex) first = data[i].next;
            else first.reset();
            data[i] = Chunk{Value, 0, value};
            return i;
        }
        data.push_back(Chunk{Value, 0, value});
        return data.size() - 1;
    }
    /// pool.rs:113-121
    void remove(size_t index) {
        Chunk c{first ? NextIndex : End, first ? *first : 0, T{}};
        first = index;
        data[index] = c;
    }
    /// pool.rs:124-138
    T* get(size_t i) { return data[i].kind == Value ? &data[i].value : nullptr; }
    const T* get(size_t i) const { return data[i].kind == Value ? &data[i].value : nullptr; }
    /// pool.rs:141-154 (panics in the reference; throws here)
    T& at(size_t i) {
        if (data[i].kind != Value) throw std::runtime_error("Pool::get_or_fail on a free slot");
        return data[i].value;
    }
    const T& at(size_t i) const {
        if (data[i].kind != Value) throw std::runtime_error("Pool::get_or_fail on a free slot");
        return data[i].value;
    }
    /// pool.rs:157-162
    size_t next_index() const { return first ? *first : data.size() - 1; }
    /// pool.rs:165-170
    bool has(size_t i) const { return data[i].kind == Value; }
    /// pool.rs:173-175 — number of slots, live or not
    size_t len() const { return data.size(); }
    /// number of live values (what PoolIterator, pool.rs:186-203, would yield)
    size_t live() const {
        size_t n = 0;
        for (const auto& c : data) n += (c.kind == Value);
        return n;
    }
};

// ------------------------------------------------------------------------------------------
// celery.rs — uniform grid + expanding search
// ------------------------------------------------------------------------------------------

/// celery.rs:27-32
struct DistanceIndex {
    double distance;
    int32_t i, j, k;
};

/// celery.rs:64-77
struct CeleryBounds {
    double x_min = 0, x_max = 0, y_min = 0, y_max = 0, z_min = 0, z_max = 0;
};

/// celery.rs:130-149
struct CeleryCellInfo {
    double x_cell_size = 0, y_cell_size = 0, z_cell_size = 0;
    double x_inverse_cell_size = 0, y_inverse_cell_size = 0, z_inverse_cell_size = 0;
    size_t cells_per_dimension = 0;
};

struct Celery {
    std::vector<Vec3> points;
    std::vector<size_t> cells;
    std::vector<size_t> delimiters;
    std::vector<size_t> sorted_indices;
    CeleryBounds bounds;
    CeleryCellInfo cell_info;
    std::vector<DistanceIndex> search_order;
    /// Not in the reference: the table holds only offsets with |i|,|j|,|k| <= table_radius and
    /// distance < search_order_complete_below; it is then a PREFIX of the full (2cpd-1)^3 table.
    /// table_radius < 0 (or >= cpd-1) builds the full table, as the reference does.
    int table_radius = -1;
    bool table_is_full = true;

    static CeleryBounds make_bounds(const std::vector<Vec3>& pts);                       // celery.rs:81-125
    static CeleryCellInfo make_cell_info(size_t n, const CeleryBounds& b);                // celery.rs:153-189
    static size_t axis_index(double v, double vmin, double vmax, double inv, size_t cpd);  // celery.rs:269-314
    size_t x_index(double x) const { return axis_index(x, bounds.x_min, bounds.x_max, cell_info.x_inverse_cell_size, cell_info.cells_per_dimension); }
    size_t y_index(double y) const { return axis_index(y, bounds.y_min, bounds.y_max, cell_info.y_inverse_cell_size, cell_info.cells_per_dimension); }
    size_t z_index(double z) const { return axis_index(z, bounds.z_min, bounds.z_max, cell_info.z_inverse_cell_size, cell_info.cells_per_dimension); }
    size_t cell_from_indices(size_t x, size_t y, size_t z) const {  // celery.rs:317-325
        const size_t cpd = cell_info.cells_per_dimension;
        return x * cpd * cpd + y * cpd + z;
    }
    size_t get_cell(const Vec3& p) const { return cell_from_indices(x_index(p.x), y_index(p.y), z_index(p.z)); }  // celery.rs:328-339

    void reset(int table_radius_);  // celery.rs:253-266
    void build_search_order();      // celery.rs:418-679

    bool check_cell_in_range(double x, double y, double z, double radius, size_t xi, size_t yi, size_t zi) const;  // celery.rs:708-743
    std::vector<size_t> find_cells_in_radius(double x, double y, double z, double radius) const;                  // celery.rs:753-797
    std::vector<size_t> find_neighbors_in_cell_radius(double x, double y, double z, double radius) const;         // celery.rs:802-819
    std::vector<size_t> find_neighbors_in_real_radius(double x, double y, double z, double radius) const;         // celery.rs:825-855
};

/// celery.rs:865-1076
struct ExpandingSearch {
    const Celery* celery;
    size_t current_search_index = 0;
    size_t x_cell_index, y_cell_index, z_cell_index;

    ExpandingSearch(const Celery& c, double x, double y, double z);       // celery.rs:882-902
    std::vector<size_t> expand(double max_radius, size_t cells_to_add);   // celery.rs:907-963
    std::vector<size_t> expand_all_no_radius();                           // celery.rs:971-1018
    std::vector<size_t> expand_all_in_radius(double max_radius);          // celery.rs:1023-1075
    /// helper shared by the three walks: append the points of one table entry (or nothing if OOB)
    void append_entry(const DistanceIndex& e, std::vector<size_t>& out) const;
};

// ------------------------------------------------------------------------------------------
// polyhedron.rs — half-edge clipper
// ------------------------------------------------------------------------------------------

using OptIdx = std::optional<size_t>;

/// polyhedron.rs:26-39
struct HalfEdge {
    OptIdx flip, next, target, face;
};
/// polyhedron.rs:44-52
struct Face {
    OptIdx point_index;  // None for the six container faces (polyhedron.rs:306-309)
    size_t starting_edge_index = 0;
    int wall = -1;       // oracle addition (D10): 0..5 for container faces F,R,B,L,U,D
};
/// polyhedron.rs:57-63
struct FaceData {
    size_t face_index;
    Vec3 weighted_normal;
};

struct CutCounters {
    uint64_t vertex_classifications = 0;  // VC
    uint64_t new_vertices = 0;            // NV
    uint64_t cuts = 0;
    uint64_t degenerate_skips = 0;        // D17 events
};

struct Polyhedron {
    OptIdx root_edge;
    Pool<HalfEdge> edges;
    Pool<Vec3> vertices;
    Pool<Face> faces;
    std::vector<FaceData> face_data;
    CutCounters counters;

    static double tolerance() { return 1e-12; }  // polyhedron.rs:221-223

    Polyhedron() = default;
    Polyhedron(double x_min, double y_min, double z_min, double x_max, double y_max, double z_max) {  // polyhedron.rs:226-244
        reset(x_min, y_min, z_min, x_max, y_max, z_max);
    }
    void reset(double x_min, double y_min, double z_min, double x_max, double y_max, double z_max);  // polyhedron.rs:247-392
    bool is_built() const { return root_edge.has_value(); }                                          // polyhedron.rs:761-763
    OptIdx find_outgoing_edge(const Plane& plane);                                                   // polyhedron.rs:396-435
    bool cut_with_plane(size_t point_index, const Plane& plane);                                     // polyhedron.rs:438-642 (+D6-D9)
    Vec3 weighted_normal(size_t face_index) const;                                                   // polyhedron.rs:776-808
    void compute_face_data();                                                                        // polyhedron.rs:812-825
    double compute_volume();                                                                         // polyhedron.rs:838-855
    void translate(const Vec3& shift);                                                               // polyhedron.rs:859-867
    std::vector<int64_t> compute_neighbors() const;                                                  // polyhedron.rs:871-881 (+D10)
    std::vector<Vec3> compute_vertices() const;                                                      // polyhedron.rs:885-893
    std::vector<Vec3> compute_face_vertices(size_t face_index) const;                                // polyhedron.rs:897-919
    double max_vertex_radius_sq() const;                                                             // NEW (security radius)

    OptIdx target_index(OptIdx e) const { return (e && edges.has(*e)) ? edges.at(*e).target : OptIdx{}; }  // polyhedron.rs:733-737
    OptIdx next_index(OptIdx e) const { return (e && edges.has(*e)) ? edges.at(*e).next : OptIdx{}; }      // polyhedron.rs:747-751
    OptIdx flip_index(OptIdx e) const { return (e && edges.has(*e)) ? edges.at(*e).flip : OptIdx{}; }      // polyhedron.rs:754-758

   private:
    void clean_up(const std::vector<size_t>& vertices_to_destroy);  // polyhedron.rs:645-730 (repaired, D7/D8)
};

// ------------------------------------------------------------------------------------------
// interface.rs — Diagram / Cell
// ------------------------------------------------------------------------------------------

enum : uint32_t {
    ORC_STATUS_OK = 0,
    ORC_STATUS_DEGENERATE_SKIP = 1u << 0,   // D17 happened at least once
    ORC_STATUS_TABLE_EXHAUSTED = 1u << 1,   // truncated search table ran out before termination
};

enum SearchMode : int {
    MODE_NO_RADIUS = 0,         // interface.rs:277  expand_all_no_radius — literal O(N) per cell
    MODE_REFERENCE_RADIUS = 1,  // interface.rs:276  expand_all_in_radius(r) — literal (D11)
    MODE_SECURITY = 2,          // NEW: results-preserving termination
};

struct CellCounters {
    uint64_t visited = 0;   // C_vis: candidates offered by the shell walk (self included)
    uint64_t tested = 0;    // C_test: candidates handed to cut_with_plane
    uint64_t vertex_classifications = 0;
    uint64_t cuts = 0;
    uint64_t new_vertices = 0;
    uint64_t table_entries = 0;  // search_order entries consumed
    uint64_t degenerate_skips = 0;
};

struct CellResult {
    double volume = 0;
    std::vector<int64_t> neighbors;  // face-slot order; walls are -1..-6
    std::vector<double> areas;       // same order
    std::vector<Vec3> vertices;      // cell-local coordinates (optional)
    std::vector<uint32_t> face_loop_sizes;  // per face (same order as neighbors): vertices of its loop (optional)
    std::vector<Vec3> face_loop_vertices;   // the loops, concatenated: compute_face_vertices (polyhedron.rs:897-919)
    uint32_t status = 0;
    double max_radius_sq = 0;        // final max |v|^2
    uint32_t pool_slots[3] = {0, 0, 0};  // slots ever used by the vertex / half-edge / face pools (Pool::len): capacity the cell needed
    CellCounters counters;
};

/// interface.rs:25-47 (D1, D3, D4 resolved as described at the top of this file)
struct Diagram {
    Celery cell_array;
    std::vector<size_t> groups;
    BoundingBox bounding_box;
    double box[6] = {0, 0, 0, 0, 0, 0};  // container: x_min,y_min,z_min,x_max,y_max,z_max
    bool initialized = false;

    void add_particle_with_group(const Vec3& p, size_t group);           // interface.rs:52-57 (+D14)
    void initialize(const double* container_box, int table_radius);      // interface.rs:60-84

    /// interface.rs:186-208 + 257-344: cell of the particle at `index`.
    /// target_group < 0 means None.  want_vertices fills CellResult::vertices.
    CellResult compute_cell_at_index(size_t index, SearchMode mode, double search_radius, int64_t target_group, bool want_vertices) const;
    /// interface.rs:211-232: cell of an arbitrary position (no self exclusion).
    CellResult compute_cell_at_point(const Vec3& position, SearchMode mode, double search_radius, int64_t target_group, bool want_vertices) const;

   private:
    CellResult compute(const Vec3& position, OptIdx self, SearchMode mode, double search_radius, int64_t target_group, bool want_vertices) const;
};

}  // namespace orc
