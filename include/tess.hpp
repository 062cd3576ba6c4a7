// tess.hpp — C++ host mirror of the reference's public API (src/interface.rs) over the C ABI tess.h.
//
// The reference is compiled code (Rust); where its toolchain is absent the host side above the C ABI is
// C++.  Type and method names, argument meaning and results follow interface.rs:
//     Diagram  (interface.rs:25)   add_particle_with_group :52, initialize :60, get_cell_at_index :186,
//                                  get_cell_at_particle :211
//     Cell     (interface.rs:237)  compute_voronoi_cell :257, compute_volume :337, compute_neighbors :342,
//                                  compute_vertices :368, compute_faces :373, original_index :387
//     VoronoiFace (interface.rs:393) compute_area :408, compute_neighbor :413
// Errors: the reference panics; here every failing ABI call throws tess::Error (code + tess_last_error()).
// Header-only; link with -ltess_b200.
#pragma once

#include <cmath>
#include <cstdint>
#include <limits>
#include <map>
#include <memory>
#include <optional>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "tess.h"

namespace tess {

struct Error : std::runtime_error {
    int code;
    Error(int c, const std::string& m) : std::runtime_error("libtess_b200 error " + std::to_string(c) + ": " + m), code(c) {}
};
inline void check(int rc) {
    if (rc != TESS_OK) throw Error(rc, tess_last_error());
}

/// vector3.rs:24-28
struct Vector3 {
    double x = 0, y = 0, z = 0;
};

/// The start shape of a cell: Polyhedron::new(x_min, y_min, z_min, x_max, y_max, z_max) (polyhedron.rs:226-233).
/// Only the axis-aligned box is supported on the GPU path.
struct Polyhedron {
    double x_min, y_min, z_min, x_max, y_max, z_max;
    bool operator==(const Polyhedron& o) const {
        return x_min == o.x_min && y_min == o.y_min && z_min == o.z_min && x_max == o.x_max && y_max == o.y_max && z_max == o.z_max;
    }
};

/// All cells of one tess_compute_all call.
class CellBatch {
   public:
    explicit CellBatch(tess_result* r) : r_(r) {
        check(tess_result_n_cells(r_, &n_cells, &n_faces));
        check(tess_result_volumes(r_, &volumes));
        check(tess_result_face_offsets(r_, &face_offsets));
        check(tess_result_neighbors(r_, &neighbors));
        check(tess_result_areas(r_, &areas));
        check(tess_result_status(r_, &status));
        // geometry is present only when TESS_OUT_VERTICES was requested
        if (tess_result_vertex_offsets(r_, &vertex_offsets) != TESS_OK || tess_result_vertices(r_, &vertices) != TESS_OK ||
            tess_result_face_vertex_offsets(r_, &face_vertex_offsets) != TESS_OK || tess_result_face_vertex_indices(r_, &face_vertex_indices) != TESS_OK)
            vertex_offsets = nullptr;
    }
    ~CellBatch() { tess_result_free(r_); }
    CellBatch(const CellBatch&) = delete;
    CellBatch& operator=(const CellBatch&) = delete;
    uint64_t n_cells = 0, n_faces = 0;
    const double* volumes = nullptr;
    const uint64_t* face_offsets = nullptr;
    const int64_t* neighbors = nullptr;
    const double* areas = nullptr;
    const uint32_t* status = nullptr;
    const uint64_t* vertex_offsets = nullptr;       // n_cells+1, or nullptr without TESS_OUT_VERTICES
    const double* vertices = nullptr;               // xyz triples, cell-local coordinates
    const uint64_t* face_vertex_offsets = nullptr;  // n_faces+1
    const uint32_t* face_vertex_indices = nullptr;  // ranks in the owning cell's vertex list
    tess_result* handle() const { return r_; }

   private:
    tess_result* r_;
};

class Cell;

/// interface.rs:25 — owns the particles, their groups, the grid and the container.
class Diagram {
   public:
    explicit Diagram(int device = 0) { check(tess_diagram_create(&d_, TESS_F64, device)); }
    ~Diagram() {
        batches_.clear();
        tess_diagram_destroy(d_);
    }
    Diagram(const Diagram&) = delete;
    Diagram& operator=(const Diagram&) = delete;

    /// interface.rs:52-57 (public here: SURVEY D3).  PointType exposes get_x/get_y/get_z (celery.rs:56-60).
    template <class PointType>
    void add_particle_with_group(const PointType& p, size_t group) {
        pending_.push_back({p.get_x(), p.get_y(), p.get_z()});
        groups_.push_back(static_cast<uint64_t>(group));
    }
    void add_particle_with_group(const Vector3& p, size_t group) {
        pending_.push_back(p);
        groups_.push_back(static_cast<uint64_t>(group));
    }
    /// batch extension: n records of `stride_bytes`, x,y,z (f64) first
    void add_particles(const void* xyz, size_t n, size_t stride_bytes = sizeof(Vector3), const uint64_t* groups = nullptr) {
        flush();
        check(tess_diagram_add_particles(d_, xyz, n, stride_bytes, groups, nullptr));
        n_ += n;
    }
    /// interface.rs:60-84; no container -> bounding box of the particles (:69-79)
    void initialize(const std::optional<Polyhedron>& container = std::nullopt) {
        flush();
        if (container) {
            const double box[6] = {container->x_min, container->y_min, container->z_min, container->x_max, container->y_max, container->z_max};
            check(tess_diagram_initialize(d_, box, nullptr));
            box_ = *container;
        } else {
            check(tess_diagram_initialize(d_, nullptr, nullptr));
            double b[6];
            check(tess_diagram_grid_info(d_, nullptr, nullptr, b, nullptr, nullptr));
            box_ = Polyhedron{b[0], b[2], b[4], b[1], b[3], b[5]};
        }
    }
    size_t len() const { return n_ + pending_.size(); }

    /// interface.rs:186-192
    Cell get_cell_at_index(size_t index, const Polyhedron& polyhedron, std::optional<double> search_radius = std::nullopt, std::optional<size_t> target_group = std::nullopt);
    /// interface.rs:211-217
    Cell get_cell_at_particle(const Vector3& point, const Polyhedron& polyhedron, std::optional<double> search_radius = std::nullopt, std::optional<size_t> target_group = std::nullopt);

    /// explicit batch fast path (extension): every cell in one call
    std::shared_ptr<CellBatch> compute_all_cells(std::optional<double> search_radius = std::nullopt, std::optional<size_t> target_group = std::nullopt,
                                                 bool with_vertices = false) {
        tess_opts o;
        tess_opts_default(&o);
        if (with_vertices) o.outputs |= TESS_OUT_VERTICES;
        if (search_radius) o.search_radius = *search_radius;
        if (target_group) o.target_group = static_cast<int64_t>(*target_group);
        tess_result* r = nullptr;
        check(tess_compute_all(d_, &o, &r));
        return std::make_shared<CellBatch>(r);
    }
    /// the same, with the results streamed into caller-owned (ideally page-locked) host arrays while later
    /// cells are still being computed (tess_compute_all_to_host); any pointer may be null
    std::shared_ptr<CellBatch> compute_all_cells_to_host(double* volumes, uint64_t* face_offsets, int64_t* neighbors, double* areas, uint32_t* status,
                                                         uint64_t cell_capacity, uint64_t face_capacity, int n_chunks = 0, std::optional<double> search_radius = std::nullopt,
                                                         std::optional<size_t> target_group = std::nullopt) {
        tess_opts o;
        tess_opts_default(&o);
        if (search_radius) o.search_radius = *search_radius;
        if (target_group) o.target_group = static_cast<int64_t>(*target_group);
        tess_result* r = nullptr;
        check(tess_compute_all_to_host(d_, &o, n_chunks, volumes, face_offsets, neighbors, (o.outputs & TESS_OUT_AREAS) ? areas : nullptr, status, cell_capacity, face_capacity, &r));
        return std::make_shared<CellBatch>(r);
    }
    std::shared_ptr<CellBatch> compute_cells_at(const Vector3* pts, size_t m, std::optional<double> search_radius, std::optional<size_t> target_group) {
        tess_opts o;
        tess_opts_default(&o);
        o.outputs |= TESS_OUT_VERTICES;
        if (search_radius) o.search_radius = *search_radius;
        if (target_group) o.target_group = static_cast<int64_t>(*target_group);
        tess_result* r = nullptr;
        check(tess_compute_at_points(d_, reinterpret_cast<const double*>(pts), m, &o, &r));
        return std::make_shared<CellBatch>(r);
    }
    const Polyhedron& container() const { return box_; }
    tess_diagram* handle() const { return d_; }

   private:
    friend class Cell;
    void flush() {
        if (pending_.empty()) return;
        check(tess_diagram_add_particles(d_, pending_.data(), pending_.size(), sizeof(Vector3), groups_.data(), nullptr));
        n_ += pending_.size();
        pending_.clear();
        groups_.clear();
    }
    void check_polyhedron(const Polyhedron& p) const {
        if (!(p == box_)) throw Error(TESS_ERR_UNSUPPORTED, "the start polyhedron of a cell must be the diagram's container box");
    }
    std::shared_ptr<CellBatch> batch(std::optional<double> radius, std::optional<size_t> group) {
        const auto key = std::make_pair(radius ? *radius : std::numeric_limits<double>::quiet_NaN(), group ? static_cast<int64_t>(*group) : int64_t(-1));
        for (auto& kv : batches_)
            if ((kv.first.first == key.first || (std::isnan(kv.first.first) && std::isnan(key.first))) && kv.first.second == key.second) return kv.second;
        auto b = compute_all_cells(radius, group, /*with_vertices=*/true);
        batches_.push_back({key, b});
        return b;
    }
    tess_diagram* d_ = nullptr;
    size_t n_ = 0;
    std::vector<Vector3> pending_;
    std::vector<uint64_t> groups_;
    Polyhedron box_{0, 0, 0, 0, 0, 0};
    std::vector<std::pair<std::pair<double, int64_t>, std::shared_ptr<CellBatch>>> batches_;
};

class VoronoiFace;

/// interface.rs:237
class Cell {
   public:
    /// interface.rs:257-313.  The first call for a given (search_radius, target_group) computes every
    /// cell of the diagram on the GPU; later cells read their row.
    void compute_voronoi_cell() {
        if (index_) {
            batch_ = diagram_->batch(radius_, group_);
            row_ = *index_;
        } else {
            batch_ = diagram_->compute_cells_at(&position_, 1, radius_, group_);
            row_ = 0;
        }
    }
    double compute_volume() { return need().volumes[row_]; }  // interface.rs:337-339
    std::vector<int64_t> compute_neighbors() {                // interface.rs:342-344 (walls: -1..-6)
        const CellBatch& b = need();
        return std::vector<int64_t>(b.neighbors + b.face_offsets[row_], b.neighbors + b.face_offsets[row_ + 1]);
    }
    std::vector<Vector3> compute_vertices() {                 // interface.rs:368-370 (cell-local coordinates)
        const CellBatch& b = need();
        std::vector<Vector3> out;
        if (!b.vertex_offsets) throw Error(TESS_ERR_STATE, "vertices were not computed");
        for (uint64_t v = b.vertex_offsets[row_]; v < b.vertex_offsets[row_ + 1]; ++v) out.push_back(Vector3{b.vertices[3 * v], b.vertices[3 * v + 1], b.vertices[3 * v + 2]});
        return out;
    }
    std::vector<VoronoiFace> compute_faces();                 // interface.rs:373-384
    std::optional<size_t> original_index() const { return index_; }  // interface.rs:387-389
    uint32_t status() { return need().status[row_]; }
    /// interface.rs:348-365: ExpandingSearch::expand_all_in_radius(radius) around the cell's position (particle ids in
    /// search-table order), filtered by group.  A cell made by index needs its position: pass the particle's coordinates
    /// (the mirror does not keep a host copy of the points).
    std::vector<size_t> compute_neighbor_cloud(const Vector3& position, double radius, std::optional<size_t> target_group = std::nullopt) const {
        tess_query* q = nullptr;
        check(tess_find_neighbors(diagram_->handle(), reinterpret_cast<const double*>(&position), 1, radius, TESS_QUERY_NEIGHBOR_CLOUD,
                                  target_group ? static_cast<int64_t>(*target_group) : int64_t(-1), nullptr, &q));
        const uint64_t* off = nullptr;
        const int64_t* idx = nullptr;
        std::vector<size_t> out;
        const int rc1 = tess_query_offsets(q, &off), rc2 = tess_query_indices(q, &idx);
        if (rc1 == TESS_OK && rc2 == TESS_OK) out.assign(idx + off[0], idx + off[1]);
        tess_query_free(q);
        check(rc1);
        check(rc2);
        return out;
    }
    /// ... for a cell made by get_cell_at_particle the position is the cell's own
    std::vector<size_t> compute_neighbor_cloud(double radius, std::optional<size_t> target_group = std::nullopt) const {
        if (index_) throw Error(TESS_ERR_STATE, "compute_neighbor_cloud(radius): a cell made by index must be given its particle's position");
        return compute_neighbor_cloud(position_, radius, target_group);
    }

   private:
    friend class Diagram;
    friend class VoronoiFace;
    Cell(Diagram* d, std::optional<size_t> index, Vector3 pos, std::optional<double> radius, std::optional<size_t> group)
        : diagram_(d), index_(index), position_(pos), radius_(radius), group_(group) {}
    const CellBatch& need() {
        if (!batch_) compute_voronoi_cell();
        return *batch_;
    }
    Diagram* diagram_;
    std::optional<size_t> index_;
    Vector3 position_;
    std::optional<double> radius_;
    std::optional<size_t> group_;
    std::shared_ptr<CellBatch> batch_;
    size_t row_ = 0;
};

/// interface.rs:393
class VoronoiFace {
   public:
    double compute_area() const { return batch_->areas[k_]; }          // interface.rs:408-410
    int64_t compute_neighbor() const { return batch_->neighbors[k_]; }  // interface.rs:413-416
    std::vector<Vector3> compute_vertices() const {                       // interface.rs:403-405: loop order
        const CellBatch& b = *batch_;
        if (!b.vertex_offsets) throw Error(TESS_ERR_STATE, "vertices were not computed");
        std::vector<Vector3> out;
        const uint64_t base = b.vertex_offsets[row_];
        for (uint64_t i = b.face_vertex_offsets[k_]; i < b.face_vertex_offsets[k_ + 1]; ++i) {
            const uint64_t v = base + b.face_vertex_indices[i];
            out.push_back(Vector3{b.vertices[3 * v], b.vertices[3 * v + 1], b.vertices[3 * v + 2]});
        }
        return out;
    }

   private:
    friend class Cell;
    VoronoiFace(std::shared_ptr<CellBatch> b, uint64_t k, uint64_t row) : batch_(std::move(b)), k_(k), row_(row) {}
    std::shared_ptr<CellBatch> batch_;
    uint64_t k_, row_;
};

inline std::vector<VoronoiFace> Cell::compute_faces() {
    const CellBatch& b = need();
    std::vector<VoronoiFace> out;
    for (uint64_t k = b.face_offsets[row_]; k < b.face_offsets[row_ + 1]; ++k) out.push_back(VoronoiFace(batch_, k, row_));
    return out;
}

inline Cell Diagram::get_cell_at_index(size_t index, const Polyhedron& polyhedron, std::optional<double> search_radius, std::optional<size_t> target_group) {
    check_polyhedron(polyhedron);
    if (index >= n_) throw Error(TESS_ERR_INVALID, "cell index out of range");
    return Cell(this, index, Vector3{}, search_radius, target_group);
}
inline Cell Diagram::get_cell_at_particle(const Vector3& point, const Polyhedron& polyhedron, std::optional<double> search_radius, std::optional<size_t> target_group) {
    check_polyhedron(polyhedron);
    return Cell(this, std::nullopt, point, search_radius, target_group);
}

/// celery.rs:865 `ExpandingSearch`: an outward walk over the grid cells around a position, resumable
/// (tess_search_create / tess_search_expand of tess.h).
class ExpandingSearch {
   public:
    /// ExpandingSearch::new (celery.rs:882-902)
    ExpandingSearch(const Diagram& d, const Vector3& position) { check(tess_search_create(d.handle(), reinterpret_cast<const double*>(&position), 1, &s_)); }
    ~ExpandingSearch() { tess_search_free(s_); }
    ExpandingSearch(const ExpandingSearch&) = delete;
    ExpandingSearch& operator=(const ExpandingSearch&) = delete;
    /// celery.rs:907-963: the particles of at most cells_to_add further grid cells, none beyond max_radius (squared cell distance)
    std::vector<size_t> expand(double max_radius, uint64_t cells_to_add) {
        tess_query* q = nullptr;
        check(tess_search_expand(s_, max_radius, cells_to_add, nullptr, &q));
        const uint64_t* off = nullptr;
        const int64_t* idx = nullptr;
        std::vector<size_t> out;
        const int rc1 = tess_query_offsets(q, &off), rc2 = tess_query_indices(q, &idx);
        if (rc1 == TESS_OK && rc2 == TESS_OK) out.assign(idx + off[0], idx + off[1]);
        tess_query_free(q);
        check(rc1);
        check(rc2);
        return out;
    }
    std::vector<size_t> expand_all_in_radius(double max_radius) { return expand(max_radius, UINT64_MAX); }                      // celery.rs:1023-1075
    std::vector<size_t> expand_all_no_radius() { return expand(std::numeric_limits<double>::infinity(), UINT64_MAX); }  // celery.rs:971-1018
    uint64_t current_search_index() const {
        const uint64_t* c = nullptr;
        check(tess_search_cursor(s_, &c));
        return c[0];
    }

   private:
    tess_search* s_ = nullptr;
};

/// Celery::find_cells_in_radius (celery.rs:753-797): ids of the grid cells within `radius` of the position
inline std::vector<size_t> find_cells_in_radius(const Diagram& d, const Vector3& position, double radius) {
    tess_query* q = nullptr;
    check(tess_find_cells_in_radius(d.handle(), reinterpret_cast<const double*>(&position), 1, radius, nullptr, &q));
    const uint64_t* off = nullptr;
    const int64_t* idx = nullptr;
    std::vector<size_t> out;
    const int rc1 = tess_query_offsets(q, &off), rc2 = tess_query_indices(q, &idx);
    if (rc1 == TESS_OK && rc2 == TESS_OK) out.assign(idx + off[0], idx + off[1]);
    tess_query_free(q);
    check(rc1);
    check(rc2);
    return out;
}

}  // namespace tess
