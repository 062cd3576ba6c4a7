// tess_oracle.cpp — CPU ORACLE (test infrastructure; see the header of tess_oracle.hpp).
// Restates /root/reference/src/{celery,polyhedron,interface}.rs; citations are file:line.
#include "tess_oracle.hpp"

#include <algorithm>
#include <stdexcept>

namespace orc {

// ==========================================================================================
// celery.rs
// ==========================================================================================

/// celery.rs:81-125 — min/max starting from the first point (panics on empty input there).
CeleryBounds Celery::make_bounds(const std::vector<Vec3>& pts) {
    if (pts.empty()) throw std::runtime_error("CeleryBounds::new on an empty point set (celery.rs:82)");
    double xmin = pts[0].x, ymin = pts[0].y, zmin = pts[0].z;
    double xmax = pts[0].x, ymax = pts[0].y, zmax = pts[0].z;
    for (const Vec3& p : pts) {
        if (p.x < xmin) xmin = p.x;
        if (p.x > xmax) xmax = p.x;
        if (p.y < ymin) ymin = p.y;
        if (p.y > ymax) ymax = p.y;
        if (p.z < zmin) zmin = p.z;
        if (p.z > zmax) zmax = p.z;
    }
    return {xmin, xmax, ymin, ymax, zmin, zmax};
}

/// celery.rs:153-189 — cpd = floor(cbrt(N / 1.25)) + 1 with libm cbrt and a saturating cast.
CeleryCellInfo Celery::make_cell_info(size_t n, const CeleryBounds& b) {
    const double num_points = static_cast<double>(n);
    const size_t cpd = to_usize(std::cbrt(num_points / 1.25)) + 1;
    const double c = static_cast<double>(cpd);
    CeleryCellInfo ci;
    ci.x_cell_size = (b.x_max - b.x_min) / c;
    ci.y_cell_size = (b.y_max - b.y_min) / c;
    ci.z_cell_size = (b.z_max - b.z_min) / c;
    ci.x_inverse_cell_size = c / (b.x_max - b.x_min);
    ci.y_inverse_cell_size = c / (b.y_max - b.y_min);
    ci.z_inverse_cell_size = c / (b.z_max - b.z_min);
    ci.cells_per_dimension = cpd;
    return ci;
}

/// celery.rs:269-314 — identical for the three axes.
size_t Celery::axis_index(double v, double vmin, double vmax, double inv, size_t cpd) {
    if (v >= vmax) return cpd - 1;
    const double index = (v - vmin) * inv;
    return std::min(to_usize(index), cpd - 1);
}

/// celery.rs:253-266 (and 231-250).
void Celery::reset(int table_radius_) {
    bounds = make_bounds(points);
    cell_info = make_cell_info(points.size(), bounds);

    // celery.rs:342-354
    cells.resize(points.size());
    for (size_t i = 0; i < points.size(); ++i) cells[i] = get_cell(points[i]);

    // celery.rs:357-369 — sort_unstable_by(cell); canonical tie-break by index (D12).
    sorted_indices.resize(points.size());
    for (size_t i = 0; i < points.size(); ++i) sorted_indices[i] = i;
    std::stable_sort(sorted_indices.begin(), sorted_indices.end(),
                     [&](size_t a, size_t b) { return cells[a] < cells[b]; });

    // celery.rs:372-414
    {
        const size_t num_points = cells.size();
        const size_t cpd = cell_info.cells_per_dimension;
        const size_t last_cell = cpd * cpd * cpd;
        delimiters.clear();
        delimiters.reserve(last_cell + 1);
        delimiters.push_back(0);
        bool returned = false;
        for (size_t i = 0; i < last_cell && !returned; ++i) {
            size_t offset = 0;
            const size_t last_delimiter = delimiters.back();
            while (i == cells[sorted_indices[last_delimiter + offset]]) {
                offset += 1;
                if (last_delimiter + offset == num_points) {
                    for (size_t k = i; k < last_cell; ++k) delimiters.push_back(num_points);
                    returned = true;
                    break;
                }
            }
            if (!returned) delimiters.push_back(last_delimiter + offset);
        }
        if (!returned) delimiters.push_back(num_points);
    }

    table_radius = table_radius_;
    build_search_order();
}

/// celery.rs:418-679.  The reference pushes the 26 sign/axis combinations in a fixed order and
/// then sort_unstable()s by distance only; ties are therefore unspecified (D12).  Here the same
/// multiset of entries is generated and sorted canonically by (distance, i, j, k).
void Celery::build_search_order() {
    const CeleryCellInfo& ci = cell_info;
    auto sq = [](double x) { return x * x; };
    auto distance = [&](int32_t i, int32_t j, int32_t k) {  // celery.rs:423-427
        return sq(static_cast<double>(i) * ci.x_cell_size) + sq(static_cast<double>(j) * ci.y_cell_size) +
               sq(static_cast<double>(k) * ci.z_cell_size);
    };
    const int32_t max_index = static_cast<int32_t>(ci.cells_per_dimension) - 1;  // celery.rs:430
    const bool full = (table_radius < 0 || table_radius >= max_index);
    const int32_t lim = full ? max_index : table_radius;
    table_is_full = full;

    search_order.clear();
    search_order.push_back({-1.0, 0, 0, 0});  // celery.rs:437-442
    const int32_t sgn[2] = {1, -1};
    // celery.rs:445-509 (three non-zero offsets), 513-615 (two), 619-673 (one).  Offset o != 0
    // carries the distance term of |o|-1 (adjacent cells are at distance 0).
    for (int32_t i = 0; i < lim; ++i)
        for (int32_t j = 0; j < lim; ++j)
            for (int32_t k = 0; k < lim; ++k) {
                const double d = distance(i, j, k);
                for (int32_t si : sgn) for (int32_t sj : sgn) for (int32_t sk : sgn)
                    search_order.push_back({d, si * (i + 1), sj * (j + 1), sk * (k + 1)});
            }
    for (int32_t i = 0; i < lim; ++i)
        for (int32_t j = 0; j < lim; ++j) {
            const double d = distance(i, j, 0);
            for (int32_t si : sgn) for (int32_t sj : sgn) search_order.push_back({d, si * (i + 1), sj * (j + 1), 0});
        }
    for (int32_t i = 0; i < lim; ++i)
        for (int32_t k = 0; k < lim; ++k) {
            const double d = distance(i, 0, k);
            for (int32_t si : sgn) for (int32_t sk : sgn) search_order.push_back({d, si * (i + 1), 0, sk * (k + 1)});
        }
    for (int32_t j = 0; j < lim; ++j)
        for (int32_t k = 0; k < lim; ++k) {
            const double d = distance(0, j, k);
            for (int32_t sj : sgn) for (int32_t sk : sgn) search_order.push_back({d, 0, sj * (j + 1), sk * (k + 1)});
        }
    for (int32_t i = 0; i < lim; ++i) {
        const double d = distance(i, 0, 0);
        for (int32_t si : sgn) search_order.push_back({d, si * (i + 1), 0, 0});
    }
    for (int32_t j = 0; j < lim; ++j) {
        const double d = distance(0, j, 0);
        for (int32_t sj : sgn) search_order.push_back({d, 0, sj * (j + 1), 0});
    }
    for (int32_t k = 0; k < lim; ++k) {
        const double d = distance(0, 0, k);
        for (int32_t sk : sgn) search_order.push_back({d, 0, 0, sk * (k + 1)});
    }

    // celery.rs:676 — sort by distance (float.rs:158-170 order); canonical tie-break (D12).
    std::sort(search_order.begin(), search_order.end(), [](const DistanceIndex& a, const DistanceIndex& b) {
        const int c = float_cmp(a.distance, b.distance);
        if (c != 0) return c < 0;
        if (a.i != b.i) return a.i < b.i;
        if (a.j != b.j) return a.j < b.j;
        return a.k < b.k;
    });

    if (!full) {
        // Offsets left out have some |o| >= lim+1, hence distance >= min_axis sq(lim*size).
        // Entries strictly below that bound form an exact prefix of the full table.
        const double lx = sq(static_cast<double>(lim) * ci.x_cell_size);
        const double ly = sq(static_cast<double>(lim) * ci.y_cell_size);
        const double lz = sq(static_cast<double>(lim) * ci.z_cell_size);
        const double bound = std::min(lx, std::min(ly, lz));
        size_t keep = 0;
        while (keep < search_order.size() && search_order[keep].distance < bound) ++keep;
        search_order.resize(keep);
    }
}

/// celery.rs:708-743
bool Celery::check_cell_in_range(double x, double y, double z, double radius, size_t x_index_, size_t y_index_, size_t z_index_) const {
    auto sq = [](double v) { return v * v; };
    auto distance_sq = [&](int32_t i, int32_t j, int32_t k) {
        return sq(static_cast<double>(i) * cell_info.x_cell_size) + sq(static_cast<double>(j) * cell_info.y_cell_size) +
               sq(static_cast<double>(k) * cell_info.z_cell_size);
    };
    auto offset = [](size_t coord, size_t index) {
        return std::max<int32_t>(0, std::abs(static_cast<int32_t>(coord) - static_cast<int32_t>(index)) - 1);
    };
    const size_t xi = x_index(x), yi = y_index(y), zi = z_index(z);
    const double ds = distance_sq(offset(xi, x_index_), offset(yi, y_index_), offset(zi, z_index_));
    return ds <= radius * radius;
}

/// celery.rs:753-797
std::vector<size_t> Celery::find_cells_in_radius(double x, double y, double z, double radius) const {
    auto max_float = [](double a, double b) { return a > b ? a : b; };  // celery.rs:691-697
    auto min_float = [](double a, double b) { return a < b ? a : b; };  // celery.rs:682-688
    const double x_low = max_float(x - radius, bounds.x_min), y_low = max_float(y - radius, bounds.y_min), z_low = max_float(z - radius, bounds.z_min);
    const double x_high = min_float(x + radius, bounds.x_max), y_high = min_float(y + radius, bounds.y_max), z_high = min_float(z + radius, bounds.z_max);
    const size_t x0 = x_index(x_low), y0 = y_index(y_low), z0 = z_index(z_low);
    const size_t x1 = x_index(x_high), y1 = y_index(y_high), z1 = z_index(z_high);
    std::vector<size_t> out;
    for (size_t i = x0; i <= x1; ++i)
        for (size_t j = y0; j <= y1; ++j)
            for (size_t k = z0; k <= z1; ++k)
                if (check_cell_in_range(x, y, z, radius, i, j, k)) out.push_back(cell_from_indices(i, j, k));
    return out;
}

/// celery.rs:802-819
std::vector<size_t> Celery::find_neighbors_in_cell_radius(double x, double y, double z, double radius) const {
    std::vector<size_t> out;
    for (size_t c : find_cells_in_radius(x, y, z, radius))
        for (size_t s = delimiters[c]; s < delimiters[c + 1]; ++s) out.push_back(sorted_indices[s]);
    return out;
}

/// celery.rs:825-855 (distance_squared: celery.rs:700-702)
std::vector<size_t> Celery::find_neighbors_in_real_radius(double x, double y, double z, double radius) const {
    std::vector<size_t> out;
    for (size_t c : find_cells_in_radius(x, y, z, radius))
        for (size_t s = delimiters[c]; s < delimiters[c + 1]; ++s) {
            const size_t pi = sorted_indices[s];
            const Vec3& p = points[pi];
            const double d2 = (x - p.x) * (x - p.x) + (y - p.y) * (y - p.y) + (z - p.z) * (z - p.z);
            if (d2 <= radius * radius) out.push_back(pi);
        }
    return out;
}

/// celery.rs:882-902
ExpandingSearch::ExpandingSearch(const Celery& c, double x, double y, double z)
    : celery(&c), x_cell_index(c.x_index(x)), y_cell_index(c.y_index(y)), z_cell_index(c.z_index(z)) {}

/// celery.rs:930-959 (== 985-1014 == 1042-1071)
void ExpandingSearch::append_entry(const DistanceIndex& e, std::vector<size_t>& out) const {
    const int32_t cpd = static_cast<int32_t>(celery->cell_info.cells_per_dimension);
    const int32_t xs = static_cast<int32_t>(x_cell_index) + e.i;
    const int32_t ys = static_cast<int32_t>(y_cell_index) + e.j;
    const int32_t zs = static_cast<int32_t>(z_cell_index) + e.k;
    if (xs < 0 || xs >= cpd || ys < 0 || ys >= cpd || zs < 0 || zs >= cpd) return;
    const size_t cell = celery->cell_from_indices(static_cast<size_t>(xs), static_cast<size_t>(ys), static_cast<size_t>(zs));
    for (size_t s = celery->delimiters[cell]; s < celery->delimiters[cell + 1]; ++s) out.push_back(celery->sorted_indices[s]);
}

/// celery.rs:907-963
std::vector<size_t> ExpandingSearch::expand(double max_radius, size_t cells_to_add) {
    std::vector<size_t> out;
    const auto& so = celery->search_order;
    for (size_t n = 0; n < cells_to_add; ++n) {
        if (current_search_index >= so.size()) return out;
        const DistanceIndex& e = so[current_search_index];
        if (e.distance > max_radius) return out;  // squared vs unsquared: D11, kept literally
        current_search_index += 1;
        append_entry(e, out);
    }
    return out;
}

/// celery.rs:971-1018
std::vector<size_t> ExpandingSearch::expand_all_no_radius() {
    std::vector<size_t> out;
    const auto& so = celery->search_order;
    while (current_search_index < so.size()) {
        const DistanceIndex& e = so[current_search_index];
        current_search_index += 1;
        append_entry(e, out);
    }
    return out;
}

/// celery.rs:1023-1075
std::vector<size_t> ExpandingSearch::expand_all_in_radius(double max_radius) {
    std::vector<size_t> out;
    const auto& so = celery->search_order;
    while (current_search_index < so.size()) {
        const DistanceIndex& e = so[current_search_index];
        if (e.distance > max_radius) return out;  // D11
        current_search_index += 1;
        append_entry(e, out);
    }
    return out;
}

// ==========================================================================================
// polyhedron.rs
// ==========================================================================================

/// polyhedron.rs:268-392 — fixed numbering of the start cube (enums at :73-199).
void Polyhedron::reset(double x_min, double y_min, double z_min, double x_max, double y_max, double z_max) {
    vertices = Pool<Vec3>{};
    faces = Pool<Face>{};
    edges = Pool<HalfEdge>{};
    face_data.clear();
    counters = CutCounters{};

    // polyhedron.rs:288-295: FDL, FDR, FUR, FUL, BDL, BDR, BUR, BUL
    vertices.add({x_min, y_min, z_min});
    vertices.add({x_max, y_min, z_min});
    vertices.add({x_max, y_min, z_max});
    vertices.add({x_min, y_min, z_max});
    vertices.add({x_min, y_max, z_min});
    vertices.add({x_max, y_max, z_min});
    vertices.add({x_max, y_max, z_max});
    vertices.add({x_min, y_max, z_max});

    // polyhedron.rs:319-383: per face one Face then four HalfEdges (face, flip, target, next).
    // Edge ids: FU0 FL1 FD2 FR3 | RU4 RF5 RD6 RB7 | BU8 BR9 BD10 BL11 | LU12 LB13 LD14 LF15 |
    //           UF16 UR17 UB18 UL19 | DF20 DL21 DB22 DR23.   Vertex ids as above.
    static const int E[24][3] = {
        // flip, target, next
        {16, 3, 1},  {15, 0, 2},  {20, 1, 3},  {5, 2, 0},    // F: FU FL FD FR   (:320-323)
        {17, 2, 5},  {3, 1, 6},   {23, 5, 7},  {9, 6, 4},    // R: RU RF RD RB   (:332-335)
        {18, 6, 9},  {7, 5, 10},  {22, 4, 11}, {13, 7, 8},   // B: BU BR BD BL   (:344-347)
        {19, 7, 13}, {11, 4, 14}, {21, 0, 15}, {1, 3, 12},   // L: LU LB LD LF   (:356-359)
        {0, 2, 17},  {4, 6, 18},  {8, 7, 19},  {12, 3, 16},  // U: UF UR UB UL   (:368-371)
        {2, 0, 21},  {14, 4, 22}, {10, 5, 23}, {6, 1, 20},   // D: DF DL DB DR   (:380-383; DR's face is D, D5)
    };
    for (int f = 0; f < 6; ++f) {
        Face face;
        face.point_index.reset();                             // polyhedron.rs:306-309
        face.starting_edge_index = static_cast<size_t>(4 * f);  // FU, RU, BU, LU, UF, DF
        face.wall = f;
        faces.add(face);
        for (int k = 0; k < 4; ++k) {
            const int e = 4 * f + k;
            HalfEdge he;
            he.flip = static_cast<size_t>(E[e][0]);
            he.target = static_cast<size_t>(E[e][1]);
            he.next = static_cast<size_t>(E[e][2]);
            he.face = static_cast<size_t>(f);
            edges.add(he);
        }
    }
    root_edge = 0;  // polyhedron.rs:391 (FU)
}

/// polyhedron.rs:396-435
OptIdx Polyhedron::find_outgoing_edge(const Plane& plane) {
    // :399-405 — any vertex Outside?  (first hit breaks the scan)
    bool need_to_cut = false;
    for (size_t i = 0; i < vertices.len(); ++i) {
        const Vec3* v = vertices.get(i);
        if (!v) continue;
        if (plane.vector_location(*v, tolerance()) == Loc::Outside) {
            need_to_cut = true;
            break;
        }
    }
    if (!need_to_cut) return {};
    // :413-432 — first edge in slot order whose target is Inside and whose flip's target is Outside
    for (size_t i = 0; i < edges.len(); ++i) {
        const HalfEdge* e = edges.get(i);
        if (!e) continue;
        const Vec3& target = vertices.at(*e->target);
        if (plane.vector_location(target, tolerance()) == Loc::Inside) {
            const Vec3& flip_target = vertices.at(*target_index(e->flip));
            if (plane.vector_location(flip_target, tolerance()) == Loc::Outside) return e->flip;
        }
    }
    return {};  // D17: material lies outside but no strictly Inside->Outside edge exists
}

/// polyhedron.rs:438-642, with the face loop closed (D6) and single frees (D7/D8).
bool Polyhedron::cut_with_plane(size_t point_index, const Plane& plane) {
    // Work counter VC: one classification per live vertex per plane offered.
    counters.vertex_classifications += vertices.live();
    const OptIdx found = find_outgoing_edge(plane);
    if (!found) {
        // Distinguish "nothing outside" from the D17 skip: the vertex scan broke early iff
        // something was Outside.
        bool any_outside = false;
        for (size_t i = 0; i < vertices.len() && !any_outside; ++i) {
            const Vec3* v = vertices.get(i);
            if (v && plane.vector_location(*v, tolerance()) == Loc::Outside) any_outside = true;
        }
        if (any_outside) counters.degenerate_skips++;
        return false;
    }
    const size_t first_outgoing_edge_index = *found;
    root_edge = first_outgoing_edge_index;  // :473

    size_t outgoing_edge_index = first_outgoing_edge_index;
    OptIdx previous_intersection;

    // :478-484
    const size_t first_outside_face_edge_index = edges.add(HalfEdge{});
    Face cap;
    cap.point_index = point_index;
    cap.starting_edge_index = first_outside_face_edge_index;
    cap.wall = -1;
    const size_t outside_face_index = faces.add(cap);
    size_t outside_face_edge_index = first_outside_face_edge_index;
    edges.at(first_outside_face_edge_index).face = outside_face_index;

    std::vector<size_t> vertices_to_destroy;  // :487 (deduplicated when consumed, D7)

    for (;;) {
        OptIdx previous_vertex_index = edges.at(outgoing_edge_index).target;  // :491
        vertices_to_destroy.push_back(*previous_vertex_index);

        OptIdx current_edge_index = edges.at(outgoing_edge_index).next;  // :506
        OptIdx current_vertex_index = target_index(current_edge_index);

        Loc previous_location = plane.vector_location(vertices.at(*previous_vertex_index), tolerance());  // :509-514
        bool need_to_cut = previous_location == Loc::Outside;                                               // :519
        Loc current_location = plane.vector_location(vertices.at(*current_vertex_index), tolerance());     // :521-526

        while (current_location != Loc::Inside) {  // :529-544
            need_to_cut = true;
            vertices_to_destroy.push_back(*current_vertex_index);
            previous_vertex_index = current_vertex_index;
            previous_location = current_location;
            current_edge_index = next_index(current_edge_index);
            current_vertex_index = target_index(current_edge_index);
            current_location = plane.vector_location(vertices.at(*current_vertex_index), tolerance());
        }

        edges.at(outgoing_edge_index).target = previous_intersection;  // :550 (None on the first face; closed below)

        if (need_to_cut) {  // :552-601
            size_t current_intersection_vertex_index;
            if (previous_location == Loc::Incident) {
                const Vec3 old_vertex = vertices.at(*previous_vertex_index);  // :555-565 copy
                current_intersection_vertex_index = vertices.add(old_vertex);
            } else {
                const Vec3 vertex = plane.intersection(vertices.at(*previous_vertex_index), vertices.at(*current_vertex_index));  // :567-572
                current_intersection_vertex_index = vertices.add(vertex);
            }
            counters.new_vertices++;

            const OptIdx current_face_index = edges.at(outgoing_edge_index).face;  // :575
            faces.at(*current_face_index).starting_edge_index = outgoing_edge_index;  // :578-580

            HalfEdge bridge;  // :582-587
            bridge.face = current_face_index;
            bridge.target = current_intersection_vertex_index;
            bridge.flip = outside_face_edge_index;
            bridge.next = current_edge_index;
            const size_t bridge_edge_index = edges.add(bridge);

            edges.at(outside_face_edge_index).flip = bridge_edge_index;  // :589
            edges.at(outgoing_edge_index).next = bridge_edge_index;      // :590

            HalfEdge cap_edge;  // :592-598
            cap_edge.face = outside_face_index;
            cap_edge.target = current_intersection_vertex_index;
            cap_edge.flip.reset();
            cap_edge.next = outside_face_edge_index;
            outside_face_edge_index = edges.add(cap_edge);

            previous_intersection = current_intersection_vertex_index;  // :600
        }

        outgoing_edge_index = *edges.at(*current_edge_index).flip;  // :603-607
        if (outgoing_edge_index == first_outgoing_edge_index) break;  // :620-622
    }

    // ---- D6: close the loop -------------------------------------------
    // The first outgoing edge never received its target, the first cap edge has no target/next,
    // and the cap edge created by the last face is the redundant twin of the first one.
    const size_t redundant_cap_edge = outside_face_edge_index;
    const OptIdx last_paired_cap_edge = edges.at(redundant_cap_edge).next;
    edges.at(first_outgoing_edge_index).target = previous_intersection;
    edges.at(first_outside_face_edge_index).target = previous_intersection;
    edges.at(first_outside_face_edge_index).next = last_paired_cap_edge;
    edges.remove(redundant_cap_edge);

    counters.cuts++;
    clean_up(vertices_to_destroy);  // :638-639
    return true;
}

/// polyhedron.rs:645-730, repaired (D7/D8): free the walked vertices once each, then free every
/// edge / face / vertex that is no longer reachable from root_edge through flip/next.
void Polyhedron::clean_up(const std::vector<size_t>& vertices_to_destroy) {
    for (size_t v : vertices_to_destroy)
        if (vertices.has(v)) vertices.remove(v);

    std::vector<char> edge_alive(edges.len(), 0);
    std::vector<size_t> stack;
    stack.push_back(*root_edge);
    edge_alive[*root_edge] = 1;
    while (!stack.empty()) {
        const size_t e = stack.back();
        stack.pop_back();
        const HalfEdge& he = edges.at(e);
        for (const OptIdx& n : {he.flip, he.next}) {
            if (n && !edge_alive[*n]) {
                edge_alive[*n] = 1;
                stack.push_back(*n);
            }
        }
    }
    std::vector<char> face_alive(faces.len(), 0), vertex_alive(vertices.len(), 0);
    for (size_t e = 0; e < edges.len(); ++e) {
        if (!edge_alive[e]) continue;
        const HalfEdge& he = edges.at(e);
        face_alive[*he.face] = 1;
        vertex_alive[*he.target] = 1;
    }
    for (size_t e = 0; e < edges.len(); ++e)
        if (edges.has(e) && !edge_alive[e]) edges.remove(e);
    for (size_t f = 0; f < faces.len(); ++f)
        if (faces.has(f) && !face_alive[f]) faces.remove(f);
    for (size_t v = 0; v < vertices.len(); ++v)
        if (vertices.has(v) && !vertex_alive[v]) vertices.remove(v);
}

/// polyhedron.rs:776-808
Vec3 Polyhedron::weighted_normal(size_t face_index) const {
    Vec3 normal{0, 0, 0};
    const size_t starting_edge_index = faces.at(face_index).starting_edge_index;
    const HalfEdge& starting_edge = edges.at(starting_edge_index);
    const Vec3& anchor = vertices.at(*starting_edge.target);

    size_t current_edge_index = *starting_edge.next;
    const HalfEdge* current_edge = &edges.at(current_edge_index);
    Vec3 current_vector = sub(vertices.at(*current_edge->target), anchor);
    current_edge_index = *current_edge->next;
    current_edge = &edges.at(current_edge_index);

    while (current_edge_index != starting_edge_index) {
        const Vec3 previous_vector = current_vector;
        current_vector = sub(vertices.at(*current_edge->target), anchor);
        normal = add(normal, cross(previous_vector, current_vector));
        current_edge_index = *current_edge->next;
        current_edge = &edges.at(current_edge_index);
    }
    return normal;
}

/// polyhedron.rs:812-825 (D13: only ever called after the last cut)
void Polyhedron::compute_face_data() {
    if (!face_data.empty()) return;
    for (size_t i = 0; i < faces.len(); ++i)
        if (faces.has(i)) face_data.push_back({i, weighted_normal(i)});
}

/// polyhedron.rs:838-855
double Polyhedron::compute_volume() {
    compute_face_data();
    double volume = 0;
    for (const FaceData& fd : face_data) {
        const Face& face = faces.at(fd.face_index);
        const HalfEdge& starting_edge = edges.at(face.starting_edge_index);
        const Vec3& target_vertex = vertices.at(*starting_edge.target);
        volume = volume + dot(target_vertex, fd.weighted_normal);
    }
    return volume / 6.0;
}

/// polyhedron.rs:859-867 (Vector3::add, vector3.rs:87-91)
void Polyhedron::translate(const Vec3& shift) {
    for (size_t i = 0; i < vertices.len(); ++i) {
        Vec3* v = vertices.get(i);
        if (v) {
            v->x = v->x + shift.x;
            v->y = v->y + shift.y;
            v->z = v->z + shift.z;
        }
    }
}

/// polyhedron.rs:871-881; container faces report -(wall+1) instead of panicking (D10).
std::vector<int64_t> Polyhedron::compute_neighbors() const {
    std::vector<int64_t> out;
    for (size_t i = 0; i < faces.len(); ++i) {
        const Face* f = faces.get(i);
        if (!f) continue;
        out.push_back(f->point_index ? static_cast<int64_t>(*f->point_index) : -static_cast<int64_t>(f->wall + 1));
    }
    return out;
}

/// polyhedron.rs:885-893
std::vector<Vec3> Polyhedron::compute_vertices() const {
    std::vector<Vec3> out;
    for (size_t i = 0; i < vertices.len(); ++i)
        if (const Vec3* v = vertices.get(i)) out.push_back(*v);
    return out;
}

/// polyhedron.rs:897-919
std::vector<Vec3> Polyhedron::compute_face_vertices(size_t face_index) const {
    std::vector<Vec3> out;
    const size_t start = faces.at(face_index).starting_edge_index;
    size_t current = start;
    for (;;) {
        const HalfEdge& e = edges.at(current);
        out.push_back(vertices.at(*e.target));
        current = *e.next;
        if (current == start) break;
    }
    return out;
}

double Polyhedron::max_vertex_radius_sq() const {
    double m = 0;
    for (size_t i = 0; i < vertices.len(); ++i)
        if (const Vec3* v = vertices.get(i)) {
            const double r2 = mag_sq(*v);
            if (r2 > m) m = r2;
        }
    return m;
}

// ==========================================================================================
// interface.rs
// ==========================================================================================

/// interface.rs:52-57; the bounding box starts at the first point (D14).
void Diagram::add_particle_with_group(const Vec3& p, size_t group) {
    if (cell_array.points.empty()) {
        bounding_box.low = p;
        bounding_box.high = p;
    }
    bounding_box.adjust_to_contain(p.x, p.y, p.z);
    cell_array.points.push_back(p);
    groups.push_back(group);
}

/// interface.rs:60-84 (D1: container = given box, else the bounding box; D4: no permutation)
void Diagram::initialize(const double* container_box, int table_radius) {
    if (container_box) {
        for (int i = 0; i < 6; ++i) box[i] = container_box[i];
    } else {
        box[0] = bounding_box.low.x;  box[1] = bounding_box.low.y;  box[2] = bounding_box.low.z;
        box[3] = bounding_box.high.x; box[4] = bounding_box.high.y; box[5] = bounding_box.high.z;
    }
    cell_array.reset(table_radius);
    initialized = true;
}

CellResult Diagram::compute_cell_at_index(size_t index, SearchMode mode, double search_radius, int64_t target_group, bool want_vertices) const {
    return compute(cell_array.points[index], index, mode, search_radius, target_group, want_vertices);  // interface.rs:193-207
}

CellResult Diagram::compute_cell_at_point(const Vec3& position, SearchMode mode, double search_radius, int64_t target_group, bool want_vertices) const {
    return compute(position, {}, mode, search_radius, target_group, want_vertices);  // interface.rs:218-231
}

/// interface.rs:257-334 + 337-344 + 408-410
CellResult Diagram::compute(const Vec3& position, OptIdx self, SearchMode mode, double search_radius, int64_t target_group, bool want_vertices) const {
    CellResult r;
    Polyhedron polyhedron(box[0], box[1], box[2], box[3], box[4], box[5]);
    polyhedron.translate(neg(position));  // interface.rs:266

    // interface.rs:316-334
    auto cut_with_point = [&](size_t q) -> bool {
        const Vec3 rel = sub(cell_array.points[q], position);
        r.counters.tested++;
        return polyhedron.cut_with_plane(q /* original index, D4 */, Plane::halfway_from_origin_to(rel));
    };
    // interface.rs:280-312 — the four (target_group, index) arms collapse to two predicates
    auto admissible = [&](size_t q) {
        if (self && q == *self) return false;
        if (target_group >= 0 && groups[q] != static_cast<size_t>(target_group)) return false;
        return true;
    };

    ExpandingSearch es(cell_array, position.x, position.y, position.z);  // interface.rs:268-273
    if (mode == MODE_NO_RADIUS || mode == MODE_REFERENCE_RADIUS) {
        const std::vector<size_t> search_points =
            (mode == MODE_REFERENCE_RADIUS) ? es.expand_all_in_radius(search_radius) : es.expand_all_no_radius();  // interface.rs:275-278
        r.counters.table_entries = es.current_search_index;
        r.counters.visited = search_points.size();
        for (size_t q : search_points)
            if (admissible(q)) cut_with_point(q);
        if (mode == MODE_NO_RADIUS && !cell_array.table_is_full) r.status |= ORC_STATUS_TABLE_EXHAUSTED;
    } else {
        double rmax2 = polyhedron.max_vertex_radius_sq();
        const auto& so = cell_array.search_order;
        bool terminated = false;
        std::vector<size_t> pts;
        for (size_t t = 0; t < so.size(); ++t) {
            if (so[t].distance > 4.0 * rmax2) {
                terminated = true;
                break;
            }
            r.counters.table_entries++;
            pts.clear();
            es.append_entry(so[t], pts);
            for (size_t q : pts) {
                r.counters.visited++;
                if (!admissible(q)) continue;
                const Vec3 rel = sub(cell_array.points[q], position);
                if (mag_sq(rel) >= 4.0 * rmax2) continue;
                if (cut_with_point(q)) rmax2 = polyhedron.max_vertex_radius_sq();
            }
        }
        if (!terminated && !cell_array.table_is_full) r.status |= ORC_STATUS_TABLE_EXHAUSTED;
    }

    r.volume = polyhedron.compute_volume();          // interface.rs:337-339
    r.neighbors = polyhedron.compute_neighbors();    // interface.rs:342-344
    for (size_t f = 0; f < polyhedron.faces.len(); ++f)
        if (polyhedron.faces.has(f)) r.areas.push_back(0.5 * mag(polyhedron.weighted_normal(f)));  // interface.rs:408-410
    if (want_vertices) {
        r.vertices = polyhedron.compute_vertices();  // interface.rs:368-370
        for (size_t f = 0; f < polyhedron.faces.len(); ++f)
            if (polyhedron.faces.has(f)) {  // VoronoiFace::compute_vertices (interface.rs:403-405)
                const std::vector<Vec3> loop = polyhedron.compute_face_vertices(f);
                r.face_loop_sizes.push_back(static_cast<uint32_t>(loop.size()));
                r.face_loop_vertices.insert(r.face_loop_vertices.end(), loop.begin(), loop.end());
            }
    }
    r.max_radius_sq = polyhedron.max_vertex_radius_sq();
    r.pool_slots[0] = static_cast<uint32_t>(polyhedron.vertices.len());
    r.pool_slots[1] = static_cast<uint32_t>(polyhedron.edges.len());
    r.pool_slots[2] = static_cast<uint32_t>(polyhedron.faces.len());
    r.counters.vertex_classifications = polyhedron.counters.vertex_classifications;
    r.counters.cuts = polyhedron.counters.cuts;
    r.counters.new_vertices = polyhedron.counters.new_vertices;
    r.counters.degenerate_skips = polyhedron.counters.degenerate_skips;
    if (polyhedron.counters.degenerate_skips) r.status |= ORC_STATUS_DEGENERATE_SKIP;
    return r;
}

}  // namespace orc
