// C++ harness over include/tess.hpp: the reference's usage pattern (interface.rs) end to end.
// Built and run by tests/test_gpu_cpp_mirror.py on the GPU box.  Prints one line per cell checked;
// the Python side compares the numbers with the CPU oracle.
#include <cstdio>
#include <cstdlib>

#include "tess.hpp"

// a user point type exposing the ToCeleryPoint getters (celery.rs:56-60)
struct MyParticle {
    double px, py, pz;
    int payload;
    double get_x() const { return px; }
    double get_y() const { return py; }
    double get_z() const { return pz; }
};

static uint64_t mix(uint64_t z) {
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
static double u01(uint64_t seed, uint64_t i, uint64_t c) { return (double)(mix(mix(seed) + 3 * i + c) >> 11) * 0x1.0p-53; }

int main(int argc, char** argv) {
    const size_t n = argc > 1 ? std::strtoull(argv[1], nullptr, 10) : 3000;
    try {
        tess::Diagram diagram;  // Diagram::default()
        for (size_t i = 0; i < n; ++i) diagram.add_particle_with_group(MyParticle{u01(62, i, 0), u01(62, i, 1), u01(62, i, 2), (int)i}, 0);
        const tess::Polyhedron box{0, 0, 0, 1, 1, 1};  // Polyhedron::new(0,0,0,1,1,1)
        diagram.initialize(box);
        double total = 0;
        for (size_t i = 0; i < n; ++i) {
            tess::Cell cell = diagram.get_cell_at_index(i, box, std::nullopt, std::nullopt);
            cell.compute_voronoi_cell();
            const double v = cell.compute_volume();
            total += v;
            if (i % 500 == 7) {
                std::printf("cell %zu volume %.17g faces", i, v);
                size_t loop_total = 0;
                for (const tess::VoronoiFace& f : cell.compute_faces()) {
                    std::printf(" %lld:%.17g", (long long)f.compute_neighbor(), f.compute_area());
                    loop_total += f.compute_vertices().size();
                }
                std::printf("\n");
                // Euler: sum of face loop lengths = 2E, V - E + F = 2
                const size_t V = cell.compute_vertices().size(), F = cell.compute_faces().size();
                if (loop_total % 2 != 0 || V + F != loop_total / 2 + 2) {
                    std::printf("ERROR: Euler characteristic violated for cell %zu (V=%zu F=%zu 2E=%zu)\n", i, V, F, loop_total);
                    return 3;
                }
            }
        }
        std::printf("total %.17g\n", total);
        // a cell around a point that is not a particle (interface.rs:211-232)
        tess::Cell q = diagram.get_cell_at_particle(tess::Vector3{0.31, 0.62, 0.44}, box);
        std::printf("query volume %.17g nfaces %zu\n", q.compute_volume(), q.compute_neighbors().size());
        // ExpandingSearch in steps (celery.rs:907-963) must visit what one sweep visits, in the same order
        {
            tess::ExpandingSearch a(diagram, tess::Vector3{0.31, 0.62, 0.44}), b(diagram, tess::Vector3{0.31, 0.62, 0.44});
            std::vector<size_t> steps;
            for (int i = 0; i < 40; ++i) {
                const auto part = a.expand(0.01, 5);
                steps.insert(steps.end(), part.begin(), part.end());
            }
            const auto all = b.expand(0.01, 200);
            std::printf("expand steps %zu sweep %zu equal %d cursor %llu cells_in_radius %zu\n", steps.size(), all.size(), (int)(steps == all),
                        (unsigned long long)a.current_search_index(), tess::find_cells_in_radius(diagram, tess::Vector3{0.31, 0.62, 0.44}, 0.1).size());
        }
        // the batch call that streams into host arrays must agree with the per-cell interface
        {
            std::vector<double> vol(n), area(64 * n);
            std::vector<uint64_t> off(n + 1);
            std::vector<int64_t> nbr(64 * n);
            std::vector<uint32_t> st(n);
            auto batch = diagram.compute_all_cells_to_host(vol.data(), off.data(), nbr.data(), area.data(), st.data(), n, 64 * n, 2);
            double streamed = 0;
            for (size_t i = 0; i < n; ++i) streamed += vol[i];
            tess::Cell c7 = diagram.get_cell_at_index(7, box);
            c7.compute_voronoi_cell();
            if (streamed != total || vol[7] != c7.compute_volume() || off[8] - off[7] != c7.compute_neighbors().size()) {
                std::printf("ERROR: streamed batch differs from the per-cell results\n");
                return 4;
            }
            std::printf("streamed total %.17g\n", streamed);
        }
        // error behaviour: wrong start polyhedron, add after initialize
        try {
            diagram.get_cell_at_index(0, tess::Polyhedron{0, 0, 0, 2, 2, 2});
            std::printf("ERROR: foreign polyhedron accepted\n");
            return 2;
        } catch (const tess::Error& e) {
            std::printf("expected error %d\n", e.code);
        }
    } catch (const std::exception& e) {
        std::printf("FAILED: %s\n", e.what());
        return 1;
    }
    return 0;
}
