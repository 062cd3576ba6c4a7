// cell_stats.cpp — capacity statistics of cells (test/tooling; links the CPU oracle, never the product).
// Prints, for a seeded point set, the distribution of the pool sizes a cell needed during its construction
// (vertex / half-edge / face slots ever used) — what sizes the thread-per-cell tables of csrc/clip_thread.cu.
//   g++ -O2 -std=c++17 -ffp-contract=off -I oracle tools/cell_stats.cpp oracle/tess_oracle.cpp -o ab_build/cell_stats
#include <algorithm>
#include <array>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "tess_oracle.hpp"

static uint64_t mix(uint64_t z) {
    z += 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
static double u01(uint64_t seed, uint64_t i, uint64_t c) { return (double)(mix(mix(seed) + 3 * i + c) >> 11) * (1.0 / 9007199254740992.0); }

int main(int argc, char** argv) {
    const size_t n = argc > 1 ? strtoull(argv[1], 0, 10) : 200000;
    const size_t sample = argc > 2 ? strtoull(argv[2], 0, 10) : 20000;
    const char* kind = argc > 3 ? argv[3] : "uniform";
    orc::Diagram d;
    if (!strcmp(kind, "bcc")) {
        const size_t m = (size_t)std::cbrt((double)n / 2.0);
        const double a = 1.0 / (double)m;
        size_t idx = 0;
        for (int s = 0; s < 2; ++s)
            for (size_t i = 0; i < m; ++i)
                for (size_t j = 0; j < m; ++j)
                    for (size_t k = 0; k < m; ++k, ++idx) {
                        const double o = s ? 0.5 : 0.0;
                        d.add_particle_with_group({(i + o) * a + a / 4 + (2 * u01(5, idx, 0) - 1) * 1e-3 * a, (j + o) * a + a / 4 + (2 * u01(5, idx, 1) - 1) * 1e-3 * a,
                                                   (k + o) * a + a / 4 + (2 * u01(5, idx, 2) - 1) * 1e-3 * a},
                                                  0);
                    }
    } else {
        for (size_t i = 0; i < n; ++i) d.add_particle_with_group({u01(2, i, 0), u01(2, i, 1), u01(2, i, 2)}, 0);
    }
    const double box[6] = {0, 0, 0, 1, 1, 1};
    d.initialize(box, 8);
    const size_t np = d.cell_array.points.size();
    std::vector<uint32_t> hv, he, hf;
    double sv = 0, se = 0, sf = 0;
    std::vector<uint32_t> histv(257, 0), histe(1025, 0), histf(257, 0);
    std::vector<std::array<uint32_t, 3>> all;
    for (size_t s = 0; s < sample; ++s) {
        const size_t i = (size_t)(mix(1234 + s) % np);
        const orc::CellResult r = d.compute_cell_at_index(i, orc::MODE_SECURITY, 0.0, -1, false);
        histv[std::min<uint32_t>(r.pool_slots[0], 256)]++;
        histe[std::min<uint32_t>(r.pool_slots[1], 1024)]++;
        histf[std::min<uint32_t>(r.pool_slots[2], 256)]++;
        all.push_back({r.pool_slots[0], r.pool_slots[1], r.pool_slots[2]});
        sv += r.pool_slots[0]; se += r.pool_slots[1]; sf += r.pool_slots[2];
    }
    printf("%s n=%zu sample=%zu  mean slots: V %.1f  E %.1f  F %.1f\n", kind, np, sample, sv / sample, se / sample, sf / sample);
    auto tail = [&](const char* name, const std::vector<uint32_t>& h, std::initializer_list<int> caps) {
        for (int c : caps) {
            size_t over = 0;
            for (size_t k = c + 1; k < h.size(); ++k) over += h[k];
            printf("  %s > %3d : %.4f %%\n", name, c, 100.0 * over / sample);
        }
    };
    tail("V", histv, {32, 36, 40, 44, 48, 52, 56, 64});
    tail("E", histe, {96, 104, 112, 120, 128, 136, 144, 160, 176, 192});
    tail("F", histf, {16, 20, 22, 24, 26, 28, 32, 40});
    // joint: cells that do NOT fit (V, E, F); bytes per thread of the lane-interleaved tables
    const int cfgs[][3] = {{40, 128, 22}, {44, 136, 24}, {44, 140, 26}, {45, 140, 24}, {46, 140, 24}, {48, 144, 26}, {48, 152, 26}, {56, 168, 28}, {64, 192, 32}};
    for (auto& c : cfgs) {
        size_t over = 0;
        for (auto& a : all) over += (a[0] > (uint32_t)c[0] || a[1] > (uint32_t)c[1] || a[2] > (uint32_t)c[2]);
        printf("  V%d E%d F%d : %5.2f %% do not fit   (%d B/thread)\n", c[0], c[1], c[2], 100.0 * over / sample, 24 * c[0] + 4 * c[1] + 5 * c[2] + c[0]);
    }
    return 0;
}
