// CPU-only test of the pure-host parts of the mirror (no context is created): tet adjacency.
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include <algorithm>
#include <vector>

#include "b200geo.h"

using namespace B200GEO;

static int fails = 0;
#define EXPECT(c)                                                     \
    do {                                                              \
        if(!(c)) {                                                    \
            printf("FAILED %s:%d: %s\n", __FILE__, __LINE__, #c);     \
            ++fails;                                                  \
        }                                                             \
    } while(0)

int main() {
    // a fan of tetrahedra around the edge (0,1): ring vertices 2..6, tets (0,1,r_k,r_{k+1})
    std::vector<global_index_t> tets;
    const int ring = 5;
    for(int k = 0; k < ring; ++k) {
        tets.push_back(0);
        tets.push_back(1);
        tets.push_back(global_index_t(2 + k));
        tets.push_back(global_index_t(2 + (k + 1) % ring));
    }
    std::vector<int64_t> adj;
    ConvexCellBatch::tet_adjacency(tets, adj);
    EXPECT(adj.size() == size_t(4 * ring));
    for(int k = 0; k < ring; ++k) {
        // facets opposite to vertices 0 and 1 are on the boundary; the other two lead around the fan
        EXPECT(adj[4 * k + 0] == -1 && adj[4 * k + 1] == -1);
        EXPECT(adj[4 * k + 2] == (k + 1) % ring);          // opposite r_k: shares (0,1,r_{k+1})
        EXPECT(adj[4 * k + 3] == (k + ring - 1) % ring);   // opposite r_{k+1}: shares (0,1,r_k)
    }
    // symmetry on a random soup of face-sharing pairs
    for(size_t c = 0; c < adj.size() / 4; ++c) {
        for(int lf = 0; lf < 4; ++lf) {
            int64_t nb = adj[4 * c + lf];
            if(nb >= 0) {
                int back = 0;
                for(int l2 = 0; l2 < 4; ++l2) {
                    back += adj[4 * size_t(nb) + l2] == int64_t(c);
                }
                EXPECT(back == 1);
            }
        }
    }
    std::vector<global_index_t> none;
    ConvexCellBatch::tet_adjacency(none, adj);
    EXPECT(adj.empty());
    // ---- OptimalTransportMap3d::extrapolate_weight: a linear weight field is reproduced exactly,
    //      a coplanar neighbourhood falls back to the nearest weight
    {
        std::vector<double> pts, w;
        std::vector<uint32_t> row;
        unsigned s = 12345u;
        for(int i = 0; i < 12; ++i) {
            double c[3];
            for(int d = 0; d < 3; ++d) {
                s = s * 1664525u + 1013904223u;
                c[d] = double(s >> 8) / double(1u << 24);
                pts.push_back(c[d]);
            }
            w.push_back(0.3 - 1.5 * c[0] + 0.25 * c[1] + 2.0 * c[2]);
            row.push_back(uint32_t(i));
        }
        const double q[3] = {0.4, 0.6, 0.1};
        const double v = OptimalTransportMap3d::extrapolate_weight(pts.data(), w.data(), q, row.data(), 12);
        EXPECT(fabs(v - (0.3 - 1.5 * q[0] + 0.25 * q[1] + 2.0 * q[2])) < 1e-10);
        for(int i = 0; i < 12; ++i) {
            pts[3 * i + 2] = 0.5;      // all neighbours in one plane
        }
        EXPECT(OptimalTransportMap3d::extrapolate_weight(pts.data(), w.data(), q, row.data(), 12) == w[0]);
        EXPECT(OptimalTransportMap3d::extrapolate_weight(pts.data(), w.data(), q, row.data(), 3) == w[0]);
    }
    // ---- compute_BRIO_order: levels of the unmodified reference (tests/golden/brio_reference.npz), Hilbert walk
    {
        const int b = 3;
        std::vector<uint64_t> seen(size_t(1) << (3 * b), 0);
        std::vector<int> at(3 * (size_t(1) << (3 * b)));
        for(uint32_t x = 0; x < (1u << b); ++x) {
            for(uint32_t y = 0; y < (1u << b); ++y) {
                for(uint32_t z = 0; z < (1u << b); ++z) {
                    const uint64_t h = hilbert_index_3d(x, y, z, b);
                    EXPECT(h < seen.size() && seen[h] == 0);
                    seen[h] = 1;
                    at[3 * h] = int(x);
                    at[3 * h + 1] = int(y);
                    at[3 * h + 2] = int(z);
                }
            }
        }
        for(size_t h = 1; h < seen.size(); ++h) {     // consecutive curve positions are lattice neighbours
            EXPECT(abs(at[3 * h] - at[3 * h - 3]) + abs(at[3 * h + 1] - at[3 * h - 2]) + abs(at[3 * h + 2] - at[3 * h - 1]) == 1);
        }
        const index_t n = 20000;
        std::vector<double> pts(3 * size_t(n));
        unsigned s = 99u;
        for(size_t i = 0; i < pts.size(); ++i) {
            s = s * 1664525u + 1013904223u;
            pts[i] = double(s >> 8) / double(1u << 24);
        }
        std::vector<index_t> order, levels;
        compute_BRIO_order(n, pts.data(), order, 3, 300, 0.125, &levels);
        const index_t ref[5] = {0, 39, 312, 2500, 20000};
        EXPECT(levels.size() == 5);
        for(size_t l = 0; l < 5 && l < levels.size(); ++l) {
            EXPECT(levels[l] == ref[l]);
        }
        std::vector<index_t> sorted(order);
        std::sort(sorted.begin(), sorted.end());
        for(index_t i = 0; i < n; ++i) {
            EXPECT(sorted[i] == i);
        }
        // the last interval is local: mean step far below the mean distance of random pairs (0.66)
        double step = 0.0;
        for(index_t i = 2501; i < n; ++i) {
            double d2 = 0.0;
            for(int d = 0; d < 3; ++d) {
                const double t = pts[3 * size_t(order[i]) + d] - pts[3 * size_t(order[i - 1]) + d];
                d2 += t * t;
            }
            step += sqrt(d2);
        }
        EXPECT(step / double(n - 2501) < 0.1);
    }
    printf(fails ? "host cpu test: %d FAILURES\n" : "host cpu test: ok\n", fails);
    return fails ? 1 : 0;
}
