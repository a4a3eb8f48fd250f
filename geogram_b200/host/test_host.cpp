// test_host.cpp -- exercises the C++ host classes the way the reference's own tests exercise
// theirs (src/tests/test_nn_search/main.cpp:150-190: every d2 row bit-equal to another
// implementation; periodic_delaunay_3d.cpp:4349-4359: sum of cell volumes).  Needs a GPU;
// run by tests/test_host_cpp.py (-m gpu).  Brute force below is the test's own checker.
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include <algorithm>
#include <random>
#include <vector>

#include "b200geo.h"

using namespace B200GEO;

static int fails = 0;
#define EXPECT(cond)                                                     \
    do {                                                                 \
        if(!(cond)) {                                                    \
            printf("FAIL %s:%d: %s\n", __FILE__, __LINE__, #cond);       \
            ++fails;                                                     \
        }                                                                \
    } while(0)

int main() {
    const index_t n = 4000, k = 12;
    std::mt19937_64 rng(12345);
    std::uniform_real_distribution<double> U(0.0, 1.0);
    std::vector<double> pts(3 * n);
    for(double& x : pts) {
        x = U(rng);
    }
    // ---- NearestNeighborSearch
    NearestNeighborSearch* nn = NearestNeighborSearch::create(3, "B200NN");
    EXPECT(nn != nullptr);
    EXPECT(NearestNeighborSearch::create(2) == nullptr);
    nn->set_points(n, pts.data());
    EXPECT(nn->nb_points() == n && nn->dimension() == 3 && nn->stride_supported());
    std::vector<index_t> nbr(k);
    std::vector<double> d2(k);
    int bad = 0;
    for(index_t q = 0; q < n; q += 37) {
        nn->get_nearest_neighbors(k, q, nbr.data(), d2.data());
        std::vector<std::pair<double, index_t> > all(n);
        for(index_t j = 0; j < n; ++j) {
            double dx = pts[3 * j] - pts[3 * q], dy = pts[3 * j + 1] - pts[3 * q + 1],
                   dz = pts[3 * j + 2] - pts[3 * q + 2];
            volatile double s = dx * dx;   // ((dx^2)+dy^2)+dz^2, no contraction (geometry_nd.h:65-73)
            volatile double t = dy * dy;
            s = s + t;
            t = dz * dz;
            s = s + t;
            all[j] = std::make_pair(double(s), j);
        }
        std::sort(all.begin(), all.end());
        for(index_t r = 0; r < k; ++r) {
            bad += (all[r].first != d2[r]) || (all[r].second != nbr[r]);
        }
        EXPECT(nbr[0] == q && d2[0] == 0.0);
    }
    EXPECT(bad == 0);
    double qp[3] = {0.5, 0.5, 0.5};
    index_t nearest = nn->get_nearest_neighbor(qp);
    nn->get_nearest_neighbors(k, qp, nbr.data(), d2.data());
    EXPECT(nbr[0] == nearest);
    EXPECT(std::is_sorted(d2.begin(), d2.end()));
    // KeepInitialValues: a fake closer candidate stays first
    std::vector<index_t> nb2(nbr);
    std::vector<double> dd2(d2);
    nb2[k - 1] = nbr[k - 1];
    nn->get_nearest_neighbors(k, qp, nb2.data(), dd2.data(), NearestNeighborSearch::KeepInitialValues());
    EXPECT(nb2 == nbr);
    delete nn;

    // ---- Delaunay_NearestNeighbors (duplicates rule)
    std::vector<double> dup(pts);
    dup.insert(dup.end(), pts.begin(), pts.begin() + 3 * 5);   // 5 exact duplicates
    Delaunay_NearestNeighbors del(3);
    del.set_default_nb_neighbors(20);
    del.set_vertices(n + 5, dup.data());
    std::vector<index_t> ne;
    del.get_neighbors(n + 2, ne);
    EXPECT(ne.empty());                     // duplicates a lower index: no neighbours
    del.get_neighbors(2, ne);
    EXPECT(ne.size() >= 19 && std::find(ne.begin(), ne.end(), n + 2) == ne.end());
    del.get_neighbors(100, ne);
    EXPECT(ne.size() == 20);
    del.enlarge_neighborhood(100, 40);
    std::vector<index_t> ne2;
    del.get_neighbors(100, ne2);
    EXPECT(ne2.size() == 40 && std::equal(ne.begin(), ne.end(), ne2.begin()));
    EXPECT(del.nearest_vertex(&pts[3 * 77]) == 77);
    bool threw = false;
    try {
        Delaunay_NearestNeighbors d2d(2);
    } catch(const std::exception&) {
        threw = true;
    }
    EXPECT(threw);

    // ---- ConvexCellBatch
    ConvexCellBatch cells;
    cells.compute(n, pts.data(), nullptr, true);
    EXPECT(fabs(cells.total_volume() - 1.0) < 1e-9);
    index_t nfsum = 0;
    bool sym = true;
    for(index_t c = 0; c < n; ++c) {
        EXPECT(cells.status(c) == 0);
        nfsum += cells.nb_v(c);
        for(index_t lv = 0; lv < cells.nb_v(c); ++lv) {
            global_index_t g = cells.v_global_index(c, lv);
            if(g >= BOX_ID) {
                continue;
            }
            bool back = false;
            for(index_t l2 = 0; l2 < cells.nb_v(g); ++l2) {
                if(cells.v_global_index(g, l2) == c) {
                    back = fabs(cells.facet_area(g, l2) - cells.facet_area(c, lv)) < 1e-12;
                }
            }
            sym = sym && back;
        }
        vec3 b = cells.barycenter(c);
        EXPECT(b.x > 0 && b.x < 1 && b.y > 0 && b.y < 1 && b.z > 0 && b.z < 1);
    }
    EXPECT(sym);                            // face (i,j) seen from both cells, same area
    printf("mean faces per cell (incl. box): %.3f\n", double(nfsum) / n);
    cells.set_periodic(1.0);
    cells.compute(n, pts.data(), nullptr, true);
    EXPECT(fabs(cells.total_volume() - 1.0) < 1e-9);
    nfsum = 0;
    for(index_t c = 0; c < n; ++c) {
        nfsum += cells.nb_v(c);
    }
    printf("mean faces per cell (periodic): %.3f  (Poisson-Voronoi: 15.535)\n", double(nfsum) / n);
    EXPECT(fabs(double(nfsum) / n - 15.535) < 0.3);

    // ---- CentroidalVoronoiTesselation
    CentroidalVoronoiTesselation cvt(3);
    cvt.set_points(n, pts.data());
    cvt.lock_point(0);
    cvt.Lloyd_iterations(3);
    EXPECT(cvt.embedding(0)[0] == pts[0] && cvt.embedding(0)[1] == pts[1]);   // locked
    double moved = 0.0;
    for(index_t i = 1; i < n; ++i) {
        moved = std::max(moved, fabs(cvt.embedding(i)[0] - pts[3 * i]));
        EXPECT(cvt.embedding(i)[0] > 0 && cvt.embedding(i)[0] < 1);
    }
    EXPECT(moved > 1e-4);
    // relaxed seeds: volumes more uniform
    ConvexCellBatch c2;
    c2.compute(n, cvt.embedding(0), nullptr, false);
    double v0 = 0, v1 = 0;
    for(index_t c = 0; c < n; ++c) {
        v0 += (cells.volume(c) - 1.0 / n) * (cells.volume(c) - 1.0 / n);
        v1 += (c2.volume(c) - 1.0 / n) * (c2.volume(c) - 1.0 / n);
    }
    EXPECT(v1 < 0.6 * v0);
    // ---- volumetric CVT of a tetrahedral mesh (one cube cut into six tetrahedra = the cube): the same
    //      Lloyd iterate as on the box domain
    {
        const double V[8][3] = {{0, 0, 0}, {1, 0, 0}, {0, 1, 0}, {1, 1, 0}, {0, 0, 1}, {1, 0, 1}, {0, 1, 1}, {1, 1, 1}};
        const index_t T[6][4] = {{0, 1, 3, 7}, {0, 1, 5, 7}, {0, 2, 3, 7}, {0, 2, 6, 7}, {0, 4, 5, 7}, {0, 4, 6, 7}};
        CentroidalVoronoiTesselation a(3), b(3);
        a.set_points(n, pts.data());
        a.Lloyd_iterations(1);
        b.set_volumetric_mesh(8, &V[0][0], 6, &T[0][0]);
        b.set_points(n, pts.data());
        b.Lloyd_iterations(1);
        double dmax = 0.0;
        for(index_t i = 0; i < n; ++i) {
            for(int k = 0; k < 3; ++k) {
                dmax = std::max(dmax, fabs(a.embedding(i)[k] - b.embedding(i)[k]));
            }
        }
        printf("Lloyd on the cube as 6 tetrahedra vs box domain: max difference %.2e\n", dmax);
        EXPECT(dmax < 1e-10);
    }
    // ---- funcgrad (CVT.cpp:323-341): f decreases under Lloyd, g = 0 on the locked point, and
    //      a finite difference of f along -g matches |g|^2 (the CVT gradient formula)
    {
        std::vector<double> x(pts), g(3 * size_t(n)), x2(3 * size_t(n)), g2(3 * size_t(n));
        double f0 = 0, f1 = 0, f2 = 0;
        cvt.funcgrad(3 * n, x.data(), f0, g.data());
        EXPECT(g[0] == 0.0 && g[1] == 0.0 && g[2] == 0.0);
        std::vector<double> xr(cvt.embedding(0), cvt.embedding(0) + 3 * size_t(n));
        cvt.funcgrad(3 * n, xr.data(), f1, g2.data());
        EXPECT(f0 > 0 && f1 < f0);
        cvt.unlock_all_points();
        cvt.funcgrad(3 * n, x.data(), f0, g.data());
        double gg = 0;
        for(size_t k = 0; k < g.size(); ++k) {
            gg += g[k] * g[k];
        }
        const double h = 1.0;   // displacement h*|g_i| ~ 1e-5 of a 0.06 cell: linear regime
        for(size_t k = 0; k < g.size(); ++k) {
            x2[k] = x[k] - h * g[k];
        }
        cvt.funcgrad(3 * n, x2.data(), f2, g2.data());
        EXPECT(fabs((f0 - f2) / h - gg) < 0.05 * gg);
    }
    // ---- Newton_iterations (L-BFGS on the device objective): lower energy than the same number of
    //      Lloyd iterations from the same start
    {
        CentroidalVoronoiTesselation a(3), b(3);
        a.set_points(n, pts.data());
        b.set_points(n, pts.data());
        a.Lloyd_iterations(8);
        b.Newton_iterations(8);
        std::vector<double> xa(a.embedding(0), a.embedding(0) + 3 * size_t(n)), ga(3 * size_t(n));
        double fa = 0, f0 = 0;
        a.funcgrad(3 * n, xa.data(), fa, ga.data());
        std::vector<double> x0(pts);
        a.funcgrad(3 * n, x0.data(), f0, ga.data());
        printf("CVT energy: start %.6e  8 Lloyd %.6e  8 Newton (L-BFGS) %.6e\n", f0, fa, b.last_energy());
        EXPECT(b.last_energy() < f0 * 0.9);
        EXPECT(b.last_energy() <= fa * 1.02);
    }
    // ---- cell geometry export: loops close, volume from the exported geometry, tetrahedra, OBJ
    {
        ConvexCellBatch pc;
        pc.set_periodic(1.0);
        pc.compute(n, pts.data(), nullptr, true, true);
        size_t sum_t = 0;
        double worst = 0.0;
        for(index_t c = 0; c < n; ++c) {
            EXPECT(pc.nb_t(c) == 2 * pc.nb_v(c) - 4);
            sum_t += pc.nb_t(c);
            if(c % 13) {
                continue;
            }
            std::vector<std::vector<index_t> > loops;
            pc.facet_loops(c, loops);
            double vol6 = 0.0;
            const double* s0 = &pts[3 * size_t(c)];
            for(size_t f = 0; f < loops.size(); ++f) {
                EXPECT(loops[f].size() >= 3);
                vec3 a = pc.triangle_point(c, loops[f][0]);
                for(size_t k = 1; k + 1 < loops[f].size(); ++k) {
                    vec3 b = pc.triangle_point(c, loops[f][k]), d = pc.triangle_point(c, loops[f][k + 1]);
                    double A[3] = {a.x - s0[0], a.y - s0[1], a.z - s0[2]};
                    double B[3] = {b.x - s0[0], b.y - s0[1], b.z - s0[2]};
                    double D[3] = {d.x - s0[0], d.y - s0[1], d.z - s0[2]};
                    vol6 += A[0] * (B[1] * D[2] - B[2] * D[1]) - A[1] * (B[0] * D[2] - B[2] * D[0]) +
                            A[2] * (B[0] * D[1] - B[1] * D[0]);
                }
            }
            worst = std::max(worst, fabs(fabs(vol6) / 6.0 - pc.volume(c)) / pc.volume(c));
        }
        EXPECT(worst < 1e-9);
        std::vector<global_index_t> tets;
        pc.delaunay_tetrahedra(tets);
        EXPECT(tets.size() == sum_t);          // 4 ids per tetrahedron, each seen from its 4 cells
        pc.save_cells("/tmp/b200geo_cells.obj", 0.1);
        FILE* fobj = fopen("/tmp/b200geo_cells.obj", "r");
        EXPECT(fobj != nullptr);
        size_t nvl = 0, nfl = 0;
        if(fobj) {
            char line[4096];
            while(fgets(line, sizeof(line), fobj)) {
                nvl += line[0] == 'v';
                nfl += line[0] == 'f';
            }
            fclose(fobj);
            remove("/tmp/b200geo_cells.obj");
        }
        size_t sum_f = 0;
        for(index_t c = 0; c < n; ++c) {
            sum_f += pc.nb_v(c);
        }
        EXPECT(nvl == sum_t && nfl == sum_f);
    }
    // ---- OptimalTransportMap3d: uniform target on the box, every cell reaches 1/n
    {
        OptimalTransportMap3d otm;
        otm.set_points(n, cvt.embedding(0));
        otm.set_epsilon(1e-4);
        otm.optimize(30);
        EXPECT(otm.converged());
        double worst = 0.0, total = 0.0;
        for(index_t i = 0; i < n; ++i) {
            worst = std::max(worst, fabs(otm.cell_measure(i) * n - 1.0));
            total += otm.cell_measure(i);
        }
        EXPECT(worst <= 1e-4);
        EXPECT(fabs(total - 1.0) < 1e-9);
        EXPECT(otm.nb_newton_iterations() > 0 && otm.nb_newton_iterations() < 30);
    }
    // ---- optimize_levels (compute_OTM/main.cpp:292-297): a smooth non-uniform target, three levels; same
    //      transport map as the single-level solve, the finest level starts next to its solution
    {
        const double* p = cvt.embedding(0);
        std::vector<double> nu(n);
        double s = 0.0;
        for(index_t i = 0; i < n; ++i) {
            nu[i] = 1.0 + 0.1 * sin(6.283185307179586 * p[3 * size_t(i)]);
            s += nu[i];
        }
        OptimalTransportMap3d one, ml;
        one.set_points(n, p);
        ml.set_points(n, p);
        for(index_t i = 0; i < n; ++i) {
            one.set_nu(i, nu[i] / s);
            ml.set_nu(i, nu[i] / s);
        }
        one.set_epsilon(1e-4);
        ml.set_epsilon(1e-4);
        one.optimize(50);
        std::vector<index_t> levels;
        levels.push_back(0);
        levels.push_back(n / 16);
        levels.push_back(n / 4);
        levels.push_back(n);
        try {
            ml.optimize_levels(levels, 50);
        } catch(const std::exception& err) {
            printf("optimize_levels threw: %s\n", err.what());
            EXPECT(false);
        }
        EXPECT(one.converged() && ml.converged());
        EXPECT(ml.level_newton_iterations().size() == 3);
        if(ml.level_newton_iterations().size() != 3) {
            printf("host test: %d FAILURES\n", fails);
            return 1;
        }
        double worst = 0.0, total = 0.0, mean1 = 0.0, meanl = 0.0, dw = 0.0, range = 0.0;
        for(index_t i = 0; i < n; ++i) {
            worst = std::max(worst, fabs(ml.cell_measure(i) * s / nu[i] - 1.0));
            total += ml.cell_measure(i);
            mean1 += one.weight(i) / n;
            meanl += ml.weight(i) / n;
        }
        for(index_t i = 0; i < n; ++i) {
            dw = std::max(dw, fabs((ml.weight(i) - meanl) - (one.weight(i) - mean1)));
            range = std::max(range, fabs(one.weight(i) - mean1));
        }
        EXPECT(worst <= 1e-4);
        EXPECT(fabs(total - 1.0) < 1e-9);
        // weights are defined up to a constant; both solves stop within epsilon of the same optimum
        EXPECT(dw <= 1e-2 * range);
        EXPECT(ml.level_newton_iterations().back() <= 10);
        printf("optimize_levels: Newton iterations per level %u %u %u (single level: %u), max weight difference %.2e of range %.2e\n",
               unsigned(ml.level_newton_iterations()[0]), unsigned(ml.level_newton_iterations()[1]),
               unsigned(ml.level_newton_iterations()[2]), unsigned(one.nb_newton_iterations()), dw, range);
    }
    printf(fails ? "host test: %d FAILURES\n" : "host test: ok\n", fails);
    return fails ? 1 : 0;
}
