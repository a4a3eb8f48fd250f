/* test_plugin.cpp -- acceptance program of the geogram plugin (b200_geo_plugin.*), run on the GPU
 * box by tests/test_gpu_plugin.py.  Links the UNMODIFIED reference (oracle/_ref/libgeoref.so) and
 * libb200vor.so.  Exit code 0 = every check passed; each check prints one line "check <name> ok".
 *
 *  1. the criterion of the reference's own src/tests/test_nn_search (main.cpp:136-181):
 *     NearestNeighborSearch::create(3, "B200NN") against "BNN", every point queried by index and
 *     by pointer, every squared distance bit-equal;
 *  2. arbitrary query points from parallel_for workers (the calling pattern of
 *     delaunay.cpp:259-274): bit-equal squared distances, and far fewer launches than queries;
 *  3. Delaunay::create(3, "B200") -> set_vertices -> update_neighbors against "NN": identical
 *     neighbour arrays; enlarge_neighborhood; the duplicates rule (delaunay_nn.cpp:122-135);
 *  4. cells of b200vor_compute_cells re-hydrated into VBW::ConvexCell (fill_convex_cell):
 *     volume / barycenter against the engine's own m, mg; against a ConvexCell clipped by the
 *     reference itself from the kNN (oracle B, generic_RVD.h:2278-2347 stop rule); ConvexCell::save
 *     and append_to_mesh of the re-hydrated cell; a further clip_by_plane on it.
 */
#include <geogram/basic/command_line.h>
#include <geogram/basic/command_line_args.h>
#include <geogram/basic/common.h>
#include <geogram/basic/logger.h>
#include <geogram/basic/process.h>
#include <geogram/delaunay/delaunay.h>
#include <geogram/mesh/mesh.h>

#include <cmath>
#include <cstdio>
#include <random>
#include <sstream>

#include "b200_geo_plugin.h"

using namespace GEO;

extern "C" int georef_init(int nthreads);

static int failures = 0;
#define CHECK(cond, name)                                                   \
    do {                                                                    \
        if(cond) {                                                          \
            std::printf("check %s ok\n", name);                             \
        } else {                                                            \
            std::printf("check %s FAILED (%s:%d)\n", name, __FILE__, __LINE__); \
            ++failures;                                                     \
        }                                                                   \
    } while(0)

int main(int argc, char** argv) {
    // the subsystems of GEO::initialize() this path needs (basic/common.cpp:103-110), as set up by
    // the reference harness of oracle/_ref (mesh-IO / image / bibliography are not linked in)
    georef_init(0);
    b200_register_plugin();
    const index_t N = argc > 1 ? index_t(std::atoi(argv[1])) : 20000;
    const index_t K = argc > 2 ? index_t(std::atoi(argv[2])) : 21;

    std::mt19937_64 rng(12345);
    std::uniform_real_distribution<double> U(0.0, 1.0);
    std::vector<double> pts(3 * size_t(N));
    for(double& x : pts) {
        x = U(rng);
    }

    // ------------------------------------------------------------------ 1. test_nn_search criterion
    {
        NearestNeighborSearch_var ref = NearestNeighborSearch::create(3, "BNN");
        NearestNeighborSearch_var gpu = NearestNeighborSearch::create(3, "B200NN");
        CHECK(dynamic_cast<NearestNeighborSearch_B200*>(gpu.get()) != nullptr, "factory_creates_B200NN");
        ref->set_points(N, pts.data());
        gpu->set_points(N, pts.data());
        std::vector<index_t> n1(K), n2(K);
        std::vector<double> d1(K), d2(K);
        bool by_index = true, by_ptr = true, self_first = true;
        for(index_t i = 0; i < N; ++i) {
            ref->get_nearest_neighbors(K, i, n1.data(), d1.data());
            gpu->get_nearest_neighbors(K, i, n2.data(), d2.data());
            for(index_t j = 0; j < K; ++j) {
                by_index = by_index && (d1[j] == d2[j]);
            }
            self_first = self_first && n2[0] == i;
            ref->get_nearest_neighbors(K, ref->point_ptr(i), n1.data(), d1.data());
            gpu->get_nearest_neighbors(K, gpu->point_ptr(i), n2.data(), d2.data());
            for(index_t j = 0; j < K; ++j) {
                by_ptr = by_ptr && (d1[j] == d2[j]);
            }
        }
        CHECK(by_index, "nn_by_index_sq_dist_bit_equal");
        CHECK(by_ptr, "nn_by_pointer_sq_dist_bit_equal");
        CHECK(self_first, "nn_self_at_rank_0");
        NearestNeighborSearch_B200* g = dynamic_cast<NearestNeighborSearch_B200*>(gpu.get());
        CHECK(g->nb_table_launches() == 1 && g->nb_point_launches() == 0, "nn_one_launch_serves_all_stored_points");

        // ---------------------------------------------------------------- 2. arbitrary points, threads
        const index_t NQ = 4000;
        std::vector<double> q(3 * size_t(NQ));
        for(double& x : q) {
            x = U(rng);
        }
        std::vector<double> dref(size_t(NQ) * K), dgpu(size_t(NQ) * K);
        std::vector<index_t> nref(size_t(NQ) * K), ngpu(size_t(NQ) * K);
        parallel_for(0, NQ, [&](index_t i) {
            ref->get_nearest_neighbors(K, &q[3 * size_t(i)], &nref[size_t(i) * K], &dref[size_t(i) * K]);
        });
        parallel_for(0, NQ, [&](index_t i) {
            gpu->get_nearest_neighbors(K, &q[3 * size_t(i)], &ngpu[size_t(i) * K], &dgpu[size_t(i) * K]);
        });
        bool eq = true;
        for(size_t e = 0; e < dref.size(); ++e) {
            eq = eq && (dref[e] == dgpu[e]);
        }
        CHECK(eq, "nn_arbitrary_points_parallel_for_bit_equal");
        std::printf("info by-point queries %u launches %u threads %u\n", unsigned(NQ), unsigned(g->nb_point_launches()),
                    unsigned(Process::maximum_concurrent_threads()));
        CHECK(g->nb_point_launches() <= NQ, "nn_point_queries_are_batched");
        index_t nearest = gpu->get_nearest_neighbor(&q[0]);
        CHECK(nearest == nref[0], "nn_get_nearest_neighbor");
    }

    // ------------------------------------------------------------------ 3. Delaunay back-end
    {
        Delaunay_var dref = Delaunay::create(3, "NN");
        Delaunay_var dgpu = Delaunay::create(3, "B200");
        CHECK(dynamic_cast<Delaunay_B200*>(dgpu.get()) != nullptr, "factory_creates_Delaunay_B200");
        CHECK(dynamic_cast<Delaunay_NearestNeighbors*>(dgpu.get()) != nullptr, "is_a_Delaunay_NearestNeighbors");
        CHECK(CmdLine::get_arg("algo:nn_search") != "B200NN", "algo_nn_search_restored");
        // two exact duplicates: the higher index gets no neighbours, the lower one skips it
        std::vector<double> v = pts;
        for(int c = 0; c < 3; ++c) {
            v[3 * 700 + c] = v[3 * 55 + c];
        }
        dref->set_vertices(N, v.data());
        dgpu->set_vertices(N, v.data());
        // same neighbourhoods: equal sizes and equal squared distances, rank by rank (the criterion of
        // test_nn_search -- indices may differ where two neighbours are exactly equidistant)
        bool same = true;
        index_t differing_ids = 0;
        GEO::vector<index_t> a, b;
        auto sqd = [&](index_t i, index_t j) {
            double s = 0.0;
            for(int c = 0; c < 3; ++c) {
                const double d = v[3 * size_t(i) + c] - v[3 * size_t(j) + c];
                s += d * d;
            }
            return s;
        };
        for(index_t i = 0; i < N; ++i) {
            dref->get_neighbors(i, a);
            dgpu->get_neighbors(i, b);
            same = same && a.size() == b.size();
            for(index_t k = 0; same && k < a.size(); ++k) {
                same = same && sqd(i, a[k]) == sqd(i, b[k]);
                differing_ids += (a[k] != b[k]);
            }
        }
        std::printf("info delaunay neighbour ids differing on exact ties: %u\n", unsigned(differing_ids));
        CHECK(same, "delaunay_neighbors_identical_to_NN");
        dgpu->get_neighbors(700, b);
        CHECK(b.empty(), "delaunay_duplicate_rule");
        Delaunay_NearestNeighbors* r = dynamic_cast<Delaunay_NearestNeighbors*>(dref.get());
        Delaunay_NearestNeighbors* g = dynamic_cast<Delaunay_NearestNeighbors*>(dgpu.get());
        r->enlarge_neighborhood(123, 47);
        g->enlarge_neighborhood(123, 47);
        dref->get_neighbors(123, a);
        dgpu->get_neighbors(123, b);
        bool enl = a.size() == b.size() && a.size() == 47;
        for(index_t k = 0; enl && k < a.size(); ++k) {
            enl = sqd(123, a[k]) == sqd(123, b[k]);
        }
        CHECK(enl, "delaunay_enlarge_neighborhood");
        CHECK(dgpu->nearest_vertex(&pts[30]) == 10, "delaunay_nearest_vertex");
    }

    // ------------------------------------------------------------------ 4. cells -> VBW::ConvexCell
    {
        const index_t NC = std::min<index_t>(N, 5000);
        b200vor_ctx* ctx = nullptr;
        bool ok = b200vor_create(&ctx, 0) == B200VOR_OK;
        double box[6] = {0, 0, 0, 1, 1, 1};
        ok = ok && b200vor_set_domain_box(ctx, box) == B200VOR_OK;
        ok = ok && b200vor_set_points(ctx, NC, pts.data(), 3, nullptr, 0) == B200VOR_OK;
        b200vor_cell_params prm = {0, 0, 2, 0, 0};
        ok = ok && b200vor_compute_cells(ctx, &prm) == B200VOR_OK;
        std::vector<double> m(NC), mg(3 * size_t(NC)), area(size_t(NC) * 64), vxyz(size_t(NC) * 64 * 3);
        std::vector<uint8_t> st(NC);
        std::vector<uint32_t> nf(NC), fn(size_t(NC) * 64), nv(NC), vf(size_t(NC) * 64);
        ok = ok && b200vor_get_cells(ctx, m.data(), mg.data(), st.data(), 0) == B200VOR_OK;
        ok = ok && b200vor_get_faces(ctx, nf.data(), fn.data(), area.data(), 64, 0) == B200VOR_OK;
        ok = ok && b200vor_get_vertices(ctx, nv.data(), vxyz.data(), vf.data(), 64, 0) == B200VOR_OK;
        CHECK(ok, "cells_computed");
        NearestNeighborSearch_var nn = NearestNeighborSearch::create(3, "BNN");
        nn->set_points(NC, pts.data());
        bool vol_ok = true, bary_ok = true, area_ok = true, refcell_ok = true, faces_ok = true;
        index_t filled = 0;
        VBW::ConvexCell C(VBW::WithVGlobal), R(VBW::WithVGlobal);
        for(index_t i = 0; i < NC; ++i) {
            if(st[i] != 0) {
                continue;
            }
            if(!B200GEO::fill_convex_cell(C, NC, pts.data(), nullptr, i, false, 0.0, box, nf[i], &fn[size_t(i) * 64],
                                          nv[i], &vf[size_t(i) * 64], 64)) {
                continue;
            }
            ++filled;
            const double vol = C.volume();
            vol_ok = vol_ok && std::fabs(vol - m[i]) <= 1e-9 * m[i];
            VBW::vec3 g = C.barycenter();
            const double s = std::cbrt(m[i]);
            bary_ok = bary_ok && std::fabs(g.x - mg[3 * size_t(i)] / m[i]) <= 1e-7 * s &&
                      std::fabs(g.y - mg[3 * size_t(i) + 1] / m[i]) <= 1e-7 * s &&
                      std::fabs(g.z - mg[3 * size_t(i) + 2] / m[i]) <= 1e-7 * s;
            for(index_t lv = 0; lv < nf[i]; ++lv) {
                area_ok = area_ok && std::fabs(C.facet_area(lv + 1) - area[size_t(i) * 64 + lv]) <= 1e-7 * s * s;
            }
            if(i % 50 == 0) {
                // the same cell clipped by the reference itself (oracle B): BNN kNN + clip_by_plane +
                // stop rule d2 > 4.1 R2 (generic_RVD.h:2322)
                const index_t kk = std::min<index_t>(NC, 120);
                std::vector<index_t> nb(kk);
                std::vector<double> d2(kk);
                nn->get_nearest_neighbors(kk, i, nb.data(), d2.data());
                R.clear();
                R.init_with_box(0, 0, 0, 1, 1, 1);
                const double* Pi = &pts[3 * size_t(i)];
                const VBW::vec3 ci = VBW::make_vec3(Pi[0], Pi[1], Pi[2]);
                for(index_t jj = 1; jj < kk; ++jj) {
                    if(d2[jj] > 4.1 * R.squared_radius(ci)) {
                        break;
                    }
                    const double* Pj = &pts[3 * size_t(nb[jj])];
                    R.clip_by_plane(VBW::make_vec4(2.0 * (Pi[0] - Pj[0]), 2.0 * (Pi[1] - Pj[1]), 2.0 * (Pi[2] - Pj[2]),
                                                   (Pj[0] * Pj[0] + Pj[1] * Pj[1] + Pj[2] * Pj[2]) -
                                                       (Pi[0] * Pi[0] + Pi[1] * Pi[1] + Pi[2] * Pi[2])),
                                    VBW::global_index_t(nb[jj]));
                }
                R.compute_geometry();
                refcell_ok = refcell_ok && std::fabs(R.volume() - vol) <= 1e-9 * vol;
                // same set of contributing neighbours
                std::vector<VBW::global_index_t> fa, fb;
                for(VBW::index_t lv = 7; lv < R.nb_v(); ++lv) {
                    if(R.vertex_is_contributing(lv)) {
                        fa.push_back(R.v_global_index(lv));
                    }
                }
                for(index_t lv = 0; lv < nf[i]; ++lv) {
                    if(fn[size_t(i) * 64 + lv] < B200VOR_BOX_ID) {
                        fb.push_back(fn[size_t(i) * 64 + lv]);
                    }
                }
                std::sort(fa.begin(), fa.end());
                std::sort(fb.begin(), fb.end());
                faces_ok = faces_ok && fa == fb;
            }
        }
        std::printf("info cells re-hydrated %u of %u\n", unsigned(filled), unsigned(NC));
        CHECK(filled > NC * 9 / 10, "convex_cell_rehydrated");
        CHECK(vol_ok, "convex_cell_volume_matches_engine");
        CHECK(bary_ok, "convex_cell_barycenter_matches_engine");
        CHECK(area_ok, "convex_cell_facet_areas_match_engine");
        CHECK(refcell_ok, "convex_cell_volume_matches_reference_clipping");
        CHECK(faces_ok, "convex_cell_faces_match_reference_clipping");

        // save / append_to_mesh / a further clip on one re-hydrated cell
        index_t pick = 0;
        while(pick < NC && (st[pick] != 0 || nv[pick] == 0 || nv[pick] > 64)) {
            ++pick;
        }
        B200GEO::fill_convex_cell(C, NC, pts.data(), nullptr, pick, false, 0.0, box, nf[pick], &fn[size_t(pick) * 64],
                                  nv[pick], &vf[size_t(pick) * 64], 64);
        std::ostringstream obj;
        const index_t nvs = C.save(obj, 1, 0.0);
        CHECK(nvs == nv[pick], "convex_cell_save_vertex_count");
        index_t nfl = 0;
        {
            std::istringstream in(obj.str());
            std::string line;
            while(std::getline(in, line)) {
                nfl += (line.size() > 1 && line[0] == 'f');
            }
        }
        CHECK(nfl == nf[pick], "convex_cell_save_facet_count");
        if(argc > 3) {
            // the OBJ the reference writes for the re-hydrated cell: diffed byte for byte by the
            // python test against the engine's own ConvexCell::save mirror
            FILE* f = std::fopen(argv[3], "w");
            std::fputs(obj.str().c_str(), f);
            std::fclose(f);
            std::printf("info saved cell %u\n", unsigned(pick));
            if(argc > 4) {
                // the seeds, for the python side of the byte-diff
                FILE* fp = std::fopen(argv[4], "wb");
                std::fwrite(pts.data(), sizeof(double), 3 * size_t(NC), fp);
                std::fclose(fp);
            }
        }
        Mesh M;
        C.append_to_mesh(&M, 0.0);
        CHECK(M.vertices.nb() == nv[pick] && M.facets.nb() == nf[pick], "convex_cell_append_to_mesh");
        // halve the cell by a plane through its barycentre: ConvexCell keeps working on it
        VBW::vec3 g = C.barycenter();
        const double v0 = C.volume();
        C.clip_by_plane(VBW::make_vec4(1.0, 0.0, 0.0, -g.x));
        C.compute_geometry();
        const double v1 = C.volume();
        CHECK(v1 > 0.2 * v0 && v1 < 0.8 * v0, "convex_cell_further_clip");
        b200vor_destroy(ctx);
    }
    std::printf(failures == 0 ? "ALL CHECKS PASSED\n" : "%d CHECK(S) FAILED\n", failures);
    return failures == 0 ? 0 : 2;
}
