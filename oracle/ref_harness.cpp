/*
 * oracle/ref_harness.cpp -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Thin C-ABI harness around the UNMODIFIED reference (geogram), whose sources
 * are compiled where they lie under /root/reference by oracle/Makefile into
 * oracle/_ref/libgeoref.so (git-ignored).  Nothing of the reference is copied
 * into this repository: this file only *calls* the reference's public API.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * `--impl reference` leg may load the resulting library, and only as the
 * checker / CPU baseline -- never on the product path.
 *
 * Oracles exposed:
 *   georef_knn            GEO::NearestNeighborSearch "BNN" (points/kd_tree.cpp:156-387)
 *   georef_cells_delaunay oracle A: PeriodicDelaunay3d + copy_Laguerre_cell_from_Delaunay
 *                         (delaunay/periodic_delaunay_3d.cpp:3098, :3634) + compute_mg
 *                         (voronoi/convex_cell.cpp:1267), driver pattern of
 *                         examples/graphics/demo_Delaunay3d/main.cpp:178-190, 432-456
 *   georef_cells_knn      oracle B: BNN kNN + VBW::ConvexCell init_with_box /
 *                         clip_by_plane / squared_radius / compute_mg with the
 *                         security-radius + k-expansion rule of
 *                         voronoi/generic_RVD.h:2278-2347
 */
#include <geogram/basic/common.h>
#include <geogram/basic/logger.h>
#include <geogram/basic/process.h>
#include <geogram/basic/progress.h>
#include <geogram/basic/command_line.h>
#include <geogram/basic/command_line_args.h>
#include <geogram/basic/file_system.h>
#include <geogram/basic/stopwatch.h>
#include <geogram/numerics/predicates.h>
#include <geogram/delaunay/delaunay.h>
#include <geogram/delaunay/periodic_delaunay_3d.h>
#include <geogram/points/nn_search.h>
#include <geogram/voronoi/convex_cell.h>
#include <geogram/mesh/mesh_reorder.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstdint>
#include <cstring>
#include <vector>

/* OpenNL entry points referenced by basic/command_line.cpp:697-698 (only from
 * CmdLine::parse, never reached here).  OpenNL is out of scope. */
extern "C" {
    void nlInitialize(int, char**) {}
    void nlPrintfFuncs(void*, void*) {}
}
/* bibliography/embedded_references.cpp is generated-data only; not needed. */
void register_embedded_bib_file(void);
void register_embedded_bib_file(void) {}

namespace {

    bool g_initialized = false;

    double now_s() {
        using clk = std::chrono::steady_clock;
        return std::chrono::duration<double>(clk::now().time_since_epoch()).count();
    }

    const uint32_t BOX_ID = 0xFFFFFFF0u;  // + 0..5 : x-,x+,y-,y+,z-,z+

    /* Collects the sorted global ids of the contributing planes of C and
     * (optionally) their facet areas.  C.compute_geometry() must have run. */
    uint32_t collect_faces(
        VBW::ConvexCell& C, uint32_t fcap, uint32_t* ids, double* areas
    ) {
        std::vector<uint8_t> used(C.nb_v(), 0);
        for(VBW::ushort t = C.first_triangle(); t != VBW::END_OF_LIST;
            t = C.next_triangle(t)) {
            for(VBW::index_t l = 0; l < 3; ++l) {
                used[C.triangle_v_local_index(t, l)] = 1;
            }
        }
        std::vector<std::pair<uint32_t, double> > F;
        for(VBW::index_t lv = 1; lv < C.nb_v(); ++lv) {  // 0 = vertex at infinity
            if(used[lv]) {
                F.push_back(std::make_pair(
                    uint32_t(C.v_global_index(lv)),
                    areas ? C.facet_area(lv) : 0.0));
            }
        }
        std::sort(F.begin(), F.end());
        uint32_t nf = uint32_t(F.size());
        for(uint32_t f = 0; f < std::min(nf, fcap); ++f) {
            ids[f] = F[f].first;
            if(areas) {
                areas[f] = F[f].second;
            }
        }
        return nf;
    }
}

extern "C" {

int georef_init(int nthreads) {
    if(g_initialized) {
        return 0;
    }
    /* Same order as GEO::initialize() (basic/common.cpp:103-110), minus the
     * mesh-IO / image / bibliography subsystems that are out of scope. */
    setenv("LC_NUMERIC", "POSIX", 1);
    GEO::FileSystem::initialize();
    GEO::Logger::initialize();
    GEO::Process::initialize(0);
    GEO::Progress::initialize();
    GEO::CmdLine::initialize();
    GEO::Stopwatch::initialize();
    GEO::PCK::initialize();
    GEO::Delaunay::initialize();
    GEO::CmdLine::import_arg_group("standard");
    GEO::CmdLine::import_arg_group("algo");
    GEO::Logger::instance()->set_quiet(true);
    if(nthreads > 0) {
        GEO::Process::set_max_threads(GEO::index_t(nthreads));
    }
    g_initialized = true;
    return 0;
}

int georef_nthreads(void) {
    return int(GEO::Process::maximum_concurrent_threads());
}

/* kNN by point index, rows of k, ascending, self at rank 0.
 * times[0] = build seconds, times[1] = query seconds. */
int georef_knn(
    uint32_t n, const double* xyz, uint32_t k,
    uint32_t* nbr, double* d2, double* times
) {
    georef_init(0);
    double t0 = now_s();
    GEO::NearestNeighborSearch_var nn = GEO::NearestNeighborSearch::create(3, "BNN");
    nn->set_points(n, xyz);
    double t1 = now_s();
    GEO::parallel_for_slice(0, n,
        [&](GEO::index_t b, GEO::index_t e) {
            for(GEO::index_t i = b; i < e; ++i) {
                nn->get_nearest_neighbors(
                    k, i, nbr + size_t(i) * k, d2 + size_t(i) * k);
            }
        });
    double t2 = now_s();
    if(times) {
        times[0] = t1 - t0;
        times[1] = t2 - t1;
    }
    return 0;
}

/* kNN for arbitrary query points (nn_search.h:129-134). */
int georef_knn_points(
    uint32_t n, const double* xyz, uint32_t nq, const double* q, uint32_t k,
    uint32_t* nbr, double* d2
) {
    georef_init(0);
    GEO::NearestNeighborSearch_var nn = GEO::NearestNeighborSearch::create(3, "BNN");
    nn->set_points(n, xyz);
    GEO::parallel_for_slice(0, nq,
        [&](GEO::index_t b, GEO::index_t e) {
            for(GEO::index_t i = b; i < e; ++i) {
                nn->get_nearest_neighbors(
                    k, q + 3 * size_t(i), nbr + size_t(i) * k, d2 + size_t(i) * k);
            }
        });
    return 0;
}

/*
 * Oracle A.  periodic != 0: domain [0,period)^3 with 27 instances, face ids are
 * real + n*instance (delaunay/periodic.h:95-99).  periodic == 0: Delaunay with
 * the vertex at infinity kept, then 6 clips by the unit-box planes (ids BOX_ID+k).
 * Outputs (caller-allocated): m[n], mg[3n], nfaces[n], face_nbr[n*fcap] (sorted),
 * face_area[n*fcap] or NULL.  times[0] = Delaunay seconds, times[1] = cells seconds.
 * Returns 0, or 1 if has_empty_cells() (diagram invalid, outputs untouched).
 */
int georef_cells_delaunay(
    uint32_t n, const double* xyz, const double* w,
    int periodic, double period, int exact_predicates,
    double* m, double* mg, uint32_t* nfaces, uint32_t* face_nbr,
    double* face_area, uint32_t fcap, double* times
) {
    georef_init(0);
    double t0 = now_s();
    GEO::SmartPointer<GEO::PeriodicDelaunay3d> del =
        new GEO::PeriodicDelaunay3d(periodic != 0, period);
    if(!periodic) {
        del->set_keeps_infinite(true);
    }
    del->set_stores_cicl(false);
    del->set_vertices(n, xyz);
    if(w != nullptr) {
        del->set_weights(w);
    }
    del->compute();
    double t1 = now_s();
    if(del->has_empty_cells()) {
        return 1;
    }
    GEO::parallel_for_slice(0, n,
        [&](GEO::index_t b, GEO::index_t e) {
            VBW::ConvexCell C;
            C.use_exact_predicates(exact_predicates != 0);
            GEO::PeriodicDelaunay3d::IncidentTetrahedra W;
            for(GEO::index_t v = b; v < e; ++v) {
                del->copy_Laguerre_cell_from_Delaunay(v, C, W);
                if(!periodic) {
                    C.clip_by_plane(VBW::vec4( 1.0, 0.0, 0.0, 0.0), BOX_ID + 0);
                    C.clip_by_plane(VBW::vec4(-1.0, 0.0, 0.0, 1.0), BOX_ID + 1);
                    C.clip_by_plane(VBW::vec4( 0.0, 1.0, 0.0, 0.0), BOX_ID + 2);
                    C.clip_by_plane(VBW::vec4( 0.0,-1.0, 0.0, 1.0), BOX_ID + 3);
                    C.clip_by_plane(VBW::vec4( 0.0, 0.0, 1.0, 0.0), BOX_ID + 4);
                    C.clip_by_plane(VBW::vec4( 0.0, 0.0,-1.0, 1.0), BOX_ID + 5);
                }
                C.compute_geometry();
                double mm;
                VBW::vec3 g;
                C.compute_mg(mm, g);
                m[v] = mm;
                mg[3 * size_t(v)] = g.x;
                mg[3 * size_t(v) + 1] = g.y;
                mg[3 * size_t(v) + 2] = g.z;
                if(nfaces != nullptr) {
                    nfaces[v] = collect_faces(
                        C, fcap, face_nbr + size_t(v) * fcap,
                        face_area ? face_area + size_t(v) * fcap : nullptr);
                }
            }
        });
    double t2 = now_s();
    if(times) {
        times[0] = t1 - t0;
        times[1] = t2 - t1;
    }
    return 0;
}

/*
 * CVT energy per cell on the exact cells of oracle A: f_i = sum over the integration
 * simplices (seed, p1, p2, p3) of the cell of
 *     (mi/10) * (U.U + V.V + W.W + U.V + V.W + W.U),  U = p1-p0, V = p2-p0, W = p3-p0,
 * the formula of ComputeCVTFuncGradVolumetric (voronoi/RVD.cpp:846-857); the simplices are
 * the fans of the cell facets as ConvexCell::compute_mg walks them (convex_cell.cpp:1267-1307),
 * with the seed as apex.  The gradient of that functor is 2*(m_i p_i - mg_i) and needs no
 * extra output.  Returns 1 if has_empty_cells().
 */
int georef_cvt_energy(
    uint32_t n, const double* xyz, const double* w, int periodic, double period,
    double* energy
) {
    georef_init(0);
    GEO::SmartPointer<GEO::PeriodicDelaunay3d> del =
        new GEO::PeriodicDelaunay3d(periodic != 0, period);
    if(!periodic) {
        del->set_keeps_infinite(true);
    }
    del->set_stores_cicl(false);
    del->set_vertices(n, xyz);
    if(w != nullptr) {
        del->set_weights(w);
    }
    del->compute();
    if(del->has_empty_cells()) {
        return 1;
    }
    GEO::parallel_for_slice(0, n,
        [&](GEO::index_t b, GEO::index_t e) {
            VBW::ConvexCell C;
            GEO::PeriodicDelaunay3d::IncidentTetrahedra W;
            for(GEO::index_t v = b; v < e; ++v) {
                del->copy_Laguerre_cell_from_Delaunay(v, C, W);
                if(!periodic) {
                    C.clip_by_plane(VBW::vec4( 1.0, 0.0, 0.0, 0.0), BOX_ID + 0);
                    C.clip_by_plane(VBW::vec4(-1.0, 0.0, 0.0, 1.0), BOX_ID + 1);
                    C.clip_by_plane(VBW::vec4( 0.0, 1.0, 0.0, 0.0), BOX_ID + 2);
                    C.clip_by_plane(VBW::vec4( 0.0,-1.0, 0.0, 1.0), BOX_ID + 3);
                    C.clip_by_plane(VBW::vec4( 0.0, 0.0, 1.0, 0.0), BOX_ID + 4);
                    C.clip_by_plane(VBW::vec4( 0.0, 0.0,-1.0, 1.0), BOX_ID + 5);
                }
                C.compute_geometry();
                const double* p0 = xyz + 3 * size_t(v);
                double f = 0.0;
                for(VBW::index_t lv = 1; lv < C.nb_v(); ++lv) {
                    VBW::index_t tfirst = C.vertex_triangle(lv);
                    if(tfirst == VBW::END_OF_LIST) {
                        continue;
                    }
                    VBW::ushort t1t2[2];
                    VBW::index_t cur = 0, t = tfirst;
                    do {
                        if(cur < 2) {
                            t1t2[cur] = VBW::ushort(t);
                        } else {
                            VBW::vec3 P[3] = {C.triangle_point(t1t2[0]), C.triangle_point(t1t2[1]),
                                              C.triangle_point(VBW::ushort(t))};
                            double U[3][3];
                            for(int k = 0; k < 3; ++k) {
                                U[k][0] = P[k].x - p0[0];
                                U[k][1] = P[k].y - p0[1];
                                U[k][2] = P[k].z - p0[2];
                            }
                            double det = U[0][0] * (U[1][1] * U[2][2] - U[1][2] * U[2][1])
                                       - U[0][1] * (U[1][0] * U[2][2] - U[1][2] * U[2][0])
                                       + U[0][2] * (U[1][0] * U[2][1] - U[1][1] * U[2][0]);
                            double mi = std::fabs(det) / 6.0;
                            double fi = 0.0;
                            for(int c = 0; c < 3; ++c) {
                                double Uc = U[0][c], Vc = U[1][c], Wc = U[2][c];
                                fi += Uc * Uc + Vc * Vc + Wc * Wc;
                                fi += Uc * Vc + Vc * Wc + Wc * Uc;
                            }
                            f += fi * (mi / 10.0);
                            t1t2[1] = VBW::ushort(t);
                        }
                        ++cur;
                        VBW::index_t le = C.triangle_find_vertex(t, lv);
                        t = C.triangle_adjacent(t, (le + 1) % 3);
                    } while(t != tfirst);
                }
                energy[v] = f;
            }
        });
    return 0;
}

/*
 * Delaunay tetrahedra of the unmodified reference (Delaunay::cell_vertex, delaunay/delaunay.h:336-485)
 * for a non-periodic point set: tets[4*k .. 4*k+3] = vertex indices of finite cell k.
 * Returns the number of finite cells (only min(count, cap) are written).
 */
uint32_t georef_delaunay_tets(uint32_t n, const double* xyz, uint32_t* tets, uint32_t cap) {
    georef_init(0);
    GEO::SmartPointer<GEO::PeriodicDelaunay3d> del = new GEO::PeriodicDelaunay3d(false, 1.0);
    del->set_keeps_infinite(false);
    del->set_stores_cicl(false);
    del->set_vertices(n, xyz);
    del->compute();
    uint32_t count = 0;
    for(GEO::index_t c = 0; c < del->nb_cells(); ++c) {
        if(del->cell_is_infinite(c)) {
            continue;
        }
        if(count < cap) {
            for(GEO::index_t lv = 0; lv < 4; ++lv) {
                tets[4 * size_t(count) + lv] = uint32_t(del->cell_vertex(c, lv));
            }
        }
        ++count;
    }
    return count;
}

/* BRIO order and levels of the unmodified reference (compute_BRIO_order, mesh/mesh_reorder.cpp:1287-1315):
 * order[n], levels[cap]; returns the number of levels written (levels[0] = 0 ... = n) */
uint32_t georef_brio_order(uint32_t n, const double* xyz, uint32_t threshold, double ratio,
                           uint32_t* order, uint32_t* levels, uint32_t cap) {
    georef_init(0);
    GEO::vector<GEO::index_t> sorted, lv;
    GEO::compute_BRIO_order(n, xyz, sorted, 3, 3, threshold, ratio, &lv);
    for(uint32_t i = 0; i < n; ++i) {
        order[i] = uint32_t(sorted[i]);
    }
    for(uint32_t l = 0; l < lv.size() && l < cap; ++l) {
        levels[l] = uint32_t(lv[l]);
    }
    return uint32_t(lv.size());
}

/*
 * Oracle B ("same algorithm").  Box = [box[0..2], box[3..5]].  k0 = initial
 * neighbour count (reference default 20, delaunay_nn.cpp:49); growth k += k/8
 * (k > 8) else k++, capped at min(kcap, n-1); each growth re-queries the kd-tree
 * from scratch (delaunay_nn.cpp:57-71).  Stop rule d2 > 4.1*R2 (generic_RVD.h:2322).
 * w (optional) turns bisectors into power planes (periodic_delaunay_3d.cpp:3719-3726
 * convention); the stop rule is then the weighted one stated in DESIGN.md.
 * kstop[n] (optional): k at stop, or 0xFFFFFFFF if the cell stayed uncertified.
 * stats[0..3] = sum P, C, V, I of the SURVEY 8(d) flop model; stats[4] = number
 * of uncertified cells; stats[5] = total kNN rounds.
 * times[0] = kd-tree build, times[1] = everything else.
 */
/* convex domain for georef_cells_knn: planes a x + b y + c z + d >= 0 clipped by the reference's own
 * ConvexCell::clip_by_plane right after init_with_box; face id of plane k = DOMAIN_ID + k */
static std::vector<double> g_domain;
static const GEO::index_t DOMAIN_ID = 0xFFFFFF00u;
void georef_set_domain_planes(uint32_t n, const double* planes) {
    g_domain.assign(planes, planes + 4 * size_t(n));
}

int georef_cells_knn(
    uint32_t n, const double* xyz, const double* w, const double* box,
    uint32_t k0, uint32_t kcap, int exact_predicates,
    double* m, double* mg, uint32_t* nfaces, uint32_t* face_nbr,
    double* face_area, uint32_t fcap, uint32_t* kstop,
    double* stats, double* times
) {
    georef_init(0);
    double t0 = now_s();
    GEO::NearestNeighborSearch_var nn = GEO::NearestNeighborSearch::create(3, "BNN");
    nn->set_points(n, xyz);
    double t1 = now_s();
    double wmax = 0.0;
    if(w != nullptr) {
        wmax = w[0];
        for(uint32_t i = 1; i < n; ++i) {
            wmax = std::max(wmax, w[i]);
        }
    }
    kcap = std::min(kcap, n - 1);
    std::atomic<uint64_t> sP(0), sC(0), sV(0), sI(0), sU(0), sR(0);
    GEO::parallel_for_slice(0, n,
        [&](GEO::index_t b, GEO::index_t e) {
            VBW::ConvexCell C(VBW::WithVGlobal);
            C.use_exact_predicates(exact_predicates != 0);
            std::vector<GEO::index_t> nb;
            std::vector<double> d2;
            uint64_t lP = 0, lC = 0, lV = 0, lI = 0, lU = 0, lR = 0;
            for(GEO::index_t i = b; i < e; ++i) {
                const double* pi = xyz + 3 * size_t(i);
                double wi = (w != nullptr) ? w[i] : 0.0;
                VBW::vec3 Pi = VBW::make_vec3(pi[0], pi[1], pi[2]);
                double Pi_len2 = pi[0]*pi[0] + pi[1]*pi[1] + pi[2]*pi[2];
                C.init_with_box(box[0], box[1], box[2], box[3], box[4], box[5]);
                for(VBW::index_t lv = 1; lv <= 6; ++lv) {
                    C.set_v_global_index(lv, BOX_ID + (lv - 1));
                }
                for(size_t kd = 0; kd < g_domain.size() / 4; ++kd) {
                    C.clip_by_plane(VBW::make_vec4(g_domain[4 * kd], g_domain[4 * kd + 1], g_domain[4 * kd + 2],
                                                   g_domain[4 * kd + 3]), DOMAIN_ID + GEO::index_t(kd));
                }
                uint32_t k = std::min(k0, kcap);
                uint32_t jj = 1;       // rank 0 is the seed itself
                bool certified = false;
                for(;;) {
                    nb.resize(k + 1);
                    d2.resize(k + 1);
                    nn->get_nearest_neighbors(k + 1, i, nb.data(), d2.data());
                    ++lR;
                    for(; jj <= k; ++jj) {
                        double R2 = C.squared_radius(Pi);
                        double bound = 4.1 * R2;
                        if(w != nullptr) {
                            // weighted stop rule: d > R + sqrt(R^2 + (wmax - wi))
                            double R = ::sqrt(R2);
                            double s = R + ::sqrt(R2 + (wmax - wi));
                            bound = 1.025 * s * s;
                        }
                        if(d2[jj] > bound) {
                            certified = true;
                            break;
                        }
                        GEO::index_t j = nb[jj];
                        const double* pj = xyz + 3 * size_t(j);
                        double wj = (w != nullptr) ? w[j] : 0.0;
                        double Pj_len2 = pj[0]*pj[0] + pj[1]*pj[1] + pj[2]*pj[2];
                        VBW::vec4 P = VBW::make_vec4(
                            2.0 * (pi[0] - pj[0]),
                            2.0 * (pi[1] - pj[1]),
                            2.0 * (pi[2] - pj[2]),
                            ((wi - Pi_len2) - (wj - Pj_len2)));
                        uint64_t live = 0;
                        for(VBW::ushort t = C.first_triangle();
                            t != VBW::END_OF_LIST; t = C.next_triangle(t)) {
                            ++live;
                        }
                        C.clip_by_plane(P, j);
                        VBW::index_t lv = C.nb_v() - 1;
                        for(VBW::ushort t = C.first_triangle();
                            t != VBW::END_OF_LIST; t = C.next_triangle(t)) {
                            lV += (C.triangle_v_local_index(t, 0) == lv);
                        }
                        ++lP;
                        lC += live;
                    }
                    if(certified || k >= kcap) {
                        break;
                    }
                    k = (k > 8) ? k + k / 8 : k + 1;
                    k = std::min(k, kcap);
                }
                if(!certified && k + 1 >= n) {
                    certified = true;  // every other seed was clipped
                }
                if(kstop != nullptr) {
                    kstop[i] = certified ? jj : 0xFFFFFFFFu;
                }
                lU += certified ? 0 : 1;
                C.compute_geometry();
                double mm;
                VBW::vec3 g;
                C.compute_mg(mm, g);
                m[i] = mm;
                mg[3 * size_t(i)] = g.x;
                mg[3 * size_t(i) + 1] = g.y;
                mg[3 * size_t(i) + 2] = g.z;
                uint32_t nf = 0;
                uint32_t tmp_ids[256];
                if(nfaces != nullptr) {
                    nf = collect_faces(
                        C, fcap, face_nbr + size_t(i) * fcap,
                        face_area ? face_area + size_t(i) * fcap : nullptr);
                    nfaces[i] = nf;
                } else {
                    nf = collect_faces(C, 256, tmp_ids, nullptr);
                }
                // integration tets: every face but the first is fanned (deg-2),
                // deg summed over faces = 3 * (#triangles) ; #triangles = 2F-4.
                if(nf >= 4) {
                    uint64_t T = 2 * uint64_t(nf) - 4;
                    lI += 3 * T - 2 * uint64_t(nf);   // sum(deg-2) over all faces
                }
            }
            sP += lP; sC += lC; sV += lV; sI += lI; sU += lU; sR += lR;
        });
    double t2 = now_s();
    if(stats) {
        stats[0] = double(sP.load());
        stats[1] = double(sC.load());
        stats[2] = double(sV.load());
        stats[3] = double(sI.load());
        stats[4] = double(sU.load());
        stats[5] = double(sR.load());
    }
    if(times) {
        times[0] = t1 - t0;
        times[1] = t2 - t1;
    }
    return 0;
}

}  // extern "C"
