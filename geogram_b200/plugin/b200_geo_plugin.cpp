/* b200_geo_plugin.cpp -- see b200_geo_plugin.h.  Compiled against the reference's headers only as a
 * test target (oracle/Makefile: plugin); links the unmodified geogram built under oracle/_ref and libb200vor.so. */
#include "b200_geo_plugin.h"

#include <geogram/basic/command_line.h>
#include <geogram/basic/logger.h>
#include <geogram/delaunay/periodic.h>

#include <algorithm>
#include <map>

namespace {
    void check(int rc, const char* what) {
        if(rc != B200VOR_OK) {
            GEO::Logger::err("B200") << what << ": " << b200vor_last_error() << std::endl;
            geo_assert_not_reached;
        }
    }
}

namespace GEO {

    /*********************************************************************** NearestNeighborSearch */

    NearestNeighborSearch_B200::NearestNeighborSearch_B200(coord_index_t dim) :
        NearestNeighborSearch(dim), ctx_(nullptr), table_k_(0), leader_active_(false),
        nb_point_launches_(0), nb_table_launches_(0) {
        geo_assert(dim == 3);
        check(b200vor_create(&ctx_, 0), "b200vor_create");   // no GPU: fails loudly, no fallback
    }

    NearestNeighborSearch_B200::~NearestNeighborSearch_B200() {
        b200vor_destroy(ctx_);
    }

    void NearestNeighborSearch_B200::set_points(index_t nb_points, const double* points) {
        set_points(nb_points, points, index_t(dimension()));
    }

    bool NearestNeighborSearch_B200::stride_supported() const {
        return true;
    }

    void NearestNeighborSearch_B200::set_points(index_t nb_points, const double* points, index_t stride) {
        // the array stays borrowed, as in the reference (nn_search.cpp:88-110); the device gets a copy
        nb_points_ = nb_points;
        points_ = points;
        stride_ = stride;
        std::lock_guard<std::mutex> g(table_lock_);
        table_k_ = 0;
        check(b200vor_set_points(ctx_, nb_points, points, stride, nullptr, 0), "b200vor_set_points");
    }

    void NearestNeighborSearch_B200::ensure_table(index_t k) const {
        // caller holds table_lock_.  One launch answers the query of EVERY stored point; a larger k
        // than the table holds doubles it (enlarge_neighborhood asks for k + k/8, again and again)
        if(table_k_ >= k) {
            return;
        }
        index_t kk = std::min(nb_points(), std::max(k, 2 * table_k_));
        table_nbr_.resize(size_t(nb_points()) * kk);
        table_d2_.resize(size_t(nb_points()) * kk);
        check(b200vor_knn(ctx_, kk, table_nbr_.data(), table_d2_.data(), 0), "b200vor_knn");
        table_k_ = kk;
        ++nb_table_launches_;
    }

    void NearestNeighborSearch_B200::get_nearest_neighbors(
        index_t nb_neighbors, index_t q, index_t* neighbors, double* neighbors_sq_dist
    ) const {
        geo_assert(q < nb_points() && nb_neighbors <= nb_points());
        std::lock_guard<std::mutex> g(table_lock_);
        ensure_table(nb_neighbors);
        std::copy_n(&table_nbr_[size_t(q) * table_k_], nb_neighbors, neighbors);
        std::copy_n(&table_d2_[size_t(q) * table_k_], nb_neighbors, neighbors_sq_dist);
    }

    void NearestNeighborSearch_B200::run_batch(std::vector<Request*>& batch) const {
        // one launch per distinct k of the batch (callers of one parallel_for all use the same k)
        std::map<index_t, std::vector<Request*> > by_k;
        for(Request* r : batch) {
            by_k[r->k].push_back(r);
        }
        for(auto& kv : by_k) {
            const index_t k = kv.first;
            std::vector<Request*>& rq = kv.second;
            const index_t nq = index_t(rq.size());
            qbuf_.resize(3 * size_t(nq));
            nbrbuf_.resize(size_t(nq) * k);
            d2buf_.resize(size_t(nq) * k);
            for(index_t i = 0; i < nq; ++i) {
                std::copy_n(rq[i]->q, 3, &qbuf_[3 * size_t(i)]);
            }
            check(b200vor_knn_points(ctx_, nq, qbuf_.data(), k, nbrbuf_.data(), d2buf_.data(), 0),
                  "b200vor_knn_points");
            ++nb_point_launches_;
            for(index_t i = 0; i < nq; ++i) {
                std::copy_n(&nbrbuf_[size_t(i) * k], k, rq[i]->nbr);
                std::copy_n(&d2buf_[size_t(i) * k], k, rq[i]->d2);
            }
        }
    }

    void NearestNeighborSearch_B200::get_nearest_neighbors(
        index_t nb_neighbors, const double* query_point, index_t* neighbors, double* neighbors_sq_dist
    ) const {
        geo_assert(nb_neighbors <= nb_points());
        // a query that IS a stored point (the reference's own acceptance loop queries point_ptr(i),
        // src/tests/test_nn_search/main.cpp:136-157) is answered from the table, no launch
        if(query_point >= points_ && query_point < points_ + size_t(nb_points_) * stride_) {
            const size_t off = size_t(query_point - points_);
            if(off % stride_ == 0) {
                get_nearest_neighbors(nb_neighbors, index_t(off / stride_), neighbors, neighbors_sq_dist);
                return;
            }
        }
        // any other point: combined with the queries of the other calling threads into one launch
        Request r = {query_point, nb_neighbors, neighbors, neighbors_sq_dist, false};
        std::unique_lock<std::mutex> lk(queue_lock_);
        pending_.push_back(&r);
        while(!r.done) {
            if(leader_active_) {
                queue_cv_.wait(lk);
                continue;
            }
            // this thread serves everything queued so far (its own request included)
            leader_active_ = true;
            std::vector<Request*> batch;
            batch.swap(pending_);
            lk.unlock();
            run_batch(batch);
            lk.lock();
            for(Request* b : batch) {
                b->done = true;
            }
            leader_active_ = false;
            queue_cv_.notify_all();
        }
    }

    void NearestNeighborSearch_B200::get_nearest_neighbors(
        index_t nb_neighbors, const double* query_point, index_t* neighbors, double* neighbors_sq_dist,
        KeepInitialValues
    ) const {
        // the initial values are a hint for tree searches (nn_search.h:137-165); the result is the same
        get_nearest_neighbors(nb_neighbors, query_point, neighbors, neighbors_sq_dist);
    }

    /*********************************************************************** Delaunay */

    namespace {
        // runs before the base-class constructor: Delaunay_NearestNeighbors creates its search
        // structure with NearestNeighborSearch::create(dimension), i.e. with algo:nn_search
        thread_local std::string g_saved_nn;
        coord_index_t select_nn(coord_index_t dim) {
            g_saved_nn = CmdLine::get_arg("algo:nn_search");
            CmdLine::set_arg("algo:nn_search", "B200NN");
            return dim;
        }
    }

    Delaunay_B200::Delaunay_B200(coord_index_t dimension) :
        Delaunay_NearestNeighbors(select_nn(dimension)) {
        CmdLine::set_arg("algo:nn_search", g_saved_nn);
        if(dimension != 3) {
            throw InvalidDimension(dimension, "Delaunay_B200", "3");
        }
        geo_assert(dynamic_cast<NearestNeighborSearch_B200*>(nn_search()) != nullptr);
    }

    Delaunay_B200::~Delaunay_B200() {
    }

    void b200_register_plugin() {
        static bool done = false;
        if(done) {
            return;
        }
        done = true;
        geo_register_NearestNeighborSearch_creator(NearestNeighborSearch_B200, "B200NN");
        geo_register_Delaunay_creator(Delaunay_B200, "B200");
    }
}

/*************************************************************************** ConvexCell shim */

namespace B200GEO {

    bool fill_convex_cell(
        VBW::ConvexCell& C, GEO::index_t n, const double* xyz, const double* w, GEO::index_t seed,
        bool periodic, double period, const double* box6,
        GEO::index_t nfaces, const uint32_t* face_nbr, GEO::index_t nverts, const uint32_t* vert_faces,
        GEO::index_t vcap
    ) {
        using namespace VBW;
        if(nverts == 0 || nverts == 0xFFFFFFFFu || nverts > vcap || nfaces == 0) {
            return false;
        }
        C.clear();
        if(!C.has_vglobal()) {
            C.create_vglobal();
        }
        // vertex 0 = vertex at infinity, as copy_Laguerre_cell_from_Delaunay creates it
        // (periodic_delaunay_3d.cpp:3634-3674)
        C.create_vertex(make_vec4(0.0, 0.0, 0.0, 0.0), global_index_t(-1));
        const double* Pi = xyz + 3 * size_t(seed);
        const double wi = w ? w[seed] : 0.0;
        const double Pi_len2 = Pi[0] * Pi[0] + Pi[1] * Pi[1] + Pi[2] * Pi[2];
        for(GEO::index_t lv = 0; lv < nfaces; ++lv) {
            const uint32_t id = face_nbr[lv];
            if(id >= B200VOR_BOX_ID) {
                // init_with_box planes (convex_cell.cpp:175-215): x >= xmin, x <= xmax, ...
                const uint32_t p = id - B200VOR_BOX_ID, axis = p >> 1;
                const bool upper = p & 1u;
                double lo = periodic ? Pi[axis] - 0.5 * period : box6[axis];
                double hi = periodic ? Pi[axis] + 0.5 * period : box6[3 + axis];
                vec4 P = make_vec4(0.0, 0.0, 0.0, 0.0);
                (axis == 0 ? P.x : (axis == 1 ? P.y : P.z)) = upper ? -1.0 : 1.0;
                P.w = upper ? hi : -lo;
                C.create_vertex(P, global_index_t(id));
                continue;
            }
            // power plane of (i, j), periodic_delaunay_3d.cpp:3719-3726, j = real + n * instance
            const GEO::index_t real = id % n, inst = id / n;
            double Pj[3];
            for(int c = 0; c < 3; ++c) {
                Pj[c] = xyz[3 * size_t(real) + c] + (periodic ? double(GEO::Periodic::translation[inst][c]) * period : 0.0);
            }
            const double wj = w ? w[real] : 0.0;
            const double Pj_len2 = Pj[0] * Pj[0] + Pj[1] * Pj[1] + Pj[2] * Pj[2];
            C.create_vertex(
                make_vec4(2.0 * (Pi[0] - Pj[0]), 2.0 * (Pi[1] - Pj[1]), 2.0 * (Pi[2] - Pj[2]),
                          (wi - Pi_len2) - (wj - Pj_len2)),
                global_index_t(id));
        }
        // one dual triangle per cell vertex; local vertex = 1 + position in the face list
        std::vector<GEO::index_t> tri(3 * size_t(nverts));
        std::map<std::pair<GEO::index_t, GEO::index_t>, GEO::index_t> owner;   // directed edge -> triangle
        for(GEO::index_t t = 0; t < nverts; ++t) {
            const uint32_t vf = vert_faces[t];
            const GEO::index_t i = 1 + (vf & 0xFFu), j = 1 + ((vf >> 8) & 0xFFu), k = 1 + ((vf >> 16) & 0xFFu);
            if(i > nfaces || j > nfaces || k > nfaces) {
                return false;
            }
            const GEO::index_t tt = C.create_triangle(i, j, k);
            geo_assert(tt == t);
            tri[3 * size_t(t)] = i; tri[3 * size_t(t) + 1] = j; tri[3 * size_t(t) + 2] = k;
            // connect_triangles convention (convex_cell.cpp:1348-1411): t owns (j,k), (k,i), (i,j)
            owner[std::make_pair(j, k)] = t;
            owner[std::make_pair(k, i)] = t;
            owner[std::make_pair(i, j)] = t;
        }
        for(GEO::index_t t = 0; t < nverts; ++t) {
            const GEO::index_t i = tri[3 * size_t(t)], j = tri[3 * size_t(t) + 1], k = tri[3 * size_t(t) + 2];
            auto a0 = owner.find(std::make_pair(k, j)), a1 = owner.find(std::make_pair(i, k)),
                 a2 = owner.find(std::make_pair(j, i));
            if(a0 == owner.end() || a1 == owner.end() || a2 == owner.end()) {
                return false;   // not a closed surface: the cell was flagged by the kernel
            }
            C.set_triangle_adjacent(t, 0, a0->second);
            C.set_triangle_adjacent(t, 1, a1->second);
            C.set_triangle_adjacent(t, 2, a2->second);
        }
        C.compute_geometry();
        return true;
    }
}
