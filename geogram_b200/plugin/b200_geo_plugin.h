/*
 * b200_geo_plugin.h -- the reference-side binding of libb200vor.so: real subclasses of
 * GEO::NearestNeighborSearch and GEO::Delaunay_NearestNeighbors, registered with geogram's own
 * factories, plus the shim that re-hydrates a GPU-computed cell into a real VBW::ConvexCell.
 *
 * This is what a geogram maintainer adds: it compiles against the reference's
 * headers and links libgeogram; nothing in geogram changes.  In this repository it is built as a
 * TEST TARGET only (oracle/Makefile, target `plugin`, against /root/reference/src/lib and the
 * unmodified reference library built under oracle/_ref) and exercised by tests/test_gpu_plugin.py
 * through tests/plugin/test_plugin.cpp; the product library (libb200vor.so) never links the reference.
 *
 *   GEO::NearestNeighborSearch_B200   "B200NN"   points/nn_search.h:69-257, registered as in
 *                                                 src/tests/test_nn_search/main.cpp:60-65
 *   GEO::Delaunay_B200                "B200"     delaunay/delaunay_nn.h:55-135, delaunay.h:798-806
 *   B200GEO::fill_convex_cell                     voronoi/convex_cell.h:677-708, :1019 (the mutators
 *                                                 PeriodicDelaunay3d::copy_Laguerre_cell_from_Delaunay uses,
 *                                                 delaunay/periodic_delaunay_3d.cpp:3634-3733)
 */
#ifndef B200_GEO_PLUGIN_H
#define B200_GEO_PLUGIN_H

#include <geogram/basic/common.h>
#include <geogram/delaunay/delaunay_nn.h>
#include <geogram/points/nn_search.h>
#include <geogram/voronoi/convex_cell.h>

#include <condition_variable>
#include <mutex>
#include <vector>

#include "../../include/b200vor.h"

namespace GEO {

    /**
     * kNN on the GPU behind geogram's NearestNeighborSearch interface.
     * Queries by stored-point index (what Delaunay::update_neighbors issues for every vertex from
     * parallel_for workers, delaunay/delaunay.cpp:259-274) are served from a table that ONE launch
     * computes for all points; queries by point are batched: concurrent callers are combined into
     * one launch (a caller that arrives while a launch is in flight joins the next batch), and a
     * query pointer that lies inside the borrowed point array is served from the table.
     * No CPU fallback: construction fails (geo_assert) without a CUDA device.
     */
    class NearestNeighborSearch_B200 : public NearestNeighborSearch {
    public:
        NearestNeighborSearch_B200(coord_index_t dim);
        void set_points(index_t nb_points, const double* points) override;
        bool stride_supported() const override;
        void set_points(index_t nb_points, const double* points, index_t stride) override;
        void get_nearest_neighbors(
            index_t nb_neighbors, const double* query_point,
            index_t* neighbors, double* neighbors_sq_dist
        ) const override;
        void get_nearest_neighbors(
            index_t nb_neighbors, const double* query_point,
            index_t* neighbors, double* neighbors_sq_dist, KeepInitialValues
        ) const override;
        void get_nearest_neighbors(
            index_t nb_neighbors, index_t query_point,
            index_t* neighbors, double* neighbors_sq_dist
        ) const override;
        b200vor_ctx* context() const { return ctx_; }
        /** launches issued so far for by-point queries / table builds (test instrumentation) */
        index_t nb_point_launches() const { return nb_point_launches_; }
        index_t nb_table_launches() const { return nb_table_launches_; }

    protected:
        ~NearestNeighborSearch_B200() override;

    private:
        struct Request {
            const double* q;
            index_t k;
            index_t* nbr;
            double* d2;
            bool done;
        };
        void run_batch(std::vector<Request*>& batch) const;
        void ensure_table(index_t k) const;

        b200vor_ctx* ctx_;
        mutable std::mutex table_lock_;
        mutable index_t table_k_;
        mutable std::vector<index_t> table_nbr_;
        mutable std::vector<double> table_d2_;
        mutable std::mutex queue_lock_;
        mutable std::condition_variable queue_cv_;
        mutable std::vector<Request*> pending_;
        mutable bool leader_active_;
        mutable index_t nb_point_launches_, nb_table_launches_;
        mutable std::vector<double> qbuf_, d2buf_;
        mutable std::vector<index_t> nbrbuf_;
    };

    /**
     * The kNN Delaunay back-end on the GPU.  RestrictedVoronoiDiagram recognises the kNN flavour by
     * dynamic_cast<Delaunay_NearestNeighbors*> (voronoi/generic_RVD.h:129-131, 206-211), so this
     * subclasses it: CentroidalVoronoiTesselation(&M, 0, "B200") works unchanged.  The base class
     * creates its search structure through the factory with algo:nn_search (delaunay_nn.cpp:50);
     * the constructor selects "B200NN" for that one call.
     */
    class Delaunay_B200 : public Delaunay_NearestNeighbors {
    public:
        Delaunay_B200(coord_index_t dimension);

    protected:
        ~Delaunay_B200() override;
    };

    /** Registers both classes with geogram's factories (idempotent; call after GEO::initialize()). */
    void b200_register_plugin();
}

namespace B200GEO {

    /**
     * One cell of b200vor_compute_cells(want_faces = 2), as returned by b200vor_get_faces /
     * b200vor_get_vertices, re-hydrated into a VBW::ConvexCell (vertex 0 = vertex at infinity,
     * vertices 1..nfaces = the contributing planes in face-list order with their global indices,
     * one triangle per cell vertex, adjacency set) so that every ConvexCell query -- volume,
     * barycenter, facet_area, save, append_to_mesh, further clip_by_plane -- runs on it.
     *
     * seed: the cell's seed; xyz / w: all seeds (w may be null); periodic / period / box6: the domain;
     * face ids >= B200VOR_BOX_ID are the box planes, ids >= n are periodic instances
     * (real + n * instance, delaunay/periodic.h:95-99).  Returns false if the cell is not available
     * (nverts == 0 or 0xFFFFFFFF or > vcap).
     */
    bool fill_convex_cell(
        VBW::ConvexCell& C, GEO::index_t n, const double* xyz, const double* w, GEO::index_t seed,
        bool periodic, double period, const double* box6,
        GEO::index_t nfaces, const uint32_t* face_nbr, GEO::index_t nverts, const uint32_t* vert_faces,
        GEO::index_t vcap
    );
}

#endif
