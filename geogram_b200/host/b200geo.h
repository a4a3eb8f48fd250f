/*
 * b200geo.h -- C++ host side of the B200 Voronoi engine, above the C ABI (include/b200vor.h).
 *
 * The classes mirror the reference's interfaces for this path -- same names, argument meaning
 * and error behaviour -- so that code written against geogram reads the same
 * (paths relative to /root/reference/src/lib/geogram):
 *
 *   B200GEO::NearestNeighborSearch       GEO::NearestNeighborSearch      points/nn_search.h:69-257
 *   B200GEO::Delaunay_NearestNeighbors   GEO::Delaunay_NearestNeighbors  delaunay/delaunay_nn.h:55-135
 *                                        (+ Delaunay::get_neighbors      delaunay/delaunay.h:421-440)
 *   B200GEO::ConvexCellBatch             VBW::ConvexCell queries         voronoi/convex_cell.h:657-909
 *   B200GEO::CentroidalVoronoiTesselation  GEO::CentroidalVoronoiTesselation::Lloyd_iterations
 *                                                                        voronoi/CVT.cpp:133-169
 *
 * Differences that follow from the device: queries are answered from batched launches (the
 * reference calls get_nearest_neighbors() one query at a time from parallel_for workers,
 * delaunay/delaunay.cpp:269-273; here the by-index overload is served from a table computed
 * for all points in one launch), and there is NO CPU fallback: constructing an object without
 * a CUDA device throws std::runtime_error.
 */
#ifndef B200GEO_H
#define B200GEO_H

#include <stdint.h>

#include <mutex>
#include <ostream>
#include <string>
#include <vector>

struct b200vor_ctx;

namespace B200GEO {

typedef uint32_t index_t;          // GEO::index_t (non-GARGANTUA build, api/defs.h:111-125)
typedef uint32_t global_index_t;   // VBW::global_index_t
typedef uint8_t coord_index_t;

struct vec3 {
    double x, y, z;
};

/* ids >= BOX_ID in face lists are the six planes of the clipping box (x-,x+,y-,y+,z-,z+) */
const global_index_t BOX_ID = 0xFFFFFFF0u;
/* ids in [DOMAIN_ID, BOX_ID) are planes of a convex domain (b200vor_set_domain_planes) */
const global_index_t DOMAIN_ID = 0xFFFFFF00u;

class NearestNeighborSearch {
public:
    /* names: "B200NN", "default".  Like the reference factory (nn_search.cpp:120-150) an unknown
       name logs a warning to stderr and falls back to the default; dimension != 3 returns nullptr. */
    static NearestNeighborSearch* create(coord_index_t dimension, const std::string& name = "default");
    virtual ~NearestNeighborSearch();

    /* points are BORROWED like in the reference (nn_search.cpp:88-94) and copied to the device */
    virtual void set_points(index_t nb_points, const double* points);
    virtual bool stride_supported() const { return true; }
    virtual void set_points(index_t nb_points, const double* points, index_t stride);

    /* ascending squared distance; exact; thread-safe (serialised on the device) */
    virtual void get_nearest_neighbors(index_t nb_neighbors, const double* query_point,
                                       index_t* neighbors, double* neighbors_sq_dist) const;
    struct KeepInitialValues {};
    virtual void get_nearest_neighbors(index_t nb_neighbors, const double* query_point,
                                       index_t* neighbors, double* neighbors_sq_dist,
                                       KeepInitialValues dummy) const;
    /* by stored point index: the point itself comes first (kd_tree.cpp:156-189) */
    virtual void get_nearest_neighbors(index_t nb_neighbors, index_t query_point,
                                       index_t* neighbors, double* neighbors_sq_dist) const;
    index_t get_nearest_neighbor(const double* query_point) const;

    /* batched forms (one launch): rows of nb_neighbors */
    void get_nearest_neighbors_all(index_t nb_neighbors, index_t* neighbors, double* sq_dist) const;
    void get_nearest_neighbors_batch(index_t nb_queries, const double* query_points,
                                     index_t nb_neighbors, index_t* neighbors, double* sq_dist) const;

    coord_index_t dimension() const { return dimension_; }
    index_t nb_points() const { return nb_points_; }
    const double* point_ptr(index_t i) const { return points_ + size_t(i) * stride_; }
    bool exact() const { return exact_; }
    virtual void set_exact(bool x) { exact_ = x; }   /* always exact on the device */
    b200vor_ctx* context() const { return ctx_; }

protected:
    explicit NearestNeighborSearch(coord_index_t dimension);
    coord_index_t dimension_;
    index_t nb_points_;
    index_t stride_;
    const double* points_;
    bool exact_;
    b200vor_ctx* ctx_;
    mutable std::mutex lock_;
    mutable index_t table_k_;
    mutable std::vector<index_t> table_nbr_;
    mutable std::vector<double> table_d2_;
};

class Delaunay_NearestNeighbors {
public:
    explicit Delaunay_NearestNeighbors(coord_index_t dimension);   /* throws InvalidDimension-like */
    ~Delaunay_NearestNeighbors();
    void set_vertices(index_t nb_vertices, const double* vertices);
    index_t nb_vertices() const { return nb_vertices_; }
    const double* vertex_ptr(index_t i) const { return vertices_ + 3 * size_t(i); }
    index_t nearest_vertex(const double* p) const;
    NearestNeighborSearch* nn_search() { return NN_; }
    void set_default_nb_neighbors(index_t x) { default_nb_neighbors_ = x; }
    index_t default_nb_neighbors() const { return default_nb_neighbors_; }
    /* neighbours of v, nearest first, v itself excluded; a point that duplicates a lower index
       gets no neighbours, higher-index duplicates are skipped (delaunay_nn.cpp:122-135) */
    void get_neighbors(index_t v, std::vector<index_t>& neighbors) const;
    /* re-queries v with nb neighbours (delaunay_nn.cpp:57-71) */
    void enlarge_neighborhood(index_t v, index_t nb);

private:
    void store_neighbors(index_t v, index_t nb, const index_t* row, const double* d2);
    NearestNeighborSearch* NN_;
    index_t nb_vertices_;
    const double* vertices_;
    index_t default_nb_neighbors_;
    std::vector<std::vector<index_t> > neighbors_;
};

/* All cells of a seed set at once; per-cell accessors named after VBW::ConvexCell. */
class ConvexCellBatch {
public:
    ConvexCellBatch();
    ~ConvexCellBatch();
    /* clipping box {xmin,ymin,zmin,xmax,ymax,zmax} (ConvexCell::init_with_box) -- default unit cube */
    void init_with_box(double xmin, double ymin, double zmin, double xmax, double ymax, double zmax);
    /* periodic cube [0,period)^3: neighbour ids are real + N*instance (periodic.h:95-99) */
    void set_periodic(double period);
    /* weights may be null (Voronoi) */
    void compute(index_t nb_seeds, const double* xyz, const double* weights, bool with_faces = true,
                 bool with_vertices = false);

    index_t nb_cells() const { return index_t(m_.size()); }
    bool empty(index_t c) const { return (status_[c] & 4u) != 0; }
    uint8_t status(index_t c) const { return status_[c]; }   /* B200VOR_CELL_* bits, 0 = exact & certified */
    double volume(index_t c) const { return m_[c]; }
    void compute_mg(index_t c, double& m, vec3& mg) const;
    vec3 barycenter(index_t c) const;
    /* contributing vertices (= faces) of the cell, ascending global index */
    index_t nb_v(index_t c) const { return nfaces_.empty() ? 0 : nfaces_[c]; }
    global_index_t v_global_index(index_t c, index_t lv) const { return face_nbr_[size_t(c) * 64 + lv]; }
    double facet_area(index_t c, index_t lv) const { return face_area_[size_t(c) * 64 + lv]; }
    double total_volume() const;

    /* ---- cell geometry (compute(..., with_vertices = true)), names of VBW::ConvexCell ---- */
    /* number of cell vertices = dual triangles (ConvexCell::nb_t); 0 when not available */
    index_t nb_t(index_t c) const;
    /* ConvexCell::triangle_point (convex_cell.h:888) */
    vec3 triangle_point(index_t c, index_t t) const;
    /* the three faces meeting at vertex t, as local face indices lv (ConvexCell::triangle_v_local_index) */
    index_t triangle_v_local_index(index_t c, index_t t, index_t lv) const;
    /* facets of the cell as oriented vertex loops, one per local face index */
    void facet_loops(index_t c, std::vector<std::vector<index_t> >& loops) const;
    /* ConvexCell::save (convex_cell.cpp:282-352): "v"/"f" records, vertices shrunk towards the
       centroid; returns the number of vertices written */
    index_t save(index_t c, std::ostream& out, global_index_t v_offset = 1, double shrink = 0.0) const;
    /* PeriodicDelaunay3d::save_cells (periodic_delaunay_3d.cpp:4363-4392): every cell, shrink 0.1 */
    void save_cells(const std::string& filename, double shrink = 0.1) const;
    /* Delaunay tetrahedra (Delaunay::cell_vertex, delaunay.h:336-485): 4 vertex ids each
       (periodic: real + N*instance, in the frame of the reporting cell), every tetrahedron once.
       Box domain: tetrahedra whose circumcentre the box clipped away are not exported.
       Returns the number of cells WITHOUT a vertex list (finished by the kNN-table pipeline, or
       more than 64 vertices): their tetrahedra are missing, the caller must check for 0. */
    index_t delaunay_tetrahedra(std::vector<global_index_t>& tets) const;
    /* Delaunay::cell_adjacent (delaunay.h:369-380) for a list of tetrahedra (4 ids each):
       adjacent[4 c + lf] = tetrahedron sharing the facet opposite to local vertex lf, or -1.
       Pure host index work (facets matched by their sorted vertex triples).  Periodic domain: pass
       the number of real seeds; facets are matched in a canonical frame (smallest vertex = instance
       0) and no entry is -1.  Box domain (nb_real_vertices = 0): -1 = convex hull, or a neighbour
       that delaunay_tetrahedra does not export. */
    static void tet_adjacency(const std::vector<global_index_t>& tets, std::vector<int64_t>& adjacent,
                              index_t nb_real_vertices = 0);

private:
    b200vor_ctx* ctx_;
    std::vector<double> m_, mg_, face_area_, vert_xyz_;
    std::vector<uint8_t> status_;
    std::vector<uint32_t> nfaces_, face_nbr_, nverts_, vert_faces_;
};

class CentroidalVoronoiTesselation {
public:
    /* volumetric CVT of the unit cube (or a periodic cube); dimension must be 3 */
    explicit CentroidalVoronoiTesselation(coord_index_t dimension = 3);
    ~CentroidalVoronoiTesselation();
    void set_domain_box(const double* box6);
    void set_periodic(double period);
    /* volumetric CVT of a tetrahedral mesh, the reference's set_volumetric(true) on a mesh with cells
       (voronoi/CVT.h, RVD.h:284-290 compute_centroids on the tetrahedra): nb_tets x 4 vertex indices */
    void set_volumetric_mesh(index_t nb_vertices, const double* vertices, index_t nb_tets, const index_t* tets);
    /* ... or of a convex polytope: nb_planes half-spaces a x + b y + c z + d >= 0 cutting the box */
    void set_domain_planes(index_t nb_planes, const double* planes);
    void set_points(index_t nb_points, const double* points);
    void resize_points(index_t nb_points) { points_.resize(3 * size_t(nb_points)); }
    index_t nb_points() const { return index_t(points_.size() / 3); }
    double* embedding(index_t p) { return &points_[3 * size_t(p)]; }
    const double* embedding(index_t p) const { return &points_[3 * size_t(p)]; }
    void lock_point(index_t i);
    void unlock_all_points() { point_is_locked_.clear(); }
    bool point_is_locked(index_t i) const { return !point_is_locked_.empty() && point_is_locked_[i]; }
    /* p_j <- mg_j / m_j where m_j > 1e-30 and j is not locked (CVT.cpp:150-163) */
    void Lloyd_iterations(index_t nb_iter);
    /* CVT objective and gradient at x (n = 3 * nb_points doubles), the callback
       Newton_iterations() hands to its optimizer (CVT.cpp:323-341):
       f = sum_i integral over cell i of |y - x_i|^2, g_i = 2 (m_i x_i - mg_i)
       (ComputeCVTFuncGradVolumetric, RVD.cpp:791-876); locked points get g = 0
       (constrain_points, CVT.cpp:309-321).  Security-radius checking is always on. */
    void funcgrad(index_t n, double* x, double& f, double* g);
    /* quasi-Newton minimisation of the CVT energy (CVT.cpp:272-307): L-BFGS with m correction pairs
       on the host, every objective evaluation = one funcgrad() on the device.  The reference hands
       funcgrad to HLBFGS (third party); this is the textbook two-loop recursion with Armijo
       backtracking, so iterates differ from HLBFGS while the fixed points are the same. */
    void Newton_iterations(index_t nb_iter, index_t m = 7);
    /* energy of the current points (last funcgrad evaluation of Newton_iterations) */
    double last_energy() const { return last_f_; }

private:
    b200vor_ctx* ctx_;
    std::vector<double> points_;
    std::vector<bool> point_is_locked_;
    double last_f_ = 0.0;
};

/* Semi-discrete optimal transport between the uniform measure of the domain (unit cube, a box or
   a periodic cube) and weighted Dirac masses at the points.  Mirrors the part of
   OptimalTransportMap3d (BrunoLevy/exploragram, not in the reference tree) that its callers use
   (src/examples/exploragram/compute_OTM/main.cpp:284-297): set_points, set_epsilon, optimize,
   weight().  The domain is a box instead of a tetrahedral mesh. */
class OptimalTransportMap3d {
public:
    OptimalTransportMap3d();
    ~OptimalTransportMap3d();
    void set_domain_box(const double* box6);
    void set_periodic(double period);
    void set_points(index_t nb_points, const double* points);
    /* target measure of point i (default: |domain| / nb_points each) */
    void set_nu(index_t i, double nu);
    /* maximum relative measure error in a cell (OptimalTransportMap::set_epsilon) */
    void set_epsilon(double eps) { epsilon_ = eps; }
    /* damped Newton iterations on the weights, all on the device */
    void optimize(index_t max_iterations);
    /* multilevel solve (OptimalTransportMap::optimize_levels, call site
       src/examples/exploragram/compute_OTM/main.cpp:292-297): `levels` = {0, e_1, ..., nb_points} as
       filled by the reference's sampling (points in BRIO order: every prefix [0, e_l) is a uniform
       sub-sample).  Level l solves the transport problem of the points [0, e_l) alone (target
       masses rescaled to the measure of the domain); every point of the next level starts from
       a linear least-squares fit of the solved weights over its 24 nearest solved points (one
       batched by-point kNN query on the device): the smooth part of the potential carries over.  A start that still leaves a cell empty is pulled towards the
       Voronoi diagram (weights halved) until every cell is non-empty (KMT hypothesis). */
    void optimize_levels(const std::vector<index_t>& levels, index_t max_iterations);
    /* value at q of the least-squares plane through (pts[row[j]], w[row[j]]), j < k (nearest weight
       when the neighbourhood is rank deficient); pure host code */
    static double extrapolate_weight(const double* pts, const double* w, const double* q,
                                     const uint32_t* row, uint32_t k);
    /* Newton iterations of every level of the last optimize_levels call */
    const std::vector<index_t>& level_newton_iterations() const { return level_iters_; }
    index_t nb_points() const { return index_t(weights_.size()); }
    double weight(index_t i) const { return weights_[i]; }
    /* measure of the Laguerre cell of point i at the current weights */
    double cell_measure(index_t i) const { return measure_[i]; }
    vec3 cell_centroid(index_t i) const;
    bool converged() const { return converged_; }
    index_t nb_newton_iterations() const { return newton_iters_; }

private:
    b200vor_ctx* ctx_;
    std::vector<double> points_, weights_, nu_, measure_, mg_;
    double epsilon_;
    bool converged_;
    index_t newton_iters_;
    std::vector<index_t> level_iters_;
    double period_;        /* > 0: periodic cube */
    double box_[6];
    bool has_box_;
    void solve_prefix(index_t e, index_t max_iterations, const char* who);
};

/* BRIO order and its levels (GEO::compute_BRIO_order, mesh/mesh_reorder.cpp:1287-1315; dimension 3):
   random shuffle, then recursively the first `ratio` of every interval longer than `threshold`; the rest of
   the interval is sorted along the 3-D Hilbert curve of a 2^16 lattice over the bounding box (the reference
   sorts by recursive median splits: same role, not the same permutation); `levels` (optional) = {0, end of
   every interval} -- identical to the reference's for the same sizes.  Feeds
   OptimalTransportMap3d::optimize_levels.  Pure host code. */
void compute_BRIO_order(index_t nb_vertices, const double* vertices, std::vector<index_t>& sorted_indices,
                        index_t stride = 3, index_t threshold = 300, double ratio = 0.125,
                        std::vector<index_t>* levels = nullptr, uint64_t seed = 0);
/* index of the lattice point (x, y, z), each below 2^bits, on the 3-D Hilbert curve (Skilling's transpose form) */
uint64_t hilbert_index_3d(uint32_t x, uint32_t y, uint32_t z, int bits);

}  // namespace B200GEO

#endif
