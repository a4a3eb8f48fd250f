/*
 * b200vor.h -- C ABI of the B200-native meshless Voronoi / Laguerre cell engine.
 *
 * This is the drop-in boundary: plain pointers and sizes,
 * caller-owned buffers, int return codes (0 = success, else see
 * b200vor_last_error()).  No C++ or torch type crosses it.  The C++ host classes in
 * geogram_b200/host/ (mirrors of GEO::NearestNeighborSearch,
 * GEO::Delaunay_NearestNeighbors, VBW::ConvexCell batch output and the Lloyd driver)
 * and the reference-side plugin shown in INTEGRATION.md call only these entry points.
 *
 * Every buffer argument comes with `on_device`: 0 = host memory (the call copies
 * across PCIe inside), 1 = device memory on the context's GPU (no copy).
 *
 * Reference interfaces replaced (paths relative to /root/reference/src/lib/geogram):
 *   b200vor_set_points       NearestNeighborSearch::set_points      points/nn_search.h:98-123
 *                            Delaunay::set_vertices                 delaunay/delaunay.h:259
 *                            PeriodicDelaunay3d::set_weights        delaunay/periodic_delaunay_3d.h:151
 *   b200vor_knn              NearestNeighborSearch::get_nearest_neighbors (by index)
 *                                                                   points/nn_search.h:180-185
 *                            Delaunay::update_neighbors             delaunay/delaunay.cpp:259-274
 *   b200vor_knn_points       NearestNeighborSearch::get_nearest_neighbors (by point)
 *                                                                   points/nn_search.h:129-134
 *   b200vor_compute_cells    clip_by_cell_SR                        voronoi/generic_RVD.h:2278-2347
 *                            ConvexCell::clip_by_plane / compute_mg voronoi/convex_cell.cpp:467, :1267
 *                            PeriodicDelaunay3d::copy_Laguerre_cell_from_Delaunay
 *                                                                   delaunay/periodic_delaunay_3d.cpp:3634
 *   b200vor_get_cells        RestrictedVoronoiDiagram::compute_centroids (mg, m)
 *                                                                   voronoi/RVD.h:284-290
 *   b200vor_get_faces        ConvexCell::v_global_index / facet_area voronoi/convex_cell.h:463, convex_cell.cpp:1160
 *   b200vor_get_vertices     ConvexCell::triangle_point / save      voronoi/convex_cell.h:888, convex_cell.cpp:282-352
 *                            PeriodicDelaunay3d::save_cells         delaunay/periodic_delaunay_3d.cpp:4363-4392
 *                            Delaunay::cell_vertex (tetrahedra)     delaunay/delaunay.h:336-485
 *   b200vor_get_cvt          ComputeCVTFuncGradVolumetric (f, g)    voronoi/RVD.cpp:791-876
 *   b200vor_otm_newton       [external] OptimalTransportMap3d::optimize (exploragram; call sites
 *                            src/examples/exploragram/compute_OTM/main.cpp:284-297)
 *   b200vor_get_spatial_order, b200vor_get_hilbert_order
 *                            compute_Hilbert_order (role)           mesh/mesh_reorder.cpp:1260-1300
 *   b200vor_lloyd            CentroidalVoronoiTesselation::Lloyd_iterations
 *                                                                   voronoi/CVT.cpp:133-169
 */
#ifndef B200VOR_H
#define B200VOR_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B200VOR_VERSION 1

typedef struct b200vor_ctx b200vor_ctx;

/* ---- return codes ---- */
#define B200VOR_OK            0
#define B200VOR_ERR_CUDA      1   /* a CUDA call failed (no GPU, OOM, launch error) */
#define B200VOR_ERR_ARG       2   /* invalid argument / call order                  */
#define B200VOR_ERR_CAPACITY  3   /* a fixed capacity was exceeded                  */

/* ---- per-cell status bits (b200vor_get_cells status[]) ---- */
#define B200VOR_CELL_OK          0u
#define B200VOR_CELL_UNCERTIFIED 1u   /* security radius not reached within kmax / ring limit */
#define B200VOR_CELL_UNCERTAIN   2u   /* a sign test fell inside the fp64 error bound          */
#define B200VOR_CELL_EMPTY       4u   /* (Laguerre) cell clipped away entirely                 */
#define B200VOR_CELL_BOXFACE     8u   /* periodic: a helper-box plane still contributes        */
#define B200VOR_CELL_OVERFLOW    16u  /* more than the kernel's triangle / plane capacity      */
#define B200VOR_CELL_DUPLICATE   32u  /* exact duplicate of a lower-index seed: no cell
                                         (delaunay/delaunay_nn.cpp:122-135)                    */
#define B200VOR_CELL_DEGENERATE  64u  /* inconsistent combinatorics after uncertain sign tests
                                         (degenerate input); m, mg are best-effort             */

/* ids of the six planes of the clipping box in face lists: B200VOR_BOX_ID + {x-,x+,y-,y+,z-,z+} */
#define B200VOR_BOX_ID 0xFFFFFFF0u

typedef struct b200vor_cell_params {
    uint32_t k0;          /* neighbours fetched in the first pass (0 = default 48)            */
    uint32_t kmax;        /* cap on k for the re-launches of uncertified cells (0 = 384)      */
    uint32_t want_faces;  /* 1: keep per-cell face lists (+ areas) for b200vor_get_faces;
                             2: also the cell vertices for b200vor_get_vertices               */
    uint32_t mode;        /* 0 = grid walk, no kNN table: the faster of the two first-pass kernels (team
                                 kernel / thread-per-cell kernel; the latter needs an unweighted input on a
                                 single-level grid), the cells it flags UNCERTIFIED / OVERFLOW are re-done
                                 by the team kernel, its large-cell instance, then the kNN-table pipeline
                             1 = team-kernel grid walk only (flags are left on the cells)
                             2 = kNN table -> clipper (re-launch failing cells with larger k)
                             3 = team kernel first (then as 0)
                             4 = thread-per-cell kernel first (then as 0)
                             In modes 0, 3 and 4 the cells a kernel flags UNCERTAIN (degenerate input)
                             are re-done with exact sign decisions when the input allows it            */
    uint32_t want_energy; /* 1: also integrate |x - p_i|^2 over every cell (CVT energy) for
                             b200vor_get_cvt                                               */
} b200vor_cell_params;

typedef struct b200vor_stats {
    /* device milliseconds of the last b200vor_set_points / b200vor_compute_cells */
    float ms_grid, ms_knn, ms_clip, ms_total;
    uint32_t passes;                 /* kNN+clip passes (1 + re-launches)            */
    uint32_t n_pass[8];              /* cells processed in each pass                 */
    uint32_t k_pass[8];              /* k of each pass                               */
    uint32_t n_uncertified, n_uncertain, n_empty, n_boxface, n_overflow, n_duplicate,
             n_degenerate;
    /* event counts of the SURVEY 8(d) flop model, summed over cells (last pass of each cell) */
    uint64_t sum_planes;             /* P: planes offered before the security-radius stop */
    uint64_t sum_tests;              /* C: triangle sign tests                            */
    uint64_t sum_created;            /* V: triangles created                              */
    uint64_t sum_tets;               /* I: integration terms                              */
    uint64_t gpu_launches;           /* kernels launched by this context so far           */
    uint32_t grid_dim[3];
    uint32_t tpc_cells;              /* cells the thread-per-cell kernel was launched on (last call) */
    uint64_t sum_scans;              /* fused walk: distance-shell scans of the grid, summed over cells */
    uint32_t exact_cells;            /* cells re-done with exact (expansion) sign decisions (last call) */
    uint32_t reserved1;
} b200vor_stats;

/* ---- life cycle ---- */
int  b200vor_create(b200vor_ctx** ctx, int device);
void b200vor_destroy(b200vor_ctx* ctx);
const char* b200vor_last_error(void);
int  b200vor_version(void);

/* ---- domain ---- */
/* cells are clipped by box = {xmin,ymin,zmin,xmax,ymax,zmax} (default: unit cube) */
int b200vor_set_domain_box(b200vor_ctx* ctx, const double* box);
/* convex polyhedral domain: every cell is also clipped by the n half-spaces a x + b y + c z + d >= 0
 * (planes: 4 doubles each, host memory; n <= 32; n = 0 removes them), before any neighbour -- what
 * VBW::ConvexCell::clip_by_plane does for one convex region of a restricted Voronoi diagram
 * (voronoi/convex_cell.cpp:467-516; first slice of the mesh-restricted cells of voronoi/generic_RVD.h:1464-1597).
 * Box domains only; a face on plane k is reported as neighbour id B200VOR_DOMAIN_ID + k.  Cells are computed
 * by the team kernels only (a cell beyond their capacities keeps its flag). */
#define B200VOR_DOMAIN_ID 0xFFFFFF00u
int b200vor_set_domain_planes(b200vor_ctx* ctx, uint32_t n, const double* planes);
/* domain = union of convex PIECES with disjoint interiors, e.g. the tetrahedra of a domain mesh (what
 * the reference's restricted Voronoi diagram integrates over, voronoi/generic_RVD.h:1464-1597 and
 * RVD_callback.cpp:794-858: every cell is clipped by every tetrahedron it meets and the integrals summed).
 * Piece p has nplanes[p] (1..32) half-spaces, stored one after the other in planes.  b200vor_compute_cells /
 * b200vor_lloyd / b200vor_get_cvt then return, per seed, the integrals over cell ∩ domain: every cell is
 * integrated over each piece in turn (cost: one pass per piece; a cell that does not meet the piece is
 * clipped away after a few planes).  No face lists; status EMPTY = the cell meets no piece.
 * b200vor_set_domain_tets builds the pieces from a tetrahedral mesh (nv vertices, nt x 4 indices). */
int b200vor_set_domain_pieces(b200vor_ctx* ctx, uint32_t npieces, const uint32_t* nplanes, const double* planes);
int b200vor_set_domain_tets(b200vor_ctx* ctx, uint32_t nv, const double* vertices, uint32_t nt, const uint32_t* tets);
/* periodic cube [0,period)^3; neighbour ids become real + n*instance (periodic.h:95-99) */
int b200vor_set_domain_periodic(b200vor_ctx* ctx, double period);

/* ---- points (copied; builds the uniform grid) ---- */
/* xyz: n points, `stride` doubles apart (>= 3).  w: NULL or n weights (Laguerre). */
int b200vor_set_points(b200vor_ctx* ctx, uint32_t n, const double* xyz, uint32_t stride,
                       const double* w, int on_device);

/* ---- kNN ---- */
/* rows of k for every point (query by index), ascending squared distance, the point
 * itself at rank 0, ties by ascending index.  d2 may be NULL. */
int b200vor_knn(b200vor_ctx* ctx, uint32_t k, uint32_t* nbr, double* d2, int on_device);
/* same for nq arbitrary query points (3 doubles each, contiguous) */
int b200vor_knn_points(b200vor_ctx* ctx, uint32_t nq, const double* q, uint32_t k,
                       uint32_t* nbr, double* d2, int on_device);

/* ---- cells ---- */
int b200vor_compute_cells(b200vor_ctx* ctx, const b200vor_cell_params* params);
/* m[n], mg[3n] (mass and mass*centroid, as ConvexCell::compute_mg), status[n]; any may be NULL */
int b200vor_get_cells(b200vor_ctx* ctx, double* m, double* mg, uint8_t* status, int on_device);
/* nfaces[n], face_nbr[n*fcap] (ascending ids), face_area[n*fcap] (may be NULL) */
int b200vor_get_faces(b200vor_ctx* ctx, uint32_t* nfaces, uint32_t* face_nbr,
                      double* face_area, uint32_t fcap, int on_device);

/* cell geometry of the last b200vor_compute_cells(want_faces=2), what ConvexCell::save
 * (voronoi/convex_cell.cpp:282-352), append_to_mesh (:371-444) and the Delaunay tetrahedra
 * (one per cell vertex: the seed and the three neighbours whose faces meet there) are built from:
 *   nverts[n]            vertices of the cell (0: empty/duplicate; 0xFFFFFFFF: not available, the
 *                        cell was finished by the kNN-table pipeline; > vcap: list truncated)
 *   vert_xyz[n*vcap*3]   ConvexCell::triangle_point, absolute coordinates (periodic: around the seed)
 *   vert_faces[n*vcap]   f0 | f1 << 8 | f2 << 16: positions in the cell's face list (b200vor_get_faces)
 *                        of the three faces meeting at the vertex, same orientation for all vertices
 * vcap must be 64. */
int b200vor_get_vertices(b200vor_ctx* ctx, uint32_t* nverts, double* vert_xyz, uint32_t* vert_faces,
                         uint32_t vcap, int on_device);

/* CVT objective of the last b200vor_compute_cells(want_energy=1)
 * (ComputeCVTFuncGradVolumetric, voronoi/RVD.cpp:791-876, summed per seed):
 *   f_cell[n]  f_i = integral over cell i of |x - p_i|^2   (sum of the functor's fi)
 *   g[3n]      g_i = 2 (m_i p_i - mg_i)                    (sum of the functor's gi)
 *   f_total    sum of f_i in a fixed order (deterministic)
 * any output may be NULL; f_total is always a host pointer */
int b200vor_get_cvt(b200vor_ctx* ctx, double* f_cell, double* g, double* f_total, int on_device);

/* ---- semi-discrete optimal transport: damped Newton on the Laguerre weights ----
 * Finds w such that vol(Lag_i(w)) = nu_i for the points of b200vor_set_points in the current
 * domain.  Stands in for OptimalTransportMap3d::optimize (BrunoLevy/exploragram, not part of
 * the reference tree; call sites src/examples/exploragram/compute_OTM/main.cpp:284-297) and
 * restates the published algorithm: gradient vol_i - nu_i and Hessian A_ij / (2 |p_i - p_j|)
 * (Levy 2015), damping of Kitagawa-Merigot-Thibert 2019 (Algorithm 1), linear systems by
 * Jacobi-preconditioned conjugate gradient on the device. */
typedef struct b200vor_otm_params {
    uint32_t max_newton;  /* 0 = 50                                                        */
    uint32_t max_cg;      /* CG iterations per Newton step, 0 = 1000                       */
    double cg_tol;        /* relative residual of the linear solve; 0 = inexact Newton: a tenth of
                             the current relative volume error, within [epsilon / 100, 1e-2]  */
    double epsilon;       /* stop when max_i |vol_i - nu_i| <= epsilon * nu_i, 0 = 0.01
                             (OptimalTransportMap::set_epsilon)                            */
    double total_mass;    /* nu == NULL: nu_i = total_mass / n, 0 = measure of the domain  */
} b200vor_otm_params;

typedef struct b200vor_otm_result {
    uint32_t newton_iters, cg_iters, cell_evals;
    int32_t converged;
    double initial_error, final_error;   /* || vol - nu ||_2 before / after                */
    double max_rel_error;                /* max_i |vol_i - nu_i| / nu_i                    */
    double min_volume, sum_volume, last_alpha;
} b200vor_otm_result;

/* nu: n target measures or NULL (uniform); w: n weights, start values in, solution out.  The cells
 * (b200vor_get_cells / b200vor_get_faces) are those of the returned weights. */
int b200vor_otm_newton(b200vor_ctx* ctx, const double* nu, double* w, const b200vor_otm_params* params,
                       b200vor_otm_result* result, int on_device);

/* ---- Lloyd relaxation: iters x (grid, cells, p <- mg/m, periodic wrap) ---- */
/* xyz_out (n x 3, may be NULL) receives the relaxed points */
int b200vor_lloyd(b200vor_ctx* ctx, uint32_t iters, const b200vor_cell_params* params,
                  double* xyz_out, int on_device);

/* ---- execution control ---- */
/* run on the caller's CUDA stream (a cudaStream_t passed as void*; e.g. torch's current
 * stream) instead of the context's own; the caller keeps ownership of the stream */
int b200vor_set_stream(b200vor_ctx* ctx, void* cuda_stream);
/* multi-GPU sharding: compute (and Lloyd-update) only the cells of the seeds whose index
 * lies in [first, last); every rank still holds all seeds (default: all) */
int b200vor_set_query_range(b200vor_ctx* ctx, uint32_t first, uint32_t last);

/* ---- spatial order ---- */
/* order[n]: the seeds in the order of the counting-sort grid (bin by bin, x fastest; inside a bin by
 * index) -- a spatial sort of the point set in the role of the reference's Hilbert / BRIO
 * pre-ordering (mesh/mesh_reorder.cpp:1260-1300, compute_Hilbert_order): consecutive entries are
 * neighbours in space.  Available after b200vor_set_points. */
int b200vor_get_spatial_order(b200vor_ctx* ctx, uint32_t* order, int on_device);
/* order[n]: the same bins visited ALONG THE 3-D HILBERT CURVE of the grid (inside a bin by index): no jump
 * at the end of a row of bins -- consecutive entries are in adjacent or next-to-adjacent bins (the g bins
 * of an axis are spread over the 2^b cells of the curve's cube).  compute_Hilbert_order (mesh/mesh_reorder.cpp:1260-1300) sorts by recursive median
 * splits; this is the curve of the uniform grid, same role (locality of a point order), not the same
 * permutation. */
int b200vor_get_hilbert_order(b200vor_ctx* ctx, uint32_t* order, int on_device);

/* ---- introspection ---- */
int b200vor_get_stats(b200vor_ctx* ctx, b200vor_stats* stats);

/* measured pipe peaks of the context's GPU (FMA chains timed with CUDA events): the denominators
 * of the flop / issue rooflines of the cell kernels (no reference counterpart; the reference's
 * Stopwatch / PredicateStats play this role on the CPU, basic/stopwatch.h:69, predicates.cpp:353) */
typedef struct b200vor_peaks {
    double fp64_tflops;     /* DFMA, 2 flop each                                   */
    double fp32_tflops;     /* FFMA, 2 flop each                                   */
    double issue_ginst_s;   /* 4 schedulers x SMs x max SM clock, warp instructions */
    double sm_clock_mhz;
    uint32_t sm_count, reserved;
} b200vor_peaks;
int b200vor_microbench(b200vor_ctx* ctx, b200vor_peaks* peaks);

/* self-test of the device's exact sign predicate (what decides the ties of degenerate inputs, the
 * role of PCK::det_4d's exact branch, numerics/predicates.cpp:2073-2110): for each of n quadruples of
 * planes (16 doubles: P0, P1, P2, P, host memory) the sign of P at the point P0 ^ P1 ^ P2
 * (-1, 0, +1; 2 = the three planes do not meet in one point) */
int b200vor_selftest_exact_sign(b200vor_ctx* ctx, uint32_t n, const double* planes, int* signs);
/* ---- multi-GPU: x-slabs with halos of ghost seeds, one process per GPU (the reference
 *      is single-address-space, its hint is delaunay/periodic_delaunay_3d.cpp:3432-3436) ----
 * Rank r owns the rows [first_global, first_global + n_own) of a global array of n_total seeds (hand
 * every rank an x-slab: sort by x, split evenly).  The domain is the one set with
 * b200vor_set_domain_box / _periodic before the first b200vor_slab_set_points.  A cell pass exchanges
 * only the seeds within a halo of the neighbouring slabs (ncclSend / ncclRecv), bins the own + ghost
 * seeds on a slab-local grid and computes the own cells; a cell whose security radius leaves the halo
 * is re-queued with a doubled halo (all ranks).  Results: b200vor_get_cells / _get_faces return the
 * n_own rows of this rank; neighbour ids are GLOBAL (index + n_total * instance on a periodic domain,
 * delaunay/periodic.h:95-113).  world == 1 needs no communicator (ghosts = periodic images only). */
#define B200VOR_UNIQUE_ID_BYTES 128
typedef struct b200vor_slab_stats {
    uint32_t n_own, n_ghost_x, n_ghost_yz, n_local;   /* seeds on the slab-local grid of the last pass   */
    uint32_t retries;      /* halo doublings of the last pass                                            */
    uint32_t n_flagged;    /* cells (all ranks) that still left the halo after the last pass             */
    double halo;           /* halo width of the last pass                                                */
    double slab_lo, slab_hi;
    float ms_exchange, ms_pack, ms_grid, ms_clip;   /* device ms, summed over the passes of the last call */
    uint64_t bytes_sent, bytes_received;            /* ghost traffic of this rank over the last call      */
    uint64_t reallocs;     /* device buffers (re-)allocated during the last call: 0 once the sizes settled */
} b200vor_slab_stats;
/* rank 0 creates the id (an ncclUniqueId), the caller hands the 128 bytes to every rank (MPI, torch, files) */
int b200vor_comm_unique_id(void* id);
/* collective: binds the context to rank `rank` of `world` (ncclCommInitRank; libnccl.so.2 is loaded on demand) */
int b200vor_comm_init(b200vor_ctx* ctx, const void* id, int rank, int world);
/* initial halo width in bins of the global seed density (default 3) */
int b200vor_slab_set_halo_bins(b200vor_ctx* ctx, double bins);
int b200vor_slab_set_points(b200vor_ctx* ctx, uint32_t n_own, uint64_t first_global, uint64_t n_total,
                            const double* xyz, const double* w, int on_device);
/* collective: one cell pass over the own seeds (params as b200vor_compute_cells; team kernel path) */
int b200vor_slab_compute_cells(b200vor_ctx* ctx, const b200vor_cell_params* params);
/* collective: `iters` Lloyd iterations (CVT.cpp:133-169); xyz_out_own: NULL or 3 n_own doubles */
int b200vor_slab_lloyd(b200vor_ctx* ctx, uint32_t iters, const b200vor_cell_params* params,
                       double* xyz_out_own, int on_device);
int b200vor_slab_get_stats(b200vor_ctx* ctx, b200vor_slab_stats* stats);

/* tuning knob: target points per grid cell (default 8.0); takes effect at the next set_points */
int b200vor_set_points_per_cell(b200vor_ctx* ctx, double ppc);

#ifdef __cplusplus
}
#endif
#endif
