// fused.cu -- the fused hot path: uniform-grid neighbour walk + half-space clipping with
// the security-radius rule + fp64 volume / centroid / facet-area integration, ONE kernel,
// no neighbour table in HBM, no host round trip.
//
// Replaces, for every seed:
//   KdTree::get_nearest_neighbors (as the neighbour provider)  points/kd_tree.cpp:156-387
//   Delaunay_NearestNeighbors::enlarge_neighborhood            delaunay/delaunay_nn.cpp:57-71
//   clip_by_cell_SR (stop rule d2 > 4.1 R2, neighbourhood growth) voronoi/generic_RVD.h:2278-2347
//   VBW::ConvexCell::init_with_box / clip_by_plane             voronoi/convex_cell.cpp:175-215, :467-516
//     triangle_is_in_conflict / triangulate_conflict_zone      voronoi/convex_cell.cpp:870-969, :761-865
//   compute_triangle_point / squared_radius                    voronoi/convex_cell.cpp:971-1082, :1312-1330
//   compute_mg / facet_area                                    voronoi/convex_cell.cpp:1267-1307, :1160-1191
//   power planes                                               delaunay/periodic_delaunay_3d.cpp:3719-3726
//
// Design (B200-first):
//  * A TEAM of G lanes (G = 16: two cells per warp) owns one cell.  The two teams of a warp run
//    in LOCKSTEP: every collective is a full-mask *_sync intrinsic of which a team reads its own
//    G-lane segment, team-level decisions are predicates, loops run while any team needs them.
//  * The cell lives in the team's slice of shared memory in DUAL form, densely packed:
//    slots [0,T) hold the cell vertices as (X,Y,Z,W) homogeneous fp64 coordinates (the
//    cofactors of the reference's 4x4 determinant), a forward error coefficient, |v|^2 and
//    the three plane ids whose intersection the vertex is.  No adjacency is stored: the
//    border of a conflict zone is found with a 64x64 bit matrix of directed dual edges
//    (an edge is interior iff its twin was marked too), and adjacency is rebuilt once,
//    for the integration, as an FxF byte table keyed by the two face ids of an edge.
//  * Neighbour candidates come straight from the counting-sort grid: the team scans the
//    (2r+1)^3 block of bins around the seed for the candidates of a distance shell
//    [lo,hi), keeps (d2 key, index, periodic image) in a 64-entry pool, sorts the pool
//    with an in-register bitonic network, and clips in ascending distance.  The first
//    shell is sized from the local density; later shells end at the current
//    security-radius bound, so a cell costs one scan (relaxed seeds) or two.  The cell is
//    certified when every candidate with d2 <= bound(R2) was offered and the scanned block
//    covers that radius; otherwise the block grows by one ring.  On a two-level grid
//    (clustered seeds, grid.cu) the block is measured in the fine units of the seed's own bin
//    and may grow until it covers the domain (template parameter TWO).
//  * Exactness: a sign test whose |value| is inside its forward error bound never cuts;
//    the cell is flagged B200VOR_CELL_UNCERTAIN (same rule as cell.cu).  Cells that exceed
//    a fixed capacity (slots, planes, ring limit) are flagged and re-done by the large-cell
//    instance below (one warp per cell), then by the kNN-table pipeline (capi.cu), never
//    silently accepted.
//  * The kernel body lives in fused_kernel.inc and is instantiated twice (capacities) x 16
//    (face lists, event counters, two-level walk, weights as template parameters).
#include "common.cuh"

#include <stdlib.h>

namespace b200vor {

#ifndef B200VOR_FILT_AFTER
#define B200VOR_FILT_AFTER 1   // two-level walk: halvings of an overflowing shell before the filtered scan (0 / 1 / 2 / 4 measured: 614 / 560 / 575 / 684 ms at 10 M clustered seeds)
#endif
#ifndef B200VOR_PYR_MIN_BINS
#define B200VOR_PYR_MIN_BINS 512   // two-level walk: boxes of more level-1 bins than this are walked super-bin by super-bin (0 / 64 / 512 measured: 701 / 644 / 639 ms at 10 M clustered seeds)
#endif
#ifndef B200VOR_FAR_RING
#define B200VOR_FAR_RING 4   // two-level walk: fine rings offered before a cell counts as far-reaching (sweep)
#endif
#ifndef B200VOR_FUSED_F32
#define B200VOR_FUSED_F32 0   // fp32 filter of the sign tests (A/B: DESIGN.md section 9)
#endif

// ---- standard cells: 16 lanes per cell, two cells per warp ----
#ifndef B200VOR_FUSED_MINBLOCKS
#define B200VOR_FUSED_MINBLOCKS 2
#endif
#ifndef B200VOR_FUSED_G
#define B200VOR_FUSED_G 16     // lanes per cell
#endif
#ifndef B200VOR_FUSED_WPB
#define B200VOR_FUSED_WPB 8    // warps per block (8 x 2 cells share their neighbourhood in L1)
#endif
#ifndef B200VOR_FUSED_WPB_TWO
#define B200VOR_FUSED_WPB_TWO 8   // warps per block on a two-level grid (2 and 8 measured: 915 vs 843 ms at 10 M clustered seeds)
#endif
#ifndef B200VOR_FUSED_CHUNK_STD
#define B200VOR_FUSED_CHUNK_STD 0   // warp-tasks per block drawn from a block-local ticket (0 = one task per warp)
#endif
#define B200VOR_FUSED_CHUNK B200VOR_FUSED_CHUNK_STD
#define B200VOR_FUSED_VARIANT fused_std
#define B200VOR_FUSED_NT 48
#define B200VOR_FUSED_PC 40
#define B200VOR_FUSED_FC 40
#include "fused_kernel.inc"
#undef B200VOR_FUSED_VARIANT
#undef B200VOR_FUSED_NT
#undef B200VOR_FUSED_PC
#undef B200VOR_FUSED_FC
#undef B200VOR_FUSED_G
#undef B200VOR_FUSED_WPB
#undef B200VOR_FUSED_WPB_TWO
#undef B200VOR_FUSED_MINBLOCKS
#undef B200VOR_FUSED_CHUNK

// ---- large cells (what the standard instance flags OVERFLOW / UNCERTIFIED: transiently large
//      cells of clustered seeds, faces with more than 16 vertices): one warp per cell ----
#define B200VOR_FUSED_VARIANT fused_big
#define B200VOR_FUSED_CHUNK 0
#define B200VOR_FUSED_G 32
#define B200VOR_FUSED_WPB 4
#define B200VOR_FUSED_WPB_TWO 4
#define B200VOR_FUSED_MINBLOCKS 2
#define B200VOR_FUSED_NT 128
#define B200VOR_FUSED_PC 64
#define B200VOR_FUSED_FC 64
#include "fused_kernel.inc"

int run_cells_fused(Ctx* c, uint32_t nq, const uint32_t* qlist, int want_faces, int collect_events,
                    float target, int want_energy, int big) {
    return big ? fused_big::run(c, nq, qlist, want_faces, collect_events, target, want_energy)
               : fused_std::run(c, nq, qlist, want_faces, collect_events, target, want_energy);
}

}  // namespace b200vor
