/*
 * oracle/voro_oracle.h -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * C ABI of this repository's own CPU restatement of the reference's meshless
 * Voronoi/Laguerre cell path (kd-tree kNN -> ConvexCell clipping with the
 * security radius -> volume / centroid).  See voro_oracle.cpp for the
 * reference file:line each function follows.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may use
 * it, and only as the checker / CPU baseline.  The product path
 * (geogram_b200/csrc) never includes or links this.
 */
#ifndef VORO_ORACLE_H
#define VORO_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* id given to the six planes of the bounding box: VO_BOX_ID + {x-,x+,y-,y+,z-,z+} */
#define VO_BOX_ID 0xFFFFFFF0u
#define VO_UNCERTIFIED 0xFFFFFFFFu

/* cell flags written to flags[i] */
#define VO_FLAG_UNCERTIFIED 1u   /* security radius not reached at k = kcap      */
#define VO_FLAG_TIE         2u   /* a sign test fell inside the fp64 error bound */
#define VO_FLAG_EMPTY       4u   /* (Laguerre) cell clipped away entirely        */
#define VO_FLAG_BOXFACE     8u   /* periodic: a helper-box plane still contributes */

void vo_set_num_threads(int nthreads);
int  vo_num_threads(void);

/* k nearest neighbours of every point (query by index): rows of k, ascending
 * squared distance, the point itself at rank 0.  */
int vo_knn(uint32_t n, const double* xyz, uint32_t k, uint32_t* nbr, double* d2);

/* k nearest neighbours of arbitrary query points. */
int vo_knn_points(uint32_t n, const double* xyz, uint32_t nq, const double* q,
                  uint32_t k, uint32_t* nbr, double* d2);

/*
 * Cells by kNN + clipping + security radius with k-expansion.
 *   w        : NULL (Voronoi) or n weights (Laguerre / power cells)
 *   periodic : 0 -> cells clipped by box[0..5] = {xmin,ymin,zmin,xmax,ymax,zmax};
 *              1 -> domain [0,period)^3, neighbours reported as real + n*instance
 *   k0, kcap : initial / maximal neighbour count (growth k += k/8, or k+1 if k<=8)
 * Outputs (caller allocated, any may be NULL except m, mg):
 *   m[n], mg[3n], nfaces[n], face_nbr[n*fcap] (ascending ids), face_area[n*fcap],
 *   kstop[n] (rank at which the security radius was reached, VO_UNCERTIFIED if not),
 *   flags[n], stats[6] = {sum P, sum C, sum V, sum I, #uncertified, #kNN rounds}.
 */
int vo_cells(uint32_t n, const double* xyz, const double* w,
             int periodic, double period, const double* box,
             uint32_t k0, uint32_t kcap,
             double* m, double* mg,
             uint32_t* nfaces, uint32_t* face_nbr, double* face_area, uint32_t fcap,
             uint32_t* kstop, uint32_t* flags, double* stats);

/* Optional extra output of the next vo_cells() calls: energy[n] receives the CVT energy of
 * every cell, the integral of |x - seed_i|^2 over the cell (the per-cell share of f in
 * ComputeCVTFuncGradVolumetric, voronoi/RVD.cpp:791-876).  NULL switches it off. */
void vo_set_energy_output(double* energy);
/* convex domain: n planes (4 doubles each, a x + b y + c z + d >= 0 inside) clipped before any neighbour on box domains; n = 0 removes them */
void vo_set_domain_planes(uint32_t n, const double* planes);

/* One Lloyd update: xyz_out = mg/m where m > 1e-30, else unchanged; periodic wrap. */
int vo_lloyd_update(uint32_t n, const double* xyz, const double* m, const double* mg,
                    int periodic, double period, double* xyz_out);

#ifdef __cplusplus
}
#endif
#endif
