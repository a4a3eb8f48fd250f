// common.cuh -- shared declarations of the B200 Voronoi engine (sm_100a only).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>
#include <vector>

#include "../../include/b200vor.h"
#include "grid_types.h"

namespace b200vor {

void set_error(const char* fmt, ...);

#define B200_CUDA_CHECK(expr)                                                       \
    do {                                                                            \
        cudaError_t err__ = (expr);                                                 \
        if(err__ != cudaSuccess) {                                                  \
            b200vor::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr,        \
                               cudaGetErrorString(err__));                          \
            return B200VOR_ERR_CUDA;                                                \
        }                                                                           \
    } while(0)

__device__ inline void cell_coords(const GridDesc& G, double x, double y, double z,
                                   int& ix, int& iy, int& iz) {
    ix = clampi(int(floor((x - G.lo[0]) * G.inv_w[0])), 0, G.g[0] - 1);
    iy = clampi(int(floor((y - G.lo[1]) * G.inv_w[1])), 0, G.g[1] - 1);
    iz = clampi(int(floor((z - G.lo[2]) * G.inv_w[2])), 0, G.g[2] - 1);
}

__device__ inline Seed load_seed(const Seed* p) {
    const double2* q = reinterpret_cast<const double2*>(p);
    double2 a = __ldg(q), b = __ldg(q + 1);
    Seed s;
    s.x = a.x; s.y = a.y; s.z = b.x; s.w = b.y;
    return s;
}

// ---------------------------------------------------------------------------------
// Context
// ---------------------------------------------------------------------------------
// device (re-)allocations of this process so far: a cudaFree + cudaMalloc in the middle of a run stalls
// the stream for milliseconds, so the slab statistics report how many a call made
inline uint64_t& devbuf_alloc_count() {
    static uint64_t count = 0;
    return count;
}

template <class T> struct DevBuf {
    T* p = nullptr;
    size_t cap = 0;   // elements
    int reserve(size_t n) {
        if(n <= cap) {
            return B200VOR_OK;
        }
        if(p) {
            cudaFree(p);
            p = nullptr;
            cap = 0;
        }
        size_t want = n + n / 8 + 64;   // (slack: sizes that drift upwards from call to call do not re-allocate each time)
        B200_CUDA_CHECK(cudaMalloc(&p, want * sizeof(T)));
        cap = want;
        devbuf_alloc_count()++;
        return B200VOR_OK;
    }
    // for sizes that keep drifting upwards (ghost counts of a relaxing seed set): nothing happens while
    // `need` fits; a buffer that has to grow is allocated with half as much again
    int reserve_growing(size_t need) {
        return need <= cap ? B200VOR_OK : reserve(need + need / 2);
    }
    void release() {
        if(p) {
            cudaFree(p);
        }
        p = nullptr;
        cap = 0;
    }
};

struct Ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev[8] = {};
    std::vector<cudaEvent_t> event_pool;   // timing events, re-used from call to call (capi.cu: Timer)

    int periodic = 0;
    double period = 1.0;
    double box[6] = {0, 0, 0, 1, 1, 1};
    double ppc = 8.0;
    int fused_sweep_mode = 0;      // (pipeline -> fused.cu) large-cell instance: 0 sweep in bands x4, 1 bands x1.25, 2 no sweep
    bool fused_bail_far = false;   // (pipeline -> fused.cu) far-reaching cells go straight to the large-cell instance
    // domain = union of convex pieces (tetrahedra of a mesh ...): planes of piece p = piece_planes[piece_off[p] .. piece_off[p+1])
    std::vector<uint32_t> piece_off;
    std::vector<double> piece_planes;
    DevBuf<double> acc_m, acc_mg, acc_e;   // sums over the pieces
    DevBuf<uint8_t> acc_st;
    int ndom = 0;              // planes of the convex domain (b200vor_set_domain_planes), clipped before any neighbour
    DevBuf<double> dom_dev;
    int open = 0;              // slab-local context: open faces of the local grid (GridDesc::open)
    struct Slab* slab = nullptr;   // multi-GPU slab state (slab.cu), or NULL

    uint32_t n = 0;
    uint32_t q0 = 0, q1 = 0xFFFFFFFFu;   // user-index range of the cells to compute
    bool own_stream = true;
    bool grid_dirty = true;
    bool has_w = false;
    double wmax = 0.0;
    DevBuf<double> xyz;        // user order, packed 3n
    DevBuf<double> wgt;        // user order
    GridDesc grid;
    DevBuf<uint32_t> cellid, rank, cell_start, perm, scan_tmp;
    DevBuf<uint32_t> sub, vstart, vcount;   // two-level grid only
    DevBuf<uint32_t> sbcount;               // two-level grid only: seeds per 8 x 8 x 8 block of bins
    DevBuf<double> wloc, wloc_tmp;          // Laguerre only: dilated per-bin weight maxima (near | far)
    DevBuf<Seed> sorted;

    // kNN table of the cell pipeline (rows in pass order, ids in grid order + n*instance)
    DevBuf<uint32_t> knn;
    DevBuf<uint8_t> knn_flag;
    DevBuf<uint32_t> qlist, qlist2, qlist3, qrows, counter;

    // cell outputs, user order
    DevBuf<double> m, mg;
    DevBuf<double> energy;     // per-cell integral of |x - p_i|^2 (only with want_energy)
    bool energy_valid = false;
    DevBuf<uint8_t> status;
    DevBuf<uint32_t> nfaces, face_nbr;
    DevBuf<double> face_area;
    uint32_t fcap = 0;
    // per-cell vertex lists (want_faces == 2): nverts[n], vert_xyz[n*vcap*3], vert_faces[n*vcap]
    DevBuf<uint32_t> nverts, vert_faces;
    DevBuf<double> vert_xyz;
    uint32_t vcap = 64;
    bool verts_valid = false;
    bool cells_valid = false, faces_valid = false;
    DevBuf<uint32_t> evcount;        // per-cell event counts (2 words)
    DevBuf<unsigned long long> sums; // reductions

    // scratch for user-facing kNN
    DevBuf<uint32_t> out_nbr;
    DevBuf<double> out_d2;
    DevBuf<double> qpts;

    b200vor_stats stats = {};
};

// grid.cu
int build_grid(Ctx* c);
int hilbert_order(Ctx* c, uint32_t* order_dev);   // 1 = grid too elongated (caller falls back to the raster order)
// knn.cu
struct KnnJob {
    uint32_t nq;
    const uint32_t* qlist;   // grid-order indices of the queries, or NULL (= 0..nq-1)
    const double* qpts;      // arbitrary query points (3 doubles), or NULL (= by index)
    uint32_t k;
    uint32_t* out_nbr;
    double* out_d2;          // may be NULL
    uint8_t* out_flag;       // may be NULL; 1 = row not guaranteed complete
    int user_ids;            // 1: ids and row order in user space; 0: grid order (+n*instance)
};
int run_knn(Ctx* c, const KnnJob& job);
// knn_team.cu
bool knn_team_applicable(const Ctx* c, const KnnJob& job);
int run_knn_team(Ctx* c, const KnnJob& job, uint32_t* nfail);
// cell.cu
struct CellJob {
    uint32_t nq;
    const uint32_t* qlist;   // grid-order indices, or NULL
    const uint32_t* rowmap;  // row of entry t in the kNN table, or NULL (= t)
    int slots;               // triangle slots per lane: 2 (64 triangles) or 4 (128)
    uint32_t k;              // row length of the kNN table
    const uint32_t* knn;
    const uint8_t* knn_flag;
    int want_faces;
    int collect_events;
    int mode;
    int all_offered;         // 1: every row lists all other seeds (k >= n-1)
    int want_energy;
};
int run_cells(Ctx* c, const CellJob& job);
// fused.cu
int run_cells_fused(Ctx* c, uint32_t nq, const uint32_t* qlist, int want_faces,
                    int collect_events, float target, int want_energy, int big);
int compact_failed(Ctx* c, const uint32_t* qlist_in, uint32_t nq_in, uint32_t* qlist_out,
                   uint32_t* rows_out, uint32_t* count_out_host, uint8_t mask);
int lloyd_update(Ctx* c);
int select_range(Ctx* c, uint32_t* qout, uint32_t* count_out_host);
int reduce_events(Ctx* c);
// tpc.cu
bool tpc_applicable(const Ctx* c);
int run_cells_tpc(Ctx* c, uint32_t nq, const uint32_t* qlist, int want_faces, int collect_events, int big);
int exact_sign_selftest(Ctx* c, uint32_t n, const double* planes_host, int* out_host);
// capi.cu (used by slab.cu)
int ctx_build_grid_timed(Ctx* c);
int ctx_cells(Ctx* c, const b200vor_cell_params* params, bool collect_events);
// slab.cu
void slab_release(Ctx* c);
uint32_t out_rows(const Ctx* c);   // rows the result getters return: n, or the own rows of a slab
bool slab_active(const Ctx* c);     // b200vor_slab_set_points was the last set_points
void slab_deactivate(Ctx* c);      // back to a plain single-GPU context (b200vor_set_points)
// microbench.cu
int microbench(Ctx* c, b200vor_peaks* out);

}  // namespace b200vor

// the opaque handle of the C ABI
struct b200vor_ctx {
    b200vor::Ctx c;
};
