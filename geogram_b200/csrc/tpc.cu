// tpc.cu -- cells_tpc_kernel: the first-pass cell kernel, ONE THREAD PER CELL (32 independent
// cells per warp).  The per-cell algorithm is tpc_cell.cuh (neighbour walk on the counting-sort
// grid in sorted distance shells, fp32-filtered half-space clipping with the security radius,
// fp64 orthoscheme integration); this file only lays the cells out in shared memory and launches.
//
// Shared-memory layout (BT threads per block, 992 bytes per thread): arrays of COLUMNS, element k
// of thread t at index k * BT + t, so that the 32 lanes of a warp always hit 32 different banks
// (16-byte vertex records: four conflict-free quarter-warp phases) whatever slot each lane is at.
//   V    [NT][BT]       16 B   x, y, z (fp32) | plane triple + error exponent
//   gid  [PC][BT]        4 B   neighbour id of the plane slot
//   pool [POOLCAP][BT]   4 B   candidates of the current distance shell
//   L    [LCAP][BT]      4 B   sort keys of the shell / plane triples of the conflict vertices of a cut
// Cells the kernel cannot finish inside these capacities (or within two rings of bins) are flagged
// OVERFLOW / UNCERTIFIED and re-done by the team kernel (fused.cu); nothing is silently accepted.
#include "common.cuh"
#include "tpc_cell.cuh"

#include <stdlib.h>

namespace b200vor {
namespace {

template <int BT, bool EV, bool FACES, int NT, int PC, int LC>
__global__ void __launch_bounds__(BT) cells_tpc_kernel(tpc::Params P) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    tpc::F4* V = reinterpret_cast<tpc::F4*>(smem_raw);
    uint32_t* gid = reinterpret_cast<uint32_t*>(V + NT * BT);
    uint32_t* pool = gid + PC * BT;
    uint32_t* L = pool + tpc::POOLCAP * BT;
    const uint32_t c = blockIdx.x * BT + threadIdx.x;   // lanes past the end idle through the warp-uniform loops
    tpc::Mem<BT> M;
    M.V = V + threadIdx.x;
    M.gid = gid + threadIdx.x;
    M.pool = pool + threadIdx.x;
    M.L = L + threadIdx.x;
    tpc::compute_cell<BT, EV, FACES, NT, PC, LC>(P, c, M);
}

template <int BT, int NT, int PC, int LC> int launch(Ctx* c, const tpc::Params& P, int collect_events, int faces) {
    constexpr size_t per_thread = sizeof(tpc::F4) * NT + 4 * PC + 4 * tpc::POOLCAP + 4 * LC;
    const size_t smem = per_thread * BT;
    const uint32_t nblk = (P.nq + BT - 1) / BT;
    typedef void (*Kern)(tpc::Params);
    static const Kern table[4] = {cells_tpc_kernel<BT, false, false, NT, PC, LC>, cells_tpc_kernel<BT, false, true, NT, PC, LC>,
                                  cells_tpc_kernel<BT, true, false, NT, PC, LC>, cells_tpc_kernel<BT, true, true, NT, PC, LC>};
    Kern k = table[(collect_events ? 2 : 0) | (faces ? 1 : 0)];
    B200_CUDA_CHECK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
    k<<<nblk, BT, smem, c->stream>>>(P);
    c->stats.gpu_launches++;
    B200_CUDA_CHECK(cudaGetLastError());
    return B200VOR_OK;
}

__global__ void exact_sign_kernel(uint32_t n, const double* __restrict__ planes, int* __restrict__ out) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if(i < n) {
        const double* p = planes + 16 * size_t(i);
        out[i] = exact::plane_vertex_sign(p, p + 4, p + 8, p + 12);
    }
}

}  // namespace

// self-test hook of the exact predicate (tests/test_gpu_tpc.py): signs of n (P0, P1, P2, P) quadruples
int exact_sign_selftest(Ctx* c, uint32_t n, const double* planes_host, int* out_host) {
    int rc;
    if((rc = c->out_d2.reserve(16 * size_t(n))) || (rc = c->out_nbr.reserve(n))) {
        return rc;
    }
    B200_CUDA_CHECK(cudaMemcpyAsync(c->out_d2.p, planes_host, 16 * size_t(n) * sizeof(double), cudaMemcpyHostToDevice, c->stream));
    exact_sign_kernel<<<(n + 63) / 64, 64, 0, c->stream>>>(n, c->out_d2.p, reinterpret_cast<int*>(c->out_nbr.p));
    c->stats.gpu_launches++;
    B200_CUDA_CHECK(cudaGetLastError());
    B200_CUDA_CHECK(cudaMemcpyAsync(out_host, c->out_nbr.p, size_t(n) * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    B200_CUDA_CHECK(cudaStreamSynchronize(c->stream));
    return B200VOR_OK;
}

// can the thread-per-cell kernel take the cells of this context?  (unweighted, single-level grid,
// at least a 3x3x3 block of bins on a periodic domain, ids that fit 27 bits)
bool tpc_applicable(const Ctx* c) {
    if(c->has_w || c->grid.two_level || c->n >= (1u << 27) || c->open || c->ndom) {
        return false;   // (a slab-local grid with open faces is walked by the team kernel only)
    }
    if(c->periodic) {
        const int gmin = std::min(c->grid.g[0], std::min(c->grid.g[1], c->grid.g[2]));
        if(gmin < 3) {
            return false;
        }
    }
    return true;
}

// big = 1: the large instance (64 vertex / plane slots), used for the cells of degenerate inputs
// that the other kernels flag UNCERTAIN: ties are decided exactly here (exact_pred.cuh)
int run_cells_tpc(Ctx* c, uint32_t nq, const uint32_t* qlist, int want_faces, int collect_events, int big) {
    if(nq == 0) {
        return B200VOR_OK;
    }
    tpc::Params P;
    P.G = c->grid;
    P.sorted = c->sorted.p;
    P.cell_start = c->cell_start.p;
    P.perm = c->perm.p;
    P.n = c->n;
    P.periodic = c->periodic;
    P.period = c->periodic ? c->period : 0.0;
    for(int a = 0; a < 6; ++a) {
        P.box[a] = c->box[a];
    }
    P.nq = nq;
    P.qlist = qlist;
    P.m = c->m.p;
    P.mg = c->mg.p;
    P.status = c->status.p;
    P.evcount = collect_events ? c->evcount.p : nullptr;
    P.nfaces = c->nfaces.p;
    P.face_nbr = c->face_nbr.p;
    P.face_area = c->face_area.p;
    P.fcap = c->fcap;
    P.target = 20.f;
    P.lockstep = 0;
    if(const char* e = getenv("B200VOR_TPC_TARGET")) {   // development knobs
        P.target = float(atof(e));
    }
    if(const char* e = getenv("B200VOR_TPC_LOCKSTEP")) {
        P.lockstep = atoi(e);
    }
    if(big) {
        return launch<64, tpc::NT_BIG, tpc::PC_BIG, tpc::LC_BIG>(c, P, collect_events, want_faces);
    }
    int bt = 224;
    if(const char* e = getenv("B200VOR_TPC_BT")) {
        bt = atoi(e);
    }
    switch(bt) {
        case 96: return launch<96, tpc::NT_STD, tpc::PC_STD, tpc::LC_STD>(c, P, collect_events, want_faces);
        case 192: return launch<192, tpc::NT_STD, tpc::PC_STD, tpc::LC_STD>(c, P, collect_events, want_faces);
        default: return launch<224, tpc::NT_STD, tpc::PC_STD, tpc::LC_STD>(c, P, collect_events, want_faces);
    }
}

}  // namespace b200vor
