// tpc_host.cpp -- TEST INFRASTRUCTURE: runs the thread-per-cell kernel body (tpc_cell.cuh, the
// very code the GPU executes) on the CPU, one cell at a time, over a counting-sort grid built
// here the way grid.cu builds it.  Lets the logic of the kernel be checked against the oracle
// without a GPU (tests/test_tpc_host.py); it is never linked into the product.
#include <math.h>
#include <stdint.h>
#include <stdlib.h>

#include <algorithm>
#include <vector>

#define TPC_HOST_STATS 1
#include "../../geogram_b200/csrc/tpc_cell.cuh"

extern "C" {
long tpc_reason[8] = {0, 0, 0, 0, 0, 0, 0, 0};
}

#include "../../geogram_b200/csrc/exact_pred.cuh"

using namespace b200vor;

// exact sign of plane P at the vertex P0 ^ P1 ^ P2 (exact_pred.cuh), for the test against rational arithmetic
extern "C" int tpc_exact_sign(const double* P0, const double* P1, const double* P2, const double* P) {
    return exact::plane_vertex_sign(P0, P1, P2, P);
}

// big != 0: the large instance (64 vertex / plane slots); nfaces / face_nbr / face_area (fcap 64) may be null
extern "C" int tpc_host_cells_ex(uint32_t n, const double* xyz, int periodic, double period, const double* box,
                                 double ppc, float target, double* m, double* mg, uint8_t* status,
                                 uint32_t* evcount, int big, uint32_t* nfaces, uint32_t* face_nbr, double* face_area);

extern "C" int tpc_host_cells(uint32_t n, const double* xyz, int periodic, double period, const double* box,
                              double ppc, float target, double* m, double* mg, uint8_t* status,
                              uint32_t* evcount) {
    return tpc_host_cells_ex(n, xyz, periodic, period, box, ppc, target, m, mg, status, evcount, 0, nullptr, nullptr,
                             nullptr);
}

extern "C" int tpc_host_cells_ex(uint32_t n, const double* xyz, int periodic, double period, const double* box,
                                 double ppc, float target, double* m, double* mg, uint8_t* status,
                                 uint32_t* evcount, int big, uint32_t* nfaces, uint32_t* face_nbr, double* face_area) {
    GridDesc G;
    G.periodic = periodic;
    G.period = period;
    G.two_level = 0;
    G.open = 0;
    double lo[3], hi[3];
    for(int a = 0; a < 3; ++a) {
        if(periodic) {
            lo[a] = 0.0;
            hi[a] = period;
        } else {
            lo[a] = 1e300;
            hi[a] = -1e300;
            for(uint32_t i = 0; i < n; ++i) {
                lo[a] = fmin(lo[a], xyz[3 * size_t(i) + a]);
                hi[a] = fmax(hi[a], xyz[3 * size_t(i) + a]);
            }
        }
    }
    double ext[3], vol = 1.0, maxext = 0.0;
    for(int a = 0; a < 3; ++a) {
        ext[a] = hi[a] - lo[a];
        maxext = fmax(maxext, ext[a]);
    }
    if(!(maxext > 0.0)) {
        maxext = 1.0;
    }
    for(int a = 0; a < 3; ++a) {
        if(!(ext[a] > 1e-9 * maxext)) {
            ext[a] = 1e-9 * maxext;
        }
        vol *= ext[a];
    }
    const double tgt = fmax(1.0, double(n) / ppc);
    const double width = cbrt(vol / tgt);
    uint64_t ncell = 1;
    for(int a = 0; a < 3; ++a) {
        int g = int(floor(ext[a] / width + 0.5));
        g = g < 1 ? 1 : (g > 2048 ? 2048 : g);
        G.g[a] = g;
        G.lo[a] = lo[a];
        G.w[a] = ext[a] / g;
        G.inv_w[a] = g / ext[a];
        ncell *= uint64_t(g);
    }
    G.ncell = uint32_t(ncell);
    G.nvirt = G.ncell;
    std::vector<uint32_t> cellid(n), start(ncell + 1, 0), perm(n);
    for(uint32_t i = 0; i < n; ++i) {
        const int ix = clampi(int(floor((xyz[3 * size_t(i)] - G.lo[0]) * G.inv_w[0])), 0, G.g[0] - 1);
        const int iy = clampi(int(floor((xyz[3 * size_t(i) + 1] - G.lo[1]) * G.inv_w[1])), 0, G.g[1] - 1);
        const int iz = clampi(int(floor((xyz[3 * size_t(i) + 2] - G.lo[2]) * G.inv_w[2])), 0, G.g[2] - 1);
        cellid[i] = uint32_t(ix) + uint32_t(G.g[0]) * (uint32_t(iy) + uint32_t(G.g[1]) * uint32_t(iz));
        start[cellid[i] + 1]++;
    }
    for(uint64_t b = 0; b < ncell; ++b) {
        start[b + 1] += start[b];
    }
    {
        std::vector<uint32_t> fill(start.begin(), start.end() - 1);
        for(uint32_t i = 0; i < n; ++i) {   // ascending user index inside a bin, as sort_bins_kernel
            perm[fill[cellid[i]]++] = i;
        }
    }
    std::vector<Seed> sorted(n);
    for(uint32_t k = 0; k < n; ++k) {
        const uint32_t i = perm[k];
        sorted[k].x = xyz[3 * size_t(i)];
        sorted[k].y = xyz[3 * size_t(i) + 1];
        sorted[k].z = xyz[3 * size_t(i) + 2];
        sorted[k].w = 0.0;
    }
    tpc::Params P;
    P.G = G;
    P.sorted = sorted.data();
    P.cell_start = start.data();
    P.perm = perm.data();
    P.n = n;
    P.periodic = periodic;
    P.period = period;
    for(int a = 0; a < 6; ++a) {
        P.box[a] = box[a];
    }
    P.nq = n;
    P.qlist = nullptr;
    P.m = m;
    P.mg = mg;
    P.status = status;
    P.evcount = evcount;
    P.target = target > 0.f ? target : 20.f;
    P.lockstep = 0;
    P.nfaces = nfaces;
    P.face_nbr = face_nbr;
    P.face_area = face_area;
    P.fcap = 64;
    const bool faces = nfaces != nullptr;
#pragma omp parallel
    {
        tpc::F4 V[tpc::NT_BIG];
        uint32_t gid[tpc::PC_BIG], pool[tpc::POOLCAP];
        uint32_t L[tpc::LC_BIG];
        tpc::Mem<1> M;
        M.V = V;
        M.gid = gid;
        M.pool = pool;
        M.L = L;
#pragma omp for schedule(dynamic, 256)
        for(uint32_t c = 0; c < n; ++c) {
            if(big) {
                if(faces) {
                    tpc::compute_cell<1, true, true, tpc::NT_BIG, tpc::PC_BIG, tpc::LC_BIG>(P, c, M);
                } else {
                    tpc::compute_cell<1, true, false, tpc::NT_BIG, tpc::PC_BIG, tpc::LC_BIG>(P, c, M);
                }
            } else if(faces) {
                tpc::compute_cell<1, true, true, tpc::NT_STD, tpc::PC_STD, tpc::LC_STD>(P, c, M);
            } else {
                tpc::compute_cell<1, true, false, tpc::NT_STD, tpc::PC_STD, tpc::LC_STD>(P, c, M);
            }
        }
    }
    return 0;
}
