// knn.cu -- exact k-nearest-neighbour search on the uniform grid.
//
// Replaces KdTree::get_nearest_neighbors (points/kd_tree.cpp:156-387) behind
// NearestNeighborSearch::get_nearest_neighbors (points/nn_search.h:129-185).
//
// One thread per query, queries in grid order so the 32 queries of a warp sit in the
// same or adjacent cells and read the same candidate spans (L1 / broadcast hits).
// Cells are visited in Chebyshev rings around the query's cell; a ring row along x is
// ONE contiguous span of the sorted Seed array.  The k best (d2, id) pairs live in a
// per-thread max-heap in shared memory (padded stride: conflict-free both for the
// per-thread heap traffic and for the transposed, coalesced write-out of the rows).
// The search stops when the k-th distance is below the distance to the unvisited
// region, so the result is the exact kNN set, not an approximation.
//
// Squared distances are ((dx*dx)+dy*dy)+dz*dz in fp64 with explicit round-to-nearest
// multiplies/adds (no FMA contraction): bit-equal to basic/geometry_nd.h:65-73 compiled
// with -ffp-contract=off, which is what test_nn_search/main.cpp:169-181 demands.
// Ties are ordered by ascending id (the reference's tie order is traversal-dependent).
#include "common.cuh"

namespace b200vor {

namespace {


struct KnnParams {
    GridDesc G;
    const Seed* sorted;
    const uint32_t* cell_start;
    const uint32_t* sub;      // two-level grid (grid.cu): per-bin sub-grid word, virtual start table
    const uint32_t* vstart;
    const uint32_t* perm;
    uint32_t n;
    KnnJob job;
};

// exact ties are ordered by USER index (grid order inside a cell depends on atomics, the
// user index does not): results are deterministic run to run
__device__ inline uint32_t user_id(const uint32_t* perm, uint32_t n, uint32_t id) {
    uint32_t inst = id / n;
    return __ldg(perm + (id - inst * n)) + inst * n;
}

template <int kStride> struct Heap {
    double* d;      // d[e * kStride]
    uint32_t* id;   // id[e * kStride]
    uint32_t count, k;
    const uint32_t* perm;
    uint32_t n;

    __device__ inline bool key_less(double da, uint32_t ia, double db, uint32_t ib) const {
        return da < db || (da == db && user_id(perm, n, ia) < user_id(perm, n, ib));
    }

    __device__ inline void sift_down(uint32_t e, uint32_t len, double dv, uint32_t iv) {
        for(;;) {
            uint32_t c = 2 * e + 1;
            if(c >= len) {
                break;
            }
            double dc = d[c * kStride];
            uint32_t ic = id[c * kStride];
            if(c + 1 < len) {
                double d2 = d[(c + 1) * kStride];
                uint32_t i2 = id[(c + 1) * kStride];
                if(key_less(dc, ic, d2, i2)) {
                    dc = d2;
                    ic = i2;
                    ++c;
                }
            }
            if(!key_less(dv, iv, dc, ic)) {
                break;
            }
            d[e * kStride] = dc;
            id[e * kStride] = ic;
            e = c;
        }
        d[e * kStride] = dv;
        id[e * kStride] = iv;
    }

    __device__ inline void offer(double dv, uint32_t iv) {
        if(count < k) {
            uint32_t e = count++;
            while(e > 0) {
                uint32_t p = (e - 1) >> 1;
                double dp = d[p * kStride];
                uint32_t ip = id[p * kStride];
                if(!key_less(dp, ip, dv, iv)) {
                    break;
                }
                d[e * kStride] = dp;
                id[e * kStride] = ip;
                e = p;
            }
            d[e * kStride] = dv;
            id[e * kStride] = iv;
        } else if(key_less(dv, iv, d[0], id[0])) {
            sift_down(0, k, dv, iv);
        }
    }

    __device__ inline void sort_ascending() {
        for(uint32_t len = count; len > 1; --len) {
            double dv = d[(len - 1) * kStride];
            uint32_t iv = id[(len - 1) * kStride];
            d[(len - 1) * kStride] = d[0];
            id[(len - 1) * kStride] = id[0];
            sift_down(0, len - 1, dv, iv);
        }
    }
};

template <int kStride>
__device__ inline void scan_span(const KnnParams& P, Heap<kStride>& H, uint32_t a, uint32_t b,
                                 double qx, double qy, double qz, double ox, double oy,
                                 double oz, uint32_t id_offset) {
    for(uint32_t j = a; j < b; ++j) {
        Seed s = load_seed(P.sorted + j);
        double dx = __dsub_rn(__dadd_rn(s.x, ox), qx);
        double dy = __dsub_rn(__dadd_rn(s.y, oy), qy);
        double dz = __dsub_rn(__dadd_rn(s.z, oz), qz);
        double d2 = __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)),
                              __dmul_rn(dz, dz));
        H.offer(d2, j + id_offset);
    }
}

// visits the cells [xa, xb] of grid row (y, z) (already wrapped), splitting at the
// periodic seam
template <int kStride>
__device__ inline void scan_row(const KnnParams& P, Heap<kStride>& H, int xa, int xb,
                                uint32_t row_base, double qx, double qy, double qz,
                                double oy, double oz, int sy, int sz) {
    const GridDesc& G = P.G;
    const int gx = G.g[0];
    int a = xa < 0 ? 0 : xa, b = xb >= gx ? gx - 1 : xb;
    if(a <= b) {
        uint32_t off = G.periodic ? P.n * translation_to_instance(0, sy, sz) : 0u;
        scan_span(P, H, __ldg(P.cell_start + row_base + a),
                  __ldg(P.cell_start + row_base + b + 1), qx, qy, qz, 0.0, oy, oz, off);
    }
    if(G.periodic) {
        if(xa < 0) {
            // cells [xa, min(xb,-1)] seen through the -x image
            int hi = (xb < 0 ? xb : -1) + gx;
            uint32_t off = P.n * translation_to_instance(-1, sy, sz);
            scan_span(P, H, __ldg(P.cell_start + row_base + (xa + gx)),
                      __ldg(P.cell_start + row_base + hi + 1), qx, qy, qz, -G.period, oy, oz,
                      off);
        }
        if(xb >= gx) {
            int lo = (xa >= gx ? xa : gx) - gx;
            uint32_t off = P.n * translation_to_instance(1, sy, sz);
            scan_span(P, H, __ldg(P.cell_start + row_base + lo),
                      __ldg(P.cell_start + row_base + (xb - gx) + 1), qx, qy, qz, G.period,
                      oy, oz, off);
        }
    }
}

// Two-level grid (clustered seeds, non-periodic): the search box grows in the FINE units of the
// query's own bin (1, 2, 4, ... sub-cells around it); every round scans the sub-cells the box meets
// and offers the seeds that lie inside the box but were outside the previous one, so that each seed
// is offered exactly once whatever the resolutions of the bins involved.  Complete when the k-th
// distance is smaller than the distance to the faces of the box.
template <int kStride>
__device__ inline bool search_two_level(const KnnParams& P, Heap<kStride>& H, double qx, double qy, double qz,
                                        uint32_t want) {
    const GridDesc& G = P.G;
    int cc[3];
    cell_coords(G, qx, qy, qz, cc[0], cc[1], cc[2]);
    const double q[3] = {qx, qy, qz};
    const uint32_t hb = uint32_t(cc[0]) + uint32_t(G.g[0]) * (uint32_t(cc[1]) + uint32_t(G.g[1]) * uint32_t(cc[2]));
    const int ls0 = int(__ldg(P.sub + hb) & 7u), s0 = 1 << ls0;
    int f0[3];
    double wl[3];
    for(int a = 0; a < 3; ++a) {
        f0[a] = cc[a] * s0 + clampi(int(floor((q[a] - G.lo[a]) * (G.inv_w[a] * double(s0)))) - cc[a] * s0, 0, s0 - 1);
        wl[a] = G.w[a] / double(s0);
    }
    double plo[3] = {1e300, 1e300, 1e300}, phi[3] = {-1e300, -1e300, -1e300};   // previous box (empty)
    for(int r = 1;; r *= 2) {
        double blo[3], bhi[3];
        int bl[3], bh[3];
        bool covers_all = true;
        for(int a = 0; a < 3; ++a) {
            const int fl = f0[a] - r, fh = f0[a] + r;
            // a box face beyond the grid is at infinity: nothing lies behind it
            // (the membership box is shrunk by eps: a seed closer than eps to a face waits for the
            //  next round, so that rounding in the stored sub-cell index can never hide a seed
            //  that the coordinate test places inside)
            const double eps = 1e-9 * G.w[a];
            blo[a] = fl <= 0 ? -1e300 : G.lo[a] + fl * wl[a] + eps;
            bhi[a] = fh >= (G.g[a] << ls0) - 1 ? 1e300 : G.lo[a] + (fh + 1) * wl[a] - eps;
            covers_all = covers_all && blo[a] <= -1e300 && bhi[a] >= 1e300;
            bl[a] = max(fl >> ls0, 0);
            bh[a] = min(fh >> ls0, G.g[a] - 1);
        }
        for(int bz = bl[2]; bz <= bh[2]; ++bz) {
            for(int by = bl[1]; by <= bh[1]; ++by) {
                for(int bx = bl[0]; bx <= bh[0]; ++bx) {
                    const uint32_t b = uint32_t(bx) + uint32_t(G.g[0]) * (uint32_t(by) + uint32_t(G.g[1]) * uint32_t(bz));
                    if(__ldg(P.cell_start + b + 1) == __ldg(P.cell_start + b)) {
                        continue;
                    }
                    const uint32_t sw = __ldg(P.sub + b);
                    const int ls = int(sw & 7u), sd = 1 << ls;
                    const int bi[3] = {bx, by, bz};
                    int ra[3], rb[3];
                    for(int a = 0; a < 3; ++a) {
                        int lof, hif;
                        if(ls >= ls0) {
                            lof = (f0[a] - r) * (1 << (ls - ls0));
                            hif = (f0[a] + r + 1) * (1 << (ls - ls0)) - 1;
                        } else {
                            lof = (f0[a] - r) >> (ls0 - ls);
                            hif = (f0[a] + r) >> (ls0 - ls);
                        }
                        const int base = bi[a] * sd;
                        ra[a] = max(lof, base) - base;
                        rb[a] = min(hif, base + sd - 1) - base;
                    }
                    for(int z = ra[2]; z <= rb[2]; ++z) {
                        for(int y = ra[1]; y <= rb[1]; ++y) {
                            const uint32_t vb = (sw >> 3) + uint32_t(sd) * (uint32_t(y) + uint32_t(sd) * uint32_t(z));
                            const uint32_t ja = __ldg(P.vstart + vb + uint32_t(ra[0]));
                            const uint32_t jb = __ldg(P.vstart + vb + uint32_t(rb[0]) + 1u);
                            for(uint32_t j = ja; j < jb; ++j) {
                                Seed sj = load_seed(P.sorted + j);
                                const bool in_new = sj.x >= blo[0] && sj.x < bhi[0] && sj.y >= blo[1] && sj.y < bhi[1] &&
                                                    sj.z >= blo[2] && sj.z < bhi[2];
                                const bool in_old = sj.x >= plo[0] && sj.x < phi[0] && sj.y >= plo[1] && sj.y < phi[1] &&
                                                    sj.z >= plo[2] && sj.z < phi[2];
                                if(in_new && !in_old) {
                                    double dx = __dsub_rn(sj.x, qx), dy = __dsub_rn(sj.y, qy), dz = __dsub_rn(sj.z, qz);
                                    double d2 = __dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz));
                                    H.offer(d2, j);
                                }
                            }
                        }
                    }
                }
            }
        }
        if(covers_all) {
            return true;    // every seed was offered
        }
        double dmin = 1e300;
        for(int a = 0; a < 3; ++a) {
            dmin = fmin(dmin, fmin(q[a] - blo[a], bhi[a] - q[a]));
        }
        dmin -= 1e-9 * G.w[0];
        if(H.count >= want && dmin > 0.0 && H.d[0] < dmin * dmin) {
            return true;
        }
        for(int a = 0; a < 3; ++a) {
            plo[a] = blo[a];
            phi[a] = bhi[a];
        }
        if(r > (1 << 30) / 2) {
            return false;
        }
    }
}

template <int kBlock>
__global__ void __launch_bounds__(kBlock) knn_kernel(KnnParams P) {
    constexpr int kStride = kBlock + 1;
    extern __shared__ __align__(16) unsigned char smem[];
    const GridDesc& G = P.G;
    const uint32_t k = P.job.k;
    double* hd = reinterpret_cast<double*>(smem);
    uint32_t* hi = reinterpret_cast<uint32_t*>(hd + size_t(k) * kStride);
    uint32_t* rows = hi + size_t(k) * kStride;   // [kBlock] output row of each thread
    uint32_t* cnts = rows + kBlock;              // [kBlock] filled entries

    const uint32_t t = blockIdx.x * kBlock + threadIdx.x;
    const bool active = t < P.job.nq;
    Heap<kStride> H;
    H.d = hd + threadIdx.x;
    H.id = hi + threadIdx.x;
    H.count = 0;
    H.k = k;
    H.perm = P.perm;
    H.n = P.n;
    bool complete = true;
    const uint32_t want_all = min(k, P.G.periodic ? 27u * P.n : P.n);
    uint32_t row = 0xFFFFFFFFu;

    if(active) {
        double qx, qy, qz;
        if(P.job.qpts) {
            qx = P.job.qpts[3 * size_t(t)];
            qy = P.job.qpts[3 * size_t(t) + 1];
            qz = P.job.qpts[3 * size_t(t) + 2];
            row = t;
        } else {
            uint32_t idx = P.job.qlist ? P.job.qlist[t] : t;
            Seed s = load_seed(P.sorted + idx);
            qx = s.x; qy = s.y; qz = s.z;
            row = P.job.user_ids ? P.perm[idx] : t;
        }
        int cx, cy, cz;
        cell_coords(G, qx, qy, qz, cx, cy, cz);
        const bool two_level = G.two_level && !G.periodic;
        if(two_level) {
            complete = search_two_level<kStride>(P, H, qx, qy, qz, want_all);
        }
        int rmax;
        if(G.periodic) {
            int gmin = min(G.g[0], min(G.g[1], G.g[2]));
            rmax = (gmin - 1) / 2;
        } else {
            rmax = max(max(cx, G.g[0] - 1 - cx),
                       max(max(cy, G.g[1] - 1 - cy), max(cz, G.g[2] - 1 - cz)));
        }
        const uint32_t want = want_all;
        for(int r = 0; !two_level; ++r) {
            for(int dz = -r; dz <= r; ++dz) {
                int z = cz + dz, sz = 0;
                if(z < 0 || z >= G.g[2]) {
                    if(!G.periodic) {
                        continue;
                    }
                    sz = z < 0 ? -1 : 1;
                    z -= sz * G.g[2];
                }
                for(int dy = -r; dy <= r; ++dy) {
                    int y = cy + dy, sy = 0;
                    if(y < 0 || y >= G.g[1]) {
                        if(!G.periodic) {
                            continue;
                        }
                        sy = y < 0 ? -1 : 1;
                        y -= sy * G.g[1];
                    }
                    uint32_t row_base = uint32_t(G.g[0]) * (uint32_t(y) + uint32_t(G.g[1]) * uint32_t(z));
                    double oy = sy * G.period, oz = sz * G.period;
                    bool shell = (dz == -r || dz == r || dy == -r || dy == r);
                    if(shell) {
                        scan_row(P, H, cx - r, cx + r, row_base, qx, qy, qz, oy, oz, sy, sz);
                    } else {
                        scan_row(P, H, cx - r, cx - r, row_base, qx, qy, qz, oy, oz, sy, sz);
                        scan_row(P, H, cx + r, cx + r, row_base, qx, qy, qz, oy, oz, sy, sz);
                    }
                }
            }
            // distance from the query to the unvisited region
            double dmin = 1e300;
            bool open = false;
            const double q[3] = {qx, qy, qz};
            const int cc[3] = {cx, cy, cz};
            for(int a = 0; a < 3; ++a) {
                if(G.periodic || cc[a] - r > 0) {
                    open = true;
                    dmin = fmin(dmin, q[a] - (G.lo[a] + (cc[a] - r) * G.w[a]));
                }
                if(G.periodic || cc[a] + r < G.g[a] - 1) {
                    open = true;
                    dmin = fmin(dmin, (G.lo[a] + (cc[a] + r + 1) * G.w[a]) - q[a]);
                }
            }
            if(!open) {
                break;   // the whole grid was visited
            }
            dmin -= 1e-9 * G.w[0];
            if(H.count >= want && dmin > 0.0 && H.d[0] < dmin * dmin) {
                break;
            }
            if(r >= rmax) {
                complete = false;   // periodic: ring limit reached before the guarantee
                break;
            }
        }
        H.sort_ascending();
    }
    rows[threadIdx.x] = row;
    cnts[threadIdx.x] = H.count;
    if(active && P.job.out_flag) {
        P.job.out_flag[P.job.user_ids ? row : t] = (complete && H.count >= want_all) ? 0 : 1;
    }
    __syncthreads();
    // transposed write-out: one warp per row, lanes across the k entries
    const uint32_t lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    for(uint32_t tt = wid; tt < kBlock; tt += kBlock / 32) {
        uint32_t rr = rows[tt];
        if(rr == 0xFFFFFFFFu) {
            continue;
        }
        uint32_t cnt = cnts[tt];
        for(uint32_t e = lane; e < k; e += 32) {
            uint32_t id = 0xFFFFFFFFu;
            double d = __longlong_as_double(0x7FF0000000000000LL);
            if(e < cnt) {
                id = hi[e * kStride + tt];
                d = hd[e * kStride + tt];
                if(P.job.user_ids) {
                    id = P.perm[id];
                }
            }
            P.job.out_nbr[size_t(rr) * k + e] = id;
            if(P.job.out_d2) {
                P.job.out_d2[size_t(rr) * k + e] = d;
            }
        }
    }
}

template <int kBlock> int launch_knn(Ctx* c, const KnnParams& P) {
    constexpr int kStride = kBlock + 1;
    size_t smem = size_t(P.job.k) * kStride * (sizeof(double) + sizeof(uint32_t)) +
                  2 * kBlock * sizeof(uint32_t);
    B200_CUDA_CHECK(cudaFuncSetAttribute(knn_kernel<kBlock>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
    uint32_t nblk = (P.job.nq + kBlock - 1) / kBlock;
    knn_kernel<kBlock><<<nblk, kBlock, smem, c->stream>>>(P);
    return B200VOR_OK;
}

}  // namespace

int run_knn(Ctx* c, const KnnJob& job) {
    if(job.nq == 0) {
        return B200VOR_OK;
    }
    if(knn_team_applicable(c, job)) {
        // team-per-query kernel (knn_team.cu); the heap kernel below takes the job again only if a
        // query met more than a pool of seeds at one distance
        uint32_t nfail = 0;
        int rc0 = run_knn_team(c, job, &nfail);
        if(rc0 || nfail == 0) {
            return rc0;
        }
    }
    KnnParams P;
    P.G = c->grid;
    P.sorted = c->sorted.p;
    P.cell_start = c->cell_start.p;
    P.sub = c->sub.p;
    P.vstart = c->vstart.p;
    P.perm = c->perm.p;
    P.n = c->n;
    P.job = job;
    // block size by k so that the heaps fit the 227 KB of shared memory
    int rc = B200VOR_OK;
    if(job.k <= 144) {
        rc = launch_knn<128>(c, P);
    } else if(job.k <= 290) {
        rc = launch_knn<64>(c, P);
    } else if(job.k <= 570) {
        rc = launch_knn<32>(c, P);
    } else {
        set_error("b200vor: k=%u exceeds the kNN kernel capacity (570)", job.k);
        return B200VOR_ERR_CAPACITY;
    }
    if(rc) {
        return rc;
    }
    c->stats.gpu_launches++;
    B200_CUDA_CHECK(cudaGetLastError());
    return B200VOR_OK;
}

}  // namespace b200vor
