// grid.cu -- uniform-grid build by counting sort (replaces BalancedKdTree::build_tree,
// points/kd_tree.cpp:429-517, as the search structure of the kNN / cell kernels).
//
// Three streaming passes, all HBM-bound:
//   bin_count   read xyz (24 B/seed), write cell id + rank (8 B), atomics on the
//               per-cell counters (L2 resident: ncell*4 B)
//   scan        exclusive prefix sum of the counters -> cell_start
//   bin_scatter read xyz + w + cellid + rank (40 B), write Seed (32 B) + perm (4 B)
// Algorithmic bytes per seed: 24 read + 24(+8 padding) + 4 written.
#include "common.cuh"

#include <math.h>
#include <stdarg.h>
#include <stdlib.h>

#include <algorithm>
#include <utility>

namespace b200vor {

namespace {

constexpr int kThreads = 256;

__global__ void bbox_kernel(const double* __restrict__ xyz, uint32_t n, double* out) {
    // out[0..2] = min, out[3..5] = max ; one block-level partial per block
    __shared__ double smin[3][kThreads / 32], smax[3][kThreads / 32];
    double mn[3] = {1e300, 1e300, 1e300}, mx[3] = {-1e300, -1e300, -1e300};
    for(uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n;
        i += gridDim.x * blockDim.x) {
        for(int c = 0; c < 3; ++c) {
            double v = xyz[3 * size_t(i) + c];
            mn[c] = fmin(mn[c], v);
            mx[c] = fmax(mx[c], v);
        }
    }
    for(int c = 0; c < 3; ++c) {
        for(int o = 16; o > 0; o >>= 1) {
            mn[c] = fmin(mn[c], __shfl_xor_sync(0xffffffffu, mn[c], o));
            mx[c] = fmax(mx[c], __shfl_xor_sync(0xffffffffu, mx[c], o));
        }
        if((threadIdx.x & 31) == 0) {
            smin[c][threadIdx.x >> 5] = mn[c];
            smax[c][threadIdx.x >> 5] = mx[c];
        }
    }
    __syncthreads();
    if(threadIdx.x < 3) {
        int c = threadIdx.x;
        double a = smin[c][0], b = smax[c][0];
        for(int w = 1; w < kThreads / 32; ++w) {
            a = fmin(a, smin[c][w]);
            b = fmax(b, smax[c][w]);
        }
        out[blockIdx.x * 6 + c] = a;
        out[blockIdx.x * 6 + 3 + c] = b;
    }
}

__global__ void bin_count_kernel(const double* __restrict__ xyz, uint32_t n, GridDesc G,
                                 uint32_t* __restrict__ cellid, uint32_t* __restrict__ rank,
                                 uint32_t* __restrict__ counts, uint32_t* __restrict__ dense_flag) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if(i >= n) {
        return;
    }
    double x = xyz[3 * size_t(i)], y = xyz[3 * size_t(i) + 1], z = xyz[3 * size_t(i) + 2];
    int ix, iy, iz;
    cell_coords(G, x, y, z, ix, iy, iz);
    uint32_t c = uint32_t(ix) + uint32_t(G.g[0]) * (uint32_t(iy) + uint32_t(G.g[1]) * uint32_t(iz));
    cellid[i] = c;
    uint32_t r = atomicAdd(&counts[c], 1u);
    rank[i] = r;
    if(r == kDenseBin) {
        *dense_flag = 1u;   // some bin holds more than kDenseBin seeds: two-level build
    }
}

// ---- two-level build (clustered seeds) ----
// sub-grid resolution of every bin and the number of virtual cells it contributes
__global__ void sub_layout_kernel(uint32_t ncell, const uint32_t* __restrict__ counts,
                                  uint32_t* __restrict__ nv, uint32_t* __restrict__ sub) {
    uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    if(b >= ncell) {
        return;
    }
    uint32_t cnt = counts[b];
    uint32_t ls = 0;
    if(cnt > kDenseBin) {
        while(ls < uint32_t(kMaxSubLog) && (uint64_t(kSubPpc) << (3 * ls)) < cnt) {
            ++ls;
        }
    }
    nv[b] = 1u << (3 * ls);
    sub[b] = ls;
}

__global__ void sub_pack_kernel(uint32_t ncell, const uint32_t* __restrict__ voff,
                                uint32_t* __restrict__ sub) {
    uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    if(b < ncell) {
        sub[b] = (voff[b] << 3) | sub[b];
    }
}

__global__ void vcell_count_kernel(const double* __restrict__ xyz, uint32_t n, GridDesc G,
                                   const uint32_t* __restrict__ sub, uint32_t* __restrict__ cellid,
                                   uint32_t* __restrict__ rank, uint32_t* __restrict__ vcounts) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if(i >= n) {
        return;
    }
    const uint32_t b = cellid[i];
    const uint32_t sw = sub[b];
    const uint32_t ls = sw & 7u;
    uint32_t v = sw >> 3;
    if(ls) {
        const int s = 1 << ls;
        const int gx = G.g[0], gy = G.g[1];
        const int ix = int(b % uint32_t(gx)), iy = int((b / uint32_t(gx)) % uint32_t(gy)),
                  iz = int(b / (uint32_t(gx) * uint32_t(gy)));
        const int bc[3] = {ix, iy, iz};
        int f[3];
        for(int a = 0; a < 3; ++a) {
            // same expression as the walk (fused.cu): fine coordinate at level ls, clamped to the bin
            int fa = int(floor((xyz[3 * size_t(i) + a] - G.lo[a]) * (G.inv_w[a] * double(s))));
            f[a] = clampi(fa - bc[a] * s, 0, s - 1);
        }
        v += uint32_t(f[0] + s * (f[1] + s * f[2]));
    }
    cellid[i] = v;
    rank[i] = atomicAdd(&vcounts[v], 1u);
}

// level-1 start table (knn.cu, cell.cu) from the virtual one
__global__ void level1_start_kernel(uint32_t ncell, uint32_t n, const uint32_t* __restrict__ sub,
                                    const uint32_t* __restrict__ vstart,
                                    uint32_t* __restrict__ cell_start) {
    uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    if(b < ncell) {
        cell_start[b] = vstart[sub[b] >> 3];
    } else if(b == ncell) {
        cell_start[b] = n;
    }
}

// ---- exclusive scan of uint32 (three phases, 4096 items per block) ----
constexpr int kScanItems = 16;
constexpr int kScanTile = kThreads * kScanItems;

__device__ inline uint32_t block_exclusive_scan(uint32_t v, uint32_t* total) {
    __shared__ uint32_t wsum[kThreads / 32];
    uint32_t lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    uint32_t inc = v;
    for(int o = 1; o < 32; o <<= 1) {
        uint32_t t = __shfl_up_sync(0xffffffffu, inc, o);
        if(lane >= uint32_t(o)) {
            inc += t;
        }
    }
    if(lane == 31) {
        wsum[wid] = inc;
    }
    __syncthreads();
    if(wid == 0) {
        uint32_t s = lane < kThreads / 32 ? wsum[lane] : 0;
        for(int o = 1; o < kThreads / 32; o <<= 1) {
            uint32_t t = __shfl_up_sync(0xffffffffu, s, o);
            if(lane >= uint32_t(o)) {
                s += t;
            }
        }
        if(lane < kThreads / 32) {
            wsum[lane] = s;
        }
    }
    __syncthreads();
    uint32_t base = wid ? wsum[wid - 1] : 0;
    if(total) {
        *total = wsum[kThreads / 32 - 1];
    }
    __syncthreads();
    return base + inc - v;
}

__global__ void scan_tile_sums(const uint32_t* __restrict__ in, uint32_t n,
                               uint32_t* __restrict__ tile_sums) {
    size_t base = size_t(blockIdx.x) * kScanTile + size_t(threadIdx.x) * kScanItems;
    uint32_t s = 0;
    for(int k = 0; k < kScanItems; ++k) {
        if(base + k < n) {
            s += in[base + k];
        }
    }
    uint32_t total;
    block_exclusive_scan(s, &total);
    if(threadIdx.x == 0) {
        tile_sums[blockIdx.x] = total;
    }
}

__global__ void scan_of_sums(uint32_t* tile_sums, uint32_t ntiles) {
    // single block: in-place exclusive scan, chunk by chunk
    __shared__ uint32_t carry;
    if(threadIdx.x == 0) {
        carry = 0;
    }
    __syncthreads();
    for(uint32_t b = 0; b < ntiles; b += kThreads) {
        uint32_t i = b + threadIdx.x;
        uint32_t v = i < ntiles ? tile_sums[i] : 0;
        uint32_t total;
        uint32_t ex = block_exclusive_scan(v, &total);
        uint32_t c = carry;
        if(i < ntiles) {
            tile_sums[i] = c + ex;
        }
        __syncthreads();
        if(threadIdx.x == 0) {
            carry = c + total;
        }
        __syncthreads();
    }
}

__global__ void scan_apply(const uint32_t* __restrict__ in, uint32_t n,
                           const uint32_t* __restrict__ tile_sums,
                           uint32_t* __restrict__ out, uint32_t total_n) {
    size_t base = size_t(blockIdx.x) * kScanTile + size_t(threadIdx.x) * kScanItems;
    uint32_t v[kScanItems];
    uint32_t s = 0;
    for(int k = 0; k < kScanItems; ++k) {
        v[k] = (base + k < n) ? in[base + k] : 0;
        s += v[k];
    }
    uint32_t ex = block_exclusive_scan(s, nullptr) + tile_sums[blockIdx.x];
    for(int k = 0; k < kScanItems; ++k) {
        if(base + k < n) {
            out[base + k] = ex;
        }
        ex += v[k];
    }
    if(blockIdx.x == 0 && threadIdx.x == 0) {
        out[n] = total_n;
    }
}

__global__ void bin_scatter_kernel(const double* __restrict__ xyz,
                                   const double* __restrict__ w, uint32_t n,
                                   const uint32_t* __restrict__ cellid,
                                   const uint32_t* __restrict__ rank,
                                   const uint32_t* __restrict__ cell_start,
                                   Seed* __restrict__ sorted, uint32_t* __restrict__ perm) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if(i >= n) {
        return;
    }
    uint32_t pos = cell_start[cellid[i]] + rank[i];
    double x = xyz[3 * size_t(i)], y = xyz[3 * size_t(i) + 1], z = xyz[3 * size_t(i) + 2];
    double ww = w ? w[i] : 0.0;
    double2* dst = reinterpret_cast<double2*>(sorted + pos);
    dst[0] = make_double2(x, y);
    dst[1] = make_double2(z, ww);
    perm[pos] = i;
}

// Orders the seeds of every bin by user index.  bin_count hands out ranks with atomics, so
// the order inside a bin would otherwise change from run to run, and with it the order in
// which equidistant neighbours are clipped (results identical to rounding, not bitwise).
// The ranks are almost sorted already: this is one read of perm plus rare moves.
__global__ void sort_bins_kernel(uint32_t ncell, const uint32_t* __restrict__ cell_start,
                                 Seed* __restrict__ sorted, uint32_t* __restrict__ perm) {
    uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
    if(c >= ncell) {
        return;
    }
    const uint32_t a = cell_start[c], b = cell_start[c + 1];
    for(uint32_t i = a + 1; i < b; ++i) {
        const uint32_t key = perm[i];
        if(perm[i - 1] <= key) {
            continue;
        }
        const double2* src = reinterpret_cast<const double2*>(sorted + i);
        const double2 s0 = src[0], s1 = src[1];
        uint32_t j = i;
        while(j > a && perm[j - 1] > key) {
            perm[j] = perm[j - 1];
            double2* dst = reinterpret_cast<double2*>(sorted + j);
            const double2* from = reinterpret_cast<const double2*>(sorted + j - 1);
            dst[0] = from[0];
            dst[1] = from[1];
            --j;
        }
        perm[j] = key;
        double2* dst = reinterpret_cast<double2*>(sorted + j);
        dst[0] = s0;
        dst[1] = s1;
    }
}

// seeds per super-bin (8 x 8 x 8 bins): the coarse occupancy level of the two-level walk
__global__ void superbin_count_kernel(GridDesc G, const uint32_t* __restrict__ cell_start,
                                      uint32_t* __restrict__ sbcount) {
    const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    if(b >= G.ncell) {
        return;
    }
    const uint32_t cnt = cell_start[b + 1] - cell_start[b];
    if(cnt == 0) {
        return;
    }
    const uint32_t gx = uint32_t(G.g[0]), gy = uint32_t(G.g[1]);
    const uint32_t ix = b % gx, iy = (b / gx) % gy, iz = b / (gx * gy);
    const uint32_t sx = (gx + (1u << kSuperLog) - 1) >> kSuperLog, sy = (gy + (1u << kSuperLog) - 1) >> kSuperLog;
    atomicAdd(&sbcount[(ix >> kSuperLog) + sx * ((iy >> kSuperLog) + sy * (iz >> kSuperLog))], cnt);
}

// ---- Hilbert order of the bins (role of compute_Hilbert_order, mesh/mesh_reorder.cpp:1260-1300) ----
// cell (x, y, z) of the 2^b cube at position h of the 3-D Hilbert curve (Skilling's transpose form)
__device__ inline void hilbert_decode(uint32_t h, int b, uint32_t& x, uint32_t& y, uint32_t& z) {
    uint32_t X[3] = {0u, 0u, 0u};
    for(int i = 0; i < b; ++i) {
        X[0] |= ((h >> (3 * i + 2)) & 1u) << i;
        X[1] |= ((h >> (3 * i + 1)) & 1u) << i;
        X[2] |= ((h >> (3 * i)) & 1u) << i;
    }
    const uint32_t N = 2u << (b - 1);
    uint32_t t = X[2] >> 1;
    X[2] ^= X[1];
    X[1] ^= X[0];
    X[0] ^= t;
    for(uint32_t Q = 2; Q != N; Q <<= 1) {
        const uint32_t P = Q - 1;
        for(int i = 2; i >= 0; --i) {
            if(X[i] & Q) {
                X[0] ^= P;
            } else {
                t = (X[0] ^ X[i]) & P;
                X[0] ^= t;
                X[i] ^= t;
            }
        }
    }
    x = X[0]; y = X[1]; z = X[2];
}

// The g bins of an axis are spread evenly over the 2^b cells of the cube axis (bin x sits in cell
// floor(x 2^b / g)), so the curve never leaves the grid for long: the bin of cube cell X, or -1.
__device__ inline int cube_to_bin(uint32_t X, int g, int b) {
    const uint32_t x = uint32_t((uint64_t(X) * uint32_t(g) + ((1u << b) - 1u)) >> b);   // ceil(X g / 2^b)
    return (x < uint32_t(g) && uint32_t((uint64_t(x) << b) / uint32_t(g)) == X) ? int(x) : -1;
}

// seeds of the bin at curve position h (0 if the cube cell holds no bin)
__global__ void hilbert_count_kernel(GridDesc G, int b, uint32_t nh, const uint32_t* __restrict__ cell_start,
                                     uint32_t* __restrict__ cnt) {
    const uint32_t h = blockIdx.x * blockDim.x + threadIdx.x;
    if(h >= nh) {
        return;
    }
    uint32_t x, y, z, c = 0;
    hilbert_decode(h, b, x, y, z);
    const int bx = cube_to_bin(x, G.g[0], b), by = cube_to_bin(y, G.g[1], b), bz = cube_to_bin(z, G.g[2], b);
    if(bx >= 0 && by >= 0 && bz >= 0) {
        const uint32_t bin = uint32_t(bx) + uint32_t(G.g[0]) * (uint32_t(by) + uint32_t(G.g[1]) * uint32_t(bz));
        c = cell_start[bin + 1] - cell_start[bin];
    }
    cnt[h] = c;
}

__global__ void hilbert_emit_kernel(GridDesc G, int b, uint32_t nh, const uint32_t* __restrict__ cell_start,
                                    const uint32_t* __restrict__ off, const uint32_t* __restrict__ perm,
                                    uint32_t* __restrict__ order) {
    const uint32_t h = blockIdx.x * blockDim.x + threadIdx.x;
    if(h >= nh) {
        return;
    }
    uint32_t x, y, z;
    hilbert_decode(h, b, x, y, z);
    const int bx = cube_to_bin(x, G.g[0], b), by = cube_to_bin(y, G.g[1], b), bz = cube_to_bin(z, G.g[2], b);
    if(bx >= 0 && by >= 0 && bz >= 0) {
        const uint32_t bin = uint32_t(bx) + uint32_t(G.g[0]) * (uint32_t(by) + uint32_t(G.g[1]) * uint32_t(bz));
        const uint32_t a = cell_start[bin], e = cell_start[bin + 1];
        for(uint32_t k = a; k < e; ++k) {
            order[off[h] + (k - a)] = perm[k];
        }
    }
}

// ---- local weight bound (Laguerre cells): wloc[b] = max weight of the seeds in the bins within
//      Chebyshev distance kWlocRing of bin b; lets the security radius use a LOCAL weight spread
//      (fused_kernel.inc, sr_bound) when the weights vary smoothly over the domain ----
__global__ void bin_wmax_kernel(uint32_t ncell, const uint32_t* __restrict__ cell_start,
                                const Seed* __restrict__ sorted, double* __restrict__ out) {
    uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    if(b >= ncell) {
        return;
    }
    double mx = -1e300;
    for(uint32_t k = cell_start[b]; k < cell_start[b + 1]; ++k) {
        mx = fmax(mx, sorted[k].w);
    }
    out[b] = mx;
}

__global__ void dilate_max_kernel(GridDesc G, int axis, int ring, const double* __restrict__ in,
                                  double* __restrict__ out) {
    uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
    if(b >= G.ncell) {
        return;
    }
    const int gx = G.g[0], gy = G.g[1];
    int idx[3] = {int(b % uint32_t(gx)), int((b / uint32_t(gx)) % uint32_t(gy)), int(b / (uint32_t(gx) * uint32_t(gy)))};
    const int stride[3] = {1, gx, gx * gy};
    const int ga = G.g[axis], base = int(b) - idx[axis] * stride[axis];
    double mx = -1e300;
    if(2 * ring + 1 >= ga) {
        for(int t = 0; t < ga; ++t) {
            mx = fmax(mx, in[base + t * stride[axis]]);
        }
    } else {
        for(int d = -ring; d <= ring; ++d) {
            int t = idx[axis] + d;
            if(t < 0 || t >= ga) {
                if(!G.periodic) {
                    continue;
                }
                t += t < 0 ? ga : -ga;
            }
            mx = fmax(mx, in[base + t * stride[axis]]);
        }
    }
    out[b] = mx;
}

}  // namespace

// order[n] (device): the seeds bin by bin ALONG THE HILBERT CURVE of the grid, inside a bin by index.
// No sort: the curve positions are enumerated (8^b of them, b = bits of the largest grid side), the
// bins inside the grid counted, scanned and emitted.  Returns 1 if the grid is too elongated for that
// (8^b > 2^27): the caller falls back to the raster order.
int hilbert_order(Ctx* c, uint32_t* order_dev) {
    const GridDesc& G = c->grid;
    int b = 1;
    while((1 << b) < std::max(G.g[0], std::max(G.g[1], G.g[2]))) {
        ++b;
    }
    if(3 * b > 27) {
        return 1;
    }
    const uint32_t nh = 1u << (3 * b);
    const uint32_t ntiles = (nh + kScanTile - 1) / kScanTile;
    int rc;
    if((rc = c->scan_tmp.reserve(size_t(ntiles) + 2 * size_t(nh) + 128))) {
        return rc;
    }
    uint32_t* tile_sums = c->scan_tmp.p;
    uint32_t* cnt = c->scan_tmp.p + ntiles + 1;
    uint32_t* off = cnt + nh;
    cudaStream_t st = c->stream;
    hilbert_count_kernel<<<(nh + kThreads - 1) / kThreads, kThreads, 0, st>>>(G, b, nh, c->cell_start.p, cnt);
    scan_tile_sums<<<ntiles, kThreads, 0, st>>>(cnt, nh, tile_sums);
    scan_of_sums<<<1, kThreads, 0, st>>>(tile_sums, ntiles);
    scan_apply<<<ntiles, kThreads, 0, st>>>(cnt, nh, tile_sums, off, c->n);   // (also writes off[nh] = n)
    hilbert_emit_kernel<<<(nh + kThreads - 1) / kThreads, kThreads, 0, st>>>(G, b, nh, c->cell_start.p, off, c->perm.p, order_dev);
    c->stats.gpu_launches += 5;
    B200_CUDA_CHECK(cudaGetLastError());
    return B200VOR_OK;
}

int build_wloc(Ctx* c) {
    // two fields: a tight one (kWlocRingNear bins) and a wide one (kWlocRingFar bins); a cell uses
    // the tightest whose far-field condition holds (fused_kernel.inc, sr_bound)
    const GridDesc& G = c->grid;
    int rc;
    if((rc = c->wloc.reserve(2 * size_t(G.ncell))) || (rc = c->wloc_tmp.reserve(2 * size_t(G.ncell)))) {
        return rc;
    }
    uint32_t nb = (G.ncell + kThreads - 1) / kThreads;
    double* near = c->wloc.p;
    double* far = c->wloc.p + G.ncell;
    double* t0 = c->wloc_tmp.p;
    double* t1 = c->wloc_tmp.p + G.ncell;
    bin_wmax_kernel<<<nb, kThreads, 0, c->stream>>>(G.ncell, c->cell_start.p, c->sorted.p, t0);
    dilate_max_kernel<<<nb, kThreads, 0, c->stream>>>(G, 0, kWlocRingNear, t0, t1);
    dilate_max_kernel<<<nb, kThreads, 0, c->stream>>>(G, 1, kWlocRingNear, t1, far);
    dilate_max_kernel<<<nb, kThreads, 0, c->stream>>>(G, 2, kWlocRingNear, far, near);
    // the wide field is the tight one dilated further (max filters compose)
    dilate_max_kernel<<<nb, kThreads, 0, c->stream>>>(G, 0, kWlocRingFar - kWlocRingNear, near, t0);
    dilate_max_kernel<<<nb, kThreads, 0, c->stream>>>(G, 1, kWlocRingFar - kWlocRingNear, t0, t1);
    dilate_max_kernel<<<nb, kThreads, 0, c->stream>>>(G, 2, kWlocRingFar - kWlocRingNear, t1, far);
    c->stats.gpu_launches += 7;
    B200_CUDA_CHECK(cudaGetLastError());
    return B200VOR_OK;
}

int build_grid_impl(Ctx* c);

int build_grid(Ctx* c) {
    int rc = build_grid_impl(c);
    if(rc == B200VOR_OK && c->has_w) {
        rc = build_wloc(c);
    }
    return rc;
}

int build_grid_impl(Ctx* c) {
    const uint32_t n = c->n;
    cudaStream_t st = c->stream;
    GridDesc& G = c->grid;
    G.periodic = c->periodic;
    G.period = c->period;
    G.open = c->open;
    double lo[3], hi[3];
    if(c->periodic) {
        for(int a = 0; a < 3; ++a) {
            lo[a] = 0.0;
            hi[a] = c->period;
        }
    } else {
        const int nb = 296;
        if(int rc0 = c->scan_tmp.reserve(size_t(nb) * 6 * 2 + 64)) {
            return rc0;
        }
        double* part = reinterpret_cast<double*>(c->scan_tmp.p);
        bbox_kernel<<<nb, kThreads, 0, st>>>(c->xyz.p, n, part);
        c->stats.gpu_launches++;
        double host[nb * 6];
        B200_CUDA_CHECK(cudaMemcpyAsync(host, part, sizeof(host), cudaMemcpyDeviceToHost, st));
        B200_CUDA_CHECK(cudaStreamSynchronize(st));
        for(int a = 0; a < 3; ++a) {
            lo[a] = 1e300;
            hi[a] = -1e300;
        }
        for(int b = 0; b < nb; ++b) {
            for(int a = 0; a < 3; ++a) {
                lo[a] = fmin(lo[a], host[b * 6 + a]);
                hi[a] = fmax(hi[a], host[b * 6 + 3 + a]);
            }
        }
    }
    double ext[3], vol = 1.0;
    double maxext = 0.0;
    for(int a = 0; a < 3; ++a) {
        ext[a] = hi[a] - lo[a];
        maxext = fmax(maxext, ext[a]);
    }
    if(!(maxext > 0.0)) {
        maxext = 1.0;
    }
    for(int a = 0; a < 3; ++a) {
        if(!(ext[a] > 1e-9 * maxext)) {
            ext[a] = 1e-9 * maxext;   // flat direction: one layer of cells
        }
        vol *= ext[a];
    }
    double target = fmax(1.0, double(n) / c->ppc);
    double width = cbrt(vol / target);
    uint64_t ncell = 1;
    for(int a = 0; a < 3; ++a) {
        int g = int(floor(ext[a] / width + 0.5));
        if(g < 1) {
            g = 1;
        }
        if(g > 2048) {
            g = 2048;
        }
        G.g[a] = g;
        G.lo[a] = lo[a];
        G.w[a] = ext[a] / g;
        G.inv_w[a] = g / ext[a];
        ncell *= uint64_t(g);
    }
    G.ncell = uint32_t(ncell);
    c->stats.grid_dim[0] = G.g[0];
    c->stats.grid_dim[1] = G.g[1];
    c->stats.grid_dim[2] = G.g[2];

    int rc;
    if((rc = c->cellid.reserve(n)) || (rc = c->rank.reserve(n)) ||
       (rc = c->perm.reserve(n)) || (rc = c->sorted.reserve(n)) ||
       (rc = c->cell_start.reserve(size_t(G.ncell) + 1))) {
        return rc;
    }
    uint32_t ntiles = (G.ncell + kScanTile - 1) / kScanTile;
    // counters live after the tile sums in scan_tmp; one word in front of them is the dense flag
    if((rc = c->scan_tmp.reserve(size_t(ntiles) + size_t(G.ncell) + 4096))) {
        return rc;
    }
    uint32_t* tile_sums = c->scan_tmp.p;
    uint32_t* dense_flag = c->scan_tmp.p + ntiles + 1;
    uint32_t* counts = c->scan_tmp.p + ntiles + 2;
    B200_CUDA_CHECK(cudaMemsetAsync(dense_flag, 0, (size_t(G.ncell) + 1) * sizeof(uint32_t), st));
    uint32_t nblk = (n + kThreads - 1) / kThreads;
    bin_count_kernel<<<nblk, kThreads, 0, st>>>(c->xyz.p, n, G, c->cellid.p, c->rank.p, counts, dense_flag);
    c->stats.gpu_launches++;
    uint32_t dense = 0;
    B200_CUDA_CHECK(cudaMemcpyAsync(&dense, dense_flag, sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    B200_CUDA_CHECK(cudaStreamSynchronize(st));
    // small periodic grids also take the two-level walk (with 1x1x1 sub-grids): it enumerates the
    // bins of its box one by one, each with its own periodic image, so the box may span up to
    // three periods -- what a cell reaching beyond half a period needs (27 instances)
    if(c->periodic && std::min(G.g[0], std::min(G.g[1], G.g[2])) < kSmallPeriodicGrid) {
        dense = 1;
    }
    G.two_level = dense ? 1 : 0;
    G.nvirt = G.ncell;
    if(!dense) {
        scan_tile_sums<<<ntiles, kThreads, 0, st>>>(counts, G.ncell, tile_sums);
        scan_of_sums<<<1, kThreads, 0, st>>>(tile_sums, ntiles);
        scan_apply<<<ntiles, kThreads, 0, st>>>(counts, G.ncell, tile_sums, c->cell_start.p, n);
        bin_scatter_kernel<<<nblk, kThreads, 0, st>>>(c->xyz.p, c->has_w ? c->wgt.p : nullptr, n,
                                                      c->cellid.p, c->rank.p, c->cell_start.p,
                                                      c->sorted.p, c->perm.p);
        sort_bins_kernel<<<(G.ncell + kThreads - 1) / kThreads, kThreads, 0, st>>>(
            G.ncell, c->cell_start.p, c->sorted.p, c->perm.p);
        c->stats.gpu_launches += 5;
        B200_CUDA_CHECK(cudaGetLastError());
        return B200VOR_OK;
    }
    // ---- clustered seeds: dense bins get their own sub-grid; the counting sort runs over
    //      the virtual cells, so that the seeds of a bin stay contiguous (level-1 table valid) ----
    if((rc = c->sub.reserve(size_t(G.ncell) + 1)) || (rc = c->vcount.reserve(size_t(G.ncell) + 1))) {
        return rc;
    }
    uint32_t cblk = (G.ncell + kThreads - 1) / kThreads;
    uint32_t* nv = c->vcount.p;         // per-bin virtual-cell counts, then re-used below
    sub_layout_kernel<<<cblk, kThreads, 0, st>>>(G.ncell, counts, nv, c->sub.p);
    // voff = exclusive scan of nv (written over `counts`, no longer needed)
    scan_tile_sums<<<ntiles, kThreads, 0, st>>>(nv, G.ncell, tile_sums);
    scan_of_sums<<<1, kThreads, 0, st>>>(tile_sums, ntiles);
    scan_apply<<<ntiles, kThreads, 0, st>>>(nv, G.ncell, tile_sums, counts, 0u);
    sub_pack_kernel<<<cblk, kThreads, 0, st>>>(G.ncell, counts, c->sub.p);
    c->stats.gpu_launches += 5;
    // total number of virtual cells = voff[last] + nv[last]
    uint32_t last[2];
    B200_CUDA_CHECK(cudaMemcpyAsync(&last[0], counts + (G.ncell - 1), sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    B200_CUDA_CHECK(cudaMemcpyAsync(&last[1], nv + (G.ncell - 1), sizeof(uint32_t), cudaMemcpyDeviceToHost, st));
    B200_CUDA_CHECK(cudaStreamSynchronize(st));
    const uint64_t nvirt64 = uint64_t(last[0]) + last[1];
    if(nvirt64 >= (1ull << 29)) {
        set_error("build_grid: %llu virtual cells exceed the two-level table", (unsigned long long)nvirt64);
        return B200VOR_ERR_CAPACITY;
    }
    G.nvirt = uint32_t(nvirt64);
    uint32_t vtiles = (G.nvirt + kScanTile - 1) / kScanTile;
    if((rc = c->vstart.reserve(size_t(G.nvirt) + 1)) ||
       (rc = c->vcount.reserve(size_t(G.nvirt) + size_t(vtiles) + 64))) {
        return rc;
    }
    uint32_t* vcounts = c->vcount.p;
    uint32_t* vtile_sums = c->vcount.p + G.nvirt + 1;
    B200_CUDA_CHECK(cudaMemsetAsync(vcounts, 0, size_t(G.nvirt) * sizeof(uint32_t), st));
    vcell_count_kernel<<<nblk, kThreads, 0, st>>>(c->xyz.p, n, G, c->sub.p, c->cellid.p, c->rank.p, vcounts);
    scan_tile_sums<<<vtiles, kThreads, 0, st>>>(vcounts, G.nvirt, vtile_sums);
    scan_of_sums<<<1, kThreads, 0, st>>>(vtile_sums, vtiles);
    scan_apply<<<vtiles, kThreads, 0, st>>>(vcounts, G.nvirt, vtile_sums, c->vstart.p, n);
    bin_scatter_kernel<<<nblk, kThreads, 0, st>>>(c->xyz.p, c->has_w ? c->wgt.p : nullptr, n,
                                                  c->cellid.p, c->rank.p, c->vstart.p,
                                                  c->sorted.p, c->perm.p);
    sort_bins_kernel<<<(G.nvirt + kThreads - 1) / kThreads, kThreads, 0, st>>>(
        G.nvirt, c->vstart.p, c->sorted.p, c->perm.p);
    level1_start_kernel<<<(G.ncell + 1 + kThreads - 1) / kThreads, kThreads, 0, st>>>(
        G.ncell, n, c->sub.p, c->vstart.p, c->cell_start.p);
    {
        const size_t nsb = size_t((G.g[0] + (1 << kSuperLog) - 1) >> kSuperLog) * size_t((G.g[1] + (1 << kSuperLog) - 1) >> kSuperLog) *
                           size_t((G.g[2] + (1 << kSuperLog) - 1) >> kSuperLog);
        if((rc = c->sbcount.reserve(nsb))) {
            return rc;
        }
        B200_CUDA_CHECK(cudaMemsetAsync(c->sbcount.p, 0, nsb * sizeof(uint32_t), st));
        superbin_count_kernel<<<(G.ncell + kThreads - 1) / kThreads, kThreads, 0, st>>>(G, c->cell_start.p, c->sbcount.p);
        c->stats.gpu_launches++;
    }
    c->stats.gpu_launches += 7;
    B200_CUDA_CHECK(cudaGetLastError());
    return B200VOR_OK;
}

}  // namespace b200vor
