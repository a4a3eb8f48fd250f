// slab.cu -- multi-GPU cell path below the C ABI: x-slabs with halos of ghost seeds.
//
// One process per GPU.  Rank r owns the rows [first, first + n_own) of the global seed array; the
// caller hands every rank an x-slab (tools and bench sort the seeds by x once, sharding.slab_order),
// but nothing below depends on it: the slab of a rank is the x-interval its own seeds span, measured
// every step.  Per cell pass:
//   1. [lo_r, hi_r] of every rank -> all ranks (one ncclAllGather of 4 doubles per rank);
//   2. every rank packs, for every peer p, its own seeds (and, on a periodic domain, their x-images
//      x +- period, delaunay/periodic.h:95-113) that fall within [lo_p - H, hi_p + H]: a STABLE
//      bucketed compaction (own-index order), so the local seed order -- and therefore every result
//      -- is reproducible; counts are exchanged, then the ghosts travel peer to peer with
//      ncclSend / ncclRecv over NVLink (40 bytes per ghost: x', y, z, w, global id | x-image);
//   3. periodic domain: the y / z images of the own + ghost seeds within H of a y / z face are
//      appended locally (second bucketed compaction), so the local problem is a plain box problem;
//   4. the rank bins its n_own + ghosts seeds only (slab-local grid, GridDesc::open marks the faces
//      behind which seeds are missing) and runs the team kernel on its own cells;
//   5. an own cell whose security radius leaves the halo comes back UNCERTIFIED (or touching an
//      open face): the counts are all-gathered with the next bounds; if any rank has one, every rank
//      doubles H and repeats the pass (re-queue) before the seeds are updated.
// The reference has no distributed counterpart; its hint is periodic_delaunay_3d.cpp:3432-3436.
#include "common.cuh"

#include <dlfcn.h>
#include <math.h>
#include <nccl.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <vector>

namespace b200vor {

namespace {

// ---- NCCL, bound at run time (a single-GPU user of libb200vor.so needs no NCCL) ----
struct NcclApi {
    void* lib = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
};
NcclApi g_nccl;

int load_nccl() {
    if(g_nccl.lib) {
        return B200VOR_OK;
    }
    // a process that already runs torch.distributed has its libnccl.so.2 loaded: same soname
    void* lib = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if(!lib) {
        lib = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    }
    if(!lib) {
        set_error("b200vor_comm: libnccl.so.2 not found (%s)", dlerror());
        return B200VOR_ERR_ARG;
    }
#define B200_SYM(field, name)                                                   \
    *reinterpret_cast<void**>(&g_nccl.field) = dlsym(lib, name);                \
    if(!g_nccl.field) {                                                         \
        set_error("b200vor_comm: symbol %s missing in libnccl", name);          \
        return B200VOR_ERR_ARG;                                                 \
    }
    B200_SYM(GetUniqueId, "ncclGetUniqueId")
    B200_SYM(CommInitRank, "ncclCommInitRank")
    B200_SYM(CommDestroy, "ncclCommDestroy")
    B200_SYM(Send, "ncclSend")
    B200_SYM(Recv, "ncclRecv")
    B200_SYM(AllGather, "ncclAllGather")
    B200_SYM(GroupStart, "ncclGroupStart")
    B200_SYM(GroupEnd, "ncclGroupEnd")
    B200_SYM(GetErrorString, "ncclGetErrorString")
#undef B200_SYM
    g_nccl.lib = lib;
    return B200VOR_OK;
}

#define B200_NCCL_CHECK(expr)                                                                   \
    do {                                                                                        \
        ncclResult_t r__ = (expr);                                                              \
        if(r__ != ncclSuccess) {                                                                \
            set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #expr, g_nccl.GetErrorString(r__)); \
            return B200VOR_ERR_CUDA;                                                            \
        }                                                                                       \
    } while(0)

constexpr int kMaxWorld = 10;        // 3 buckets per peer in a 32-bit mask
constexpr int kPackThreads = 256;
constexpr int kPackItems = 8;        // items per thread
constexpr int kPackBlock = kPackThreads * kPackItems;
constexpr int kGhostDoubles = 5;     // x', y, z, w, (global id | x-image << 32)

}  // namespace

struct Slab {
    ncclComm_t comm = nullptr;
    int rank = 0, world = 1;
    bool has_comm = false;
    // the caller's domain (the context itself holds the slab-local box)
    bool dom_captured = false;
    int dom_periodic = 0;
    double dom_period = 1.0;
    double dom_box[6] = {0, 0, 0, 1, 1, 1};
    bool active = false;
    uint32_t n_own = 0;
    uint64_t first = 0, n_total = 0;
    bool has_w = false;
    DevBuf<double> own, own_w, next;       // own seeds (3 n_own), weights, staged Lloyd update
    DevBuf<double> sendbuf, recvbuf;       // ghosts, kGhostDoubles each
    DevBuf<uint32_t> blkcnt;               // bucketed compaction: [bucket][block] counts -> offsets
    DevBuf<double> gather_dev;             // all-gather scratch (4 doubles per rank) + local part
    DevBuf<uint32_t> count_dev;            // count matrix (world x world) + own row
    DevBuf<uint32_t> lgid;                 // global index of every local seed
    DevBuf<uint8_t> ltc;                   // packed translation (tx+1 | ty+1 << 2 | tz+1 << 4)
    DevBuf<double> red;                    // reductions
    double halo_bins = 3.0;                // initial halo, in bins of the global density
    double H = 0.0;
    std::vector<double> lo, hi;            // slab bounds of every rank
    double wmax_global = 0.0;
    double center = 0.0;                   // periodic domain: the own seeds are taken at their image next to it
    bool bounds_valid = false;             // lo / hi describe the current own seeds
    b200vor_slab_stats st = {};
    uint64_t alloc0 = 0;   // devbuf_alloc_count() when the current call began
    cudaEvent_t ev[4] = {};
};

namespace {

// ---------------------------------------------------------------------------------------------
// bucketed stable compaction: item i goes to every bucket b whose bit is set in mask(i); inside a
// bucket the entries keep the order of the items.  count -> exclusive scan (bucket-major) -> scatter
// ---------------------------------------------------------------------------------------------
// image of the x-period that brings x next to `center` (0 on a box domain)
__device__ inline int home_image(double x, double center, double period, int periodic) {
    return periodic ? -int(rint((x - center) / period)) : 0;
}

struct XPred {
    // buckets q = 3 p + (s + 1): own seed i, image s of the x-period, wanted by peer p
    const double* own;
    uint32_t n;
    int world, rank, periodic;
    double period, H, center;
    double lo[kMaxWorld], hi[kMaxWorld];
    __device__ uint32_t mask(uint32_t i) const {
        const double x = own[3 * size_t(i)];
        const int s0 = home_image(x, center, period, periodic);   // the image this rank keeps as its own row
        uint32_t m = 0;
        for(int p = 0; p < world; ++p) {
            for(int s = -1; s <= 1; ++s) {
                if((s != 0 && !periodic) || (s == s0 && p == rank)) {
                    continue;
                }
                const double xs = x + double(s) * period;
                if(xs >= lo[p] - H && xs <= hi[p] + H) {
                    m |= 1u << (3 * p + s + 1);
                }
            }
        }
        return m;
    }
};

struct XEmit {
    const double* own;
    const double* own_w;
    uint64_t first;
    double period;
    double* out;   // kGhostDoubles per entry
    __device__ void emit(uint32_t i, uint32_t b, uint32_t pos) const {
        const int s = int(b % 3u) - 1;
        double* o = out + size_t(pos) * kGhostDoubles;
        o[0] = own[3 * size_t(i)] + double(s) * period;
        o[1] = own[3 * size_t(i) + 1];
        o[2] = own[3 * size_t(i) + 2];
        o[3] = own_w ? own_w[i] : 0.0;
        const unsigned long long tag = (unsigned long long)(first + i) | ((unsigned long long)(s + 1) << 32);
        o[4] = __longlong_as_double((long long)tag);
    }
};

struct YZPred {
    // buckets b = (ty + 1) + 3 (tz + 1), b != 4: local seed i, image (ty, tz) of the y / z periods
    const double* xyz;
    uint32_t n;
    double period, H;
    __device__ uint32_t mask(uint32_t i) const {
        const double y = xyz[3 * size_t(i) + 1], z = xyz[3 * size_t(i) + 2];
        // image t of coordinate v lies in [-H, period + H]
        const uint32_t ym = (y >= period - H ? 1u : 0u) | 2u | (y <= H ? 4u : 0u);   // ty = -1, 0, +1
        const uint32_t zm = (z >= period - H ? 1u : 0u) | 2u | (z <= H ? 4u : 0u);
        uint32_t m = 0;
        for(int tz = 0; tz < 3; ++tz) {
            for(int ty = 0; ty < 3; ++ty) {
                if(((ym >> ty) & 1u) && ((zm >> tz) & 1u) && !(ty == 1 && tz == 1)) {
                    m |= 1u << (ty + 3 * tz);
                }
            }
        }
        return m;
    }
};

struct YZEmit {
    double* xyz;        // local arrays, appended at base + pos
    double* w;          // or NULL
    uint32_t* lgid;
    uint8_t* ltc;
    uint32_t base;
    double period;
    __device__ void emit(uint32_t i, uint32_t b, uint32_t pos) const {
        const int ty = int(b % 3u) - 1, tz = int(b / 3u) - 1;
        const size_t o = size_t(base) + pos;
        xyz[3 * o] = xyz[3 * size_t(i)];
        xyz[3 * o + 1] = xyz[3 * size_t(i) + 1] + double(ty) * period;
        xyz[3 * o + 2] = xyz[3 * size_t(i) + 2] + double(tz) * period;
        if(w) {
            w[o] = w[i];
        }
        lgid[o] = lgid[i];
        ltc[o] = uint8_t((ltc[i] & 3u) | (uint32_t(ty + 1) << 2) | (uint32_t(tz + 1) << 4));
    }
};

template <class Pred>
__global__ void bucket_count_kernel(Pred P, uint32_t nbuckets, uint32_t nblocks, uint32_t* __restrict__ blkcnt) {
    __shared__ uint32_t cnt[32];
    if(threadIdx.x < 32) {
        cnt[threadIdx.x] = 0;
    }
    __syncthreads();
    const uint32_t base = blockIdx.x * kPackBlock;
    for(int k = 0; k < kPackItems; ++k) {
        const uint32_t i = base + k * kPackThreads + threadIdx.x;
        uint32_t m = i < P.n ? P.mask(i) : 0u;
        while(m) {
            const int b = __ffs(int(m)) - 1;
            m &= m - 1u;
            atomicAdd(&cnt[b], 1u);
        }
    }
    __syncthreads();
    if(threadIdx.x < nbuckets) {
        blkcnt[size_t(threadIdx.x) * nblocks + blockIdx.x] = cnt[threadIdx.x];
    }
}

// in-place exclusive scan of v[0..len) by ONE block (len = buckets x blocks: a few 10^4); v[len] = total
__global__ void scan_small_kernel(uint32_t* __restrict__ v, uint32_t len) {
    __shared__ uint32_t part[1024];
    const uint32_t T = blockDim.x, t = threadIdx.x;
    const uint32_t per = (len + T - 1) / T;
    const uint32_t b = min(len, t * per), e = min(len, b + per);
    uint32_t s = 0;
    for(uint32_t i = b; i < e; ++i) {
        s += v[i];
    }
    part[t] = s;
    __syncthreads();
    // Hillis-Steele over the per-thread sums
    for(uint32_t o = 1; o < T; o <<= 1) {
        uint32_t add = t >= o ? part[t - o] : 0u;
        __syncthreads();
        part[t] += add;
        __syncthreads();
    }
    uint32_t run = part[t] - s;
    for(uint32_t i = b; i < e; ++i) {
        const uint32_t x = v[i];
        v[i] = run;
        run += x;
    }
    if(t == T - 1) {
        v[len] = part[t];
    }
}

template <class Pred, class Emit>
__global__ void bucket_scatter_kernel(Pred P, Emit E, uint32_t nbuckets, uint32_t nblocks,
                                      const uint32_t* __restrict__ blkoff) {
    __shared__ uint32_t wcnt[kPackThreads / 32];
    __shared__ uint32_t run[32];
    __shared__ uint32_t allm;
    const uint32_t lane = threadIdx.x & 31u, wid = threadIdx.x >> 5;
    if(threadIdx.x < 32) {
        run[threadIdx.x] = threadIdx.x < nbuckets ? blkoff[size_t(threadIdx.x) * nblocks + blockIdx.x] : 0u;
    }
    __syncthreads();
    const uint32_t base = blockIdx.x * kPackBlock;
    for(int k = 0; k < kPackItems; ++k) {
        const uint32_t i = base + k * kPackThreads + threadIdx.x;
        const uint32_t m = i < P.n ? P.mask(i) : 0u;
        // union of the masks of the block (__syncthreads_or only tells whether any is non-zero)
        if(threadIdx.x == 0) {
            allm = 0u;
        }
        __syncthreads();
        const uint32_t wm = __reduce_or_sync(0xffffffffu, m);
        if(lane == 0 && wm) {
            atomicOr(&allm, wm);
        }
        __syncthreads();
        uint32_t all = allm;
        __syncthreads();   // (the next round resets allm)
        while(all) {
            const int b = __ffs(int(all)) - 1;
            all &= all - 1u;
            const bool in = (m >> b) & 1u;
            const uint32_t bal = __ballot_sync(0xffffffffu, in);
            if(lane == 0) {
                wcnt[wid] = __popc(bal);
            }
            __syncthreads();
            uint32_t before = 0, total = 0;
            for(uint32_t w = 0; w < kPackThreads / 32; ++w) {
                const uint32_t c = wcnt[w];
                before += w < wid ? c : 0u;
                total += c;
            }
            const uint32_t start = run[b];
            if(in) {
                E.emit(i, uint32_t(b), start + before + __popc(bal & ((1u << lane) - 1u)));
            }
            __syncthreads();
            if(threadIdx.x == 0) {
                run[b] = start + total;
            }
            __syncthreads();
        }
    }
}

// [min x, max x] of the seeds `xyz`, each taken at its image next to `center` (a slab of a periodic
// domain is an arc: seeds that Lloyd wrapped around the period stay next to their slab), + max weight,
// + number of own cells whose status meets `mask`: one partial (4 doubles) per block
__global__ void bounds_kernel(const double* __restrict__ xyz, const double* __restrict__ w,
                              const uint8_t* __restrict__ status, uint8_t mask, uint32_t n, int periodic,
                              double period, double center, double* __restrict__ part) {
    __shared__ double sm[4][8];
    double mn = 1e300, mx = -1e300, wm = -1e300, fl = 0.0;
    for(uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        double x = xyz[3 * size_t(i)];
        x += double(home_image(x, center, period, periodic)) * period;
        mn = fmin(mn, x);
        mx = fmax(mx, x);
        if(w) {
            wm = fmax(wm, w[i]);
        }
        if(status && (status[i] & mask)) {
            fl += 1.0;
        }
    }
    for(int o = 16; o > 0; o >>= 1) {
        mn = fmin(mn, __shfl_xor_sync(0xffffffffu, mn, o));
        mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        wm = fmax(wm, __shfl_xor_sync(0xffffffffu, wm, o));
        fl += __shfl_xor_sync(0xffffffffu, fl, o);
    }
    if((threadIdx.x & 31) == 0) {
        sm[0][threadIdx.x >> 5] = mn;
        sm[1][threadIdx.x >> 5] = mx;
        sm[2][threadIdx.x >> 5] = wm;
        sm[3][threadIdx.x >> 5] = fl;
    }
    __syncthreads();
    if(threadIdx.x == 0) {
        for(int k = 1; k < 8; ++k) {
            mn = fmin(mn, sm[0][k]);
            mx = fmax(mx, sm[1][k]);
            wm = fmax(wm, sm[2][k]);
            fl += sm[3][k];
        }
        part[4 * blockIdx.x] = mn;
        part[4 * blockIdx.x + 1] = mx;
        part[4 * blockIdx.x + 2] = wm;
        part[4 * blockIdx.x + 3] = fl;
    }
}

// circular mean of the own x: partial sums of (cos, sin)(2 pi x / period), fixed order
__global__ void circular_mean_kernel(const double* __restrict__ xyz, uint32_t n, double period,
                                     double* __restrict__ part) {
    __shared__ double sm[2][8];
    double c = 0.0, sn = 0.0;
    for(uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        double sv, cv;
        sincospi(2.0 * xyz[3 * size_t(i)] / period, &sv, &cv);
        c += cv;
        sn += sv;
    }
    for(int o = 16; o > 0; o >>= 1) {
        c += __shfl_xor_sync(0xffffffffu, c, o);
        sn += __shfl_xor_sync(0xffffffffu, sn, o);
    }
    if((threadIdx.x & 31) == 0) {
        sm[0][threadIdx.x >> 5] = c;
        sm[1][threadIdx.x >> 5] = sn;
    }
    __syncthreads();
    if(threadIdx.x == 0) {
        for(int k = 1; k < 8; ++k) {
            c += sm[0][k];
            sn += sm[1][k];
        }
        part[2 * blockIdx.x] = c;
        part[2 * blockIdx.x + 1] = sn;
    }
}

// partials -> out[0..3] = min, max, wmax, flagged (one thread: a few hundred values)
__global__ void bounds_final_kernel(const double* __restrict__ part, int nb, double* __restrict__ out) {
    double mn = 1e300, mx = -1e300, wm = -1e300, fl = 0.0;
    for(int k = 0; k < nb; ++k) {
        mn = fmin(mn, part[4 * k]);
        mx = fmax(mx, part[4 * k + 1]);
        wm = fmax(wm, part[4 * k + 2]);
        fl += part[4 * k + 3];
    }
    out[0] = mn;
    out[1] = mx;
    out[2] = wm;
    out[3] = fl;
}

// local arrays: rows [0, n_own) = the own seeds (at their image next to the slab centre), rows
// [n_own, n_own + n_ghost) = received ghosts
__global__ void assemble_kernel(uint32_t n_own, uint32_t n_ghost, const double* __restrict__ own,
                                const double* __restrict__ own_w, const double* __restrict__ ghosts,
                                uint64_t first, int periodic, double period, double center,
                                double* __restrict__ xyz, double* __restrict__ w,
                                uint32_t* __restrict__ lgid, uint8_t* __restrict__ ltc) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if(i < n_own) {
        const double x = own[3 * size_t(i)];
        const int s0 = home_image(x, center, period, periodic);
        xyz[3 * size_t(i)] = x + double(s0) * period;
        xyz[3 * size_t(i) + 1] = own[3 * size_t(i) + 1];
        xyz[3 * size_t(i) + 2] = own[3 * size_t(i) + 2];
        if(w) {
            w[i] = own_w[i];
        }
        lgid[i] = uint32_t(first + i);
        ltc[i] = uint8_t(uint32_t(s0 + 1) | (1u << 2) | (1u << 4));
    } else if(i < n_own + n_ghost) {
        const double* g = ghosts + size_t(i - n_own) * kGhostDoubles;
        xyz[3 * size_t(i)] = g[0];
        xyz[3 * size_t(i) + 1] = g[1];
        xyz[3 * size_t(i) + 2] = g[2];
        if(w) {
            w[i] = g[3];
        }
        const unsigned long long tag = (unsigned long long)__double_as_longlong(g[4]);
        lgid[i] = uint32_t(tag & 0xffffffffull);
        ltc[i] = uint8_t(uint32_t(tag >> 32) | (1u << 2) | (1u << 4));
    }
}

// face lists of the own cells: local neighbour index -> global index + n_total * instance
// (delaunay/periodic.h:95-113), rows re-sorted by the new ids (one thread per row, <= 64 faces)
__global__ void remap_faces_kernel(uint32_t n_own, uint32_t fcap, const uint32_t* __restrict__ nfaces,
                                   uint32_t* __restrict__ face_nbr, double* __restrict__ face_area,
                                   const uint32_t* __restrict__ lgid, const uint8_t* __restrict__ ltc,
                                   uint32_t n_total, int periodic) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if(i >= n_own) {
        return;
    }
    const uint32_t nf = min(nfaces[i], fcap);
    uint32_t* ids = face_nbr + size_t(i) * fcap;
    double* ar = face_area + size_t(i) * fcap;
    for(uint32_t k = 0; k < nf; ++k) {
        const uint32_t j = ids[k];
        if(j < B200VOR_DOMAIN_ID) {
            uint32_t inst = 0;
            if(periodic) {
                // translation of the neighbour RELATIVE to the image the cell's own seed is kept at
                const uint32_t t = ltc[j];
                const int tx = clampi(int(t & 3u) - int(ltc[i] & 3u), -1, 1);
                inst = translation_to_instance(tx, int((t >> 2) & 3u) - 1, int((t >> 4) & 3u) - 1);
            }
            ids[k] = lgid[j] + inst * n_total;
        }
    }
    for(uint32_t k = 1; k < nf; ++k) {   // insertion sort (rows are short and nearly sorted)
        const uint32_t id = ids[k];
        const double a = ar[k];
        uint32_t q = k;
        while(q > 0 && ids[q - 1] > id) {
            ids[q] = ids[q - 1];
            ar[q] = ar[q - 1];
            --q;
        }
        ids[q] = id;
        ar[q] = a;
    }
}

// m g of a cell is computed at the image its own seed is kept at: back to the caller's frame
__global__ void mg_frame_kernel(uint32_t n_own, const uint8_t* __restrict__ ltc, const double* __restrict__ m,
                                double period, double* __restrict__ mg) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if(i < n_own) {
        const int s0 = int(ltc[i] & 3u) - 1;
        if(s0 != 0) {
            mg[3 * size_t(i)] -= double(s0) * period * m[i];
        }
    }
}

__global__ void lloyd_own_kernel(uint32_t n_own, const double* __restrict__ own, const double* __restrict__ m,
                                 const double* __restrict__ mg, int periodic, double period,
                                 double* __restrict__ next) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if(i >= n_own) {
        return;
    }
    const double mm = m[i];
    for(int a = 0; a < 3; ++a) {
        double x = own[3 * size_t(i) + a];
        if(mm > 1e-30) {   // voronoi/CVT.cpp:154-163
            const double s = 1.0 / mm;
            x = s * mg[3 * size_t(i) + a];
            if(periodic) {   // demo_Delaunay3d/main.cpp:443-451
                if(x < 0.0) {
                    x += period;
                }
                if(x > period) {
                    x -= period;
                }
            }
        }
        next[3 * size_t(i) + a] = x;
    }
}

// development trace (B200VOR_SLAB_DEBUG=1): stage markers with a stream sync each
#define SLAB_TRACE(c, ...)                                           \
    do {                                                             \
        if(getenv("B200VOR_SLAB_DEBUG")) {                           \
            cudaError_t e__ = cudaStreamSynchronize((c)->stream);    \
            fprintf(stderr, "[slab] " __VA_ARGS__);                  \
            fprintf(stderr, " (%s)\n", cudaGetErrorString(e__));     \
            fflush(stderr);                                          \
        }                                                            \
    } while(0)

float elapsed(cudaEvent_t a, cudaEvent_t b) {
    float ms = 0.f;
    cudaEventElapsedTime(&ms, a, b);
    return ms;
}

// bounds of the seeds `xyz` (+ flag count of the last pass, + max weight) of every rank -> every rank
int exchange_bounds(Ctx* c, const double* xyz, const uint8_t* status, uint64_t* flagged_total) {
    Slab* S = c->slab;
    const int nb = 296;
    int rc;
    if((rc = S->red.reserve(4 * nb + 8)) || (rc = S->gather_dev.reserve(4 * size_t(S->world) + 8))) {
        return rc;
    }
    bounds_kernel<<<nb, 256, 0, c->stream>>>(xyz, S->has_w ? S->own_w.p : nullptr, status,
                                             uint8_t(B200VOR_CELL_UNCERTIFIED | B200VOR_CELL_BOXFACE), S->n_own,
                                             S->dom_periodic, S->dom_period, S->center, S->red.p);
    double* mine = S->gather_dev.p + 4 * size_t(S->world);
    bounds_final_kernel<<<1, 1, 0, c->stream>>>(S->red.p, nb, mine);
    c->stats.gpu_launches += 2;
    B200_CUDA_CHECK(cudaGetLastError());
    std::vector<double> host(4 * size_t(S->world));
    if(S->world > 1) {
        B200_NCCL_CHECK(g_nccl.AllGather(mine, S->gather_dev.p, 4, ncclDouble, S->comm, c->stream));
        B200_CUDA_CHECK(cudaMemcpyAsync(host.data(), S->gather_dev.p, host.size() * sizeof(double),
                                        cudaMemcpyDeviceToHost, c->stream));
    } else {
        B200_CUDA_CHECK(cudaMemcpyAsync(host.data(), mine, 4 * sizeof(double), cudaMemcpyDeviceToHost, c->stream));
    }
    B200_CUDA_CHECK(cudaStreamSynchronize(c->stream));
    S->lo.assign(S->world, 0.0);
    S->hi.assign(S->world, 0.0);
    uint64_t fl = 0;
    double wm = -1e300;
    for(int p = 0; p < S->world; ++p) {
        S->lo[p] = host[4 * p];
        S->hi[p] = host[4 * p + 1];
        wm = fmax(wm, host[4 * p + 2]);
        fl += uint64_t(host[4 * p + 3]);
    }
    S->wmax_global = wm;
    if(flagged_total) {
        *flagged_total = fl;
    }
    return B200VOR_OK;
}

// halo exchange at width S->H, local arrays, slab-local grid
int build_local(Ctx* c) {
    Slab* S = c->slab;
    cudaStream_t st = c->stream;
    const int W = S->world;
    const uint32_t n_own = S->n_own;
    int rc;
    cudaEventRecord(S->ev[0], st);
    // ---- phase X: own seeds wanted by the peers ----
    const uint32_t nbx = 3u * uint32_t(W);
    const uint32_t nblk = (n_own + kPackBlock - 1) / kPackBlock;
    if((rc = S->blkcnt.reserve(size_t(std::max(nbx, 9u)) * (size_t(nblk) + 8) * 2 + 64))) {
        return rc;
    }
    XPred xp;
    xp.own = S->own.p;
    xp.n = n_own;
    xp.world = W;
    xp.rank = S->rank;
    xp.periodic = S->dom_periodic;
    xp.period = S->dom_periodic ? S->dom_period : 0.0;
    xp.H = S->H;
    xp.center = S->center;
    for(int p = 0; p < W; ++p) {
        xp.lo[p] = S->lo[p];
        xp.hi[p] = S->hi[p];
    }
    bucket_count_kernel<XPred><<<nblk, kPackThreads, 0, st>>>(xp, nbx, nblk, S->blkcnt.p);
    scan_small_kernel<<<1, 1024, 0, st>>>(S->blkcnt.p, nbx * nblk);
    c->stats.gpu_launches += 2;
    // send offsets of the peers = offsets of the buckets 3p at block 0; total behind the last entry
    std::vector<uint32_t> off(size_t(W) + 1);
    for(int p = 0; p < W; ++p) {
        B200_CUDA_CHECK(cudaMemcpyAsync(&off[p], S->blkcnt.p + size_t(3 * p) * nblk, sizeof(uint32_t),
                                        cudaMemcpyDeviceToHost, st));
    }
    B200_CUDA_CHECK(cudaMemcpyAsync(&off[W], S->blkcnt.p + size_t(nbx) * nblk, sizeof(uint32_t),
                                    cudaMemcpyDeviceToHost, st));
    B200_CUDA_CHECK(cudaStreamSynchronize(st));
    const uint32_t nsend = off[W];
    SLAB_TRACE(c, "x-pack: %u to send, H %.4f", nsend, S->H);
    if((rc = S->sendbuf.reserve_growing(size_t(nsend) * kGhostDoubles + 8))) {   // (ghost counts grow as the seeds relax)
        return rc;
    }
    XEmit xe;
    xe.own = S->own.p;
    xe.own_w = S->has_w ? S->own_w.p : nullptr;
    xe.first = S->first;
    xe.period = xp.period;
    xe.out = S->sendbuf.p;
    if(nsend) {
        bucket_scatter_kernel<XPred, XEmit><<<nblk, kPackThreads, 0, st>>>(xp, xe, nbx, nblk, S->blkcnt.p);
        c->stats.gpu_launches++;
    }
    // ---- counts: row `rank` of the world x world matrix ----
    std::vector<uint32_t> mat(size_t(W) * W, 0u), rowv(W);
    for(int p = 0; p < W; ++p) {
        rowv[p] = off[p + 1] - off[p];
    }
    if(W > 1) {
        if((rc = S->count_dev.reserve(size_t(W) * W + W + 8))) {
            return rc;
        }
        uint32_t* mine = S->count_dev.p + size_t(W) * W;
        B200_CUDA_CHECK(cudaMemcpyAsync(mine, rowv.data(), W * sizeof(uint32_t), cudaMemcpyHostToDevice, st));
        B200_NCCL_CHECK(g_nccl.AllGather(mine, S->count_dev.p, W, ncclUint32, S->comm, st));
        B200_CUDA_CHECK(cudaMemcpyAsync(mat.data(), S->count_dev.p, mat.size() * sizeof(uint32_t),
                                        cudaMemcpyDeviceToHost, st));
        B200_CUDA_CHECK(cudaStreamSynchronize(st));
    } else {
        mat[0] = rowv[0];
    }
    std::vector<uint32_t> roff(size_t(W) + 1, 0u);
    for(int p = 0; p < W; ++p) {
        roff[p + 1] = roff[p] + mat[size_t(p) * W + S->rank];   // what peer p sends to me
    }
    const uint32_t nrecv = roff[W];
    if((rc = S->recvbuf.reserve_growing(size_t(nrecv) * kGhostDoubles + 8))) {
        return rc;
    }
    cudaEventRecord(S->ev[1], st);
    // ---- ghosts, peer to peer (the images a rank wants of its own seeds: device copy) ----
    if(W > 1) {
        B200_NCCL_CHECK(g_nccl.GroupStart());
        for(int p = 0; p < W; ++p) {
            if(p == S->rank) {
                continue;
            }
            if(rowv[p]) {
                B200_NCCL_CHECK(g_nccl.Send(S->sendbuf.p + size_t(off[p]) * kGhostDoubles,
                                            size_t(rowv[p]) * kGhostDoubles, ncclDouble, p, S->comm, st));
                S->st.bytes_sent += uint64_t(rowv[p]) * kGhostDoubles * sizeof(double);
            }
            const uint32_t cnt = roff[p + 1] - roff[p];
            if(cnt) {
                B200_NCCL_CHECK(g_nccl.Recv(S->recvbuf.p + size_t(roff[p]) * kGhostDoubles,
                                            size_t(cnt) * kGhostDoubles, ncclDouble, p, S->comm, st));
                S->st.bytes_received += uint64_t(cnt) * kGhostDoubles * sizeof(double);
            }
        }
        B200_NCCL_CHECK(g_nccl.GroupEnd());
    }
    if(rowv[S->rank]) {
        B200_CUDA_CHECK(cudaMemcpyAsync(S->recvbuf.p + size_t(roff[S->rank]) * kGhostDoubles,
                                        S->sendbuf.p + size_t(off[S->rank]) * kGhostDoubles,
                                        size_t(rowv[S->rank]) * kGhostDoubles * sizeof(double),
                                        cudaMemcpyDeviceToDevice, st));
    }
    cudaEventRecord(S->ev[2], st);
    SLAB_TRACE(c, "exchange done: %u received", nrecv);
    // ---- local arrays; periodic domain: y / z images ----
    const uint32_t nx = n_own + nrecv;
    // upper bound of the y / z images, refined below: reserve generously, then exactly
    uint32_t nyz = 0;
    const uint32_t nblk2 = (nx + kPackBlock - 1) / kPackBlock;
    // room for the y / z images as well (as many as last time + 15 %: the ghost counts drift by a per
    // cent per Lloyd step, and growing a GB-sized buffer in the middle of a run costs 100+ ms)
    const size_t want = size_t(nx) + size_t(S->st.n_ghost_yz) + (size_t(nx) + S->st.n_ghost_yz) / 7 + 4096;
    if(3 * want <= c->xyz.cap && want <= S->lgid.cap && want <= S->ltc.cap && (!S->has_w || want <= c->wgt.cap)) {
        // fits
    } else if((rc = c->xyz.reserve(3 * (want + want / 8))) || (rc = S->lgid.reserve(want + want / 8)) ||
              (rc = S->ltc.reserve(want + want / 8)) || (S->has_w && (rc = c->wgt.reserve(want + want / 8)))) {
        return rc;
    }
    assemble_kernel<<<(nx + 255) / 256, 256, 0, st>>>(n_own, nrecv, S->own.p, S->has_w ? S->own_w.p : nullptr,
                                                      S->recvbuf.p, S->first, S->dom_periodic, S->dom_period,
                                                      S->center, c->xyz.p,
                                                      S->has_w ? c->wgt.p : nullptr, S->lgid.p, S->ltc.p);
    c->stats.gpu_launches++;
    if(S->dom_periodic) {
        if((rc = S->blkcnt.reserve(9 * (size_t(nblk2) + 8) + 64))) {
            return rc;
        }
        YZPred yp;
        yp.xyz = c->xyz.p;
        yp.n = nx;
        yp.period = S->dom_period;
        yp.H = S->H;
        bucket_count_kernel<YZPred><<<nblk2, kPackThreads, 0, st>>>(yp, 9, nblk2, S->blkcnt.p);
        scan_small_kernel<<<1, 1024, 0, st>>>(S->blkcnt.p, 9 * nblk2);
        c->stats.gpu_launches += 2;
        B200_CUDA_CHECK(cudaMemcpyAsync(&nyz, S->blkcnt.p + size_t(9) * nblk2, sizeof(uint32_t),
                                        cudaMemcpyDeviceToHost, st));
        B200_CUDA_CHECK(cudaStreamSynchronize(st));
        SLAB_TRACE(c, "yz images: %u", nyz);
        if(nyz) {
            // grow the local arrays, keeping their contents
            const size_t nl = size_t(nx) + nyz;
            if(c->xyz.cap < 3 * nl || S->lgid.cap < nl || S->ltc.cap < nl || (S->has_w && c->wgt.cap < nl)) {
                DevBuf<double> nxyz, nw;
                DevBuf<uint32_t> ng;
                DevBuf<uint8_t> nt;
                if((rc = nxyz.reserve(3 * nl)) || (rc = ng.reserve(nl)) || (rc = nt.reserve(nl)) ||
                   (S->has_w && (rc = nw.reserve(nl)))) {
                    return rc;
                }
                B200_CUDA_CHECK(cudaMemcpyAsync(nxyz.p, c->xyz.p, 3 * size_t(nx) * sizeof(double), cudaMemcpyDeviceToDevice, st));
                B200_CUDA_CHECK(cudaMemcpyAsync(ng.p, S->lgid.p, size_t(nx) * sizeof(uint32_t), cudaMemcpyDeviceToDevice, st));
                B200_CUDA_CHECK(cudaMemcpyAsync(nt.p, S->ltc.p, size_t(nx), cudaMemcpyDeviceToDevice, st));
                if(S->has_w) {
                    B200_CUDA_CHECK(cudaMemcpyAsync(nw.p, c->wgt.p, size_t(nx) * sizeof(double), cudaMemcpyDeviceToDevice, st));
                }
                B200_CUDA_CHECK(cudaStreamSynchronize(st));
                c->xyz.release();
                S->lgid.release();
                S->ltc.release();
                c->xyz = nxyz;
                S->lgid = ng;
                S->ltc = nt;
                if(S->has_w) {
                    c->wgt.release();
                    c->wgt = nw;
                }
                yp.xyz = c->xyz.p;
            }
            YZEmit ye;
            ye.xyz = c->xyz.p;
            ye.w = S->has_w ? c->wgt.p : nullptr;
            ye.lgid = S->lgid.p;
            ye.ltc = S->ltc.p;
            ye.base = nx;
            ye.period = S->dom_period;
            bucket_scatter_kernel<YZPred, YZEmit><<<nblk2, kPackThreads, 0, st>>>(yp, ye, 9, nblk2, S->blkcnt.p);
            c->stats.gpu_launches++;
        }
    }
    B200_CUDA_CHECK(cudaGetLastError());
    cudaEventRecord(S->ev[3], st);
    // ---- the slab-local context ----
    const uint32_t nl = nx + nyz;
    c->n = nl;
    c->periodic = 0;
    c->has_w = S->has_w;
    c->wmax = S->wmax_global;
    c->q0 = 0;
    c->q1 = n_own;
    const double xl = S->lo[S->rank] - S->H, xh = S->hi[S->rank] + S->H;
    if(S->dom_periodic) {
        c->box[0] = xl;
        c->box[3] = xh;
        c->box[1] = c->box[2] = -S->H;
        c->box[4] = c->box[5] = S->dom_period + S->H;
        c->open = 0x3F;
    } else {
        for(int a = 0; a < 6; ++a) {
            c->box[a] = S->dom_box[a];
        }
        c->open = 0;
        // a face of the slab is open unless the halo reaches the domain face behind it
        if(xl > S->dom_box[0]) {
            c->box[0] = xl;
            c->open |= 1;
        }
        if(xh < S->dom_box[3]) {
            c->box[3] = xh;
            c->open |= 2;
        }
    }
    c->grid_dirty = true;
    c->cells_valid = c->faces_valid = false;
    S->st.n_own = n_own;
    S->st.n_ghost_x = nrecv;
    S->st.n_ghost_yz = nyz;
    S->st.n_local = nl;
    S->st.halo = S->H;
    S->st.slab_lo = S->lo[S->rank];
    S->st.slab_hi = S->hi[S->rank];
    SLAB_TRACE(c, "local arrays: %u seeds, open %d", nl, c->open);
    c->stats.ms_grid = 0.f;
    if((rc = ctx_build_grid_timed(c))) {
        return rc;
    }
    SLAB_TRACE(c, "grid %d x %d x %d, two-level %d", c->grid.g[0], c->grid.g[1], c->grid.g[2], c->grid.two_level);
    S->st.ms_pack += elapsed(S->ev[0], S->ev[1]) + elapsed(S->ev[2], S->ev[3]);
    S->st.ms_exchange += elapsed(S->ev[1], S->ev[2]);
    S->st.ms_grid += c->stats.ms_grid;
    return B200VOR_OK;
}

double initial_halo(const Slab* S, double ppc) {
    // halo_bins bins of the grid a single GPU would build over the whole domain
    double vol = 1.0;
    if(S->dom_periodic) {
        vol = S->dom_period * S->dom_period * S->dom_period;
    } else {
        for(int a = 0; a < 3; ++a) {
            vol *= fmax(S->dom_box[3 + a] - S->dom_box[a], 1e-300);
        }
    }
    const double bin = cbrt(vol * ppc / double(std::max<uint64_t>(S->n_total, 1)));
    return S->halo_bins * bin;
}

double max_halo(const Slab* S) {
    if(S->dom_periodic) {
        return 0.49 * S->dom_period;   // an image must stay distinguishable from the seed itself
    }
    double ext = 0.0;
    for(int a = 0; a < 3; ++a) {
        ext = fmax(ext, S->dom_box[3 + a] - S->dom_box[a]);
    }
    return ext;
}

// One cell pass over the own seeds; doubles the halo while any rank has a cell that left it.
// lloyd: the own seeds move to their centroids (CVT.cpp:154-163).  The bounds of the moved seeds and
// the flag counts of the pass travel in ONE all-gather: if no rank flagged a cell the move is
// committed and the bounds already serve the next pass, else the move is dropped and the pass
// repeated (same seeds, same bounds, wider halo).
int slab_pass(Ctx* c, const b200vor_cell_params* params, bool lloyd) {
    Slab* S = c->slab;
    int rc;
    if(!S->bounds_valid) {
        if((rc = exchange_bounds(c, S->own.p, nullptr, nullptr))) {
            return rc;
        }
        S->bounds_valid = true;
    }
    S->st.retries = 0;
    for(;;) {
        if((rc = build_local(c))) {
            return rc;
        }
        c->stats.ms_knn = c->stats.ms_clip = 0.f;
        if((rc = ctx_cells(c, params, false))) {
            return rc;
        }
        S->st.ms_clip += c->stats.ms_clip;
        if(S->dom_periodic) {
            mg_frame_kernel<<<(S->n_own + 255) / 256, 256, 0, c->stream>>>(S->n_own, S->ltc.p, c->m.p, S->dom_period,
                                                                           c->mg.p);
            c->stats.gpu_launches++;
        }
        SLAB_TRACE(c, "cells done");
        if(lloyd) {
            lloyd_own_kernel<<<(S->n_own + 255) / 256, 256, 0, c->stream>>>(
                S->n_own, S->own.p, c->m.p, c->mg.p, S->dom_periodic, S->dom_period, S->next.p);
            c->stats.gpu_launches++;
            B200_CUDA_CHECK(cudaGetLastError());
        }
        const std::vector<double> lo0 = S->lo, hi0 = S->hi;
        uint64_t flagged = 0;
        if((rc = exchange_bounds(c, lloyd ? S->next.p : S->own.p, c->status.p, &flagged))) {
            return rc;
        }
        S->st.n_flagged = uint32_t(std::min<uint64_t>(flagged, 0xffffffffull));
        SLAB_TRACE(c, "flagged %llu", (unsigned long long)flagged);
        if(flagged == 0 || S->H >= max_halo(S)) {
            if(lloyd) {
                std::swap(S->own.p, S->next.p);
                std::swap(S->own.cap, S->next.cap);
            }
            break;
        }
        S->lo = lo0;
        S->hi = hi0;
        S->H = fmin(2.0 * S->H, max_halo(S));   // re-queue: every rank repeats the pass with a wider halo
        S->st.retries++;
    }
    if(params && params->want_faces) {
        remap_faces_kernel<<<(S->n_own + 127) / 128, 128, 0, c->stream>>>(
            S->n_own, c->fcap, c->nfaces.p, c->face_nbr.p, c->face_area.p, S->lgid.p, S->ltc.p,
            uint32_t(S->n_total), S->dom_periodic);
        c->stats.gpu_launches++;
        B200_CUDA_CHECK(cudaGetLastError());
    }
    return B200VOR_OK;
}

void reset_step_stats(Slab* S) {
    S->st.ms_exchange = S->st.ms_pack = S->st.ms_grid = S->st.ms_clip = 0.f;
    S->st.bytes_sent = S->st.bytes_received = 0;
    S->alloc0 = devbuf_alloc_count();
}

}  // namespace

bool slab_active(const Ctx* c) {
    return c->slab && c->slab->active;
}

uint32_t out_rows(const Ctx* c) {
    return (c->slab && c->slab->active) ? c->slab->n_own : c->n;
}

void slab_deactivate(Ctx* c) {
    Slab* S = c->slab;
    if(!S || !S->active) {
        return;
    }
    S->active = false;
    c->open = 0;
    c->periodic = S->dom_periodic;
    c->period = S->dom_period;
    memcpy(c->box, S->dom_box, sizeof(S->dom_box));
    S->dom_captured = false;
    c->q0 = 0;
    c->q1 = 0xFFFFFFFFu;
}

void slab_release(Ctx* c) {
    Slab* S = c->slab;
    if(!S) {
        return;
    }
    if(S->has_comm && g_nccl.lib) {
        g_nccl.CommDestroy(S->comm);
    }
    S->own.release(); S->own_w.release(); S->next.release(); S->sendbuf.release(); S->recvbuf.release();
    S->blkcnt.release(); S->gather_dev.release(); S->count_dev.release(); S->lgid.release(); S->ltc.release();
    S->red.release();
    for(cudaEvent_t e : S->ev) {
        if(e) {
            cudaEventDestroy(e);
        }
    }
    delete S;
    c->slab = nullptr;
}

}  // namespace b200vor

using namespace b200vor;

extern "C" {

int b200vor_comm_unique_id(void* id) {
    if(!id) {
        set_error("b200vor_comm_unique_id: null argument");
        return B200VOR_ERR_ARG;
    }
    int rc = load_nccl();
    if(rc) {
        return rc;
    }
    static_assert(sizeof(ncclUniqueId) == B200VOR_UNIQUE_ID_BYTES, "ncclUniqueId size");
    ncclUniqueId u;
    B200_NCCL_CHECK(g_nccl.GetUniqueId(&u));
    memcpy(id, &u, sizeof(u));
    return B200VOR_OK;
}

static Slab* slab_of(Ctx* c) {
    if(!c->slab) {
        c->slab = new Slab();
        for(cudaEvent_t& e : c->slab->ev) {
            cudaEventCreate(&e);
        }
    }
    return c->slab;
}

int b200vor_comm_init(b200vor_ctx* h, const void* id, int rank, int world) {
    if(!h || world < 1 || world > kMaxWorld || rank < 0 || rank >= world || (world > 1 && !id)) {
        set_error("b200vor_comm_init: invalid argument (rank %d of %d, at most %d ranks)", rank, world, kMaxWorld);
        return B200VOR_ERR_ARG;
    }
    Ctx* c = &h->c;
    B200_CUDA_CHECK(cudaSetDevice(c->device));
    Slab* S = slab_of(c);
    if(S->has_comm) {
        set_error("b200vor_comm_init: the context already has a communicator");
        return B200VOR_ERR_ARG;
    }
    S->rank = rank;
    S->world = world;
    if(world > 1) {
        int rc = load_nccl();
        if(rc) {
            return rc;
        }
        ncclUniqueId u;
        memcpy(&u, id, sizeof(u));
        B200_NCCL_CHECK(g_nccl.CommInitRank(&S->comm, world, u, rank));
        S->has_comm = true;
    }
    return B200VOR_OK;
}

int b200vor_slab_set_halo_bins(b200vor_ctx* h, double bins) {
    if(!h || !(bins > 0.0)) {
        set_error("b200vor_slab_set_halo_bins: invalid argument");
        return B200VOR_ERR_ARG;
    }
    slab_of(&h->c)->halo_bins = bins;
    return B200VOR_OK;
}

int b200vor_slab_set_points(b200vor_ctx* h, uint32_t n_own, uint64_t first_global, uint64_t n_total,
                            const double* xyz, const double* w, int on_device) {
    if(!h || !xyz || n_own == 0 || first_global + n_own > n_total) {
        set_error("b200vor_slab_set_points: invalid argument");
        return B200VOR_ERR_ARG;
    }
    Ctx* c = &h->c;
    B200_CUDA_CHECK(cudaSetDevice(c->device));
    Slab* S = slab_of(c);
    if(S->world > 1 && !S->has_comm) {
        set_error("b200vor_slab_set_points: call b200vor_comm_init first");
        return B200VOR_ERR_ARG;
    }
    if(!S->dom_captured) {
        // the caller's domain; from here on the context itself describes the slab-local box
        S->dom_periodic = c->periodic;
        S->dom_period = c->period;
        memcpy(S->dom_box, c->box, sizeof(S->dom_box));
        S->dom_captured = true;
    }
    if(n_total * (S->dom_periodic ? 27ull : 1ull) >= 0xFFFFFFF0ull) {
        set_error("b200vor_slab_set_points: n_total too large for 32-bit neighbour ids");
        return B200VOR_ERR_ARG;
    }
    int rc;
    if((rc = S->own.reserve(3 * size_t(n_own))) || (rc = S->next.reserve(3 * size_t(n_own))) ||
       (w && (rc = S->own_w.reserve(n_own)))) {
        return rc;
    }
    const cudaMemcpyKind kind = on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
    B200_CUDA_CHECK(cudaMemcpyAsync(S->own.p, xyz, 3 * size_t(n_own) * sizeof(double), kind, c->stream));
    if(w) {
        B200_CUDA_CHECK(cudaMemcpyAsync(S->own_w.p, w, size_t(n_own) * sizeof(double), kind, c->stream));
    }
    S->has_w = w != nullptr;
    S->n_own = n_own;
    S->first = first_global;
    S->n_total = n_total;
    S->H = initial_halo(S, c->ppc);
    S->center = 0.0;
    if(S->dom_periodic) {
        // centre of the arc the own seeds occupy on the period: their circular mean (kept for the run)
        const int nb = 296;
        if((rc = S->red.reserve(4 * nb + 8))) {
            return rc;
        }
        circular_mean_kernel<<<nb, 256, 0, c->stream>>>(S->own.p, n_own, S->dom_period, S->red.p);
        c->stats.gpu_launches++;
        double part[2 * nb];
        B200_CUDA_CHECK(cudaMemcpyAsync(part, S->red.p, sizeof(part), cudaMemcpyDeviceToHost, c->stream));
        B200_CUDA_CHECK(cudaStreamSynchronize(c->stream));
        double cs = 0.0, sn = 0.0;
        for(int b = 0; b < nb; ++b) {
            cs += part[2 * b];
            sn += part[2 * b + 1];
        }
        double a = atan2(sn, cs) / (2.0 * M_PI);   // in (-1/2, 1/2]
        if(a < 0.0) {
            a += 1.0;
        }
        S->center = a * S->dom_period;
    }
    S->active = true;
    S->bounds_valid = false;
    c->cells_valid = c->faces_valid = false;
    return B200VOR_OK;
}

int b200vor_slab_compute_cells(b200vor_ctx* h, const b200vor_cell_params* params) {
    if(!h || !h->c.slab || !h->c.slab->active) {
        set_error("b200vor_slab_compute_cells: no slab points (b200vor_slab_set_points)");
        return B200VOR_ERR_ARG;
    }
    Ctx* c = &h->c;
    B200_CUDA_CHECK(cudaSetDevice(c->device));
    reset_step_stats(c->slab);
    b200vor_cell_params p = {0, 0, 0, 0, 0};
    if(params) {
        p = *params;
    }
    p.mode = 3;   // team kernel, then its large-cell instance; what they flag stays flagged
    int rc = slab_pass(c, &p, false);
    if(rc) {
        return rc;
    }
    B200_CUDA_CHECK(cudaStreamSynchronize(c->stream));
    return B200VOR_OK;
}

int b200vor_slab_lloyd(b200vor_ctx* h, uint32_t iters, const b200vor_cell_params* params, double* xyz_out_own,
                       int on_device) {
    if(!h || !h->c.slab || !h->c.slab->active) {
        set_error("b200vor_slab_lloyd: no slab points (b200vor_slab_set_points)");
        return B200VOR_ERR_ARG;
    }
    Ctx* c = &h->c;
    Slab* S = c->slab;
    B200_CUDA_CHECK(cudaSetDevice(c->device));
    reset_step_stats(S);
    b200vor_cell_params p = {0, 0, 0, 0, 0};
    if(params) {
        p = *params;
    }
    p.mode = 3;
    p.want_faces = 0;
    p.want_energy = 0;
    int rc;
    for(uint32_t it = 0; it < iters; ++it) {
        if((rc = slab_pass(c, &p, true))) {
            return rc;
        }
        c->cells_valid = false;
    }
    if(xyz_out_own) {
        B200_CUDA_CHECK(cudaMemcpyAsync(xyz_out_own, S->own.p, 3 * size_t(S->n_own) * sizeof(double),
                                        on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, c->stream));
    }
    B200_CUDA_CHECK(cudaStreamSynchronize(c->stream));
    return B200VOR_OK;
}

int b200vor_slab_get_stats(b200vor_ctx* h, b200vor_slab_stats* out) {
    if(!h || !out || !h->c.slab) {
        set_error("b200vor_slab_get_stats: no slab state");
        return B200VOR_ERR_ARG;
    }
    *out = h->c.slab->st;
    out->reallocs = devbuf_alloc_count() - h->c.slab->alloc0;
    return B200VOR_OK;
}

}  // extern "C"
