// tpc_cell.cuh -- one Voronoi cell by ONE THREAD: the body of cells_tpc_kernel (tpc.cu).
//
// Same path as the team kernel (fused_kernel.inc): neighbour walk on the counting-sort grid in
// distance shells, half-space clipping with the security-radius rule, fp64 integration.  It
// replaces, per seed,
//   KdTree::get_nearest_neighbors (neighbour provider)            points/kd_tree.cpp:156-387
//   clip_by_cell_SR (stop rule d2 > 4.1 R2, neighbourhood growth)  voronoi/generic_RVD.h:2278-2347
//   VBW::ConvexCell::init_with_box / clip_by_plane                 voronoi/convex_cell.cpp:175-215, :467-516
//     triangle_is_in_conflict / triangulate_conflict_zone          voronoi/convex_cell.cpp:870-969, :761-865
//   compute_triangle_point / squared_radius / compute_mg           voronoi/convex_cell.cpp:971-1082, :1312-1330, :1267-1307
// but is organised for 32 independent cells per warp:
//  * the cell is stored in the thread's COLUMN of shared memory as 16-byte records
//    (x, y, z in fp32 | three 6-bit plane ids | an 8-bit exponent bounding the position error);
//    planes are not stored, only the neighbour id per plane slot (the plane is re-derived in fp64
//    from the seed pair whenever a vertex is created);
//  * sign tests run in fp32 with a rigorous forward error bound (plane rounding + vertex
//    position error + evaluation); a test inside the bound is decided in fp64 from the three
//    planes of the vertex (same arithmetic and bound as the team kernel), and only if that is
//    uncertain too the cell is flagged UNCERTAIN (reference fast-mode tie rule,
//    convex_cell.cpp:958-968).  A certified fp32 sign equals the certified fp64 sign, so both
//    kernels produce the same combinatorial cell;
//  * no adjacency: the border of a conflict zone is found by matching the directed dual edges
//    of the (few) conflict vertices against each other;
//  * volume and first moment are computed ONCE, at the end, in fp64, from the plane triples of
//    the final vertices by the orthoscheme (flag) decomposition: every (vertex, edge, face)
//    incidence contributes the signed simplex (seed, foot on the face, foot on the edge,
//    vertex) -- no adjacency, no face loops, and a fixed evaluation order (deterministic).
// Everything here is __host__ __device__: tests/host/tpc_host.cpp runs the very same code on
// the CPU against the oracle (logic check without a GPU; the product only ever runs it on the GPU).
#pragma once

#include <math.h>
#include <stdint.h>
#include <string.h>

#include "../../include/b200vor.h"
#include "exact_pred.cuh"
#include "grid_types.h"

namespace b200vor {
namespace tpc {

constexpr int NT_STD = 40;    // cell-vertex slots of the standard instance
constexpr int PC_STD = 40;    // plane slots (0..5 = clipping box)
constexpr int NT_BIG = 64;    // the large instance: cells of degenerate inputs re-done with exact ties
constexpr int PC_BIG = 64;
constexpr int POOLCAP = 32;   // candidates of one distance shell
constexpr int LC_STD = 16;    // conflict vertices of one cut (2 * LC >= POOLCAP: the words also hold sort keys)
constexpr int LC_BIG = 32;
constexpr int kMaxRing = 2;
constexpr float kU = 5.9604645e-8f;               // 2^-24
constexpr double kEps = 1.1102230246251565e-16;   // 2^-53
constexpr uint32_t kImg0 = 13u;                   // base-3 code of the zero translation
B200_HD constexpr unsigned long long live_init(int pc) {
    return 0x3Full | (pc < 64 ? ~((1ull << (pc & 63)) - 1ull) : 0ull);
}

struct Params {
    GridDesc G;
    const Seed* sorted;
    const uint32_t* cell_start;
    const uint32_t* perm;
    uint32_t n;
    int periodic;
    double period;
    double box[6];
    uint32_t nq;
    const uint32_t* qlist;   // grid-order indices of the cells to compute, or NULL (= all)
    double* m;
    double* mg;
    uint8_t* status;
    uint32_t* evcount;
    float target;            // expected candidates per distance shell
    int lockstep;            // 1: the lanes of a warp offer their k-th plane together (A/B knob)
    // face lists (FACES instances): neighbour ids ascending, areas (convex_cell.cpp:1160-1191)
    uint32_t* nfaces;
    uint32_t* face_nbr;
    double* face_area;
    uint32_t fcap;
};

#if defined(TPC_HOST_STATS)
#define TPC_REASON(k) __atomic_fetch_add(&tpc_reason[k], 1, __ATOMIC_RELAXED)
extern "C" long tpc_reason[8];
#else
#define TPC_REASON(k)
#endif

struct alignas(16) F4 {
    float x, y, z;
    uint32_t w;              // p0 | p1 << 6 | p2 << 12 | error exponent << 18
};

// the thread's columns of the block's shared memory (host emulation: STRIDE = 1)
template <int STRIDE> struct Mem {
    F4* V;            // [NT]
    uint32_t* gid;    // [PC]     neighbour (grid-order index | image << 27) of the plane slot
    uint32_t* pool;   // [POOLCAP]
    uint32_t* L;      // [LC]     plane triples of the conflict vertices of a cut; the same words hold
                      //          the 16-bit sort keys of a shell (two per word) while it is sorted
    B200_HD F4& v(uint32_t s) const { return V[s * STRIDE]; }
    B200_HD uint32_t& g(uint32_t p) const { return gid[p * STRIDE]; }
    B200_HD uint32_t& pl(uint32_t k) const { return pool[k * STRIDE]; }
    B200_HD uint16_t& ky(uint32_t k) const { return reinterpret_cast<uint16_t*>(&L[(k >> 1) * STRIDE])[k & 1u]; }
    B200_HD uint32_t& l(uint32_t k) const { return L[k * STRIDE]; }
};

// ---------------------------------------------------------------------------- portable helpers
B200_HD inline uint32_t f2u(float f) {
#if defined(__CUDA_ARCH__)
    return __float_as_uint(f);
#else
    uint32_t u;
    memcpy(&u, &f, 4);
    return u;
#endif
}
B200_HD inline float u2f(uint32_t u) {
#if defined(__CUDA_ARCH__)
    return __uint_as_float(u);
#else
    float f;
    memcpy(&f, &u, 4);
    return f;
#endif
}
B200_HD inline int ctz64(unsigned long long m) {
#if defined(__CUDA_ARCH__)
    return __ffsll((long long)m) - 1;
#else
    return __builtin_ctzll(m);
#endif
}
B200_HD inline int popc64(unsigned long long m) {
#if defined(__CUDA_ARCH__)
    return __popcll(m);
#else
    return __builtin_popcountll(m);
#endif
}
B200_HD inline float d2f_ru(double x) {
#if defined(__CUDA_ARCH__)
    return __double2float_ru(x);
#else
    float f = float(x);
    return double(f) < x ? nextafterf(f, INFINITY) : f;
#endif
}
B200_HD inline float d2f_rd(double x) {
#if defined(__CUDA_ARCH__)
    return __double2float_rd(x);
#else
    float f = float(x);
    return double(f) > x ? nextafterf(f, -INFINITY) : f;
#endif
}
B200_HD inline Seed ld_seed(const Seed* p) {
#if defined(__CUDA_ARCH__)
    const double2* q = reinterpret_cast<const double2*>(p);
    double2 a = __ldg(q), b = __ldg(q + 1);
    Seed s;
    s.x = a.x; s.y = a.y; s.z = b.x; s.w = b.y;
    return s;
#else
    return *p;
#endif
}
B200_HD inline uint32_t ldu(const uint32_t* p) {
#if defined(__CUDA_ARCH__)
    return __ldg(p);
#else
    return *p;
#endif
}
B200_HD inline float fmaxf3(float a, float b, float c) { return fmaxf(a, fmaxf(b, c)); }
// warp-level predicates: the 32 cells of a warp run the loops of compute_cell() together, so
// that they can be brought back in step before the expensive part of every round
B200_HD inline bool warp_any(bool p) {
#if defined(__CUDA_ARCH__)
    return __any_sync(0xffffffffu, p) != 0;
#else
    return p;
#endif
}
B200_HD inline void warp_sync() {
#if defined(__CUDA_ARCH__)
    __syncwarp();
#endif
}

struct PlaneD {
    double a, b, c, d;
    double lam;   // foot scale: the foot of the seed on the plane is lam * (a,b,c)
};
struct PlaneF {
    float a, b, c, d;
    float n1;     // >= |a| + |b| + |c| of the fp64 plane
    float nm;     // >= max |normal component|
    float da;     // >= |d|
};

B200_HD inline PlaneF to_f(const PlaneD& p) {
    PlaneF f;
    f.a = float(p.a); f.b = float(p.b); f.c = float(p.c); f.d = float(p.d);
    const float aa = fabsf(f.a), ab = fabsf(f.b), ac = fabsf(f.c);
    f.n1 = (aa + ab + ac) * (1.0f + 8.0f * kU);
    f.nm = fmaxf3(aa, ab, ac) * (1.0f + 4.0f * kU);
    f.da = fabsf(f.d) * (1.0f + 4.0f * kU);
    return f;
}

// per-cell constants
struct Frame {
    Seed si;
    double blo[3], bhi[3];   // clipping box, seed-centred
    double period;
};

B200_HD inline PlaneD box_plane(const Frame& F, uint32_t p) {
    const uint32_t axis = p >> 1;
    const bool upper = p & 1u;
    const double sgn = upper ? -1.0 : 1.0;
    PlaneD P;
    P.a = axis == 0 ? sgn : 0.0;
    P.b = axis == 1 ? sgn : 0.0;
    P.c = axis == 2 ? sgn : 0.0;
    P.d = upper ? F.bhi[axis] : -F.blo[axis];
    P.lam = -P.d;
    return P;
}

// bisector of the seed and neighbour `id` (grid-order index | image << 27), seed-centred frame:
// (-2q, |q|^2), q = (p_j + translation) - p_i  (periodic_delaunay_3d.cpp:3719-3726 in that frame)
B200_HD inline PlaneD nbr_plane(const Params& P, const Frame& F, uint32_t id, double& d2) {
    const uint32_t j = id & 0x7FFFFFFu, im = id >> 27;
    const Seed sj = ld_seed(P.sorted + j);
    const double shx = double(int(im % 3u) - 1) * F.period, shy = double(int((im / 3u) % 3u) - 1) * F.period,
                 shz = double(int(im / 9u) - 1) * F.period;
    const double qx = (sj.x + shx) - F.si.x, qy = (sj.y + shy) - F.si.y, qz = (sj.z + shz) - F.si.z;
    d2 = fma(qz, qz, fma(qy, qy, qx * qx));
    PlaneD Q;
    Q.a = -2.0 * qx; Q.b = -2.0 * qy; Q.c = -2.0 * qz;
    Q.d = d2;
    Q.lam = -0.25;   // -d / |n|^2 = -|q|^2 / (4 |q|^2), exactly
    return Q;
}

template <int STRIDE>
B200_HD inline PlaneD slot_plane(const Params& P, const Frame& F, const Mem<STRIDE>& M, uint32_t p) {
    if(p < 6u) {
        return box_plane(F, p);
    }
    double d2;
    return nbr_plane(P, F, M.g(p), d2);
}

// ---------------------------------------------------------------------------- fp32 vertex
// Intersection of three planes in fp32 (Cramer, cross-product form) with a bound rho on the
// absolute error of every coordinate against the EXACT intersection of the fp64 planes:
//   |X~ - X| <= 64u Q,  Q = sum |d_i| nm_j nm_k     (input rounding 6u Q + evaluation 10u Q, x4 slack)
//   |W~ - W| <= 72u nm_0 nm_1 nm_2                  (input rounding 18u + evaluation 30u)
//   x = X / W, |x| <= Rmax (the vertex lies on an edge of the current cell)
//   rho = (64u Q + 72u n012 Rmax) / (|W~| - 72u n012) + 3u Rmax
// Returns the error exponent code (255 = degenerate: unknown position).
B200_HD inline uint32_t make_vertex_f(const PlaneF& P0, const PlaneF& P1, const PlaneF& P2, float Rmax, float& x,
                                      float& y, float& z) {
    const float ux = P1.b * P2.c - P1.c * P2.b, uy = P1.c * P2.a - P1.a * P2.c, uz = P1.a * P2.b - P1.b * P2.a;
    const float vx = P2.b * P0.c - P2.c * P0.b, vy = P2.c * P0.a - P2.a * P0.c, vz = P2.a * P0.b - P2.b * P0.a;
    const float wx = P0.b * P1.c - P0.c * P1.b, wy = P0.c * P1.a - P0.a * P1.c, wz = P0.a * P1.b - P0.b * P1.a;
    const float W = P0.a * ux + P0.b * uy + P0.c * uz;
    const float X = -(P0.d * ux + P1.d * vx + P2.d * wx);
    const float Y = -(P0.d * uy + P1.d * vy + P2.d * wy);
    const float Z = -(P0.d * uz + P1.d * vz + P2.d * wz);
    const float n012 = P0.nm * P1.nm * P2.nm;
    const float Q = P0.da * P1.nm * P2.nm + P1.da * P2.nm * P0.nm + P2.da * P0.nm * P1.nm;
    const float aW = fabsf(W);
    const float eW = 72.0f * kU * n012;
    if(!(aW > 2.0f * eW)) {
        x = y = z = 0.0f;
        return 255u;
    }
    const float inv = 1.0f / W;
    x = X * inv; y = Y * inv; z = Z * inv;
    const float rho = (64.0f * kU * Q + eW * Rmax) / (aW - eW) * 1.0001f + 3.0f * kU * Rmax;
    // power of two >= rho
    uint32_t code = (f2u(rho) >> 23) + 1u;
    return code > 254u ? 255u : (code < 1u ? 1u : code);
}

// ---------------------------------------------------------------------------- fp64 vertex
// Homogeneous intersection of three planes with its forward error coefficient (the arithmetic
// and the bound of make_vertex in fused_kernel.inc / cell.cu).
struct VertexD {
    double x, y, z, w;
    float e;
    bool degenerate;
};
B200_HD inline VertexD make_vertex_d(const PlaneD& P0, const PlaneD& P1, const PlaneD& P2) {
    const double a0 = P0.a, b0 = P0.b, c0 = P0.c, d0 = P0.d;
    const double a1 = P1.a, b1 = P1.b, c1 = P1.c, d1 = P1.d;
    const double a2 = P2.a, b2 = P2.b, c2 = P2.c, d2 = P2.d;
    const float n0 = d2f_ru(fmax(fabs(a0), fmax(fabs(b0), fabs(c0))));
    const float n1 = d2f_ru(fmax(fabs(a1), fmax(fabs(b1), fabs(c1))));
    const float n2 = d2f_ru(fmax(fabs(a2), fmax(fabs(b2), fabs(c2))));
    const double ux = b1 * c2 - c1 * b2, uy = c1 * a2 - a1 * c2, uz = a1 * b2 - b1 * a2;
    const double vx = b2 * c0 - c2 * b0, vy = c2 * a0 - a2 * c0, vz = a2 * b0 - b2 * a0;
    const double wx = b0 * c1 - c0 * b1, wy = c0 * a1 - a0 * c1, wz = a0 * b1 - b0 * a1;
    double w = a0 * ux + b0 * uy + c0 * uz;
    double x = -(d0 * ux + d1 * vx + d2 * wx);
    double y = -(d0 * uy + d1 * vy + d2 * wy);
    double z = -(d0 * uz + d1 * vz + d2 * wz);
    if(w < 0.0) {
        w = -w; x = -x; y = -y; z = -z;
    }
    VertexD V;
    V.x = x; V.y = y; V.z = z; V.w = w;
    const float D0 = d2f_ru(fabs(d0)), D1 = d2f_ru(fabs(d1)), D2 = d2f_ru(fabs(d2));
    const float Q = D0 * n1 * n2 + D1 * n2 * n0 + D2 * n0 * n1;
    const float sx = d2f_ru(fabs(x) + fabs(y) + fabs(z));
    const float sw = d2f_ru(w);
    const float e48 = float(48.0 * kEps) * 1.001f, e4 = float(4.0 * kEps) * 1.001f;
    const float n012 = n0 * n1 * n2;
    const float ea = (e48 * Q + e4 * sx) * 1.001f;
    const float eb = (e48 * n012 + e4 * sw) * 1.001f;
    V.e = (ea + eb) * 1.0001f;
    V.degenerate = !(sw > 2.0f * e48 * n012);
    return V;
}

// ---------------------------------------------------------------------------- coverage
// squared distance from the seed to the nearest face of the scanned block of bins that has
// unscanned space behind it (1e300 when the block covers the whole grid)
B200_HD inline double coverage2(const GridDesc& Gd, const Seed& s, int cx, int cy, int cz, int r) {
    double dmin = 1e300;
    const double q[3] = {s.x, s.y, s.z};
    const int cc[3] = {cx, cy, cz};
    for(int a = 0; a < 3; ++a) {
        if(Gd.periodic || cc[a] - r > 0) {
            dmin = fmin(dmin, q[a] - (Gd.lo[a] + (cc[a] - r) * Gd.w[a]));
        }
        if(Gd.periodic || cc[a] + r < Gd.g[a] - 1) {
            dmin = fmin(dmin, (Gd.lo[a] + (cc[a] + r + 1) * Gd.w[a]) - q[a]);
        }
    }
    if(dmin >= 1e300) {
        return 1e300;
    }
    dmin -= 1e-9 * Gd.w[0];
    return dmin > 0.0 ? dmin * dmin : 0.0;
}

// ---------------------------------------------------------------------------- the cell
// exact decision of a tie of the fp64 filter (out of line: rare, and it carries a large frame)
#if defined(__CUDA_ARCH__)
__device__ __noinline__
#else
inline
#endif
int exact_plane_vertex_sign(const PlaneD& P0, const PlaneD& P1, const PlaneD& P2, const PlaneD& Q) {
    const double a0[4] = {P0.a, P0.b, P0.c, P0.d}, a1[4] = {P1.a, P1.b, P1.c, P1.d}, a2[4] = {P2.a, P2.b, P2.c, P2.d},
                 q[4] = {Q.a, Q.b, Q.c, Q.d};
    return exact::plane_vertex_sign(a0, a1, a2, q);
}

// EV: event counters of the SURVEY 8(d) flop model.  FACES: face lists + areas.  NT / PC: capacities.
template <int STRIDE, bool EV, bool FACES, int NT, int PC, int LCAP>
B200_HD inline void compute_cell(const Params& P, uint32_t c, const Mem<STRIDE>& M) {
    constexpr unsigned long long kLiveInit = live_init(PC);
    const GridDesc& Gd = P.G;
    // every lane of a warp goes through the (warp-uniform) loops below; a lane without a cell
    // (tail of the launch) idles through them and writes nothing
    const bool alive = c < P.nq;
    if(!alive) {
        c = 0;
    }
    const uint32_t i = P.qlist ? ldu(P.qlist + c) : c;
    Frame F;
    F.si = ld_seed(P.sorted + i);
    F.period = P.periodic ? P.period : 0.0;
    const Seed& si = F.si;
    const uint32_t row = ldu(P.perm + i);
    const double sxyz[3] = {si.x, si.y, si.z};
    for(int a = 0; a < 3; ++a) {
        F.blo[a] = P.periodic ? -0.5 * P.period : P.box[a] - sxyz[a];
        F.bhi[a] = P.periodic ? 0.5 * P.period : P.box[3 + a] - sxyz[a];
    }

    // ------------------------------------------------------------------ init_with_box
    // (convex_cell.cpp:175-215): 8 vertices, triples oriented like the team kernel
    float R2 = 0.0f;
    for(uint32_t s = 0; s < 8; ++s) {
        const uint32_t bxb = s & 1u, byb = (s >> 1) & 1u, bzb = (s >> 2) & 1u;
        const uint32_t px = bxb, py = 2u + byb, pz = 4u + bzb;
        const bool odd = (bxb + byb + bzb) & 1u;
        const uint32_t v1 = odd ? pz : py, v2 = odd ? py : pz;
        F4 V;
        V.x = float(bxb ? F.bhi[0] : F.blo[0]);
        V.y = float(byb ? F.bhi[1] : F.blo[1]);
        V.z = float(bzb ? F.bhi[2] : F.blo[2]);
        const float mx = fmaxf3(fabsf(V.x), fabsf(V.y), fabsf(V.z));
        const float rho = 2.0f * kU * mx + 1e-37f;
        uint32_t code = (f2u(rho) >> 23) + 1u;
        code = code > 254u ? 254u : code;
        V.w = px | (v1 << 6) | (v2 << 12) | (code << 18);
        M.v(s) = V;
        const float r2 = (V.x * V.x + V.y * V.y + V.z * V.z) * (1.0f + 16.0f * kU);
        R2 = fmaxf(R2, r2);
    }
    uint32_t T = 8;
    unsigned long long live = kLiveInit;
    uint32_t status = 0, extra = 0;
    uint32_t nP = 0, nC = 0, nV = 0, nscan = 0;
    float Rmax = sqrtf(R2) * (1.0f + 4.0f * kU);
    float bnd = 4.1f * R2;

    // ------------------------------------------------------------------ neighbour walk
    int cx = clampi(int(floor((si.x - Gd.lo[0]) * Gd.inv_w[0])), 0, Gd.g[0] - 1);
    int cy = clampi(int(floor((si.y - Gd.lo[1]) * Gd.inv_w[1])), 0, Gd.g[1] - 1);
    int cz = clampi(int(floor((si.z - Gd.lo[2]) * Gd.inv_w[2])), 0, Gd.g[2] - 1);
    int ring = 1, rmax;
    if(Gd.periodic) {
        const int gmin = Gd.g[0] < Gd.g[1] ? (Gd.g[0] < Gd.g[2] ? Gd.g[0] : Gd.g[2]) : (Gd.g[1] < Gd.g[2] ? Gd.g[1] : Gd.g[2]);
        rmax = (gmin - 1) / 2;
        rmax = rmax < kMaxRing ? rmax : kMaxRing;
    } else {
        rmax = kMaxRing;
    }
    bool certified = false;
    bool done = rmax < 1 || !alive;
    double cov2 = done ? 0.0 : coverage2(Gd, si, cx, cy, cz, ring);
    double lo = 0.0, hi = 0.0;
    // shell width: the ball expected to hold `target` more seeds at the density of the 27 bins
    // around the seed
    float c3;
    {
        uint32_t cnt27 = 0, bins27 = 0;
        for(int rowi = 0; rowi < 9; ++rowi) {
            int y = cy + rowi % 3 - 1, z = cz + rowi / 3 - 1;
            if(Gd.periodic) {
                y = y < 0 ? y + Gd.g[1] : (y >= Gd.g[1] ? y - Gd.g[1] : y);
                z = z < 0 ? z + Gd.g[2] : (z >= Gd.g[2] ? z - Gd.g[2] : z);
            } else if(y < 0 || y >= Gd.g[1] || z < 0 || z >= Gd.g[2]) {
                continue;
            }
            const uint32_t base = uint32_t(Gd.g[0]) * (uint32_t(y) + uint32_t(Gd.g[1]) * uint32_t(z));
            const int xa = cx - 1 < 0 ? 0 : cx - 1, xb = cx + 1 >= Gd.g[0] ? Gd.g[0] - 1 : cx + 1;
            cnt27 += ldu(P.cell_start + base + xb + 1) - ldu(P.cell_start + base + xa);
            bins27 += uint32_t(xb - xa + 1);
        }
        const float vol = float(bins27) * float(Gd.w[0] * Gd.w[1] * Gd.w[2]);
        c3 = P.target * 0.2387324146378430f * vol / float(cnt27 > 0u ? cnt27 : 1u);
    }

    while(warp_any(!done)) {
        // ---- next distance shell [lo, hi) ----
        for(; !done;) {
            if(lo > double(bnd)) {
                certified = true;
                done = true;
                break;
            }
            if(lo >= cov2) {
                if(ring >= rmax) {
                    TPC_REASON(4);
                    certified = (cov2 >= 1e300);
                    done = true;
                    break;
                }
                ++ring;
                cov2 = coverage2(Gd, si, cx, cy, cz, ring);
                continue;
            }
            const float rlo = sqrtf(float(lo));
            const float r3 = rlo * rlo * rlo + c3;
            const float rhi = cbrtf(r3) * 1.0001f;
            hi = fmax(double(rhi) * double(rhi), lo * (1.0 + 1e-6) + 1e-300);
            hi = fmin(hi, fmin(double(bnd) * (1.0 + 1e-6) + 1e-300, cov2));
            break;
        }
        // ================================================================ scan
        // seeds of the (2 ring + 1)^3 block of bins with lo <= d2 < hi -> pool
        uint32_t cnt = 0;
        bool overflow = false;
        if(!done) {
            const int R1 = 2 * ring + 1;
            const int nrows = R1 * R1;
            const double slack = 1e-9 * Gd.w[0];
            const float kscale = 65535.0f / float(hi - lo);
            for(int rowi = 0; rowi < nrows; ++rowi) {
                const int dy = rowi % R1 - ring, dz = rowi / R1 - ring;
                int y = cy + dy, z = cz + dz, sy = 0, sz = 0;
                // the row of bins is scanned only where it meets the ball d2 < hi
                double dyy = dy == 0 ? 0.0
                                     : (dy < 0 ? si.y - (Gd.lo[1] + (cy + dy + 1) * Gd.w[1])
                                               : (Gd.lo[1] + (cy + dy) * Gd.w[1]) - si.y);
                double dzz = dz == 0 ? 0.0
                                     : (dz < 0 ? si.z - (Gd.lo[2] + (cz + dz + 1) * Gd.w[2])
                                               : (Gd.lo[2] + (cz + dz) * Gd.w[2]) - si.z);
                dyy = fmax(dyy - slack, 0.0);
                dzz = fmax(dzz - slack, 0.0);
                const double rem = hi - (dyy * dyy + dzz * dzz);
                bool ok = rem > 0.0;
                int xa = cx - ring, xb = cx + ring;
                if(ok) {
                    const double rx = sqrt(rem) + slack;
                    const int fa = int(floor((si.x - rx - Gd.lo[0]) * Gd.inv_w[0]));
                    const int fb = int(floor((si.x + rx - Gd.lo[0]) * Gd.inv_w[0]));
                    xa = xa > fa ? xa : fa;
                    xb = xb < fb ? xb : fb;
                    ok = xa <= xb;
                }
                if(y < 0 || y >= Gd.g[1]) {
                    if(Gd.periodic) {
                        sy = y < 0 ? -1 : 1;
                        y -= sy * Gd.g[1];
                    } else {
                        ok = false;
                    }
                }
                if(z < 0 || z >= Gd.g[2]) {
                    if(Gd.periodic) {
                        sz = z < 0 ? -1 : 1;
                        z -= sz * Gd.g[2];
                    } else {
                        ok = false;
                    }
                }
                uint32_t a0 = 0, n0 = 0, a1 = 0, n1 = 0;
                int sx1 = 0;
                if(ok) {
                    const int gx = Gd.g[0];
                    const uint32_t base = uint32_t(gx) * (uint32_t(y) + uint32_t(Gd.g[1]) * uint32_t(z));
                    const int ma = xa < 0 ? 0 : xa, mb = xb >= gx ? gx - 1 : xb;
                    if(ma <= mb) {
                        a0 = ldu(P.cell_start + base + ma);
                        n0 = ldu(P.cell_start + base + mb + 1) - a0;
                    }
                    if(Gd.periodic) {
                        if(xa < 0) {
                            sx1 = -1;
                            a1 = ldu(P.cell_start + base + (xa + gx));
                            n1 = ldu(P.cell_start + base + gx) - a1;
                        } else if(xb >= gx) {
                            sx1 = 1;
                            a1 = ldu(P.cell_start + base);
                            n1 = ldu(P.cell_start + base + (xb - gx) + 1) - a1;
                        }
                    }
                }
                const double oy = double(sy) * F.period, oz = double(sz) * F.period;
                for(int sp = 0; sp < 2; ++sp) {
                    const uint32_t ja = sp ? a1 : a0, jn = sp ? n1 : n0;
                    const int sx = sp ? sx1 : 0;
                    const double ox = double(sx) * F.period;
                    const uint32_t im = uint32_t(sx + 1) + 3u * uint32_t(sy + 1) + 9u * uint32_t(sz + 1);
                    auto consider = [&](const Seed& sj, uint32_t j) {
                        const double qx = (sj.x + ox) - si.x, qy = (sj.y + oy) - si.y, qz = (sj.z + oz) - si.z;
                        const double d2 = fma(qz, qz, fma(qy, qy, qx * qx));
                        if(d2 >= lo && d2 < hi && !(j == i && im == kImg0)) {
                            if(d2 == 0.0) {
                                // exact duplicate (delaunay_nn.cpp:122-135): the lower user index keeps
                                // the cell; the duplicate itself is never a neighbour
                                if(im == kImg0 && ldu(P.perm + j) < row) {
                                    extra |= B200VOR_CELL_DUPLICATE;
                                }
                            } else if(cnt < uint32_t(POOLCAP)) {
                                M.pl(cnt) = j | (im << 27);
                                M.ky(cnt) = uint16_t(float(d2 - lo) * kscale);
                                ++cnt;
                            } else {
                                overflow = true;
                            }
                        }
                    };
                    // four seeds per round: the loads are issued together (few warps per scheduler:
                    // the latency of a lone load is not hidden by other warps)
                    const uint32_t je = ja + jn;
                    uint32_t j = ja;
                    for(; j + 4u <= je; j += 4u) {
                        const Seed s0 = ld_seed(P.sorted + j), s1 = ld_seed(P.sorted + j + 1u),
                                   s2 = ld_seed(P.sorted + j + 2u), s3 = ld_seed(P.sorted + j + 3u);
                        consider(s0, j);
                        consider(s1, j + 1u);
                        consider(s2, j + 2u);
                        consider(s3, j + 3u);
                    }
                    for(; j < je; ++j) {
                        consider(ld_seed(P.sorted + j), j);
                    }
                }
            }
        }
        if(EV && !done) {
            ++nscan;
        }
        if(overflow) {
            // more candidates in the shell than the pool holds: left to the team kernel
            TPC_REASON(0);
            status |= B200VOR_CELL_OVERFLOW;
            done = true;
            cnt = 0;
        }
        // ================================================================ sort (insertion, by key)
        for(uint32_t a = 1; a < cnt; ++a) {
            const uint32_t ida = M.pl(a);
            const uint16_t ka = M.ky(a);
            uint32_t b = a;
            while(b > 0 && M.ky(b - 1) > ka) {
                M.pl(b) = M.pl(b - 1);
                M.ky(b) = M.ky(b - 1);
                --b;
            }
            M.pl(b) = ida;
            M.ky(b) = ka;
        }
        // ================================================================ clip
        uint32_t k = 0;
        bool clip = !done;   // this lane still has planes of the shell to offer
        while(warp_any(clip)) {
          // ---- the next plane of the shell that cuts the cell.  Every lane advances through its
          //      own pool here (cheap, regular work); the lanes of the warp are brought back in step
          //      before the cut, so that the expensive, irregular part runs with most lanes active ----
          unsigned long long mask = 0ull;
          uint32_t id = 0;
          PlaneD pd = {0.0, 0.0, 0.0, 0.0, 0.0};
          PlaneF pf = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
          uint32_t tried = 0;
          while(clip && k < cnt && mask == 0ull && !(P.lockstep && tried)) {
            id = M.pl(k);
            ++k;
            ++tried;
            double d2;
            pd = nbr_plane(P, F, id, d2);
            if(d2 > double(bnd)) {
                continue;   // beyond the security radius (generic_RVD.h:2322): cannot cut
            }
            pf = to_f(pd);
            // ---- sign tests (triangle_is_in_conflict, convex_cell.cpp:870-969) in fp32:
            //      |s~ - s| <= n1 rho_v + 4.5u (n1 Rmax + |d|)
            const float K = 4.5f * kU * (pf.n1 * Rmax + pf.da) * (1.0f + 8.0f * kU);
            unsigned long long unc = 0ull;
            for(uint32_t s = 0; s < T; ++s) {
                const F4 V = M.v(s);
                const float val = fmaf(pf.a, V.x, fmaf(pf.b, V.y, fmaf(pf.c, V.z, pf.d)));
                const float rho = u2f(((V.w >> 18) & 0xFFu) << 23);
                const float E = fmaf(pf.n1, rho, K);
                if(val < -E) {
                    mask |= 1ull << s;
                } else if(!(val > E)) {
                    unc |= 1ull << s;
                }
            }
            if(EV) {
                ++nP;
                nC += T;
            }
            // ---- tests inside the fp32 bound: decided in fp64 from the vertex's three planes ----
            while(unc) {
                const uint32_t s = uint32_t(ctz64(unc));
                unc &= unc - 1ull;
                const uint32_t t = M.v(s).w;
                const PlaneD Q0 = slot_plane<STRIDE>(P, F, M, t & 63u), Q1 = slot_plane<STRIDE>(P, F, M, (t >> 6) & 63u),
                             Q2 = slot_plane<STRIDE>(P, F, M, (t >> 12) & 63u);
                const VertexD V = make_vertex_d(Q0, Q1, Q2);
                const double dot = fma(pd.a, V.x, fma(pd.b, V.y, fma(pd.c, V.z, pd.d * V.w)));
                const float Mq = d2f_ru(fmax(fmax(fabs(pd.a), fabs(pd.b)), fmax(fabs(pd.c), fabs(pd.d)))) * 1.0001f;
                const bool tie = V.degenerate || d2f_rd(fabs(dot)) <= Mq * V.e;
                TPC_REASON(5);
                if(tie) {
                    TPC_REASON(6);
                    // inside the fp64 bound as well: decided EXACTLY (expansions), with the rule of the
                    // reference's default mode -- a vertex ON the plane is in conflict
                    // (convex_cell.cpp:905-911: det_4d >= 0)
                    const int sg = exact_plane_vertex_sign(Q0, Q1, Q2, pd);
                    if(sg == 2) {
                        status |= B200VOR_CELL_UNCERTAIN;   // the three planes do not meet in a point
                    } else if(sg <= 0) {
                        mask |= 1ull << s;
                    }
                } else if(dot < 0.0) {
                    mask |= 1ull << s;
                }
            }
          }
          warp_sync();
          if(mask == 0ull) {
              clip = false;   // shell used up
          }
          if(clip) do {
            // ================================================================ cut
            // (triangulate_conflict_zone, convex_cell.cpp:761-865, without adjacency)
            const uint32_t nconf = uint32_t(popc64(mask));
            if(nconf == T) {
                T = 0;
                status |= B200VOR_CELL_EMPTY;
                certified = true;
                done = true;
                break;
            }
            if(nconf > uint32_t(LCAP) || 3u * nconf > 64u) {
                TPC_REASON(1);
                status |= B200VOR_CELL_OVERFLOW;
                done = true;
                break;
            }
            {
                uint32_t q = 0;
                unsigned long long mm = mask;
                while(mm) {
                    const uint32_t s = uint32_t(ctz64(mm));
                    mm &= mm - 1ull;
                    M.l(q++) = M.v(s).w & 0x3FFFFu;
                }
            }
            if(live == ~0ull) {
                // plane table full: drop the planes no surviving vertex refers to
                unsigned long long used = 0ull;
                for(uint32_t s = 0; s < T; ++s) {
                    if(!((mask >> s) & 1ull)) {
                        const uint32_t t = M.v(s).w;
                        used |= (1ull << (t & 63u)) | (1ull << ((t >> 6) & 63u)) | (1ull << ((t >> 12) & 63u));
                    }
                }
                live = used | kLiveInit;
                if(live == ~0ull) {
                    TPC_REASON(2);
                    status |= B200VOR_CELL_OVERFLOW;
                    done = true;
                    break;
                }
            }
            const uint32_t pn = uint32_t(ctz64(~live));
            live |= 1ull << pn;
            M.g(pn) = id;
            // ---- border edges: directed dual edges of the conflict vertices whose twin is not among
            //      them (bit 3 a + e = edge e of conflict vertex a) ----
            unsigned long long border = 0ull;
            for(uint32_t a = 0; a < nconf; ++a) {
                const uint32_t t = M.l(a);
                const uint32_t pa = t & 63u, pb = (t >> 6) & 63u, pc = (t >> 12) & 63u;
                // twins of the directed edges (pa->pb), (pb->pc), (pc->pa), as first | second << 6
                const uint32_t tw0 = pb | (pa << 6), tw1 = pc | (pb << 6), tw2 = pa | (pc << 6);
                bool f0 = false, f1 = false, f2 = false;
                for(uint32_t b = 0; b < nconf; ++b) {
                    const uint32_t tj = M.l(b);
                    const uint32_t r0 = tj & 0xFFFu, r1 = (tj >> 6) & 0xFFFu, r2 = ((tj >> 12) & 63u) | ((tj & 63u) << 6);
                    f0 |= (r0 == tw0) | (r1 == tw0) | (r2 == tw0);
                    f1 |= (r0 == tw1) | (r1 == tw1) | (r2 == tw1);
                    f2 |= (r0 == tw2) | (r1 == tw2) | (r2 == tw2);
                }
                // (a vertex never holds the twin of its own edge: b == a needs no exclusion)
                border |= (unsigned long long)((f0 ? 0u : 1u) | (f1 ? 0u : 2u) | (f2 ? 0u : 4u)) << (3u * a);
            }
            // ---- one new vertex (pn, u, v) per border edge (u -> v).  A separate loop over the set
            //      bits: at step k every lane of the warp that has a k-th border edge is active ----
            unsigned long long holes = mask;
            uint32_t Ttop = T, nnew = 0;
            bool ovf = false;
            while(border) {
                const uint32_t bit = uint32_t(ctz64(border));
                border &= border - 1ull;
                const uint32_t a = bit / 3u, e = bit - 3u * a;
                const uint32_t t = M.l(a);
                const uint32_t u = (t >> (6u * e)) & 63u, v = (t >> (6u * (e == 2u ? 0u : e + 1u))) & 63u;
                uint32_t s;
                if(holes) {
                    s = uint32_t(ctz64(holes));
                    holes &= holes - 1ull;
                } else {
                    s = Ttop++;
                    if(s >= uint32_t(NT)) {
                        ovf = true;
                        break;
                    }
                }
                const PlaneF fu = to_f(slot_plane<STRIDE>(P, F, M, u));
                const PlaneF fv = to_f(slot_plane<STRIDE>(P, F, M, v));
                F4 V;
                const uint32_t code = make_vertex_f(pf, fu, fv, Rmax, V.x, V.y, V.z);
                V.w = pn | (u << 6) | (v << 12) | (code << 18);
                M.v(s) = V;
                ++nnew;
            }
            if(ovf) {
                TPC_REASON(3);
                status |= B200VOR_CELL_OVERFLOW;
                done = true;
                break;
            }
            const uint32_t Tnew = T - nconf + nnew;
            if(holes) {
                // fewer new vertices than holes: the holes below Tnew take the survivors stored
                // at or above Tnew
                const unsigned long long open = holes;
                uint32_t top = T - 1u;
                while(holes) {
                    const uint32_t h = uint32_t(ctz64(holes));
                    if(h >= Tnew) {
                        break;
                    }
                    holes &= holes - 1ull;
                    while((open >> top) & 1ull) {
                        --top;
                    }
                    M.v(h) = M.v(top);
                    --top;
                }
            }
            T = Tnew;
            if(EV) {
                nV += nnew;
            }
            // ---- squared radius of the cell (convex_cell.cpp:1312-1330), rounded up:
            //      |x|^2 <= (|x~| + sqrt(3) rho)^2 <= r~^2 + 4 rho (Rmax + 2 rho) ----
            {
                float rb = 0.0f;
                for(uint32_t s = 0; s < T; ++s) {
                    const F4 V = M.v(s);
                    const float rho = u2f(((V.w >> 18) & 0xFFu) << 23);
                    const float r2 = fmaf(V.x, V.x, fmaf(V.y, V.y, V.z * V.z));
                    rb = fmaxf(rb, fmaf(4.0f * rho, Rmax + 2.0f * rho, r2));
                }
                rb *= 1.0f + 16.0f * kU;
                R2 = fminf(R2, rb);   // the cell only shrinks (also absorbs rho = inf)
                Rmax = sqrtf(R2) * (1.0f + 4.0f * kU);
                bnd = 4.1f * R2;
            }
          } while(false);
          if(done) {
              clip = false;
          }
        }
        lo = hi;
    }
    status |= extra;
    if(!certified) {
        status |= B200VOR_CELL_UNCERTIFIED;
    }

    // ------------------------------------------------------------------ integration (fp64)
    const bool dead = (status & (B200VOR_CELL_EMPTY | B200VOR_CELL_DUPLICATE)) != 0;
    const bool failed = (status & (B200VOR_CELL_OVERFLOW | B200VOR_CELL_UNCERTIFIED)) != 0 || !alive;
    double m6 = 0.0, gx = 0.0, gy = 0.0, gz = 0.0;
    unsigned long long used = 0ull;
    bool degenerate = false;
    // FACES: 6 x signed volume of the orthoschemes of every face, accumulated per plane slot in the
    // (dead) coordinate words of the vertex records -- only the plane triples are read from here on
    auto facc = [&](uint32_t p) -> double& { return *reinterpret_cast<double*>(&M.v(p).x); };
    if(FACES) {
        static_assert(PC <= NT, "face accumulators live in the vertex records");
        for(uint32_t p = 0; p < uint32_t(PC); ++p) {
            facc(p) = 0.0;
        }
    }
    if(!dead && !failed) {
        for(uint32_t s = 0; s < T; ++s) {
            const uint32_t t = M.v(s).w;
            const uint32_t p0 = t & 63u, p1 = (t >> 6) & 63u, p2 = (t >> 12) & 63u;
            used |= (1ull << p0) | (1ull << p1) | (1ull << p2);
            const PlaneD P0 = slot_plane<STRIDE>(P, F, M, p0), P1 = slot_plane<STRIDE>(P, F, M, p1),
                         P2 = slot_plane<STRIDE>(P, F, M, p2);
            // vertex (compute_triangle_point, convex_cell.cpp:971-1082)
            const double ux = P1.b * P2.c - P1.c * P2.b, uy = P1.c * P2.a - P1.a * P2.c, uz = P1.a * P2.b - P1.b * P2.a;
            const double vx = P2.b * P0.c - P2.c * P0.b, vy = P2.c * P0.a - P2.a * P0.c, vz = P2.a * P0.b - P2.b * P0.a;
            const double wx = P0.b * P1.c - P0.c * P1.b, wy = P0.c * P1.a - P0.a * P1.c, wz = P0.a * P1.b - P0.b * P1.a;
            const double W = P0.a * ux + P0.b * uy + P0.c * uz;
            if(W == 0.0) {
                degenerate = true;
                continue;
            }
            const double iw = -1.0 / W;
            const double X = (P0.d * ux + P1.d * vx + P2.d * wx) * iw;
            const double Y = (P0.d * uy + P1.d * vy + P2.d * wy) * iw;
            const double Z = (P0.d * uz + P1.d * vz + P2.d * wz) * iw;
            const double sg = W < 0.0 ? -1.0 : 1.0;
            // feet of the seed on the three planes
            const double A[3][3] = {{P0.lam * P0.a, P0.lam * P0.b, P0.lam * P0.c},
                                    {P1.lam * P1.a, P1.lam * P1.b, P1.lam * P1.c},
                                    {P2.lam * P2.a, P2.lam * P2.b, P2.lam * P2.c}};
            // edges (f, a) = (0,1), (1,2), (2,0): direction n_f x n_a
            const double D[3][3] = {{wx, wy, wz}, {ux, uy, uz}, {vx, vy, vz}};
#pragma unroll
            for(int e = 0; e < 3; ++e) {
                const int f = e, a = (e + 1) % 3;
                const double dx = D[e][0], dy = D[e][1], dz = D[e][2];
                const double id2 = 1.0 / (dx * dx + dy * dy + dz * dz);
                // g = d x vertex
                const double g0 = dy * Z - dz * Y, g1 = dz * X - dx * Z, g2 = dx * Y - dy * X;
                const double tf = ((X - A[f][0]) * dx + (Y - A[f][1]) * dy + (Z - A[f][2]) * dz) * id2;
                const double ta = ((X - A[a][0]) * dx + (Y - A[a][1]) * dy + (Z - A[a][2]) * dz) * id2;
                // signed orthoschemes (seed, foot on the face, foot on the edge, vertex), x6
                const double vf = sg * tf * (A[f][0] * g0 + A[f][1] * g1 + A[f][2] * g2);
                const double va = -sg * ta * (A[a][0] * g0 + A[a][1] * g1 + A[a][2] * g2);
                m6 += vf + va;
                if(FACES) {
                    const uint32_t pf_ = e == 0 ? p0 : (e == 1 ? p1 : p2), pa_ = e == 0 ? p1 : (e == 1 ? p2 : p0);
                    facc(pf_) += vf;
                    facc(pa_) += va;
                }
                // first moment: volume x (foot_face + foot_edge + vertex) / 4, foot_edge = vertex - t d
                gx += vf * (A[f][0] + 2.0 * X - tf * dx) + va * (A[a][0] + 2.0 * X - ta * dx);
                gy += vf * (A[f][1] + 2.0 * Y - tf * dy) + va * (A[a][1] + 2.0 * Y - ta * dy);
                gz += vf * (A[f][2] + 2.0 * Z - tf * dz) + va * (A[a][2] + 2.0 * Z - ta * dz);
            }
        }
        const uint32_t nF = uint32_t(popc64(used));
        if(P.periodic && (used & 0x3Full)) {
            status |= B200VOR_CELL_BOXFACE;
        }
        // Euler's relation T = 2F - 4
        if(T > 0 && (degenerate || T + 4u != 2u * nF)) {
            status |= B200VOR_CELL_DEGENERATE;
        }
    }
    if(!alive) {
        return;
    }
    const double mass = (dead || failed) ? 0.0 : m6 / 6.0;
    P.m[row] = mass;
    P.mg[3 * size_t(row)] = (dead || failed) ? 0.0 : gx / 24.0 + mass * si.x;
    P.mg[3 * size_t(row) + 1] = (dead || failed) ? 0.0 : gy / 24.0 + mass * si.y;
    P.mg[3 * size_t(row) + 2] = (dead || failed) ? 0.0 : gz / 24.0 + mass * si.z;
    P.status[row] = uint8_t(status);
    if(FACES) {
        // face list: ids of the contributing planes, ascending (ConvexCell::v_global_index; periodic
        // images as real + n * instance, periodic.h:95-99), areas from the orthoschemes:
        // area_f = (6 vol_f) / (2 h_f), h_f = distance seed - plane = -lam_f |n_f|
        uint32_t nf = 0;
        if(!dead && !failed) {
            unsigned long long rest = used;
            while(rest) {
                const uint32_t p = uint32_t(ctz64(rest));
                rest &= rest - 1ull;
                uint32_t id;
                if(p < 6u) {
                    id = B200VOR_BOX_ID + p;
                } else {
                    const uint32_t g = M.g(p), im = g >> 27;
                    const uint32_t inst = translation_to_instance(int(im % 3u) - 1, int((im / 3u) % 3u) - 1, int(im / 9u) - 1);
                    id = ldu(P.perm + (g & 0x7FFFFFFu)) + inst * P.n;
                }
                // ids staged in the pool / conflict-list words (dead by now): entry nf
                if(nf < uint32_t(POOLCAP)) {
                    M.pl(nf) = id;
                } else {
                    M.l(nf - uint32_t(POOLCAP)) = id;
                }
                ++nf;
            }
            static_assert(PC <= POOLCAP + LCAP, "face ids are staged in the pool and conflict-list words");
            auto fid = [&](uint32_t k) -> uint32_t { return k < uint32_t(POOLCAP) ? M.pl(k) : M.l(k - uint32_t(POOLCAP)); };
            rest = used;
            uint32_t k = 0;
            while(rest) {
                const uint32_t p = uint32_t(ctz64(rest));
                rest &= rest - 1ull;
                const uint32_t id = fid(k);
                uint32_t rk = 0;
                for(uint32_t q = 0; q < nf; ++q) {
                    rk += fid(q) < id ? 1u : 0u;
                }
                if(rk < P.fcap) {
                    const PlaneD Pp = slot_plane<STRIDE>(P, F, M, p);
                    const double nn = sqrt(Pp.a * Pp.a + Pp.b * Pp.b + Pp.c * Pp.c);
                    const double h2 = -2.0 * Pp.lam * nn;
                    P.face_nbr[size_t(row) * P.fcap + rk] = id;
                    P.face_area[size_t(row) * P.fcap + rk] = h2 != 0.0 ? fabs(facc(p) / h2) : 0.0;
                }
                ++k;
            }
        }
        P.nfaces[row] = nf;
    }
    if(EV) {
        const uint32_t cP = nP < 0xffffu ? nP : 0xffffu, cV = nV < 0xfffu ? nV : 0xfffu, cS = nscan < 15u ? nscan : 15u;
        P.evcount[2 * size_t(row)] = cP | (cV << 16) | (cS << 28);
        P.evcount[2 * size_t(row) + 1] = (nC < 0xfffffu ? nC : 0xfffffu) | (T << 20);
    }
}

}  // namespace tpc
}  // namespace b200vor
