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
//  * A TEAM of G lanes (G = 16: two cells per warp) owns one cell.  All team collectives
//    are *_sync intrinsics on the team's member mask, so the two teams of a warp run
//    converged when their control flow agrees and diverge (hardware-managed) when not.
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
//    covers that radius; otherwise the block grows by one ring.
//  * Exactness: a sign test whose |value| is inside its forward error bound never cuts;
//    the cell is flagged B200VOR_CELL_UNCERTAIN (same rule as cell.cu).  Cells that exceed
//    a fixed capacity (slots, planes, ring limit) are flagged and re-done by the legacy
//    kNN-table pipeline (capi.cu), never silently accepted.
#include "common.cuh"

#include <stdlib.h>

namespace b200vor {

namespace {

#ifndef B200VOR_FUSED_MINBLOCKS
#define B200VOR_FUSED_MINBLOCKS 2
#endif
#ifndef B200VOR_FUSED_G
#define B200VOR_FUSED_G 16     // lanes per cell
#endif
#ifndef B200VOR_FUSED_WPB
#define B200VOR_FUSED_WPB 8    // warps per block (8 x 2 cells share their neighbourhood in L1)
#endif
constexpr int NT = 48;      // cell-vertex (dual triangle) slots
constexpr int PC = 40;      // plane slots (0..5 = clipping box)
constexpr int POOL = 64;    // candidate pool
constexpr int FC = 40;      // faces of a finished cell that the twin table can index
constexpr int kMaxRing = 4;
constexpr double kEps = 1.1102230246251565e-16;   // 2^-53
constexpr uint32_t kImg0 = 1u | (1u << 2) | (1u << 4);   // packed zero translation
constexpr uint32_t kImgBox = 0xFFu;
constexpr unsigned long long kLiveInit = 0x3Full | ~((1ull << PC) - 1ull);

struct FusedParams {
    GridDesc G;
    const Seed* sorted;
    const uint32_t* cell_start;
    const uint32_t* perm;
    uint32_t n;
    int periodic;
    double period;
    double box[6];
    int has_w;
    double wmax;
    uint32_t nq;
    const uint32_t* qlist;   // grid-order indices of the cells to compute, or NULL (= all)
    double* m;
    double* mg;
    double* energy;          // NULL, or the per-cell integral of |x - seed|^2
    uint8_t* status;
    uint32_t* nfaces;
    uint32_t* face_nbr;
    double* face_area;
    uint32_t fcap;
    uint32_t* evcount;
    float target;            // expected candidates of the first shell
};

template <int G> struct __align__(16) TeamShared {
    double vx[NT][4];        // (X,Y,Z,W), W > 0; Cartesian (x,y,z,-) during integration
    double pl[PC][4];        // half-space a x + b y + c z + d >= 0, seed-centred frame
    // ---- from here to `pool_img` the area is re-used as the FC x FC twin table ----
    float ve[NT];            // forward error coefficient of the vertex
    float vr[NT];            // |v|^2, rounded up
    uint32_t mark[PC][2];    // directed dual edges of the conflict zone (bit matrix)
    union {
        struct {
            uint32_t pool_key[POOL];   // (d2 as float, rounded down) & ~63 | pool index
            uint32_t desc[G][5];       // scan rows: prefix end, a0, n0, a1, images
        } scan;
        struct {
            double cand[G][4];         // staged planes of the current batch of candidates
            float cand_m[G];           // max(|normal|max, |d|) of the plane, rounded up
            float cand_n[G];           // max |normal component| of the plane, rounded up
        } stage;
        uint32_t pad_[144];            // keeps the overlay below >= FC*FC bytes for G < 16
    } u;
    uint32_t pool_j[POOL];
    uint8_t pool_img[POOL];
    // ----
    double cand_d2[G];
    uint32_t cand_j[G];
    uint32_t tri[NT];        // p0 | p1 << 8 | p2 << 16
    uint32_t ctri[G];        // `tri` of the conflict vertices of a cut, compacted onto lanes
    uint32_t gid[PC];        // grid-order index of the plane's neighbour seed
    float pnm[PC];           // max |normal component| of the plane, rounded up
    uint16_t rec[NT];        // border edges of a cut: a | b << 8
    uint8_t order[POOL];     // pool entries in ascending distance
    uint8_t gimg[PC];        // packed periodic image of the neighbour, kImgBox = box plane
    uint8_t cslot[NT];       // conflict slots of a cut, in slot order (= holes)
    uint8_t cand_img[G];
};

constexpr uint32_t kFull = 0xffffffffu;

// Team collectives in LOCKSTEP form: every lane of the warp executes them with the full
// mask; a team only looks at its own G-lane segment.
template <int G> struct Team {
    uint32_t tbase, lt;
    static constexpr uint32_t kGMask = (G == 32) ? 0xffffffffu : ((1u << (G & 31)) - 1u);
    __device__ inline uint32_t ballot(bool p) const {
        return (__ballot_sync(kFull, p) >> tbase) & kGMask;
    }
    __device__ inline bool any(bool p) const { return ballot(p) != 0u; }
    __device__ inline uint32_t lt_mask() const { return (1u << lt) - 1u; }
    template <class T> __device__ inline T bcast(T v, uint32_t src) const {
        return __shfl_sync(kFull, v, int(src), G);
    }
    __device__ inline double sum(double v) const {
#pragma unroll
        for(int o = G / 2; o > 0; o >>= 1) {
            v += __shfl_xor_sync(kFull, v, o, G);
        }
        return v;
    }
    __device__ inline uint32_t max_u32(uint32_t v) const {
#pragma unroll
        for(int o = G / 2; o > 0; o >>= 1) {
            v = max(v, __shfl_xor_sync(kFull, v, o, G));
        }
        return v;
    }
    __device__ inline uint32_t or_u32(uint32_t v) const {
#pragma unroll
        for(int o = G / 2; o > 0; o >>= 1) {
            v |= __shfl_xor_sync(kFull, v, o, G);
        }
        return v;
    }
    // exclusive prefix sum over the team, total in `total`
    __device__ inline uint32_t exscan(uint32_t v, uint32_t& total) const {
        uint32_t inc = v;
#pragma unroll
        for(int o = 1; o < G; o <<= 1) {
            uint32_t t = __shfl_up_sync(kFull, inc, o, G);
            if(lt >= uint32_t(o)) {
                inc += t;
            }
        }
        total = __shfl_sync(kFull, inc, G - 1, G);
        return inc - v;
    }
};

__device__ inline bool wany(bool p) { return __any_sync(kFull, p) != 0; }

// translation (-period, 0, +period) of one axis of a packed periodic image (2 bits: 0,1,2)
__device__ inline double img_shift(uint32_t code, double period) {
    code &= 3u;
    return code == 1u ? 0.0 : (code == 0u ? -period : period);
}

struct Plane {
    double a, b, c, d;
};

__device__ inline Plane load_plane(const double* q) {
    const double2* p = reinterpret_cast<const double2*>(q);
    double2 ab = p[0], cd = p[1];
    Plane P;
    P.a = ab.x; P.b = ab.y; P.c = cd.x; P.d = cd.y;
    return P;
}

struct Vertex {
    double x, y, z, w;
    float e, r2;
    bool degenerate;
};

// Homogeneous intersection of three planes (Cramer in cross-product form), normalised to
// W > 0, with its forward error coefficient and |v|^2 (same arithmetic as cell.cu).
// n0,n1,n2: max |normal component| of each plane, rounded up.
__device__ inline Vertex make_vertex(const Plane& P0, const Plane& P1, const Plane& P2, float n0,
                                     float n1, float n2) {
    double a0 = P0.a, b0 = P0.b, c0 = P0.c, d0 = P0.d;
    double a1 = P1.a, b1 = P1.b, c1 = P1.c, d1 = P1.d;
    double a2 = P2.a, b2 = P2.b, c2 = P2.c, d2 = P2.d;
    double ux = b1 * c2 - c1 * b2, uy = c1 * a2 - a1 * c2, uz = a1 * b2 - b1 * a2;
    double vx = b2 * c0 - c2 * b0, vy = c2 * a0 - a2 * c0, vz = a2 * b0 - b2 * a0;
    double wx = b0 * c1 - c0 * b1, wy = c0 * a1 - a0 * c1, wz = a0 * b1 - b0 * a1;
    double w = a0 * ux + b0 * uy + c0 * uz;
    double x = -(d0 * ux + d1 * vx + d2 * wx);
    double y = -(d0 * uy + d1 * vy + d2 * wy);
    double z = -(d0 * uz + d1 * vz + d2 * wz);
    if(w < 0.0) {
        w = -w; x = -x; y = -y; z = -z;
    }
    Vertex V;
    V.x = x; V.y = y; V.z = z; V.w = w;
    float D0 = __double2float_ru(fabs(d0)), D1 = __double2float_ru(fabs(d1)),
          D2 = __double2float_ru(fabs(d2));
    float Q = D0 * n1 * n2 + D1 * n2 * n0 + D2 * n0 * n1;
    float sx = __double2float_ru(fabs(x) + fabs(y) + fabs(z));
    float sw = __double2float_ru(w);
    const float e48 = float(48.0 * kEps) * 1.001f, e4 = float(4.0 * kEps) * 1.001f;
    float n012 = n0 * n1 * n2;
    float ea = (e48 * Q + e4 * sx) * 1.001f;
    float eb = (e48 * n012 + e4 * sw) * 1.001f;
    // one coefficient: nmax*ea + |d|*eb <= max(nmax,|d|) * (ea + eb)
    V.e = (ea + eb) * 1.0001f;
    V.degenerate = !(sw > 2.0f * e48 * n012);
    float num = __double2float_ru(x * x + y * y + z * z);
    float den = __double2float_rd(w * w);
    V.r2 = __fdiv_ru(num, den);
    return V;
}

template <int G>
__device__ inline void store_vertex(TeamShared<G>* sh, uint32_t s, const Vertex& V, uint32_t t) {
    double2* p = reinterpret_cast<double2*>(sh->vx[s]);
    p[0] = make_double2(V.x, V.y);
    p[1] = make_double2(V.z, V.w);
    sh->ve[s] = V.e;
    sh->vr[s] = V.r2;
    sh->tri[s] = t;
}

// In-register bitonic sort of RK*G keys, element e = r*G + lane.
template <int G, int RK>
__device__ inline void bitonic_sort(uint32_t lt, uint32_t (&key)[RK]) {
    constexpr int N = RK * G;
#pragma unroll
    for(int k = 2; k <= N; k <<= 1) {
#pragma unroll
        for(int j = k >> 1; j > 0; j >>= 1) {
            if(j >= G) {
                const int jr = j / G;
#pragma unroll
                for(int r = 0; r < RK; ++r) {
                    if((r & jr) == 0) {
                        bool asc = (((r * G) & k) == 0);
                        uint32_t lo = min(key[r], key[r | jr]), hi = max(key[r], key[r | jr]);
                        key[r] = asc ? lo : hi;
                        key[r | jr] = asc ? hi : lo;
                    }
                }
            } else {
#pragma unroll
                for(int r = 0; r < RK; ++r) {
                    uint32_t other = __shfl_xor_sync(kFull, key[r], j, G);
                    uint32_t e = uint32_t(r * G) + lt;
                    bool keep_min = (((lt & uint32_t(j)) == 0) == ((e & uint32_t(k)) == 0));
                    key[r] = keep_min ? min(key[r], other) : max(key[r], other);
                }
            }
        }
    }
}

// squared distance from the seed to the nearest face of the scanned block that has
// unscanned space behind it (1e300 when the block covers the whole grid)
__device__ inline double coverage2(const GridDesc& Gd, double sx, double sy, double sz, int cx,
                                   int cy, int cz, int r) {
    double dmin = 1e300;
    const double q[3] = {sx, sy, sz};
    const int cc[3] = {cx, cy, cz};
#pragma unroll
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

template <int G, bool FACES, int WPB, bool EV>
__global__ void __launch_bounds__(32 * WPB, B200VOR_FUSED_MINBLOCKS) cells_fused_kernel(FusedParams P) {
    constexpr int TPW = 32 / G;
    constexpr int R = NT / G;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const uint32_t lane = threadIdx.x & 31u, wid = threadIdx.x >> 5;
    const uint32_t tiw = lane / G;
    Team<G> tm;
    tm.lt = lane % G;
    tm.tbase = tiw * G;
    const uint32_t lt = tm.lt;
    TeamShared<G>* sh = reinterpret_cast<TeamShared<G>*>(smem_raw) + (wid * TPW + tiw);
    const uint32_t c = (blockIdx.x * WPB + wid) * TPW + tiw;
    // a team without a cell (tail of the grid) computes cell 0 again and writes nothing:
    // the warp stays converged on every collective
    const bool alive = c < P.nq;
    const uint32_t i = alive ? (P.qlist ? P.qlist[c] : c) : (P.qlist ? P.qlist[0] : 0u);
    const Seed si = load_seed(P.sorted + i);
    const uint32_t row = P.perm[i];
    const GridDesc& Gd = P.G;

    // ------------------------------------------------------------------ init_with_box
    // (convex_cell.cpp:175-215) clipping box in the seed-centred frame
    for(uint32_t p = lt; p < 6; p += G) {
        int axis = p >> 1;
        bool upper = p & 1;
        double sgn = upper ? -1.0 : 1.0;
        double lo_a = P.periodic ? -0.5 * P.period
                                 : (axis == 0 ? P.box[0] - si.x : (axis == 1 ? P.box[1] - si.y : P.box[2] - si.z));
        double hi_a = P.periodic ? 0.5 * P.period
                                 : (axis == 0 ? P.box[3] - si.x : (axis == 1 ? P.box[4] - si.y : P.box[5] - si.z));
        double2* q = reinterpret_cast<double2*>(sh->pl[p]);
        q[0] = make_double2(axis == 0 ? sgn : 0.0, axis == 1 ? sgn : 0.0);
        q[1] = make_double2(axis == 2 ? sgn : 0.0, upper ? hi_a : -lo_a);
        sh->gid[p] = B200VOR_BOX_ID + p;
        sh->gimg[p] = uint8_t(kImgBox);
        sh->pnm[p] = 1.0f;
    }
    for(uint32_t p = lt; p < uint32_t(PC); p += G) {
        sh->mark[p][0] = 0;
        sh->mark[p][1] = 0;
    }
    __syncwarp();
    bool degenerate = false;
    for(uint32_t s = lt; s < 8; s += G) {
        uint32_t bxb = s & 1, byb = (s >> 1) & 1, bzb = (s >> 2) & 1;
        uint32_t px = bxb, py = 2 + byb, pz = 4 + bzb;
        bool odd = (bxb + byb + bzb) & 1;
        uint32_t v1 = odd ? pz : py, v2 = odd ? py : pz;
        Vertex V = make_vertex(load_plane(sh->pl[px]), load_plane(sh->pl[v1]),
                               load_plane(sh->pl[v2]), 1.0f, 1.0f, 1.0f);
        store_vertex<G>(sh, s, V, px | (v1 << 8) | (v2 << 16));
    }
    __syncwarp();
    uint32_t T = 8;                       // live slots [0,T)
    unsigned long long live = kLiveInit;  // allocated plane slots (ids >= PC never free)
    uint32_t status = 0, extra = 0;
    uint32_t nP = 0, nC = 0, nV = 0, nscan = 0;
    float R2;
    {
        uint32_t rb = lt < 8 ? __float_as_uint(sh->vr[lt]) : 0u;
        R2 = __uint_as_float(tm.max_u32(rb));
    }

    // ------------------------------------------------------------------ neighbour walk
    int cx, cy, cz;
    cell_coords(Gd, si.x, si.y, si.z, cx, cy, cz);
    const double dw = P.has_w ? (P.wmax - si.w) : 0.0;
    int ring = 1, rmax;
    if(Gd.periodic) {
        int gmin = min(Gd.g[0], min(Gd.g[1], Gd.g[2]));
        rmax = min(kMaxRing, (gmin - 1) / 2);
    } else {
        rmax = max(max(cx, Gd.g[0] - 1 - cx), max(max(cy, Gd.g[1] - 1 - cy), max(cz, Gd.g[2] - 1 - cz)));
        rmax = max(min(rmax, kMaxRing), 1);
    }
    bool certified = false;
    bool done = rmax < 1;     // periodic grid too coarse for a 3x3x3 block: legacy pipeline
    double cov2 = done ? 0.0 : coverage2(Gd, si.x, si.y, si.z, cx, cy, cz, ring);
    double lo = 0.0, hi = 0.0;
    bool first = true, retry = false;
    int shrink = 0;
    uint32_t cnt = 0, pos = 0;
    // d2 beyond which no bisector can cut the cell (generic_RVD.h:2322; Laguerre: a power
    // plane cuts only if |q| <= R + sqrt(R^2 + wmax - wi))
    auto sr_bound = [&](float r2) -> double {
        double b = 4.1 * double(r2);
        if(P.has_w) {
            double Rr = sqrt(double(r2));
            double sr = Rr + sqrt(double(r2) + dw);
            b = 1.025 * sr * sr;
        }
        return b;
    };
    double bnd = sr_bound(R2);

    while(wany(!done)) {
        // ---- next distance shell [lo,hi) for the teams whose pool is used up ----
        const bool need = !done && pos >= cnt;
        bool auto_hi = false;
        if(need) {
            if(retry) {
                // more than POOL candidates in the shell: halve it
                hi = lo + (hi - lo) * 0.5;
                if(++shrink > 48) {
                    done = true;   // > POOL candidates at (nearly) one distance: legacy pipeline
                }
            } else {
                lo = hi;
                for(;;) {
                    const double bound = bnd;
                    if(lo > bound) {
                        certified = true;
                        done = true;
                        break;
                    }
                    if(lo >= cov2) {
                        if(ring >= rmax) {
                            // no face of the block has unscanned space behind it: all seen
                            certified = (cov2 >= 1e300);
                            done = true;
                            break;
                        }
                        ++ring;
                        cov2 = coverage2(Gd, si.x, si.y, si.z, cx, cy, cz, ring);
                        continue;
                    }
                    hi = fmin(bound * (1.0 + 1e-12) + 1e-300, cov2);
                    auto_hi = first;
                    break;
                }
            }
        }
        const bool sc = need && !done;
        if(wany(sc)) {
            // ================================================================ scan
            // candidates of the (2 ring + 1)^3 block of bins with lo <= d2 < hi -> pool
            const int R1 = 2 * ring + 1;
            const int nrows = R1 * R1;
            uint32_t count = 0;
            for(int rb = 0; wany(sc && rb < nrows); rb += G) {
                const bool rowact = sc && rb < nrows;
                int rowi = rb + int(lt);
                uint32_t a0 = 0, n0 = 0, a1 = 0, n1 = 0, img = 0;
                uint32_t nbins = 0;
                if(rowact && rowi < nrows) {
                    int dy = rowi % R1 - ring, dz = rowi / R1 - ring;
                    int y = cy + dy, z = cz + dz, sy = 0, sz = 0;
                    // the row of bins is scanned only where it meets the ball d2 < hi
                    const double slack = 1e-9 * Gd.w[0];
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
                        xa = max(xa, int(floor((si.x - rx - Gd.lo[0]) * Gd.inv_w[0])));
                        xb = min(xb, int(floor((si.x + rx - Gd.lo[0]) * Gd.inv_w[0])));
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
                    if(ok) {
                        const int gx = Gd.g[0];
                        uint32_t base = uint32_t(gx) * (uint32_t(y) + uint32_t(Gd.g[1]) * uint32_t(z));
                        int ma = xa < 0 ? 0 : xa, mb = xb >= gx ? gx - 1 : xb;
                        nbins = uint32_t(xb - xa + 1);
                        if(ma <= mb) {
                            a0 = __ldg(P.cell_start + base + ma);
                            n0 = __ldg(P.cell_start + base + mb + 1) - a0;
                        }
                        int sx1 = 0;
                        if(Gd.periodic) {
                            if(xa < 0) {
                                sx1 = -1;
                                a1 = __ldg(P.cell_start + base + (xa + gx));
                                n1 = __ldg(P.cell_start + base + gx) - a1;
                            } else if(xb >= gx) {
                                sx1 = 1;
                                a1 = __ldg(P.cell_start + base);
                                n1 = __ldg(P.cell_start + base + (xb - gx) + 1) - a1;
                            }
                        }
                        img = uint32_t(1) | (uint32_t(sy + 1) << 2) | (uint32_t(sz + 1) << 4) |
                              (uint32_t(sx1 + 1) << 8);
                    }
                }
                uint32_t total;
                uint32_t pre = tm.exscan(n0 + n1, total);
                sh->u.scan.desc[lt][0] = pre + n0 + n1;
                sh->u.scan.desc[lt][1] = a0;
                sh->u.scan.desc[lt][2] = n0;
                sh->u.scan.desc[lt][3] = a1;
                sh->u.scan.desc[lt][4] = img;
                if(wany(auto_hi && rb == 0)) {
                    // first shell: radius of the ball expected to hold P.target seeds at
                    // the density of the scanned bins
                    uint32_t bins = nbins;
#pragma unroll
                    for(int o = G / 2; o > 0; o >>= 1) {
                        bins += __shfl_xor_sync(kFull, bins, o, G);
                    }
                    double vol = double(bins) * Gd.w[0] * Gd.w[1] * Gd.w[2];
                    if(auto_hi && rb == 0 && total > 0) {
                        float r3 = P.target * 0.2387324146378430f * float(vol) / float(total);
                        float rr = cbrtf(r3);
                        hi = fmin(double(rr * rr), hi);
                    }
                }
                __syncwarp();
                uint32_t cur = 0, cur_end = sh->u.scan.desc[0][0], cur_beg = 0;
                for(uint32_t c0 = 0; wany(rowact && c0 < total); c0 += G) {
                    uint32_t cc = c0 + lt;
                    bool in = false;
                    uint32_t j = 0, im = 0;
                    double d2 = 0.0;
                    if(rowact && cc < total) {
                        while(cc >= cur_end) {
                            cur_beg = cur_end;
                            ++cur;
                            cur_end = sh->u.scan.desc[cur][0];
                        }
                        uint32_t off = cc - cur_beg;
                        uint32_t dn0 = sh->u.scan.desc[cur][2];
                        uint32_t dimg = sh->u.scan.desc[cur][4];
                        if(off < dn0) {
                            j = sh->u.scan.desc[cur][1] + off;
                            im = dimg & 0x3Fu;
                        } else {
                            j = sh->u.scan.desc[cur][3] + (off - dn0);
                            im = (dimg & 0x3Cu) | ((dimg >> 8) & 3u);
                        }
                        Seed sj = load_seed(P.sorted + j);
                        double qx = (sj.x + img_shift(im, P.period)) - si.x;
                        double qy = (sj.y + img_shift(im >> 2, P.period)) - si.y;
                        double qz = (sj.z + img_shift(im >> 4, P.period)) - si.z;
                        d2 = qx * qx + qy * qy + qz * qz;
                        in = d2 >= lo && d2 < hi && !(j == i && im == kImg0);
                    }
                    uint32_t bal = tm.ballot(in);
                    if(in) {
                        uint32_t ppos = count + __popc(bal & tm.lt_mask());
                        if(ppos < uint32_t(POOL)) {
                            uint32_t kb = __float_as_uint(__double2float_rd(d2));
                            sh->u.scan.pool_key[ppos] = (kb & ~63u) | ppos;
                            sh->pool_j[ppos] = j;
                            sh->pool_img[ppos] = uint8_t(im);
                        }
                    }
                    count += __popc(bal);
                }
                __syncwarp();
            }
            if(sc) {
                if(EV) {
                    ++nscan;
                }
                first = false;
                pos = 0;
                if(count > uint32_t(POOL)) {
                    retry = true;
                    cnt = 0;
                } else {
                    retry = false;
                    shrink = 0;
                    cnt = count;
                }
            }
            // ================================================================ sort
            const bool so = sc && !retry;
            constexpr int RP = POOL / G;
            uint32_t key[RP];
#pragma unroll
            for(int r = 0; r < RP; ++r) {
                uint32_t e = lt + G * r;
                key[r] = (so && e < cnt) ? sh->u.scan.pool_key[e] : 0xffffffffu;
            }
            if(wany(so && cnt > uint32_t(G))) {
                bitonic_sort<G, RP>(lt, key);
            } else {
                uint32_t k1[1] = {key[0]};
                bitonic_sort<G, 1>(lt, k1);
                key[0] = k1[0];
            }
            if(so) {
#pragma unroll
                for(int r = 0; r < RP; ++r) {
                    uint32_t e = lt + G * r;
                    if(e < cnt) {
                        sh->order[e] = uint8_t(key[r] & 63u);
                    }
                }
            }
            __syncwarp();
        }

        // ==================================================================== clip
        while(wany(!done && pos < cnt)) {
            const bool A = !done && pos < cnt;
            const uint32_t b0 = pos, bend = min(pos + uint32_t(G), cnt);
            // ---- stage the planes of the next G candidates (ascending distance) ----
            const bool mine = A && b0 + lt < bend;
            Plane pq;
            pq.a = pq.b = pq.c = pq.d = 0.0;
            double myd2 = 1e300;
            float Mq = 0.f;
            bool dup = false;
            if(mine) {
                uint32_t idx = sh->order[b0 + lt];
                uint32_t j = sh->pool_j[idx];
                uint32_t im = sh->pool_img[idx];
                Seed sj = load_seed(P.sorted + j);
                double qx = (sj.x + img_shift(im, P.period)) - si.x;
                double qy = (sj.y + img_shift(im >> 2, P.period)) - si.y;
                double qz = (sj.z + img_shift(im >> 4, P.period)) - si.z;
                myd2 = qx * qx + qy * qy + qz * qz;
                // plane (i,j) = (-2q, |q|^2 + wi - wj) (periodic_delaunay_3d.cpp:3719-3726,
                // seed-centred frame)
                pq.a = -2.0 * qx; pq.b = -2.0 * qy; pq.c = -2.0 * qz;
                pq.d = myd2 + (si.w - sj.w);
                Mq = __double2float_ru(fmax(fmax(fabs(pq.a), fabs(pq.b)), fmax(fabs(pq.c), fabs(pq.d)))) * 1.0001f;
                double2* cp = reinterpret_cast<double2*>(sh->u.stage.cand[lt]);
                cp[0] = make_double2(pq.a, pq.b);
                cp[1] = make_double2(pq.c, pq.d);
                sh->u.stage.cand_m[lt] = Mq;
                sh->u.stage.cand_n[lt] =
                    __double2float_ru(fmax(fabs(pq.a), fmax(fabs(pq.b), fabs(pq.c))));
                sh->cand_d2[lt] = myd2;
                sh->cand_j[lt] = j;
                sh->cand_img[lt] = uint8_t(im);
                if(myd2 == 0.0) {
                    // exact duplicate (delaunay_nn.cpp:122-135): the lower user index
                    // keeps the cell; the duplicate itself is never a neighbour
                    dup = im == kImg0 && __ldg(P.perm + j) < row;
                }
            }
            // ---- pre-filter, lanes = candidate planes: a plane that cuts nothing of the
            //      current cell can never cut the (smaller) later cell ----
            bool cuts = false, tie = false;
            {
                const bool tst = mine && myd2 != 0.0 && myd2 <= bnd;
                const uint32_t Tp = __reduce_max_sync(kFull, tst ? T : 0u);
#pragma unroll 2
                for(uint32_t s = 0; s < Tp; ++s) {
                    if(tst && s < T) {
                        const double2* p = reinterpret_cast<const double2*>(sh->vx[s]);
                        double2 xy = p[0], zw = p[1];
                        double dot = fma(pq.a, xy.x, fma(pq.b, xy.y, fma(pq.c, zw.x, pq.d * zw.y)));
                        float E = Mq * sh->ve[s];
                        // conflict only when CERTAINLY outside; a tie leaves the vertex
                        // alone (reference fast mode, convex_cell.cpp:968) and flags the cell
                        bool t1 = __double2float_rd(fabs(dot)) <= E;
                        cuts |= !t1 && (dot < 0.0);
                        tie |= t1;
                    }
                }
                const uint32_t nt = __popc(tm.ballot(tst));
                if(EV) {
                    nP += nt;
                    nC += nt * T;
                }
            }
            uint32_t flags = tm.ballot(cuts);
            // candidates beyond the security radius end the walk (sorted: a suffix)
            const uint32_t beyond = tm.ballot(mine && myd2 > bnd);
            if(tm.any(tie)) {
                status |= B200VOR_CELL_UNCERTAIN;
            }
            if(tm.any(dup)) {
                extra |= B200VOR_CELL_DUPLICATE;
            }
            __syncwarp();
            for(;;) {
                // ---- next plane of the batch that may cut ----
                bool F = A && !done && flags != 0u;
                if(!wany(F)) {
                    break;
                }
                uint32_t fk = 0;
                if(F) {
                    fk = uint32_t(__ffs(int(flags))) - 1u;
                    flags &= flags - 1u;
                }
                // ================================================================ cut
                // (triangulate_conflict_zone, convex_cell.cpp:761-865, without adjacency)
                const Plane pe = load_plane(sh->u.stage.cand[fk]);
                // ---- sign tests, lanes = vertices; the same pass refreshes the squared
                //      radius of the current cell (convex_cell.cpp:1312-1330) ----
                const uint32_t Tw = __reduce_max_sync(kFull, F ? T : 0u);
                uint32_t cm[R], tr[R];
                uint32_t cfbits = 0, nconf = 0;
                {
                    const float M = sh->u.stage.cand_m[fk];
                    bool unc = false;
                    uint32_t rb = 0;
#pragma unroll
                    for(int r = 0; r < R; ++r) {
                        cm[r] = 0;
                        tr[r] = 0;
                        if(uint32_t(G * r) < Tw) {   // warp-uniform: rounds past the last vertex are skipped
                            uint32_t s = lt + G * r;
                            bool cf = false;
                            if(F && s < T) {
                                const double2* p = reinterpret_cast<const double2*>(sh->vx[s]);
                                double2 xy = p[0], zw = p[1];
                                double dot = fma(pe.a, xy.x, fma(pe.b, xy.y, fma(pe.c, zw.x, pe.d * zw.y)));
                                float E = M * sh->ve[s];
                                bool t1 = __double2float_rd(fabs(dot)) <= E;
                                cf = !t1 && (dot < 0.0);
                                unc |= t1;
                                rb = max(rb, __float_as_uint(sh->vr[s]));
                                if(cf) {
                                    tr[r] = sh->tri[s];
                                }
                            }
                            cm[r] = tm.ballot(cf);
                            cfbits |= cf ? (1u << r) : 0u;
                            nconf += __popc(cm[r]);
                        }
                    }
                    rb = tm.max_u32(rb);
                    const bool anyunc = tm.any(unc);
                    if(F) {
                        R2 = __uint_as_float(rb);
                        bnd = sr_bound(R2);
                        if(sh->cand_d2[fk] > bnd) {
                            certified = true;   // security radius reached
                            done = true;
                            F = false;
                        }
                    }
                    if(anyunc && F) {
                        status |= B200VOR_CELL_UNCERTAIN;
                    }
                    if(EV) {
                        nC += F ? T : 0u;
                    }
                }
                if(F && nconf == 0) {
                    F = false;   // the cell shrank since the pre-filter: no longer cut
                }
                if(F && nconf == T) {
                    T = 0;
                    status |= B200VOR_CELL_EMPTY;
                    R2 = 0.f;
                    certified = true;
                    done = true;
                    F = false;
                }
                if(F && nconf > uint32_t(G)) {
                    // more conflict vertices than lanes: left to the kNN-table pipeline
                    status |= B200VOR_CELL_OVERFLOW;
                    done = true;
                    F = false;
                }
                if(wany(F && live == ~0ull)) {
                    // plane table full: drop the planes no vertex refers to any more
                    uint32_t ulo = 0, uhi = 0;
#pragma unroll
                    for(int r = 0; r < R; ++r) {
                        uint32_t s = lt + G * r;
                        if(s < T) {
                            uint32_t t = sh->tri[s];
#pragma unroll
                            for(int e = 0; e < 3; ++e) {
                                uint32_t v = (t >> (8 * e)) & 0xffu;
                                if(v < 32) {
                                    ulo |= 1u << v;
                                } else {
                                    uhi |= 1u << (v - 32);
                                }
                            }
                        }
                    }
                    ulo = tm.or_u32(ulo);
                    uhi = tm.or_u32(uhi);
                    if(F && live == ~0ull) {
                        live = ((unsigned long long)ulo | ((unsigned long long)uhi << 32)) | kLiveInit;
                        if(live == ~0ull) {
                            status |= B200VOR_CELL_OVERFLOW;
                            done = true;
                            F = false;
                        }
                    }
                }
                const uint32_t pn = uint32_t(__ffsll((long long)~live)) - 1u;
                // ---- conflict vertices compacted onto lanes, in slot order (= the holes the
                //      new vertices fill) ----
                {
                    uint32_t hbase = 0;
#pragma unroll
                    for(int r = 0; r < R; ++r) {
                        if(F && ((cfbits >> r) & 1u)) {
                            uint32_t rk = hbase + __popc(cm[r] & tm.lt_mask());
                            sh->cslot[rk] = uint8_t(lt + G * r);
                            sh->ctri[rk] = tr[r];
                        }
                        hbase += __popc(cm[r]);
                    }
                }
                __syncwarp();
                // ---- mark the directed edges of the conflict triangles ----
                const bool Q = F && lt < nconf;
                uint32_t p0 = 0, p1 = 0, p2 = 0;
                if(Q) {
                    uint32_t t = sh->ctri[lt];
                    p0 = t & 0xffu; p1 = (t >> 8) & 0xffu; p2 = (t >> 16) & 0xffu;
                    atomicOr(&sh->mark[p0][p1 >> 5], 1u << (p1 & 31u));
                    atomicOr(&sh->mark[p1][p2 >> 5], 1u << (p2 & 31u));
                    atomicOr(&sh->mark[p2][p0 >> 5], 1u << (p0 & 31u));
                }
                __syncwarp();
                // ---- border edges (twin not marked) -> records ----
                uint32_t bits = 0;
                if(Q) {
                    bits = (((sh->mark[p1][p0 >> 5] >> (p0 & 31u)) & 1u) ^ 1u) |
                           ((((sh->mark[p2][p1 >> 5] >> (p1 & 31u)) & 1u) ^ 1u) << 1) |
                           ((((sh->mark[p0][p2 >> 5] >> (p2 & 31u)) & 1u) ^ 1u) << 2);
                }
                uint32_t nnew;
                {
                    uint32_t rk = tm.exscan(__popc(bits), nnew);
                    if(F && nnew > uint32_t(G)) {
                        // a new face with more vertices than lanes: kNN-table pipeline
                        status |= B200VOR_CELL_OVERFLOW;
                        done = true;
                        F = false;
                    }
                    if(F) {
                        if(bits & 1u) {
                            sh->rec[rk++] = uint16_t(p0 | (p1 << 8));
                        }
                        if(bits & 2u) {
                            sh->rec[rk++] = uint16_t(p1 | (p2 << 8));
                        }
                        if(bits & 4u) {
                            sh->rec[rk] = uint16_t(p2 | (p0 << 8));
                        }
                    }
                }
                __syncwarp();
                // the whole bit matrix is cleared with two vector stores per lane
                for(uint32_t k = lt; k < uint32_t(PC * 8 / 16); k += G) {
                    reinterpret_cast<uint4*>(&sh->mark[0][0])[k] = make_uint4(0u, 0u, 0u, 0u);
                }
                uint32_t Tnew = T - nconf + nnew;
                if(F && Tnew > uint32_t(NT)) {
                    status |= B200VOR_CELL_OVERFLOW;
                    done = true;
                    F = false;
                }
                const float npn = sh->u.stage.cand_n[fk];
                if(F) {
                    live |= 1ull << pn;
                    if(lt == 0) {
                        double2* p = reinterpret_cast<double2*>(sh->pl[pn]);
                        p[0] = make_double2(pe.a, pe.b);
                        p[1] = make_double2(pe.c, pe.d);
                        sh->gid[pn] = sh->cand_j[fk];
                        sh->gimg[pn] = sh->cand_img[fk];
                        sh->pnm[pn] = npn;
                    }
                }
                // ---- fewer new vertices than holes (two faces vanished at once): the
                //      holes left below Tnew take the survivors stored at or above Tnew ----
                if(wany(F && nnew < nconf)) {
                    const bool FR = F && nnew < nconf;
                    // open holes: rank >= nnew, slot < Tnew (cslot is in slot order)
                    // survivors to move: slots in [Tnew, T) that are not holes
                    uint32_t mvbase = 0;
                    uint32_t dst[R];
#pragma unroll
                    for(int r = 0; r < R; ++r) {
                        uint32_t s = lt + G * r;
                        bool mover = FR && s >= Tnew && s < T && !((cfbits >> r) & 1u);
                        uint32_t mb = tm.ballot(mover);
                        dst[r] = mover ? mvbase + __popc(mb & tm.lt_mask()) : 0xffffffffu;
                        mvbase += __popc(mb);
                    }
                    double2 mxy[R], mzw[R];
                    float me[R], mr[R];
                    uint32_t mt[R];
#pragma unroll
                    for(int r = 0; r < R; ++r) {
                        if(dst[r] != 0xffffffffu) {
                            uint32_t s = lt + G * r;
                            const double2* p = reinterpret_cast<const double2*>(sh->vx[s]);
                            mxy[r] = p[0]; mzw[r] = p[1];
                            me[r] = sh->ve[s]; mr[r] = sh->vr[s]; mt[r] = sh->tri[s];
                        }
                    }
                    __syncwarp();
#pragma unroll
                    for(int r = 0; r < R; ++r) {
                        if(dst[r] != 0xffffffffu) {
                            // the (nnew + k)-th hole is the k-th open one: holes below it
                            // are all < Tnew as well
                            uint32_t s = sh->cslot[nnew + dst[r]];
                            double2* p = reinterpret_cast<double2*>(sh->vx[s]);
                            p[0] = mxy[r]; p[1] = mzw[r];
                            sh->ve[s] = me[r]; sh->vr[s] = mr[r]; sh->tri[s] = mt[r];
                        }
                    }
                }
                // ---- one new vertex per border edge, one lane each ----
                if(F && lt < nnew) {
                    uint32_t rr = sh->rec[lt];
                    uint32_t u = rr & 0xffu, v = rr >> 8;
                    uint32_t s = lt < nconf ? uint32_t(sh->cslot[lt]) : T + (lt - nconf);
                    Vertex V = make_vertex(pe, load_plane(sh->pl[u]), load_plane(sh->pl[v]), npn,
                                           sh->pnm[u], sh->pnm[v]);
                    degenerate |= V.degenerate;
                    store_vertex<G>(sh, s, V, pn | (u << 8) | (v << 16));
                }
                if(F) {
                    if(EV) {
                        nV += nnew;
                    }
                    T = Tnew;
                }
                __syncwarp();
            }
            // ---- squared radius of the cell as the batch left it ----
            {
                const bool Aa = A && !done;
                const uint32_t Tw = __reduce_max_sync(kFull, Aa ? T : 0u);
                uint32_t rb = 0;
#pragma unroll
                for(int r = 0; r < R; ++r) {
                    uint32_t s = lt + G * r;
                    if(uint32_t(G * r) < Tw && Aa && s < T) {
                        rb = max(rb, __float_as_uint(sh->vr[s]));
                    }
                }
                rb = tm.max_u32(rb);
                if(Aa) {
                    R2 = __uint_as_float(rb);
                    bnd = sr_bound(R2);
                }
            }
            if(A) {
                pos = bend;
                if(beyond != 0u && !done) {
                    certified = true;   // every candidate up to the security radius was offered
                    done = true;
                }
            }
        }
    }
    status |= extra;
    if(!certified) {
        status |= B200VOR_CELL_UNCERTIFIED;
    }

    // ---------------------------------------------------------------- integration
    const bool dead = (status & (B200VOR_CELL_EMPTY | B200VOR_CELL_DUPLICATE)) != 0;
    unsigned long long used;
    {
        uint32_t ulo = 0, uhi = 0;
#pragma unroll
        for(int r = 0; r < R; ++r) {
            uint32_t s = lt + G * r;
            if(s < T) {
                uint32_t t = sh->tri[s];
#pragma unroll
                for(int e = 0; e < 3; ++e) {
                    uint32_t v = (t >> (8 * e)) & 0xffu;
                    if(v < 32) {
                        ulo |= 1u << v;
                    } else {
                        uhi |= 1u << (v - 32);
                    }
                }
            }
        }
        ulo = tm.or_u32(ulo);
        uhi = tm.or_u32(uhi);
        used = (unsigned long long)ulo | ((unsigned long long)uhi << 32);
    }
    const uint32_t F = uint32_t(__popcll(used));
    // foot scale -d/|n|^2 of every contributing plane, once per face: it replaces d in place
    for(uint32_t p = lt; p < uint32_t(PC); p += G) {
        if((used >> p) & 1ull) {
            Plane pf = load_plane(sh->pl[p]);
            sh->pl[p][3] = -pf.d / (pf.a * pf.a + pf.b * pf.b + pf.c * pf.c);
        }
    }
    {
        bool mine = false;
        for(uint32_t p = lt; p < uint32_t(PC); p += G) {
            mine = mine || (((used >> p) & 1ull) && sh->gimg[p] == kImgBox);
        }
        if(tm.any(mine) && P.periodic && !(status & (B200VOR_CELL_EMPTY | B200VOR_CELL_OVERFLOW))) {
            status |= B200VOR_CELL_BOXFACE;
        }
    }
    const bool tableok = F <= uint32_t(FC) && T > 0;
    if(T > 0 && F > uint32_t(FC)) {
        status |= B200VOR_CELL_OVERFLOW;
    }
    // Cartesian vertices in place
#pragma unroll
    for(int r = 0; r < R; ++r) {
        uint32_t s = lt + G * r;
        if(s < T) {
            double2* p = reinterpret_cast<double2*>(sh->vx[s]);
            double2 xy = p[0], zw = p[1];
            double iw = 1.0 / zw.y;
            p[0] = make_double2(xy.x * iw, xy.y * iw);
            p[1] = make_double2(zw.x * iw, 1.0);
        }
    }
    uint8_t* table = reinterpret_cast<uint8_t*>(sh->ve);
    static_assert(offsetof(TeamShared<G>, cand_d2) - offsetof(TeamShared<G>, ve) >= FC * FC,
                  "twin table does not fit its overlay");
    __syncwarp();
    for(uint32_t k = lt; k < uint32_t(FC * FC) / 4u; k += G) {
        reinterpret_cast<uint32_t*>(table)[k] = 0xffffffffu;
    }
    __syncwarp();
    uint32_t ci[R][3];
    bool ok = true;
#pragma unroll
    for(int r = 0; r < R; ++r) {
        uint32_t s = lt + G * r;
        ci[r][0] = ci[r][1] = ci[r][2] = 0;
        if(s < T && tableok) {
            uint32_t t = sh->tri[s];
#pragma unroll
            for(int e = 0; e < 3; ++e) {
                uint32_t p = (t >> (8 * e)) & 0xffu;
                ci[r][e] = uint32_t(__popcll(used & ((1ull << p) - 1ull)));
            }
#pragma unroll
            for(int e = 0; e < 3; ++e) {
                table[ci[r][e] * FC + ci[r][(e + 1) % 3]] = uint8_t(s);
            }
        }
    }
    __syncwarp();
    double m6 = 0.0, gx = 0.0, gy = 0.0, gz = 0.0, e6 = 0.0;
    const bool want_e = P.energy != nullptr;
    double ncr[FACES ? R : 1][3];
#pragma unroll
    for(int r = 0; r < R; ++r) {
        if(FACES) {
            ncr[FACES ? r : 0][0] = ncr[FACES ? r : 0][1] = ncr[FACES ? r : 0][2] = 0.0;
        }
        uint32_t s = lt + G * r;
        if(s < T && tableok) {
            uint32_t t = sh->tri[s];
            const double2* up = reinterpret_cast<const double2*>(sh->vx[s]);
            double2 uxy = up[0], uzw = up[1];
            double ux = uxy.x, uy = uxy.y, uz = uzw.x;
#pragma unroll
            for(int l = 0; l < 3; ++l) {
                uint32_t f = (t >> (8 * l)) & 0xffu;
                // next vertex around face t[l]: the triangle that owns the directed edge
                // (t[l] -> t[l+2]), twin of this triangle's edge (t[l+2] -> t[l])
                uint32_t nb = table[ci[r][l] * FC + ci[r][(l + 2) % 3]];
                if(nb >= T) {
                    ok = false;
                    nb = s;
                }
                const double2* vp = reinterpret_cast<const double2*>(sh->vx[nb]);
                double2 vxy = vp[0], vzw = vp[1];
                double vx = vxy.x, vy = vxy.y, vz = vzw.x;
                Plane pf = load_plane(sh->pl[f]);
                double cxp = uy * vz - uz * vy, cyp = uz * vx - ux * vz, czp = ux * vy - uy * vx;
                double nc = pf.a * cxp + pf.b * cyp + pf.c * czp;
                double inv = pf.d;   // foot scale, see above
                double vol6 = inv * nc;
                m6 += vol6;
                const double fx = pf.a * inv, fy = pf.b * inv, fz = pf.c * inv;   // foot of the face plane
                gx += vol6 * (fx + ux + vx);
                gy += vol6 * (fy + uy + vy);
                gz += vol6 * (fz + uz + vz);
                if(want_e) {
                    // RVD.cpp:846-857 on the simplex (seed, foot, u, v), seed at the origin:
                    // U.U + V.V + W.W + U.V + V.W + W.U
                    double q = (fx * fx + fy * fy + fz * fz) + (ux * ux + uy * uy + uz * uz) +
                               (vx * vx + vy * vy + vz * vz) + (fx * ux + fy * uy + fz * uz) +
                               (ux * vx + uy * vy + uz * vz) + (vx * fx + vy * fy + vz * fz);
                    e6 += vol6 * q;
                }
                if(FACES) {
                    ncr[FACES ? r : 0][l] = nc;
                }
            }
        }
    }
    m6 = tm.sum(m6);
    gx = tm.sum(gx);
    gy = tm.sum(gy);
    gz = tm.sum(gz);
    if(want_e) {
        e6 = tm.sum(e6);
    }
    // Euler's relation T = 2F - 4 and every directed edge has its twin
    const bool consistent = !tm.any(!ok) && !tm.any(degenerate) && (T + 4 == 2 * F);
    if(T > 0 && tableok && !consistent) {
        status |= B200VOR_CELL_DEGENERATE;
    }
    const double sgn = m6 < 0.0 ? -1.0 : 1.0;
    const double mass = sgn * m6 / 6.0;
    if(alive && lt == 0) {
        double mm = dead ? 0.0 : mass;
        P.m[row] = mm;
        P.mg[3 * size_t(row)] = dead ? 0.0 : sgn * gx / 24.0 + mm * si.x;
        P.mg[3 * size_t(row) + 1] = dead ? 0.0 : sgn * gy / 24.0 + mm * si.y;
        P.mg[3 * size_t(row) + 2] = dead ? 0.0 : sgn * gz / 24.0 + mm * si.z;
        P.status[row] = uint8_t(status);
        if(want_e) {
            P.energy[row] = dead ? 0.0 : sgn * e6 / 60.0;
        }
        if(EV) {
            P.evcount[2 * size_t(row)] = min(nP, 0xffffu) | (min(nV, 0xfffu) << 16) | (min(nscan, 15u) << 28);
            P.evcount[2 * size_t(row) + 1] = min(nC, 0xfffffu) | (T << 20);
        }
    }
    if(FACES) {
        __syncwarp();
        // facet areas (convex_cell.cpp:1160-1191) and neighbour ids; staged in the vertex
        // array, which is no longer read
        double* area = &sh->vx[0][0];                                  // [PC]
        uint32_t* ids = reinterpret_cast<uint32_t*>(&sh->vx[32][0]);   // [PC]
        uint32_t nf = 0;
        unsigned long long rest = (tableok && !dead) ? used : 0ull;
        while(wany(rest != 0ull)) {
            const bool act = rest != 0ull;
            uint32_t f = act ? uint32_t(__ffsll((long long)rest)) - 1u : 0xffu;
            rest &= rest - 1ull;
            double part = 0.0;
#pragma unroll
            for(int r = 0; r < R; ++r) {
                uint32_t s = lt + G * r;
                if(act && s < T) {
                    uint32_t t = sh->tri[s];
#pragma unroll
                    for(int l = 0; l < 3; ++l) {
                        if(((t >> (8 * l)) & 0xffu) == f) {
                            part += ncr[FACES ? r : 0][l];
                        }
                    }
                }
            }
            part = tm.sum(part);
            if(act && lt == 0) {
                Plane pf = load_plane(sh->pl[f]);
                double nn = sqrt(pf.a * pf.a + pf.b * pf.b + pf.c * pf.c);
                area[nf] = 0.5 * fabs(part) / nn;
                uint32_t im = sh->gimg[f];
                uint32_t id;
                if(im == kImgBox) {
                    id = sh->gid[f];
                } else {
                    uint32_t inst = translation_to_instance(int(im & 3u) - 1, int((im >> 2) & 3u) - 1,
                                                            int((im >> 4) & 3u) - 1);
                    id = __ldg(P.perm + sh->gid[f]) + inst * P.n;
                }
                ids[nf] = id;
            }
            nf += act ? 1u : 0u;
        }
        __syncwarp();
        if(alive) {
            for(uint32_t k = lt; k < nf; k += G) {
                uint32_t id = ids[k];
                uint32_t rk = 0;
                for(uint32_t q = 0; q < nf; ++q) {
                    rk += ids[q] < id;
                }
                if(rk < P.fcap) {
                    P.face_nbr[size_t(row) * P.fcap + rk] = id;
                    P.face_area[size_t(row) * P.fcap + rk] = area[k];
                }
            }
            if(lt == 0) {
                P.nfaces[row] = nf;
            }
        }
    }
}

}  // namespace

int run_cells_fused(Ctx* c, uint32_t nq, const uint32_t* qlist, int want_faces, int collect_events,
                    float target, int want_energy) {
    if(nq == 0) {
        return B200VOR_OK;
    }
    FusedParams P;
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
    P.has_w = c->has_w ? 1 : 0;
    P.wmax = c->wmax;
    P.nq = nq;
    P.qlist = qlist;
    P.m = c->m.p;
    P.mg = c->mg.p;
    P.energy = want_energy ? c->energy.p : nullptr;
    P.status = c->status.p;
    P.nfaces = c->nfaces.p;
    P.face_nbr = c->face_nbr.p;
    P.face_area = c->face_area.p;
    P.fcap = c->fcap;
    P.evcount = collect_events ? c->evcount.p : nullptr;
    P.target = target > 0.f ? target : 40.f;
    if(const char* e = getenv("B200VOR_TARGET")) {   // development knob
        P.target = float(atof(e));
    }
    constexpr int G = B200VOR_FUSED_G, WPB = B200VOR_FUSED_WPB, TPW = 32 / G;
    size_t smem = sizeof(TeamShared<G>) * TPW * WPB;
    uint32_t nblk = (nq + TPW * WPB - 1) / (TPW * WPB);
    // four instances: face lists or not, event counters (SURVEY 8(d) flop model) or not
    void (*kern)(FusedParams) =
        want_faces ? (collect_events ? cells_fused_kernel<G, true, WPB, true> : cells_fused_kernel<G, true, WPB, false>)
                   : (collect_events ? cells_fused_kernel<G, false, WPB, true> : cells_fused_kernel<G, false, WPB, false>);
    B200_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
    kern<<<nblk, 32 * WPB, smem, c->stream>>>(P);
    c->stats.gpu_launches++;
    B200_CUDA_CHECK(cudaGetLastError());
    return B200VOR_OK;
}

}  // namespace b200vor
