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

namespace b200vor {

namespace {

constexpr int NT = 64;      // cell-vertex (dual triangle) slots
constexpr int PC = 64;      // plane slots (0..5 = clipping box)
constexpr int POOL = 64;    // candidate pool
constexpr int FC = 40;      // faces of a finished cell that the twin table can index
constexpr int kMaxRing = 4;
constexpr double kEps = 1.1102230246251565e-16;   // 2^-53
constexpr uint32_t kImg0 = 1u | (1u << 2) | (1u << 4);   // packed zero translation
constexpr uint32_t kImgBox = 0xFFu;

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
    // ---- from here to `desc` the area is re-used as the FC x FC twin table ----
    float ve[NT];            // forward error coefficient of the vertex
    float vr[NT];            // |v|^2, rounded up
    uint32_t mark[PC][2];    // directed dual edges of the conflict zone (bit matrix)
    uint32_t pool_key[POOL];
    uint32_t pool_j[POOL];
    uint32_t desc[G][5];     // scan rows: prefix end, a0, n0, a1, images
    // ----
    uint32_t tri[NT];        // p0 | p1 << 8 | p2 << 16
    uint32_t gid[PC];        // grid-order index of the plane's neighbour seed
    uint16_t rec[NT];        // border edges of a cut: a | b << 8
    uint8_t pool_img[POOL];
    uint8_t gimg[PC];        // packed periodic image of the neighbour, kImgBox = box plane
    uint8_t slot_of_rank[NT];
};

template <int G> struct Team {
    uint32_t tmask, tbase, lt;
    static constexpr uint32_t kGMask = (G == 32) ? 0xffffffffu : ((1u << (G & 31)) - 1u);
    __device__ inline uint32_t ballot(bool p) const {
        return (__ballot_sync(tmask, p) >> tbase) & kGMask;
    }
    __device__ inline bool any(bool p) const { return __any_sync(tmask, p) != 0; }
    __device__ inline void sync() const { __syncwarp(tmask); }
    __device__ inline uint32_t lt_mask() const { return (1u << lt) - 1u; }
    template <class T> __device__ inline T bcast(T v, uint32_t src) const {
        return __shfl_sync(tmask, v, int(src), G);
    }
    __device__ inline double sum(double v) const {
#pragma unroll
        for(int o = G / 2; o > 0; o >>= 1) {
            v += __shfl_xor_sync(tmask, v, o, G);
        }
        return v;
    }
    __device__ inline uint32_t max_u32(uint32_t v) const { return __reduce_max_sync(tmask, v); }
    __device__ inline uint32_t or_u32(uint32_t v) const { return __reduce_or_sync(tmask, v); }
    __device__ inline uint32_t add_u32(uint32_t v) const { return __reduce_add_sync(tmask, v); }
    // exclusive prefix sum over the team, total in `total`
    __device__ inline uint32_t exscan(uint32_t v, uint32_t& total) const {
        uint32_t inc = v;
#pragma unroll
        for(int o = 1; o < G; o <<= 1) {
            uint32_t t = __shfl_up_sync(tmask, inc, o, G);
            if(lt >= uint32_t(o)) {
                inc += t;
            }
        }
        total = __shfl_sync(tmask, inc, G - 1, G);
        return inc - v;
    }
};

struct Vertex {
    double x, y, z, w;
    float e, r2;
    bool degenerate;
};

// Homogeneous intersection of three planes (Cramer in cross-product form), normalised to
// W > 0, with its forward error coefficient and |v|^2 (same arithmetic as cell.cu).
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

__device__ inline Vertex make_vertex(const Plane& P0, const Plane& P1, const Plane& P2) {
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
    float n0 = __double2float_ru(fmax(fabs(a0), fmax(fabs(b0), fabs(c0))));
    float n1 = __double2float_ru(fmax(fabs(a1), fmax(fabs(b1), fabs(c1))));
    float n2 = __double2float_ru(fmax(fabs(a2), fmax(fabs(b2), fabs(c2))));
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

template <int G> struct Cell {
    static constexpr int R = NT / G;     // slot rows per lane
    Team<G> tm;
    TeamShared<G>* sh;
    uint32_t T;                          // live slots [0,T)
    unsigned long long live;             // allocated plane slots
    float R2;
    uint32_t status;
    uint32_t nP, nC, nV;
    bool degenerate;

    __device__ inline void store_vertex(uint32_t s, const Vertex& V, uint32_t t) {
        double2* p = reinterpret_cast<double2*>(sh->vx[s]);
        p[0] = make_double2(V.x, V.y);
        p[1] = make_double2(V.z, V.w);
        sh->ve[s] = V.e;
        sh->vr[s] = V.r2;
        sh->tri[s] = t;
    }

    __device__ inline void update_R2() {
        uint32_t rb = 0;
#pragma unroll
        for(int r = 0; r < R; ++r) {
            uint32_t s = tm.lt + G * r;
            if(s < T) {
                rb = max(rb, __float_as_uint(sh->vr[s]));
            }
        }
        R2 = __uint_as_float(tm.max_u32(rb));
    }

    __device__ inline void init_box(const double* bx) {
        // bx: clipping box in the seed-centred frame {xmin,ymin,zmin,xmax,ymax,zmax}
        const uint32_t lt = tm.lt;
        for(uint32_t p = lt; p < 6; p += G) {
            int axis = p >> 1;
            bool upper = p & 1;
            double sgn = upper ? -1.0 : 1.0;
            double* q = sh->pl[p];
            q[0] = axis == 0 ? sgn : 0.0;
            q[1] = axis == 1 ? sgn : 0.0;
            q[2] = axis == 2 ? sgn : 0.0;
            q[3] = upper ? bx[3 + axis] : -bx[axis];
            sh->gid[p] = B200VOR_BOX_ID + p;
            sh->gimg[p] = uint8_t(kImgBox);
        }
        for(uint32_t p = lt; p < PC; p += G) {
            sh->mark[p][0] = 0;
            sh->mark[p][1] = 0;
        }
        tm.sync();
        degenerate = false;
        for(uint32_t s = lt; s < 8; s += G) {
            uint32_t bxb = s & 1, byb = (s >> 1) & 1, bzb = (s >> 2) & 1;
            uint32_t px = bxb, py = 2 + byb, pz = 4 + bzb;
            bool odd = (bxb + byb + bzb) & 1;
            uint32_t v1 = odd ? pz : py, v2 = odd ? py : pz;
            Vertex V = make_vertex(load_plane(sh->pl[px]), load_plane(sh->pl[v1]),
                                   load_plane(sh->pl[v2]));
            store_vertex(s, V, px | (v1 << 8) | (v2 << 16));
        }
        tm.sync();
        T = 8;
        live = 0x3Full;
        status = 0;
        nP = nC = nV = 0;
        update_R2();
    }

    __device__ inline unsigned long long used_planes() const {
        uint32_t lo = 0, hi = 0;
#pragma unroll
        for(int r = 0; r < R; ++r) {
            uint32_t s = tm.lt + G * r;
            if(s < T) {
                uint32_t t = sh->tri[s];
#pragma unroll
                for(int e = 0; e < 3; ++e) {
                    uint32_t v = (t >> (8 * e)) & 0xffu;
                    if(v < 32) {
                        lo |= 1u << v;
                    } else {
                        hi |= 1u << (v - 32);
                    }
                }
            }
        }
        lo = tm.or_u32(lo);
        hi = tm.or_u32(hi);
        return (unsigned long long)lo | ((unsigned long long)hi << 32);
    }

    // Clips by a x + b y + c z + d >= 0.  gj/gi: neighbour seed index and image.
    __device__ inline void clip(double a, double b, double c, double d, uint32_t gj, uint32_t gi) {
        const uint32_t lt = tm.lt;
        ++nP;
        nC += T;
        const float nmaxf = __double2float_ru(fmax(fabs(a), fmax(fabs(b), fabs(c))));
        const float dabs = __double2float_ru(fabs(d));
        const float M = fmaxf(nmaxf, dabs) * 1.0001f;
        uint32_t cm[R];
        uint32_t cfbits = 0;
        bool unc = false;
        uint32_t nconf = 0;
#pragma unroll
        for(int r = 0; r < R; ++r) {
            cm[r] = 0;
            if(uint32_t(G * r) < T) {
                uint32_t s = lt + G * r;
                bool valid = s < T;
                bool cf = false;
                if(valid) {
                    const double2* p = reinterpret_cast<const double2*>(sh->vx[s]);
                    double2 xy = p[0], zw = p[1];
                    double dot = fma(a, xy.x, fma(b, xy.y, fma(c, zw.x, d * zw.y)));
                    float E = M * sh->ve[s];
                    // conflict only when CERTAINLY outside; a tie leaves the vertex alone
                    // (reference fast mode, convex_cell.cpp:968) and flags the cell
                    bool tie = __double2float_rd(fabs(dot)) <= E;
                    cf = !tie && (dot < 0.0);
                    unc |= tie;
                }
                cm[r] = tm.ballot(cf);
                cfbits |= cf ? (1u << r) : 0u;
                nconf += __popc(cm[r]);
            }
        }
        if(tm.any(unc)) {
            status |= B200VOR_CELL_UNCERTAIN;
        }
        if(nconf == 0) {
            return;
        }
        if(nconf == T) {
            T = 0;
            status |= B200VOR_CELL_EMPTY;
            R2 = 0.f;
            return;
        }
        if(live == ~0ull) {
            live = used_planes() | 0x3Full;
            if(live == ~0ull) {
                status |= B200VOR_CELL_OVERFLOW;
                return;
            }
        }
        const uint32_t pn = uint32_t(__ffsll((long long)~live)) - 1u;
        // ---- mark the directed edges of the conflict triangles ----
        uint32_t myt[R];
#pragma unroll
        for(int r = 0; r < R; ++r) {
            myt[r] = 0;
            if(cfbits & (1u << r)) {
                uint32_t t = sh->tri[lt + G * r];
                myt[r] = t;
                uint32_t p0 = t & 0xffu, p1 = (t >> 8) & 0xffu, p2 = (t >> 16) & 0xffu;
                atomicOr(&sh->mark[p0][p1 >> 5], 1u << (p1 & 31u));
                atomicOr(&sh->mark[p1][p2 >> 5], 1u << (p2 & 31u));
                atomicOr(&sh->mark[p2][p0 >> 5], 1u << (p0 & 31u));
            }
        }
        tm.sync();
        // ---- border edges: those whose twin is not marked ----
        uint32_t bits = 0;
#pragma unroll
        for(int r = 0; r < R; ++r) {
            if(cfbits & (1u << r)) {
                uint32_t t = myt[r];
#pragma unroll
                for(int e = 0; e < 3; ++e) {
                    uint32_t u = (t >> (8 * e)) & 0xffu, v = (t >> (8 * ((e + 1) % 3))) & 0xffu;
                    uint32_t twin = (sh->mark[v][u >> 5] >> (u & 31u)) & 1u;
                    if(!twin) {
                        bits |= 1u << (3 * r + e);
                    }
                }
            }
        }
        tm.sync();
#pragma unroll
        for(int r = 0; r < R; ++r) {
            if(cfbits & (1u << r)) {
                uint32_t t = myt[r];
                uint32_t p0 = t & 0xffu, p1 = (t >> 8) & 0xffu, p2 = (t >> 16) & 0xffu;
                sh->mark[p0][0] = 0; sh->mark[p0][1] = 0;
                sh->mark[p1][0] = 0; sh->mark[p1][1] = 0;
                sh->mark[p2][0] = 0; sh->mark[p2][1] = 0;
            }
        }
        uint32_t nnew;
        const uint32_t cnt = __popc(bits);
        uint32_t rk = tm.exscan(cnt, nnew);
        const uint32_t Tnew = T - nconf + nnew;
        if(Tnew > uint32_t(NT) || nnew > uint32_t(NT)) {
            tm.sync();
            status |= B200VOR_CELL_OVERFLOW;
            return;
        }
        live |= 1ull << pn;
        if(lt == 0) {
            double2* p = reinterpret_cast<double2*>(sh->pl[pn]);
            p[0] = make_double2(a, b);
            p[1] = make_double2(c, d);
            sh->gid[pn] = gj;
            sh->gimg[pn] = uint8_t(gi);
        }
        // ---- border-edge records, holes ranked in slot order ----
        uint32_t hbase = 0;
#pragma unroll
        for(int r = 0; r < R; ++r) {
            if(cfbits & (1u << r)) {
                uint32_t t = myt[r];
#pragma unroll
                for(int e = 0; e < 3; ++e) {
                    if(bits & (1u << (3 * r + e))) {
                        uint32_t u = (t >> (8 * e)) & 0xffu, v = (t >> (8 * ((e + 1) % 3))) & 0xffu;
                        sh->rec[rk++] = uint16_t(u | (v << 8));
                    }
                }
                uint32_t hr = hbase + __popc(cm[r] & tm.lt_mask());
                if(hr < nnew) {
                    sh->slot_of_rank[hr] = uint8_t(lt + G * r);
                }
            }
            hbase += __popc(cm[r]);
        }
        tm.sync();
        // ---- fewer new triangles than holes (two faces vanished at once): the holes
        //      left below Tnew take the surviving entries stored at or above Tnew ----
        if(nnew < nconf) {
            uint32_t hb = 0, tb = 0;
            uint32_t dst[R];
#pragma unroll
            for(int r = 0; r < R; ++r) {
                uint32_t s = lt + G * r;
                bool ishole = (cfbits >> r) & 1u;
                uint32_t hr = 0;
                // recompute the hole rank of this slot
                {
                    uint32_t before = 0;
#pragma unroll
                    for(int q = 0; q < R; ++q) {
                        if(q < r) {
                            before += __popc(cm[q]);
                        }
                    }
                    hr = before + __popc(cm[r] & tm.lt_mask());
                }
                bool openhole = ishole && hr >= nnew && s < Tnew;
                uint32_t ob = tm.ballot(openhole);
                if(openhole) {
                    sh->slot_of_rank[hb + __popc(ob & tm.lt_mask())] = uint8_t(s);
                }
                hb += __popc(ob);
                bool mover = s >= Tnew && s < T && !ishole;
                uint32_t mb = tm.ballot(mover);
                dst[r] = mover ? tb + __popc(mb & tm.lt_mask()) : 0xffffffffu;
                tb += __popc(mb);
            }
            tm.sync();
            // read the movers, then write them (sources are all >= Tnew, targets < Tnew)
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
#pragma unroll
            for(int r = 0; r < R; ++r) {
                if(dst[r] != 0xffffffffu) {
                    uint32_t s = sh->slot_of_rank[dst[r]];
                    double2* p = reinterpret_cast<double2*>(sh->vx[s]);
                    p[0] = mxy[r]; p[1] = mzw[r];
                    sh->ve[s] = me[r]; sh->vr[s] = mr[r]; sh->tri[s] = mt[r];
                }
            }
            tm.sync();
            // the first nnew holes still take the new triangles: restore their slots
            uint32_t hbase2 = 0;
#pragma unroll
            for(int r = 0; r < R; ++r) {
                if(cfbits & (1u << r)) {
                    uint32_t hr = hbase2 + __popc(cm[r] & tm.lt_mask());
                    if(hr < nnew) {
                        sh->slot_of_rank[hr] = uint8_t(lt + G * r);
                    }
                }
                hbase2 += __popc(cm[r]);
            }
            tm.sync();
        }
        // ---- one new vertex per border edge ----
        for(uint32_t k = lt; k < nnew; k += G) {
            uint32_t rr = sh->rec[k];
            uint32_t u = rr & 0xffu, v = rr >> 8;
            uint32_t s = k < nconf ? uint32_t(sh->slot_of_rank[k]) : T + (k - nconf);
            Plane pe;
            pe.a = a; pe.b = b; pe.c = c; pe.d = d;
            Vertex V = make_vertex(pe, load_plane(sh->pl[u]), load_plane(sh->pl[v]));
            degenerate |= V.degenerate;
            store_vertex(s, V, pn | (u << 8) | (v << 16));
        }
        nV += nnew;
        T = Tnew;
        tm.sync();
        update_R2();
    }
};

// In-register bitonic sort of R*G keys, element e = r*G + lane.
template <int G, int RK>
__device__ inline void bitonic_sort(const Team<G>& tm, uint32_t (&key)[RK]) {
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
                        const int e_hi = r * G;          // (e & k) is lane-independent here
                        bool asc = ((e_hi & k) == 0);
                        uint32_t lo = min(key[r], key[r | jr]), hi = max(key[r], key[r | jr]);
                        key[r] = asc ? lo : hi;
                        key[r | jr] = asc ? hi : lo;
                    }
                }
            } else {
#pragma unroll
                for(int r = 0; r < RK; ++r) {
                    uint32_t other = __shfl_xor_sync(tm.tmask, key[r], j, G);
                    uint32_t e = uint32_t(r * G) + tm.lt;
                    bool asc = ((e & uint32_t(k)) == 0);
                    bool lower = ((tm.lt & uint32_t(j)) == 0);
                    bool keep_min = (lower == asc);
                    key[r] = keep_min ? min(key[r], other) : max(key[r], other);
                }
            }
        }
    }
}

template <int G> struct Scanner {
    // scan state of one team
    int cx, cy, cz;
    double six, siy, siz;
    uint32_t self;
};

// Collects into the pool the candidates of the (2r+1)^3 block of bins with
// lo <= d2 < hi.  Returns their number (entries beyond POOL are dropped), and the number
// of seeds in the block in nblock.
template <int G>
__device__ inline uint32_t scan_shell(const FusedParams& P, const Team<G>& tm, TeamShared<G>* sh,
                                      const Scanner<G>& S, int r, double lo, double hi,
                                      bool auto_hi, double hi_cap, double& hi_used,
                                      uint32_t& nblock) {
    const GridDesc& Gd = P.G;
    const int R1 = 2 * r + 1;
    const int nrows = R1 * R1;
    uint32_t count = 0;
    nblock = 0;
    hi_used = hi;
    for(int rb = 0; rb < nrows; rb += G) {
        // ---- row descriptors: lane = one (y,z) row of the block ----
        int row = rb + int(tm.lt);
        uint32_t a0 = 0, n0 = 0, a1 = 0, n1 = 0, img = 0;
        if(row < nrows) {
            int dy = row % R1 - r, dz = row / R1 - r;
            int y = S.cy + dy, z = S.cz + dz, sy = 0, sz = 0;
            bool ok = true;
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
                int xa = S.cx - r, xb = S.cx + r;
                int ma = xa < 0 ? 0 : xa, mb = xb >= gx ? gx - 1 : xb;
                a0 = __ldg(P.cell_start + base + ma);
                n0 = __ldg(P.cell_start + base + mb + 1) - a0;
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
        sh->desc[tm.lt][0] = pre + n0 + n1;
        sh->desc[tm.lt][1] = a0;
        sh->desc[tm.lt][2] = n0;
        sh->desc[tm.lt][3] = a1;
        sh->desc[tm.lt][4] = img;
        nblock += total;
        if(auto_hi && rb == 0) {
            // first shell: radius of the ball expected to hold P.target seeds at the
            // density of the scanned rows
            double cells = double(min(nrows, G)) * double(R1);
            double vol = cells * Gd.w[0] * Gd.w[1] * Gd.w[2];
            double h = hi_cap;
            if(total > 0) {
                double r3 = double(P.target) * 0.2387324146378430 * vol / double(total);
                double rr = cbrt(r3);
                h = fmin(rr * rr, hi_cap);
            }
            hi = h;
            hi_used = h;
        }
        tm.sync();
        // ---- candidates of these rows, balanced over the lanes ----
        uint32_t cur = 0, cur_end = sh->desc[0][0], cur_beg = 0;
        for(uint32_t c0 = 0; c0 < total; c0 += G) {
            uint32_t c = c0 + tm.lt;
            bool in = false;
            uint32_t j = 0, im = 0;
            double d2 = 0.0;
            if(c < total) {
                while(c >= cur_end) {
                    cur_beg = cur_end;
                    ++cur;
                    cur_end = sh->desc[cur][0];
                }
                uint32_t off = c - cur_beg;
                uint32_t dn0 = sh->desc[cur][2];
                uint32_t dimg = sh->desc[cur][4];
                if(off < dn0) {
                    j = sh->desc[cur][1] + off;
                    im = dimg & 0x3Fu;
                } else {
                    j = sh->desc[cur][3] + (off - dn0);
                    im = (dimg & 0x3Cu) | ((dimg >> 8) & 3u);
                }
                Seed sj = load_seed(P.sorted + j);
                double tx = double(int(im & 3u) - 1), ty = double(int((im >> 2) & 3u) - 1),
                       tz = double(int((im >> 4) & 3u) - 1);
                double qx = (sj.x + tx * P.period) - S.six;
                double qy = (sj.y + ty * P.period) - S.siy;
                double qz = (sj.z + tz * P.period) - S.siz;
                d2 = qx * qx + qy * qy + qz * qz;
                in = d2 >= lo && d2 < hi && !(j == S.self && im == kImg0);
            }
            uint32_t bal = tm.ballot(in);
            if(in) {
                uint32_t pos = count + __popc(bal & tm.lt_mask());
                if(pos < uint32_t(POOL)) {
                    uint32_t kb = __float_as_uint(__double2float_rd(d2));
                    sh->pool_key[pos] = (kb & ~63u) | pos;
                    sh->pool_j[pos] = j;
                    sh->pool_img[pos] = uint8_t(im);
                }
            }
            count += __popc(bal);
        }
        tm.sync();
    }
    return count;
}

template <int G> __device__ inline double coverage2(const GridDesc& Gd, const Scanner<G>& S, int r) {
    // squared distance from the seed to the nearest face of the scanned block that has
    // unscanned space behind it (infinity when the block covers the whole grid)
    double dmin = 1e300;
    const double q[3] = {S.six, S.siy, S.siz};
    const int cc[3] = {S.cx, S.cy, S.cz};
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

template <int G, bool FACES, int WPB>
__global__ void __launch_bounds__(32 * WPB) cells_fused_kernel(FusedParams P) {
    constexpr int TPW = 32 / G;
    constexpr int R = NT / G;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const uint32_t lane = threadIdx.x & 31u, wid = threadIdx.x >> 5;
    const uint32_t tiw = lane / G;
    Team<G> tm;
    tm.lt = lane % G;
    tm.tbase = tiw * G;
    tm.tmask = (G == 32) ? 0xffffffffu : (Team<G>::kGMask << tm.tbase);
    TeamShared<G>* sh = reinterpret_cast<TeamShared<G>*>(smem_raw) + (wid * TPW + tiw);
    const uint32_t c = (blockIdx.x * WPB + wid) * TPW + tiw;
    if(c >= P.nq) {
        return;
    }
    const uint32_t i = P.qlist ? P.qlist[c] : c;
    const Seed si = load_seed(P.sorted + i);
    const uint32_t row = P.perm[i];

    Cell<G> C;
    C.tm = tm;
    C.sh = sh;
    double bx[6];
    if(P.periodic) {
#pragma unroll
        for(int a = 0; a < 3; ++a) {
            bx[a] = -0.5 * P.period;
            bx[3 + a] = 0.5 * P.period;
        }
    } else {
        bx[0] = P.box[0] - si.x; bx[1] = P.box[1] - si.y; bx[2] = P.box[2] - si.z;
        bx[3] = P.box[3] - si.x; bx[4] = P.box[4] - si.y; bx[5] = P.box[5] - si.z;
    }
    C.init_box(bx);

    Scanner<G> S;
    cell_coords(P.G, si.x, si.y, si.z, S.cx, S.cy, S.cz);
    S.six = si.x; S.siy = si.y; S.siz = si.z;
    S.self = i;
    const double dw = P.has_w ? (P.wmax - si.w) : 0.0;

    int ring = 1;
    int rmax;
    if(P.G.periodic) {
        int gmin = min(P.G.g[0], min(P.G.g[1], P.G.g[2]));
        rmax = min(kMaxRing, (gmin - 1) / 2);
    } else {
        rmax = max(max(S.cx, P.G.g[0] - 1 - S.cx),
                   max(max(S.cy, P.G.g[1] - 1 - S.cy), max(S.cz, P.G.g[2] - 1 - S.cz)));
        rmax = min(rmax, kMaxRing);
        rmax = max(rmax, 1);
    }
    bool certified = false;
    uint32_t extra = 0;
    if(rmax < 1) {
        ring = 0;   // periodic grid too coarse for a 3x3x3 block: legacy pipeline
    }
    double cov2 = ring ? coverage2(P.G, S, ring) : 0.0;
    double lo = 0.0;
    bool first = true;
    bool stop = (ring == 0);
    for(int guard = 0; guard < 64 && !stop; ++guard) {
        // d2 beyond which no bisector can cut the cell (generic_RVD.h:2322; Laguerre:
        // a power plane cuts only if |q| <= R + sqrt(R^2 + wmax - wi))
        double bound = 4.1 * double(C.R2);
        if(P.has_w) {
            double Rr = sqrt(double(C.R2));
            double sr = Rr + sqrt(double(C.R2) + dw);
            bound = 1.025 * sr * sr;
        }
        if(lo > bound) {
            certified = true;
            break;
        }
        if(lo >= cov2) {
            if(ring >= rmax) {
                // the whole grid is inside the block when no face has space behind it
                certified = (cov2 >= 1e300);
                break;
            }
            ++ring;
            cov2 = coverage2(P.G, S, ring);
            continue;
        }
        double hi_cap = fmin(bound * (1.0 + 1e-12) + 1e-300, cov2);
        double hi = hi_cap, hi_used;
        uint32_t nblock;
        uint32_t cnt = scan_shell<G>(P, tm, sh, S, ring, lo, hi, first, hi_cap, hi_used, nblock);
        hi = hi_used;
        int shrink = 0;
        while(cnt > uint32_t(POOL) && shrink < 48) {
            hi = lo + (hi - lo) * 0.5;
            cnt = scan_shell<G>(P, tm, sh, S, ring, lo, hi, false, hi_cap, hi_used, nblock);
            ++shrink;
        }
        if(cnt > uint32_t(POOL)) {
            break;   // more than POOL candidates at (nearly) one distance: legacy pipeline
        }
        first = false;
        // ---- sort the pool, clip in ascending distance ----
        uint32_t key[R];
#pragma unroll
        for(int r = 0; r < R; ++r) {
            uint32_t e = tm.lt + G * r;
            key[r] = e < cnt ? sh->pool_key[e] : 0xffffffffu;
        }
        if(cnt > uint32_t(G)) {
            bitonic_sort<G, R>(tm, key);
        } else {
            uint32_t k1[1] = {key[0]};
            bitonic_sort<G, 1>(tm, k1);
            key[0] = k1[0];
        }
#pragma unroll
        for(int r = 0; r < R; ++r) {
            if(uint32_t(G * r) < cnt && !stop) {
                // this lane's candidate of the batch
                uint32_t kk = key[r];
                bool valid = kk != 0xffffffffu;
                uint32_t idx = kk & 63u;
                uint32_t j = 0, im = kImg0;
                double pa = 0, pb = 0, pc = 0, pd = 0, d2 = 0;
                if(valid) {
                    j = sh->pool_j[idx];
                    im = sh->pool_img[idx];
                    Seed sj = load_seed(P.sorted + j);
                    double tx = double(int(im & 3u) - 1), ty = double(int((im >> 2) & 3u) - 1),
                           tz = double(int((im >> 4) & 3u) - 1);
                    double qx = (sj.x + tx * P.period) - si.x;
                    double qy = (sj.y + ty * P.period) - si.y;
                    double qz = (sj.z + tz * P.period) - si.z;
                    d2 = qx * qx + qy * qy + qz * qz;
                    pa = -2.0 * qx; pb = -2.0 * qy; pc = -2.0 * qz;
                    pd = d2 + (si.w - sj.w);
                }
                uint32_t nb = min(uint32_t(G), cnt - uint32_t(G * r));
                for(uint32_t t = 0; t < nb; ++t) {
                    double d2t = tm.bcast(d2, t);
                    uint32_t jt = tm.bcast(j, t), it = tm.bcast(im, t);
                    if(d2t == 0.0) {
                        // exact duplicate (delaunay_nn.cpp:122-135): the lower user index
                        // keeps the cell
                        if(it == kImg0 && __ldg(P.perm + jt) < row) {
                            extra |= B200VOR_CELL_DUPLICATE;
                        }
                        continue;
                    }
                    double bnd = 4.1 * double(C.R2);
                    if(P.has_w) {
                        double Rr = sqrt(double(C.R2));
                        double sr = Rr + sqrt(double(C.R2) + dw);
                        bnd = 1.025 * sr * sr;
                    }
                    if(d2t > bnd) {
                        certified = true;
                        stop = true;
                        break;
                    }
                    double a = tm.bcast(pa, t), b = tm.bcast(pb, t), cc = tm.bcast(pc, t),
                           d = tm.bcast(pd, t);
                    C.clip(a, b, cc, d, jt, it);
                    if(C.status & (B200VOR_CELL_OVERFLOW | B200VOR_CELL_EMPTY)) {
                        stop = true;
                        certified = (C.status & B200VOR_CELL_EMPTY) != 0;
                        break;
                    }
                }
            }
        }
        lo = hi;
    }
    C.status |= extra;
    if(!certified) {
        C.status |= B200VOR_CELL_UNCERTIFIED;
    }

    // ---------------------------------------------------------------- integration
    const uint32_t T = C.T;
    const bool dead = (C.status & (B200VOR_CELL_EMPTY | B200VOR_CELL_DUPLICATE)) != 0;
    unsigned long long used = C.used_planes();
    const uint32_t F = uint32_t(__popcll(used));
    if(P.periodic && !(C.status & B200VOR_CELL_EMPTY)) {
        bool mine = false;
        for(uint32_t p = tm.lt; p < uint32_t(PC); p += G) {
            mine = mine || (((used >> p) & 1ull) && sh->gimg[p] == kImgBox);
        }
        if(tm.any(mine)) {
            C.status |= B200VOR_CELL_BOXFACE;
        }
    }
    bool tableok = F <= uint32_t(FC) && T > 0;
    if(T > 0 && F > uint32_t(FC)) {
        C.status |= B200VOR_CELL_OVERFLOW;
    }
    // Cartesian vertices in place
#pragma unroll
    for(int r = 0; r < R; ++r) {
        uint32_t s = tm.lt + G * r;
        if(s < T) {
            double2* p = reinterpret_cast<double2*>(sh->vx[s]);
            double2 xy = p[0], zw = p[1];
            double iw = 1.0 / zw.y;
            p[0] = make_double2(xy.x * iw, xy.y * iw);
            p[1] = make_double2(zw.x * iw, 1.0);
        }
    }
    uint8_t* table = reinterpret_cast<uint8_t*>(sh->ve);
    static_assert(offsetof(TeamShared<G>, tri) - offsetof(TeamShared<G>, ve) >= FC * FC,
                  "twin table does not fit its overlay");
    tm.sync();
    if(tableok) {
        for(uint32_t k = tm.lt; k < uint32_t(FC * FC) / 4u; k += G) {
            reinterpret_cast<uint32_t*>(table)[k] = 0xffffffffu;
        }
    }
    tm.sync();
    uint32_t ci[R][3];
    bool ok = true;
#pragma unroll
    for(int r = 0; r < R; ++r) {
        uint32_t s = tm.lt + G * r;
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
    tm.sync();
    double m6 = 0.0, gx = 0.0, gy = 0.0, gz = 0.0;
    double ncr[R][3];
#pragma unroll
    for(int r = 0; r < R; ++r) {
#pragma unroll
        for(int l = 0; l < 3; ++l) {
            ncr[r][l] = 0.0;
        }
        uint32_t s = tm.lt + G * r;
        if(s < T && tableok) {
            uint32_t t = sh->tri[s];
            const double* u = sh->vx[s];
            double ux = u[0], uy = u[1], uz = u[2];
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
                const double* v = sh->vx[nb];
                double vx = v[0], vy = v[1], vz = v[2];
                const double* pf = sh->pl[f];
                double a = pf[0], b = pf[1], cq = pf[2], d = pf[3];
                double cxp = uy * vz - uz * vy, cyp = uz * vx - ux * vz, czp = ux * vy - uy * vx;
                double nc = a * cxp + b * cyp + cq * czp;
                double inv = -d / (a * a + b * b + cq * cq);
                double vol6 = inv * nc;
                m6 += vol6;
                gx += vol6 * (a * inv + ux + vx);
                gy += vol6 * (b * inv + uy + vy);
                gz += vol6 * (cq * inv + uz + vz);
                ncr[r][l] = nc;
            }
        }
    }
    m6 = tm.sum(m6);
    gx = tm.sum(gx);
    gy = tm.sum(gy);
    gz = tm.sum(gz);
    // Euler's relation T = 2F - 4 and every directed edge has its twin
    bool consistent = !tm.any(!ok) && !tm.any(C.degenerate) && (T + 4 == 2 * F);
    if(T > 0 && tableok && !consistent) {
        C.status |= B200VOR_CELL_DEGENERATE;
    }
    double sgn = m6 < 0.0 ? -1.0 : 1.0;
    double mass = sgn * m6 / 6.0;
    if(tm.lt == 0) {
        double mm = dead ? 0.0 : mass;
        P.m[row] = mm;
        P.mg[3 * size_t(row)] = dead ? 0.0 : sgn * gx / 24.0 + mm * si.x;
        P.mg[3 * size_t(row) + 1] = dead ? 0.0 : sgn * gy / 24.0 + mm * si.y;
        P.mg[3 * size_t(row) + 2] = dead ? 0.0 : sgn * gz / 24.0 + mm * si.z;
        P.status[row] = uint8_t(C.status);
        if(P.evcount) {
            P.evcount[2 * size_t(row)] = min(C.nP, 0xffffu) | (min(C.nV, 0xffffu) << 16);
            P.evcount[2 * size_t(row) + 1] = min(C.nC, 0xfffffu) | (T << 20);
        }
    }
    if(FACES) {
        tm.sync();
        // areas and ids are staged in the (now free) vertex error / pool area after the
        // table: use rec-sized scratch inside vx is not possible (vertices are read above,
        // but no longer below), so stage in vx
        double* area = &sh->vx[0][0];                          // [PC]
        uint32_t* ids = reinterpret_cast<uint32_t*>(&sh->vx[32][0]);   // [PC]
        uint32_t nf = 0;
        unsigned long long rest = (tableok && !dead) ? used : 0ull;
        // all reads of vx are complete (ncr is in registers)
        while(rest) {
            uint32_t f = uint32_t(__ffsll((long long)rest)) - 1u;
            rest &= rest - 1ull;
            double part = 0.0;
#pragma unroll
            for(int r = 0; r < R; ++r) {
                uint32_t s = tm.lt + G * r;
                if(s < T) {
                    uint32_t t = sh->tri[s];
#pragma unroll
                    for(int l = 0; l < 3; ++l) {
                        if(((t >> (8 * l)) & 0xffu) == f) {
                            part += ncr[r][l];
                        }
                    }
                }
            }
            part = tm.sum(part);
            if(tm.lt == 0) {
                const double* pf = sh->pl[f];
                double nn = sqrt(pf[0] * pf[0] + pf[1] * pf[1] + pf[2] * pf[2]);
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
            ++nf;
        }
        tm.sync();
        for(uint32_t k = tm.lt; k < nf; k += G) {
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
        if(tm.lt == 0) {
            P.nfaces[row] = nf;
        }
    }
}

}  // namespace

int run_cells_fused(Ctx* c, uint32_t nq, const uint32_t* qlist, int want_faces, int collect_events,
                    float target) {
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
    P.status = c->status.p;
    P.nfaces = c->nfaces.p;
    P.face_nbr = c->face_nbr.p;
    P.face_area = c->face_area.p;
    P.fcap = c->fcap;
    P.evcount = collect_events ? c->evcount.p : nullptr;
    P.target = target > 0.f ? target : 40.f;
    constexpr int G = 16, WPB = 4, TPW = 32 / G;
    size_t smem = sizeof(TeamShared<G>) * TPW * WPB;
    uint32_t nblk = (nq + TPW * WPB - 1) / (TPW * WPB);
    if(want_faces) {
        B200_CUDA_CHECK(cudaFuncSetAttribute(cells_fused_kernel<G, true, WPB>,
                                             cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
        cells_fused_kernel<G, true, WPB><<<nblk, 32 * WPB, smem, c->stream>>>(P);
    } else {
        B200_CUDA_CHECK(cudaFuncSetAttribute(cells_fused_kernel<G, false, WPB>,
                                             cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem)));
        cells_fused_kernel<G, false, WPB><<<nblk, 32 * WPB, smem, c->stream>>>(P);
    }
    c->stats.gpu_launches++;
    B200_CUDA_CHECK(cudaGetLastError());
    return B200VOR_OK;
}

}  // namespace b200vor
