// cell.cu -- one-warp-per-cell half-space clipper with the security-radius rule and the
// fused volume / centroid / facet-area integration.
//
// Replaces, for one seed at a time:
//   VBW::ConvexCell::init_with_box            voronoi/convex_cell.cpp:175-215
//   VBW::ConvexCell::clip_by_plane            voronoi/convex_cell.cpp:467-516
//     triangle_is_in_conflict                 voronoi/convex_cell.cpp:870-969
//     triangulate_conflict_zone               voronoi/convex_cell.cpp:761-865
//   compute_triangle_point (Cramer)           voronoi/convex_cell.cpp:971-1082
//   squared_radius                            voronoi/convex_cell.cpp:1312-1330
//   compute_mg / facet_area                   voronoi/convex_cell.cpp:1267-1307, :1160-1191
//   clip_by_cell_SR (stop rule d2 > 4.1 R2)   voronoi/generic_RVD.h:2278-2347
//   power planes                              delaunay/periodic_delaunay_3d.cpp:3719-3726
//
// Design (B200-first, not a transcription):
//  * The cell is kept in DUAL form like the reference (a cell vertex = a triangle of
//    three plane ids, with the three adjacent triangles), but the triangle table lives
//    in REGISTERS: lane l owns triangle slots l, l+32, ...  A clip is one fused
//    multiply-add chain per lane against the lane's cached homogeneous vertex, one
//    ballot per slot row, and -- only when the plane cuts -- a warp-parallel
//    re-triangulation: border edges are found by every surviving triangle from the
//    conflict ballots, ranked by a shuffle scan, handed to the free slots through a
//    64-entry shared record list, and stitched around the new plane with two
//    byte tables indexed by plane id.  There is no linked list and no serial walk.
//  * Everything is computed in a seed-centred frame: plane (i,j) = (-2q, |q|^2 + wi - wj)
//    with q = pj - pi formed in fp64, which removes the |pj|^2-|pi|^2 cancellation of
//    the reference's global-frame formula.
//  * The cached vertex (X,Y,Z,W) is exactly the cofactor vector of the reference's 4x4
//    determinant, so "E.(X,Y,Z,W) < 0" IS the reference's sign test; a forward error
//    bound is carried per vertex and a test that falls inside it flags the cell
//    B200VOR_CELL_UNCERTAIN (never silently accepted).
//  * The security radius is tracked incrementally: each lane keeps |v|^2 of its
//    vertices (fp32, rounded up), R2 is one REDUX max after each effective clip.
//  * Integration: every cell vertex has exactly three outgoing directed face edges;
//    each contributes the signed tetrahedron (seed, foot of the face plane, u, v).
//    3 terms per lane, one fixed-order butterfly reduction: deterministic fp64.
#include "common.cuh"

namespace b200vor {

namespace {

constexpr uint32_t kFull = 0xffffffffu;
constexpr int kPlaneCap = 64;          // live planes per cell (local ids are bytes)
constexpr int kWarpsPerBlock = 8;
constexpr double kEps = 1.1102230246251565e-16;   // 2^-53

__constant__ int c_translation[27][3] = {
    // delaunay/periodic.cpp:44-72
    {0, 0, 0},   {-1, -1, -1}, {-1, -1, 0}, {-1, -1, 1}, {-1, 0, -1}, {-1, 0, 0},
    {-1, 0, 1},  {-1, 1, -1},  {-1, 1, 0},  {-1, 1, 1},  {0, -1, -1}, {0, -1, 0},
    {0, -1, 1},  {0, 0, -1},   {0, 0, 1},   {0, 1, -1},  {0, 1, 0},   {0, 1, 1},
    {1, -1, -1}, {1, -1, 0},   {1, -1, 1},  {1, 0, -1},  {1, 0, 0},   {1, 0, 1},
    {1, 1, -1},  {1, 1, 0},    {1, 1, 1}};

struct CellParams {
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
    CellJob job;
    double* m;
    double* mg;
    double* energy;
    uint8_t* status;
    uint32_t* nfaces;
    uint32_t* face_nbr;
    double* face_area;
    uint32_t fcap;
    uint32_t* evcount;
};

template <int S> struct SharedLayout {
    static constexpr int NT = 32 * S;
    double pl[kPlaneCap * 4];
    double vtx[NT * 3];
    float pnm[kPlaneCap];
    uint32_t gid[kPlaneCap];
    uint32_t rec[NT];
    uint8_t slot_of[NT];
    uint8_t first_by[kPlaneCap];
    uint8_t second_by[kPlaneCap];
};

template <int S> struct WarpCell {
    static constexpr int NT = 32 * S;
    // per lane
    double X[S], Y[S], Z[S], W[S];
    float ea[S], eb[S], r2[S];
    uint32_t tri[S], adj[S];
    bool degenerate;
    // warp-uniform
    uint32_t vmask[S];
    unsigned long long live;
    float R2;
    uint32_t status;
    uint32_t nP, nC, nV;
    uint32_t lane;
    SharedLayout<S>* sh;

    __device__ inline uint32_t lanemask_lt() const { return (1u << lane) - 1u; }

    // homogeneous intersection of planes p0,p1,p2 (Cramer, in cross-product form),
    // normalised to W > 0, with forward error coefficients and |v|^2
    __device__ inline void make_vertex(int s, uint32_t p0, uint32_t p1, uint32_t p2) {
        const double* q0 = sh->pl + 4 * p0;
        const double* q1 = sh->pl + 4 * p1;
        const double* q2 = sh->pl + 4 * p2;
        double a0 = q0[0], b0 = q0[1], c0 = q0[2], d0 = q0[3];
        double a1 = q1[0], b1 = q1[1], c1 = q1[2], d1 = q1[3];
        double a2 = q2[0], b2 = q2[1], c2 = q2[2], d2 = q2[3];
        // n1 x n2, n2 x n0, n0 x n1
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
        X[s] = x; Y[s] = y; Z[s] = z; W[s] = w;
        // error model: |cross comp| <= 2 Nj Nk with relative error 4 eps; three products
        // and two sums on top.  c = 8 covers it with margin.
        float n0 = sh->pnm[p0], n1 = sh->pnm[p1], n2 = sh->pnm[p2];
        float D0 = __double2float_ru(fabs(d0)), D1 = __double2float_ru(fabs(d1)),
              D2 = __double2float_ru(fabs(d2));
        float Q = D0 * n1 * n2 + D1 * n2 * n0 + D2 * n0 * n1;
        float sx = __double2float_ru(fabs(x) + fabs(y) + fabs(z));
        float sw = __double2float_ru(w);
        const float e48 = float(48.0 * kEps) * 1.001f, e4 = float(4.0 * kEps) * 1.001f;
        ea[s] = (e48 * Q + e4 * sx) * 1.001f;
        eb[s] = (e48 * (n0 * n1 * n2) + e4 * sw) * 1.001f;
        if(!(sw > 2.0f * e48 * (n0 * n1 * n2))) {
            degenerate = true;   // three (nearly) linearly dependent normals: no reliable vertex
        }
        // |v|^2 rounded up (only the security radius uses it; 4.1 vs 4 leaves 2.5% slack)
        float num = __double2float_ru(x * x + y * y + z * z);
        float den = __double2float_rd(w * w);
        r2[s] = __fdiv_ru(num, den);
    }

    __device__ inline void init_box(const double* bx) {
        // bx: box in the seed-centred frame {xmin,ymin,zmin,xmax,ymax,zmax}
#pragma unroll
        for(int s = 0; s < S; ++s) {
            X[s] = Y[s] = Z[s] = W[s] = 0.0;
            ea[s] = eb[s] = r2[s] = 0.f;
            tri[s] = adj[s] = 0;
            vmask[s] = 0;
        }
        degenerate = false;
        if(lane < 6) {
            int axis = lane >> 1;
            bool upper = lane & 1;
            double sgn = upper ? -1.0 : 1.0;
            double* p = sh->pl + 4 * lane;
            p[0] = axis == 0 ? sgn : 0.0;
            p[1] = axis == 1 ? sgn : 0.0;
            p[2] = axis == 2 ? sgn : 0.0;
            p[3] = upper ? bx[3 + axis] : -bx[axis];
            sh->pnm[lane] = 1.0f;
            sh->gid[lane] = B200VOR_BOX_ID + lane;
        }
        __syncwarp();
        live = 0x3Full;
        vmask[0] = 0xFFu;
        if(lane < 8) {
            uint32_t bxb = lane & 1, byb = (lane >> 1) & 1, bzb = (lane >> 2) & 1;
            uint32_t px = bxb, py = 2 + byb, pz = 4 + bzb;
            bool odd = (bxb + byb + bzb) & 1;
            uint32_t v1 = odd ? pz : py, v2 = odd ? py : pz;
            tri[0] = px | (v1 << 8) | (v2 << 16);
            // neighbour across the edge opposite a plane = corner with that axis flipped
            uint32_t n0 = lane ^ 1u, n1 = lane ^ (odd ? 4u : 2u), n2 = lane ^ (odd ? 2u : 4u);
            adj[0] = n0 | (n1 << 8) | (n2 << 16);
            make_vertex(0, px, v1, v2);
        }
        uint32_t rb = lane < 8 ? __float_as_uint(r2[0]) : 0u;
        R2 = __uint_as_float(__reduce_max_sync(kFull, rb));
        status = 0;
        nP = nC = nV = 0;
    }

    __device__ inline unsigned long long used_planes() const {
        uint32_t lo = 0, hi = 0;
#pragma unroll
        for(int s = 0; s < S; ++s) {
            if((vmask[s] >> lane) & 1u) {
#pragma unroll
                for(int e = 0; e < 3; ++e) {
                    uint32_t v = (tri[s] >> (8 * e)) & 0xffu;
                    if(v < 32) {
                        lo |= 1u << v;
                    } else {
                        hi |= 1u << (v - 32);
                    }
                }
            }
        }
        lo = __reduce_or_sync(kFull, lo);
        hi = __reduce_or_sync(kFull, hi);
        return (unsigned long long)lo | ((unsigned long long)hi << 32);
    }

    __device__ inline bool conflict_bit(const uint32_t* cm, uint32_t slot) const {
        uint32_t m = cm[0];
#pragma unroll
        for(int s = 1; s < S; ++s) {
            if((slot >> 5) == uint32_t(s)) {
                m = cm[s];
            }
        }
        return (m >> (slot & 31u)) & 1u;
    }

    // Clips by the half-space a x + b y + c z + d >= 0 (seed-centred frame).
    // Returns true when the cell changed.
    __device__ inline bool clip(double a, double b, double c, double d, uint32_t g) {
        ++nP;
        const float nmaxf = __double2float_ru(fmax(fabs(a), fmax(fabs(b), fabs(c))));
        const float dabs = __double2float_ru(fabs(d));
        uint32_t cm[S];
        bool cf[S];
        bool unc = false;
        uint32_t any = 0;
#pragma unroll
        for(int s = 0; s < S; ++s) {
            bool valid = (vmask[s] >> lane) & 1u;
            double dot = fma(a, X[s], fma(b, Y[s], fma(c, Z[s], d * W[s])));
            // conflict only when CERTAINLY outside: a tie (|dot| inside the error bound)
            // leaves the vertex alone, like the reference's fast mode where ties are not
            // conflicts (convex_cell.cpp:968); the cell is flagged UNCERTAIN
            float E = fmaf(nmaxf, ea[s], dabs * eb[s]);
            bool tie = __double2float_rd(fabs(dot)) <= E;
            cf[s] = valid && !tie && (dot < 0.0);
            unc |= valid && tie;
            cm[s] = __ballot_sync(kFull, cf[s]);
            any |= cm[s];
            nC += __popc(vmask[s]);
        }
        if(__any_sync(kFull, unc)) {
            status |= B200VOR_CELL_UNCERTAIN;
        }
        if(any == 0) {
            return false;
        }
        bool all = true;
#pragma unroll
        for(int s = 0; s < S; ++s) {
            all = all && (cm[s] == vmask[s]);
        }
        if(all) {
#pragma unroll
            for(int s = 0; s < S; ++s) {
                vmask[s] = 0;
            }
            status |= B200VOR_CELL_EMPTY;
            R2 = 0.f;
            return true;
        }
        // plane slot (garbage-collect the plane table when full)
        if(live == ~0ull) {
            live = used_planes();
            if(live == ~0ull) {
                status |= B200VOR_CELL_OVERFLOW;
                return false;
            }
        }
        const uint32_t pn = uint32_t(__ffsll((long long)~live)) - 1u;
        live |= 1ull << pn;
        if(lane == 0) {
            double* p = sh->pl + 4 * pn;
            p[0] = a; p[1] = b; p[2] = c; p[3] = d;
            sh->pnm[pn] = nmaxf;
            sh->gid[pn] = g;
        }
        // border edges, seen from the surviving side
        uint32_t bits = 0;
#pragma unroll
        for(int s = 0; s < S; ++s) {
            bool valid = (vmask[s] >> lane) & 1u;
            if(valid && !cf[s]) {
#pragma unroll
                for(int e = 0; e < 3; ++e) {
                    uint32_t o = (adj[s] >> (8 * e)) & uint32_t(NT - 1);
                    if(conflict_bit(cm, o)) {
                        bits |= 1u << (3 * s + e);
                    }
                }
            }
        }
        const uint32_t cnt = __popc(bits);
        uint32_t inc = cnt;
#pragma unroll
        for(int o = 1; o < 32; o <<= 1) {
            uint32_t t = __shfl_up_sync(kFull, inc, o);
            if(lane >= uint32_t(o)) {
                inc += t;
            }
        }
        const uint32_t total = __shfl_sync(kFull, inc, 31);
        const uint32_t pre = inc - cnt;
        uint32_t fm[S];
        uint32_t nfree = 0;
#pragma unroll
        for(int s = 0; s < S; ++s) {
            fm[s] = ~vmask[s] | cm[s];
            nfree += __popc(fm[s]);
        }
        if(total > nfree) {
            status |= B200VOR_CELL_OVERFLOW;
            return false;
        }
        uint32_t r = pre;
#pragma unroll
        for(int s = 0; s < S; ++s) {
#pragma unroll
            for(int e = 0; e < 3; ++e) {
                if(bits & (1u << (3 * s + e))) {
                    uint32_t v1 = (tri[s] >> (8 * ((e + 1) % 3))) & 0xffu;
                    uint32_t v2 = (tri[s] >> (8 * ((e + 2) % 3))) & 0xffu;
                    sh->rec[r++] = (lane + 32u * s) | (v2 << 16) | (v1 << 24);
                }
            }
        }
        __syncwarp();
        bool take[S];
        uint32_t newm[S];
        uint32_t base = 0;
#pragma unroll
        for(int s = 0; s < S; ++s) {
            bool isfree = (fm[s] >> lane) & 1u;
            uint32_t rank = base + __popc(fm[s] & lanemask_lt());
            take[s] = isfree && rank < total;
            if(take[s]) {
                uint32_t rr = sh->rec[rank];
                uint32_t o = rr & 0xffu, x = (rr >> 16) & 0xffu, y = rr >> 24;
                uint32_t me = lane + 32u * s;
                tri[s] = pn | (x << 8) | (y << 16);
                adj[s] = o;
                sh->slot_of[rank] = uint8_t(me);
                sh->first_by[x] = uint8_t(me);
                sh->second_by[y] = uint8_t(me);
            }
            newm[s] = __ballot_sync(kFull, take[s]);
            base += __popc(fm[s]);
        }
        __syncwarp();
#pragma unroll
        for(int s = 0; s < S; ++s) {
            if(take[s]) {
                uint32_t x = (tri[s] >> 8) & 0xffu, y = (tri[s] >> 16) & 0xffu;
                adj[s] |= (uint32_t(sh->first_by[y]) << 8) | (uint32_t(sh->second_by[x]) << 16);
            }
        }
        r = pre;
#pragma unroll
        for(int s = 0; s < S; ++s) {
#pragma unroll
            for(int e = 0; e < 3; ++e) {
                if(bits & (1u << (3 * s + e))) {
                    uint32_t ns = sh->slot_of[r++];
                    adj[s] = (adj[s] & ~(0xffu << (8 * e))) | (ns << (8 * e));
                }
            }
        }
        uint32_t rb = 0;
#pragma unroll
        for(int s = 0; s < S; ++s) {
            vmask[s] = (vmask[s] & ~cm[s]) | newm[s];
            if(take[s]) {
                make_vertex(s, pn, (tri[s] >> 8) & 0xffu, (tri[s] >> 16) & 0xffu);
            }
            if((vmask[s] >> lane) & 1u) {
                rb = max(rb, __float_as_uint(r2[s]));
            }
        }
        nV += total;
        R2 = __uint_as_float(__reduce_max_sync(kFull, rb));
        __syncwarp();
        return true;
    }
};

__device__ inline double warp_sum(double v) {
#pragma unroll
    for(int o = 16; o > 0; o >>= 1) {
        v += __shfl_xor_sync(kFull, v, o);
    }
    return v;
}

// Volume, mass*centroid, faces.  Returns through lane-uniform outputs.
template <int S, bool FACES>
__device__ inline void integrate(WarpCell<S>& C, const CellParams& P, const Seed& si,
                                 uint32_t row, uint32_t& ntri) {
    const uint32_t lane = C.lane;
    SharedLayout<S>* sh = C.sh;
    ntri = 0;
#pragma unroll
    for(int s = 0; s < S; ++s) {
        ntri += __popc(C.vmask[s]);
        if((C.vmask[s] >> lane) & 1u) {
            double iw = 1.0 / C.W[s];
            double* v = sh->vtx + 3 * (lane + 32 * s);
            v[0] = C.X[s] * iw;
            v[1] = C.Y[s] * iw;
            v[2] = C.Z[s] * iw;
        }
    }
    __syncwarp();
    double m6 = 0.0, gx = 0.0, gy = 0.0, gz = 0.0, e6 = 0.0;
    const bool want_e = P.energy != nullptr;
    double ncr[S][3];   // n_f . (u x v) per outgoing directed edge (for the areas)
#pragma unroll
    for(int s = 0; s < S; ++s) {
#pragma unroll
        for(int l = 0; l < 3; ++l) {
            ncr[s][l] = 0.0;
        }
        if((C.vmask[s] >> lane) & 1u) {
            const double* u = sh->vtx + 3 * (lane + 32 * s);
            double ux = u[0], uy = u[1], uz = u[2];
#pragma unroll
            for(int l = 0; l < 3; ++l) {
                uint32_t f = (C.tri[s] >> (8 * l)) & 0xffu;
                uint32_t nb = (C.adj[s] >> (8 * ((l + 1) % 3))) & uint32_t(32 * S - 1);
                const double* v = sh->vtx + 3 * nb;
                double vx = v[0], vy = v[1], vz = v[2];
                const double* pf = sh->pl + 4 * f;
                double a = pf[0], b = pf[1], c = pf[2], d = pf[3];
                double cx = uy * vz - uz * vy, cy = uz * vx - ux * vz, cz = ux * vy - uy * vx;
                double nc = a * cx + b * cy + c * cz;
                double inv = -d / (a * a + b * b + c * c);
                double vol6 = inv * nc;                  // foot . (u x v), foot = inv * n
                m6 += vol6;
                const double fx = a * inv, fy = b * inv, fz = c * inv;
                gx += vol6 * (fx + ux + vx);
                gy += vol6 * (fy + uy + vy);
                gz += vol6 * (fz + uz + vz);
                if(want_e) {   // RVD.cpp:846-857 on the simplex (seed, foot, u, v)
                    double q = (fx * fx + fy * fy + fz * fz) + (ux * ux + uy * uy + uz * uz) +
                               (vx * vx + vy * vy + vz * vz) + (fx * ux + fy * uy + fz * uz) +
                               (ux * vx + uy * vy + uz * vz) + (vx * fx + vy * fy + vz * fz);
                    e6 += vol6 * q;
                }
                ncr[s][l] = nc;
            }
        }
    }
    m6 = warp_sum(m6);
    gx = warp_sum(gx);
    gy = warp_sum(gy);
    gz = warp_sum(gz);
    if(want_e) {
        e6 = warp_sum(e6);
    }
    double sgn = m6 < 0.0 ? -1.0 : 1.0;
    double mass = sgn * m6 / 6.0;
    if(lane == 0) {
        bool dead = (C.status & (B200VOR_CELL_EMPTY | B200VOR_CELL_DUPLICATE)) != 0;
        double mm = dead ? 0.0 : mass;
        P.m[row] = mm;
        P.mg[3 * size_t(row)] = dead ? 0.0 : sgn * gx / 24.0 + mm * si.x;
        P.mg[3 * size_t(row) + 1] = dead ? 0.0 : sgn * gy / 24.0 + mm * si.y;
        P.mg[3 * size_t(row) + 2] = dead ? 0.0 : sgn * gz / 24.0 + mm * si.z;
        if(want_e) {
            P.energy[row] = dead ? 0.0 : sgn * e6 / 60.0;
        }
    }
    if(FACES) {
        unsigned long long used = C.used_planes();
        uint32_t nf = 0;
        // reuse rec[] for the ids and pl rows are still needed: areas go to vtx after
        // a barrier (vertices are no longer read once ncr is in registers)
        __syncwarp();
        double* area = sh->vtx;        // [kPlaneCap]
        uint32_t* ids = sh->rec;       // [>= kPlaneCap]
        unsigned long long rest = used;
        while(rest) {
            uint32_t f = uint32_t(__ffsll((long long)rest)) - 1u;
            rest &= rest - 1ull;
            double part = 0.0;
#pragma unroll
            for(int s = 0; s < S; ++s) {
                if((C.vmask[s] >> lane) & 1u) {
#pragma unroll
                    for(int l = 0; l < 3; ++l) {
                        if(((C.tri[s] >> (8 * l)) & 0xffu) == f) {
                            part += ncr[s][l];
                        }
                    }
                }
            }
            part = warp_sum(part);
            if(lane == 0) {
                const double* pf = sh->pl + 4 * f;
                double nn = sqrt(pf[0] * pf[0] + pf[1] * pf[1] + pf[2] * pf[2]);
                area[nf] = 0.5 * fabs(part) / nn;
                ids[nf] = sh->gid[f];
            }
            ++nf;
        }
        __syncwarp();
        if(C.status & (B200VOR_CELL_EMPTY | B200VOR_CELL_DUPLICATE)) {
            nf = 0;
        }
        // rank sort by id, write rows of fcap
        for(uint32_t k = lane; k < nf; k += 32) {
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
        if(lane == 0) {
            P.nfaces[row] = nf;
        }
        __syncwarp();
    }
}

template <int S>
__device__ inline bool any_box_face(WarpCell<S>& C) {
    // a used plane slot whose global id is one of the six helper-box planes
    unsigned long long used = C.used_planes();
    bool mine = false;
#pragma unroll
    for(int h = 0; h < 2; ++h) {
        uint32_t p = C.lane + 32u * h;
        mine = mine || (((used >> p) & 1ull) && C.sh->gid[p] >= B200VOR_BOX_ID);
    }
    return __any_sync(kFull, mine);
}

// Integrity of the dual triangulation: Euler's relation T = 2F - 4 and symmetric
// adjacency.  Fails only when uncertain sign tests produced a conflict zone that is
// not a disk (degenerate input).
template <int S>
__device__ inline bool consistent(WarpCell<S>& C) {
    constexpr uint32_t NT = 32 * S;
    uint32_t T = 0;
#pragma unroll
    for(int s = 0; s < S; ++s) {
        T += __popc(C.vmask[s]);
        C.sh->rec[C.lane + 32 * s] = ((C.vmask[s] >> C.lane) & 1u) ? C.adj[s] : 0xffffffffu;
    }
    __syncwarp();
    uint32_t F = __popcll(C.used_planes());
    bool ok = true;
#pragma unroll
    for(int s = 0; s < S; ++s) {
        if((C.vmask[s] >> C.lane) & 1u) {
            uint32_t me = C.lane + 32 * s;
#pragma unroll
            for(int e = 0; e < 3; ++e) {
                uint32_t nb = (C.adj[s] >> (8 * e)) & 0xffu;
                uint32_t back = nb < NT ? C.sh->rec[nb] : 0xffffffffu;
                bool sym = back != 0xffffffffu &&
                           ((back & 0xffu) == me || ((back >> 8) & 0xffu) == me ||
                            ((back >> 16) & 0xffu) == me);
                ok = ok && sym;
            }
        }
    }
    ok = __all_sync(kFull, ok && !C.degenerate);
    __syncwarp();
    return ok && (T + 4 == 2 * F);
}

// ---- mode 0: candidates come from the kNN table, ascending distance ----
template <int S, bool FACES, int WPB>
__global__ void __launch_bounds__(32 * WPB) cells_knn_kernel(CellParams P) {
    __shared__ SharedLayout<S> shared[WPB];
    const uint32_t lane = threadIdx.x & 31u, wid = threadIdx.x >> 5;
    const uint32_t c = blockIdx.x * WPB + wid;
    if(c >= P.job.nq) {
        return;
    }
    const uint32_t i = P.job.qlist ? P.job.qlist[c] : c;
    const Seed si = load_seed(P.sorted + i);
    const uint32_t row = P.perm[i];
    WarpCell<S> C;
    C.lane = lane;
    C.sh = &shared[wid];
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

    const uint32_t k = P.job.k;
    const uint32_t trow = P.job.rowmap ? P.job.rowmap[c] : c;
    const uint32_t* nrow = P.job.knn + size_t(trow) * k;
    const double dw = P.has_w ? (P.wmax - si.w) : 0.0;
    bool certified = false, done = false;
    uint32_t extra = 0;
    for(uint32_t base = 0; base < k && !done; base += 32) {
        uint32_t jr = base + lane;
        uint32_t id = jr < k ? __ldg(nrow + jr) : 0xFFFFFFFFu;
        double pa = 0, pb = 0, pc = 0, pd = 0, d2 = 0;
        uint32_t uid = 0xFFFFFFFFu;
        if(id != 0xFFFFFFFFu) {
            uint32_t j = id, inst = 0;
            if(P.periodic) {
                inst = id / P.n;
                j = id - inst * P.n;
            }
            Seed sj = load_seed(P.sorted + j);
            double qx = sj.x - si.x, qy = sj.y - si.y, qz = sj.z - si.z;
            if(inst) {
                qx = (sj.x + c_translation[inst][0] * P.period) - si.x;
                qy = (sj.y + c_translation[inst][1] * P.period) - si.y;
                qz = (sj.z + c_translation[inst][2] * P.period) - si.z;
            }
            d2 = qx * qx + qy * qy + qz * qz;
            pa = -2.0 * qx; pb = -2.0 * qy; pc = -2.0 * qz;
            pd = d2 + (si.w - sj.w);
            uid = __ldg(P.perm + j) + inst * P.n;
        }
        for(int t = 0; t < 32; ++t) {
            uint32_t idt = __shfl_sync(kFull, id, t);
            if(idt == 0xFFFFFFFFu) {
                if(base + t < k) {
                    certified = true;   // table row shorter than k: every seed was offered
                }
                done = true;
                break;
            }
            double d2t = __shfl_sync(kFull, d2, t);
            if(idt == i) {
                continue;
            }
            if(d2t == 0.0) {
                // exact duplicate (delaunay_nn.cpp:122-135): the lower user index keeps the cell
                uint32_t ut = __shfl_sync(kFull, uid, t);
                if(ut < row) {
                    extra |= B200VOR_CELL_DUPLICATE;
                }
                continue;
            }
            double bound = 4.1 * double(C.R2);
            if(P.has_w) {
                double R = sqrt(double(C.R2));
                double sr = R + sqrt(double(C.R2) + dw);
                bound = 1.025 * sr * sr;
            }
            if(d2t > bound) {
                certified = true;
                done = true;
                break;
            }
            double a = __shfl_sync(kFull, pa, t), b = __shfl_sync(kFull, pb, t),
                   cc = __shfl_sync(kFull, pc, t), d = __shfl_sync(kFull, pd, t);
            uint32_t ut = __shfl_sync(kFull, uid, t);
            C.clip(a, b, cc, d, ut);
            if(C.status & (B200VOR_CELL_OVERFLOW | B200VOR_CELL_EMPTY)) {
                done = true;
                certified = (C.status & B200VOR_CELL_EMPTY) != 0;
                break;
            }
        }
    }
    C.status |= extra;
    if(P.job.all_offered) {
        certified = true;    // the table row holds every other seed
    }
    if(P.job.knn_flag && P.job.knn_flag[trow]) {
        certified = false;
    }
    if(!certified) {
        C.status |= B200VOR_CELL_UNCERTIFIED;
    }
    if(P.periodic && !(C.status & B200VOR_CELL_EMPTY) && any_box_face(C)) {
        C.status |= B200VOR_CELL_BOXFACE;
    }
    if(!(C.status & B200VOR_CELL_EMPTY) && !consistent(C)) {
        C.status |= B200VOR_CELL_DEGENERATE;
    }
    uint32_t ntri;
    integrate<S, FACES>(C, P, si, row, ntri);
    if(lane == 0) {
        P.status[row] = uint8_t(C.status);
        if(P.evcount) {
            P.evcount[2 * size_t(row)] = min(C.nP, 0xffffu) | (min(C.nV, 0xfffu) << 16);
            P.evcount[2 * size_t(row) + 1] = min(C.nC, 0xfffffu) | (ntri << 20);
        }
    }
}

// ---- helpers: failed-cell compaction, Lloyd update, event reduction ----
__global__ void compact_kernel(const uint32_t* __restrict__ qin, uint32_t nq,
                               const uint32_t* __restrict__ perm,
                               const uint8_t* __restrict__ status, uint8_t mask,
                               uint32_t* __restrict__ qout, uint32_t* __restrict__ rows_out,
                               const uint32_t* __restrict__ rows_in,
                               uint32_t* __restrict__ counter) {
    uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    bool fail = false;
    uint32_t i = 0;
    if(t < nq) {
        i = qin ? qin[t] : t;
        fail = (status[perm[i]] & mask) != 0;
    }
    // warp-aggregated append (order inside the list does not affect any result)
    uint32_t bal = __ballot_sync(kFull, fail);
    uint32_t lane = threadIdx.x & 31u;
    uint32_t base = 0;
    if(lane == 0 && bal) {
        base = atomicAdd(counter, __popc(bal));
    }
    base = __shfl_sync(kFull, base, 0);
    if(fail) {
        uint32_t pos = base + __popc(bal & ((1u << lane) - 1u));
        qout[pos] = i;
        if(rows_out) {
            rows_out[pos] = rows_in ? rows_in[t] : t;
        }
    }
}

// grid-order indices of the seeds whose USER index lies in [u0, u1) (multi-GPU shard)
__global__ void select_range_kernel(uint32_t n, const uint32_t* __restrict__ perm, uint32_t u0,
                                    uint32_t u1, uint32_t* __restrict__ qout,
                                    uint32_t* __restrict__ counter) {
    // one block-wide ordered compaction per 1024 seeds keeps the list in grid order
    __shared__ uint32_t wcount[32];
    __shared__ uint32_t base;
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    bool in = false;
    if(i < n) {
        uint32_t u = perm[i];
        in = u >= u0 && u < u1;
    }
    uint32_t lane = threadIdx.x & 31u, wid = threadIdx.x >> 5;
    uint32_t bal = __ballot_sync(kFull, in);
    if(lane == 0) {
        wcount[wid] = __popc(bal);
    }
    __syncthreads();
    if(threadIdx.x == 0) {
        uint32_t tot = 0;
        for(uint32_t w = 0; w < blockDim.x / 32; ++w) {
            uint32_t cw = wcount[w];
            wcount[w] = tot;
            tot += cw;
        }
        // blocks append in arbitrary order; the per-block runs stay in grid order, and the
        // result of every cell is independent of the list order
        base = atomicAdd(counter, tot);
    }
    __syncthreads();
    if(in) {
        qout[base + wcount[wid] + __popc(bal & ((1u << lane) - 1u))] = i;
    }
}

__global__ void lloyd_kernel(uint32_t n, uint32_t u0, double* __restrict__ xyz,
                             const double* __restrict__ m, const double* __restrict__ mg,
                             int periodic, double period) {
    uint32_t i = u0 + blockIdx.x * blockDim.x + threadIdx.x;
    if(i >= n) {
        return;
    }
    double mm = m[i];
    if(mm > 1e-30) {   // voronoi/CVT.cpp:154-163: s = 1 / m, p = s * mg (same rounding as the reference)
        const double s = 1.0 / mm;
#pragma unroll
        for(int a = 0; a < 3; ++a) {
            double x = s * mg[3 * size_t(i) + a];
            if(periodic) {   // demo_Delaunay3d/main.cpp:443-451
                if(x < 0.0) {
                    x += period;
                }
                if(x > period) {
                    x -= period;
                }
            }
            xyz[3 * size_t(i) + a] = x;
        }
    }
}

__global__ void reduce_events_kernel(uint32_t n, const uint32_t* __restrict__ ev,
                                     const uint8_t* __restrict__ status,
                                     unsigned long long* __restrict__ sums) {
    // sums: 0..3 = P, C, V, I ; 4..10 = status bit counts ; 11 = grid scans (fused walk)
    unsigned long long acc[12] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    for(uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        if(ev) {
            uint32_t a = ev[2 * size_t(i)], b = ev[2 * size_t(i) + 1];
            acc[0] += a & 0xffffu;
            acc[2] += (a >> 16) & 0xfffu;
            acc[11] += a >> 28;
            acc[1] += b & 0xfffffu;
            acc[3] += 3ull * (b >> 20);
        }
        uint8_t s = status[i];
        for(int k = 0; k < 7; ++k) {
            acc[4 + k] += (s >> k) & 1u;
        }
    }
    for(int k = 0; k < 12; ++k) {
        unsigned long long v = acc[k];
        for(int o = 16; o > 0; o >>= 1) {
            v += __shfl_xor_sync(kFull, v, o);
        }
        if((threadIdx.x & 31) == 0 && v) {
            atomicAdd(&sums[k], v);
        }
    }
}

}  // namespace

int run_cells(Ctx* c, const CellJob& job) {
    if(job.nq == 0) {
        return B200VOR_OK;
    }
    CellParams P;
    P.G = c->grid;
    P.sorted = c->sorted.p;
    P.cell_start = c->cell_start.p;
    P.perm = c->perm.p;
    P.n = c->n;
    P.periodic = c->periodic;
    P.period = c->period;
    for(int a = 0; a < 6; ++a) {
        P.box[a] = c->box[a];
    }
    P.has_w = c->has_w ? 1 : 0;
    P.wmax = c->wmax;
    P.job = job;
    P.m = c->m.p;
    P.mg = c->mg.p;
    P.energy = job.want_energy ? c->energy.p : nullptr;
    P.status = c->status.p;
    P.nfaces = c->nfaces.p;
    P.face_nbr = c->face_nbr.p;
    P.face_area = c->face_area.p;
    P.fcap = c->fcap;
    P.evcount = job.collect_events ? c->evcount.p : nullptr;
    if(job.slots == 4) {
        constexpr int W4 = 4;
        uint32_t nblk = (job.nq + W4 - 1) / W4;
        if(job.want_faces) {
            cells_knn_kernel<4, true, W4><<<nblk, 32 * W4, 0, c->stream>>>(P);
        } else {
            cells_knn_kernel<4, false, W4><<<nblk, 32 * W4, 0, c->stream>>>(P);
        }
    } else {
        uint32_t nblk = (job.nq + kWarpsPerBlock - 1) / kWarpsPerBlock;
        if(job.want_faces) {
            cells_knn_kernel<2, true, kWarpsPerBlock><<<nblk, 32 * kWarpsPerBlock, 0, c->stream>>>(P);
        } else {
            cells_knn_kernel<2, false, kWarpsPerBlock><<<nblk, 32 * kWarpsPerBlock, 0, c->stream>>>(P);
        }
    }
    c->stats.gpu_launches++;
    B200_CUDA_CHECK(cudaGetLastError());
    return B200VOR_OK;
}

int compact_failed(Ctx* c, const uint32_t* qlist_in, uint32_t nq_in, uint32_t* qlist_out,
                   uint32_t* rows_out, uint32_t* count_out_host, uint8_t mask) {
    int rc = c->counter.reserve(4);
    if(rc) {
        return rc;
    }
    B200_CUDA_CHECK(cudaMemsetAsync(c->counter.p, 0, sizeof(uint32_t), c->stream));
    uint32_t nblk = (nq_in + 255) / 256;
    compact_kernel<<<nblk, 256, 0, c->stream>>>(qlist_in, nq_in, c->perm.p, c->status.p, mask,
                                                qlist_out, rows_out, nullptr, c->counter.p);
    c->stats.gpu_launches++;
    B200_CUDA_CHECK(cudaMemcpyAsync(count_out_host, c->counter.p, sizeof(uint32_t),
                                    cudaMemcpyDeviceToHost, c->stream));
    B200_CUDA_CHECK(cudaStreamSynchronize(c->stream));
    return B200VOR_OK;
}

int select_range(Ctx* c, uint32_t* qout, uint32_t* count_out_host) {
    int rc = c->counter.reserve(4);
    if(rc) {
        return rc;
    }
    B200_CUDA_CHECK(cudaMemsetAsync(c->counter.p, 0, sizeof(uint32_t), c->stream));
    uint32_t nblk = (c->n + 1023) / 1024;
    select_range_kernel<<<nblk, 1024, 0, c->stream>>>(c->n, c->perm.p, c->q0, c->q1, qout,
                                                      c->counter.p);
    c->stats.gpu_launches++;
    B200_CUDA_CHECK(cudaMemcpyAsync(count_out_host, c->counter.p, sizeof(uint32_t),
                                    cudaMemcpyDeviceToHost, c->stream));
    B200_CUDA_CHECK(cudaStreamSynchronize(c->stream));
    return B200VOR_OK;
}

int lloyd_update(Ctx* c) {
    uint32_t u1 = c->q1 < c->n ? c->q1 : c->n;
    uint32_t cnt = u1 > c->q0 ? u1 - c->q0 : 0;
    if(cnt == 0) {
        return B200VOR_OK;
    }
    uint32_t nblk = (cnt + 255) / 256;
    lloyd_kernel<<<nblk, 256, 0, c->stream>>>(u1, c->q0, c->xyz.p, c->m.p, c->mg.p, c->periodic,
                                              c->period);
    c->stats.gpu_launches++;
    B200_CUDA_CHECK(cudaGetLastError());
    return B200VOR_OK;
}

int reduce_events(Ctx* c) {
    int rc = c->sums.reserve(16);
    if(rc) {
        return rc;
    }
    B200_CUDA_CHECK(cudaMemsetAsync(c->sums.p, 0, 16 * sizeof(unsigned long long), c->stream));
    reduce_events_kernel<<<296, 256, 0, c->stream>>>(c->n, c->evcount.cap ? c->evcount.p : nullptr,
                                                     c->status.p, c->sums.p);
    c->stats.gpu_launches++;
    unsigned long long h[16];
    B200_CUDA_CHECK(cudaMemcpyAsync(h, c->sums.p, sizeof(h), cudaMemcpyDeviceToHost, c->stream));
    B200_CUDA_CHECK(cudaStreamSynchronize(c->stream));
    c->stats.sum_planes = h[0];
    c->stats.sum_tests = h[1];
    c->stats.sum_created = h[2];
    c->stats.sum_tets = h[3];
    c->stats.n_uncertified = uint32_t(h[4]);
    c->stats.n_uncertain = uint32_t(h[5]);
    c->stats.n_empty = uint32_t(h[6]);
    c->stats.n_boxface = uint32_t(h[7]);
    c->stats.n_overflow = uint32_t(h[8]);
    c->stats.n_duplicate = uint32_t(h[9]);
    c->stats.n_degenerate = uint32_t(h[10]);
    c->stats.sum_scans = h[11];
    return B200VOR_OK;
}

}  // namespace b200vor
