/*
 * oracle/voro_oracle.cpp -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement (own structure, not a copy) of the reference algorithm for the
 * meshless Voronoi/Laguerre cell path.  Citations are relative to
 * /root/reference/src/lib/geogram/.
 *
 *   KdIndex            balanced kd-tree kNN       points/kd_tree.cpp:429-517 (build),
 *                                                 :244-387 (query), kd_tree.h:173-197 (k-list)
 *   sq_dist            ((dx^2)+dy^2)+dz^2, no FMA  basic/geometry_nd.h:65-73
 *   DualCell           half-space cell, dual form  voronoi/convex_cell.cpp:175-215 (box),
 *                                                 :467-516 + :761-865 (clip), :971-1082 (vertex),
 *                                                 :1267-1307 (mass / centroid), :1160-1191 (facet area),
 *                                                 :1312-1330 (squared radius)
 *   vo_cells           security radius + k growth  voronoi/generic_RVD.h:2278-2347,
 *                      neighbour glue              delaunay/delaunay_nn.cpp:105-145
 *   power planes       (2(pi-pj), (wi-|pi|^2)-(wj-|pj|^2))   delaunay/periodic_delaunay_3d.cpp:3719-3726
 *   periodic ids       real + n*instance           delaunay/periodic.h:95-113, periodic.cpp:44-86
 *   vo_lloyd_update    p <- mg/m if m > 1e-30      voronoi/CVT.cpp:154-163, wrap of
 *                                                 examples/graphics/demo_Delaunay3d/main.cpp:443-451
 *
 * Pinning: tests/test_oracle_vs_reference.py checks this file against the
 * unmodified reference (oracle/_ref/libgeoref.so) where that library exists, and
 * against tests/golden/ fixtures generated from the reference otherwise.
 *
 * Differences from the reference, on purpose:
 *   - the sign test is the reference's plain fp64 4x4 determinant
 *     (convex_cell.cpp:958-968); a test whose |det| is inside the semi-static filter
 *     bound of numerics/predicates/det4d.h is re-evaluated in long double and the
 *     cell is flagged VO_FLAG_TIE (exact arithmetic is not restated);
 *   - no vertex at infinity: every cell starts from a finite box.
 *
 * Built with -frounding-math -ffp-contract=off like the reference
 * (cmake/platforms/Linux-gcc.cmake:43-44).
 */
#include "voro_oracle.h"

#include <algorithm>
#include <cmath>
#include <cstring>
#include <limits>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

namespace {

int g_threads = 0;
double* g_energy = nullptr;   /* vo_set_energy_output */

inline double sq(double x) { return x * x; }

/* basic/geometry_nd.h:65-73 */
inline double sq_dist(const double* p, const double* q) {
    double r = 0.0;
    r += sq(q[0] - p[0]);
    r += sq(q[1] - p[1]);
    r += sq(q[2] - p[2]);
    return r;
}

/*************************** kd-tree kNN ***************************/

struct KList {
    uint32_t kmax, k;
    uint32_t* id;
    double* d2;
    double furthest() const {
        return k == kmax ? d2[k - 1] : std::numeric_limits<double>::max();
    }
    /* kd_tree.h:173-197 : a later point with an equal distance goes first */
    void insert(uint32_t j, double d) {
        int i;
        for(i = int(k); i > 0; --i) {
            if(d2[i - 1] < d) {
                break;
            }
            if(uint32_t(i) < kmax) {
                id[i] = id[i - 1];
                d2[i] = d2[i - 1];
            }
        }
        if(uint32_t(i) < kmax) {
            id[i] = j;
            d2[i] = d;
        }
        if(k < kmax) {
            ++k;
        }
    }
};

struct KdIndex {
    static const uint32_t LEAF = 16;   /* kd_tree.h:334 */
    uint32_t n;
    const double* pts;
    std::vector<uint32_t> perm;
    std::vector<uint8_t> coord;
    std::vector<double> val;
    double lo[3], hi[3];

    uint32_t max_node(uint32_t node, uint32_t b, uint32_t e) const {
        if(e - b <= LEAF) {
            return node;
        }
        uint32_t m = b + (e - b) / 2;
        return std::max(max_node(2 * node, b, m), max_node(2 * node + 1, m, e));
    }

    void build(uint32_t n_in, const double* p) {
        n = n_in;
        pts = p;
        perm.resize(n);
        for(uint32_t i = 0; i < n; ++i) {
            perm[i] = i;
        }
        for(int c = 0; c < 3; ++c) {
            lo[c] = std::numeric_limits<double>::max();
            hi[c] = -std::numeric_limits<double>::max();
        }
        for(uint32_t i = 0; i < n; ++i) {
            for(int c = 0; c < 3; ++c) {
                lo[c] = std::min(lo[c], p[3 * size_t(i) + c]);
                hi[c] = std::max(hi[c], p[3 * size_t(i) + c]);
            }
        }
        uint32_t sz = max_node(1, 0, n) + 1;
        coord.assign(sz, 0);
        val.assign(sz, 0.0);
        split(1, 0, n);
    }

    void split(uint32_t node, uint32_t b, uint32_t e) {
        if(e - b <= LEAF) {
            return;
        }
        /* coordinate of largest spread, kd_tree.cpp:519-538 */
        int best = 0;
        double best_spread = -1.0;
        for(int c = 0; c < 3; ++c) {
            double mn = std::numeric_limits<double>::max(), mx = -mn;
            for(uint32_t i = b; i < e; ++i) {
                double x = pts[3 * size_t(perm[i]) + c];
                mn = std::min(mn, x);
                mx = std::max(mx, x);
            }
            if(mx - mn > best_spread) {
                best_spread = mx - mn;
                best = c;
            }
        }
        uint32_t m = b + (e - b) / 2;
        const double* P = pts;
        std::nth_element(
            perm.begin() + b, perm.begin() + m, perm.begin() + e,
            [P, best](uint32_t i, uint32_t j) {
                return P[3 * size_t(i) + best] < P[3 * size_t(j) + best];
            });
        coord[node] = uint8_t(best);
        val[node] = pts[3 * size_t(perm[m]) + best];
        split(2 * node, b, m);
        split(2 * node + 1, m, e);
    }

    void leaf(uint32_t b, uint32_t e, const double* q, KList& L) const {
        double d[LEAF];
        for(uint32_t i = b; i < e; ++i) {
            d[i - b] = sq_dist(q, pts + 3 * size_t(perm[i]));
        }
        double R = L.furthest();
        for(uint32_t i = b; i < e; ++i) {
            if(d[i - b] <= R) {
                L.insert(perm[i], d[i - b]);
                R = L.furthest();
            }
        }
    }

    /* kd_tree.cpp:244-341: near child first, far child iff box_dist <= furthest */
    void visit(uint32_t node, uint32_t b, uint32_t e, double* bmin, double* bmax,
               double box_dist, const double* q, KList& L) const {
        if(e - b <= LEAF) {
            leaf(b, e, q, L);
            return;
        }
        int c = coord[node];
        double v = val[node];
        uint32_t m = b + (e - b) / 2;
        double cut = q[c] - v;
        if(cut < 0.0) {
            double save = bmax[c];
            bmax[c] = v;
            visit(2 * node, b, m, bmin, bmax, box_dist, q, L);
            bmax[c] = save;
            double diff = bmin[c] - q[c];
            if(diff > 0.0) {
                box_dist -= sq(diff);
            }
            box_dist += sq(cut);
            if(box_dist <= L.furthest()) {
                save = bmin[c];
                bmin[c] = v;
                visit(2 * node + 1, m, e, bmin, bmax, box_dist, q, L);
                bmin[c] = save;
            }
        } else {
            double save = bmin[c];
            bmin[c] = v;
            visit(2 * node + 1, m, e, bmin, bmax, box_dist, q, L);
            bmin[c] = save;
            double diff = q[c] - bmax[c];
            if(diff > 0.0) {
                box_dist -= sq(diff);
            }
            box_dist += sq(cut);
            if(box_dist <= L.furthest()) {
                save = bmax[c];
                bmax[c] = v;
                visit(2 * node, b, m, bmin, bmax, box_dist, q, L);
                bmax[c] = save;
            }
        }
    }

    void query(const double* q, uint32_t k, uint32_t* id, double* d2) const {
        double bmin[3], bmax[3], box_dist = 0.0;
        for(int c = 0; c < 3; ++c) {
            bmin[c] = lo[c];
            bmax[c] = hi[c];
            if(q[c] < lo[c]) {
                box_dist += sq(lo[c] - q[c]);
            } else if(q[c] > hi[c]) {
                box_dist += sq(hi[c] - q[c]);
            }
        }
        KList L = {k, 0, id, d2};
        visit(1, 0, n, bmin, bmax, box_dist, q, L);
    }
};

/*************************** dual-form cell ***************************/

struct Plane4 {
    double a, b, c, d;
};

struct DTri {
    uint16_t v[3];     /* the three planes meeting at this cell vertex        */
    uint16_t adj[3];   /* adj[e]: triangle across the edge opposite to v[e]  */
    bool alive;
};

inline double det3(double a11, double a12, double a13, double a21, double a22,
                   double a23, double a31, double a32, double a33) {
    return a11 * (a22 * a33 - a23 * a32) - a21 * (a12 * a33 - a13 * a32) +
           a31 * (a12 * a23 - a13 * a22);
}

struct DualCell {
    std::vector<Plane4> plane;
    std::vector<uint32_t> gid;
    std::vector<DTri> tri;
    std::vector<uint16_t> free_slots;
    std::vector<int> first_by, second_by;   /* scratch for the border cycle */
    bool tie = false;
    uint64_t nP = 0, nC = 0, nV = 0;

    /* convex_cell.cpp:175-215 (without the vertex at infinity); the 8 corner
     * triangles are generated, each ordered so that det(normals) > 0. */
    void init_box(const double* bx) {
        plane.clear();
        gid.clear();
        tri.clear();
        free_slots.clear();
        tie = false;
        plane.push_back({1, 0, 0, -bx[0]});
        plane.push_back({-1, 0, 0, bx[3]});
        plane.push_back({0, 1, 0, -bx[1]});
        plane.push_back({0, -1, 0, bx[4]});
        plane.push_back({0, 0, 1, -bx[2]});
        plane.push_back({0, 0, -1, bx[5]});
        for(uint32_t f = 0; f < 6; ++f) {
            gid.push_back(VO_BOX_ID + f);
        }
        for(int c = 0; c < 8; ++c) {
            uint16_t px = uint16_t(c & 1), py = uint16_t(2 + ((c >> 1) & 1)),
                     pz = uint16_t(4 + ((c >> 2) & 1));
            int flips = (c & 1) + ((c >> 1) & 1) + ((c >> 2) & 1);
            DTri T;
            T.v[0] = px;
            T.v[1] = (flips & 1) ? pz : py;
            T.v[2] = (flips & 1) ? py : pz;
            T.alive = true;
            T.adj[0] = T.adj[1] = T.adj[2] = 0xFFFF;
            tri.push_back(T);
        }
        connect();
    }

    /* convex_cell.cpp:1348-1411: adjacency from the oriented-edge table */
    void connect() {
        size_t nv = plane.size();
        std::vector<int> e2t(nv * nv, -1);
        for(size_t t = 0; t < tri.size(); ++t) {
            if(!tri[t].alive) {
                continue;
            }
            for(int e = 0; e < 3; ++e) {
                e2t[tri[t].v[(e + 1) % 3] * nv + tri[t].v[(e + 2) % 3]] = int(t);
            }
        }
        for(size_t t = 0; t < tri.size(); ++t) {
            if(!tri[t].alive) {
                continue;
            }
            for(int e = 0; e < 3; ++e) {
                tri[t].adj[e] =
                    uint16_t(e2t[tri[t].v[(e + 2) % 3] * nv + tri[t].v[(e + 1) % 3]]);
            }
        }
    }

    /* homogeneous vertex by Cramer, convex_cell.cpp:1023-1081 */
    void vertex_h(const DTri& T, double* X) const {
        const Plane4& p = plane[T.v[0]];
        const Plane4& q = plane[T.v[1]];
        const Plane4& r = plane[T.v[2]];
        X[0] = -det3(p.d, p.b, p.c, q.d, q.b, q.c, r.d, r.b, r.c);
        X[1] = -det3(p.a, p.d, p.c, q.a, q.d, q.c, r.a, r.d, r.c);
        X[2] = -det3(p.a, p.b, p.d, q.a, q.b, q.d, r.a, r.b, r.d);
        X[3] = det3(p.a, p.b, p.c, q.a, q.b, q.c, r.a, r.b, r.c);
    }

    void vertex(const DTri& T, double* x) const {
        double X[4];
        vertex_h(T, X);
        x[0] = X[0] / X[3];
        x[1] = X[1] / X[3];
        x[2] = X[2] / X[3];
    }

    /* convex_cell.cpp:870-969, finite case.  The 4x4 determinant of the four
     * plane equations, developed along the new plane, is E . (X,Y,Z,W); the vertex
     * X/W is outside when that has the sign opposite to W.  Ties count as
     * conflict (exact-mode convention, :909-911). */
    bool in_conflict(const DTri& T, const Plane4& E) {
        double X[4];
        vertex_h(T, X);
        double t0 = E.a * X[0], t1 = E.b * X[1], t2 = E.c * X[2], t3 = E.d * X[3];
        double det = ((t0 + t1) + t2) + t3;
        /* semi-static filter in the spirit of predicates/det4d.h */
        const Plane4* P[3] = {&plane[T.v[0]], &plane[T.v[1]], &plane[T.v[2]]};
        double mx = std::fabs(E.a), my = std::fabs(E.b), mz = std::fabs(E.c),
               mw = std::fabs(E.d);
        for(int i = 0; i < 3; ++i) {
            mx = std::max(mx, std::fabs(P[i]->a));
            my = std::max(my, std::fabs(P[i]->b));
            mz = std::max(mz, std::fabs(P[i]->c));
            mw = std::max(mw, std::fabs(P[i]->d));
        }
        double eps = 2.1113540660531680e-14 * (mx * my * mz * mw);
        double s;
        if(det > eps || det < -eps) {
            s = det;
        } else {
            tie = true;
            long double lX[4];
            const Plane4 &p = *P[0], &q = *P[1], &r = *P[2];
            auto d3 = [](long double a11, long double a12, long double a13,
                         long double a21, long double a22, long double a23,
                         long double a31, long double a32, long double a33) {
                return a11 * (a22 * a33 - a23 * a32) - a21 * (a12 * a33 - a13 * a32) +
                       a31 * (a12 * a23 - a13 * a22);
            };
            lX[0] = -d3(p.d, p.b, p.c, q.d, q.b, q.c, r.d, r.b, r.c);
            lX[1] = -d3(p.a, p.d, p.c, q.a, q.d, q.c, r.a, r.d, r.c);
            lX[2] = -d3(p.a, p.b, p.d, q.a, q.b, q.d, r.a, r.b, r.d);
            lX[3] = d3(p.a, p.b, p.c, q.a, q.b, q.c, r.a, r.b, r.c);
            long double ld = (long double)E.a * lX[0] + (long double)E.b * lX[1] +
                             (long double)E.c * lX[2] + (long double)E.d * lX[3];
            s = (ld > 0) ? 1.0 : ((ld < 0) ? -1.0 : 0.0);
        }
        if(s == 0.0) {
            return true;
        }
        return (s > 0.0) != (X[3] > 0.0);
    }

    size_t live_count() const {
        size_t c = 0;
        for(const DTri& T : tri) {
            c += T.alive;
        }
        return c;
    }

    /* convex_cell.cpp:467-516 (classification) and :761-865 (re-triangulation).
     * Returns true when the plane cut something. */
    bool clip(const Plane4& E, uint32_t id) {
        ++nP;
        std::vector<uint8_t> conflict(tri.size(), 0);
        size_t nconf = 0, nlive = 0;
        for(size_t t = 0; t < tri.size(); ++t) {
            if(tri[t].alive) {
                ++nlive;
                if(in_conflict(tri[t], E)) {
                    conflict[t] = 1;
                    ++nconf;
                }
            }
        }
        nC += nlive;
        if(nconf == 0) {
            return false;
        }
        uint16_t pn = uint16_t(plane.size());
        plane.push_back(E);
        gid.push_back(id);
        if(nconf == nlive) {
            for(DTri& T : tri) {
                T.alive = false;
            }
            return true;
        }
        /* one new triangle (pn, v2, v1) per border edge (valid t, edge e) */
        struct Border {
            uint16_t t, e;
        };
        std::vector<Border> border;
        for(size_t t = 0; t < tri.size(); ++t) {
            if(!tri[t].alive || conflict[t]) {
                continue;
            }
            for(uint16_t e = 0; e < 3; ++e) {
                if(conflict[tri[t].adj[e]]) {
                    border.push_back({uint16_t(t), e});
                }
            }
        }
        for(size_t t = 0; t < tri.size(); ++t) {
            if(conflict[t]) {
                tri[t].alive = false;
                free_slots.push_back(uint16_t(t));
            }
        }
        first_by.assign(plane.size(), -1);
        second_by.assign(plane.size(), -1);
        std::vector<uint16_t> created;
        for(const Border& B : border) {
            uint16_t v1 = tri[B.t].v[(B.e + 1) % 3], v2 = tri[B.t].v[(B.e + 2) % 3];
            uint16_t s;
            if(!free_slots.empty()) {
                s = free_slots.back();
                free_slots.pop_back();
            } else {
                s = uint16_t(tri.size());
                tri.push_back(DTri());
            }
            DTri& T = tri[s];
            T.v[0] = pn;
            T.v[1] = v2;
            T.v[2] = v1;
            T.alive = true;
            T.adj[0] = B.t;
            tri[B.t].adj[B.e] = s;
            first_by[v2] = s;
            second_by[v1] = s;
            created.push_back(s);
            ++nV;
        }
        /* around pn: (pn,x,y) meets (pn,y,*) across the edge opposite x and
         * (pn,*,x) across the edge opposite y */
        for(uint16_t s : created) {
            tri[s].adj[1] = uint16_t(first_by[tri[s].v[2]]);
            tri[s].adj[2] = uint16_t(second_by[tri[s].v[1]]);
        }
        return true;
    }

    /* convex_cell.cpp:1312-1330 */
    double squared_radius(const double* c) const {
        double R2 = 0.0;
        for(const DTri& T : tri) {
            if(T.alive) {
                double x[3];
                vertex(T, x);
                R2 = std::max(R2, sq_dist(c, x));
            }
        }
        return R2;
    }

    /* cyclic list of the triangles around plane v starting at t0 */
    void ring(uint16_t v, uint16_t t0, std::vector<uint16_t>& out) const {
        out.clear();
        uint16_t t = t0;
        do {
            out.push_back(t);
            int l = (tri[t].v[0] == v) ? 0 : ((tri[t].v[1] == v) ? 1 : 2);
            t = tri[t].adj[(l + 1) % 3];
        } while(t != t0 && out.size() < 4096);
    }

    /* convex_cell.cpp:1267-1307 (mass, mass*centroid) and :1160-1191 (areas).
     * ids/areas: contributing planes sorted by global id.  Returns #faces,
     * nI = number of integration tetrahedra. */
    /* seed/energy (optional): CVT energy of the cell, the integral of |x - seed|^2
     * (voronoi/RVD.cpp:846-857 restated for a tetrahedron that does not have the
     * seed as a vertex: vol/10 * (sum_i q_i.q_i + sum_{i<j} q_i.q_j), q_i = v_i - seed) */
    uint32_t integrate(double& m, double* mg, uint32_t fcap, uint32_t* ids,
                       double* areas, uint64_t& nI, const double* seed = nullptr,
                       double* energy = nullptr) const {
        if(energy) {
            *energy = 0.0;
        }
        m = 0.0;
        mg[0] = mg[1] = mg[2] = 0.0;
        std::vector<double> P(3 * tri.size());
        std::vector<int> v2t(plane.size(), -1);
        int origin = -1;
        for(size_t t = 0; t < tri.size(); ++t) {
            if(tri[t].alive) {
                vertex(tri[t], &P[3 * t]);
                for(int l = 0; l < 3; ++l) {
                    v2t[tri[t].v[l]] = int(t);
                }
            }
        }
        std::vector<std::pair<uint32_t, double> > F;
        std::vector<uint16_t> loop;
        for(size_t v = 0; v < plane.size(); ++v) {
            if(v2t[v] < 0) {
                continue;
            }
            ring(uint16_t(v), uint16_t(v2t[v]), loop);
            double area = 0.0;
            const double* a = &P[3 * loop[0]];
            for(size_t i = 2; i < loop.size(); ++i) {
                const double* b = &P[3 * loop[i - 1]];
                const double* c = &P[3 * loop[i]];
                double U[3] = {b[0] - a[0], b[1] - a[1], b[2] - a[2]};
                double V[3] = {c[0] - a[0], c[1] - a[1], c[2] - a[2]};
                double N[3] = {U[1] * V[2] - U[2] * V[1], U[2] * V[0] - U[0] * V[2],
                               U[0] * V[1] - U[1] * V[0]};
                area += 0.5 * std::sqrt(N[0] * N[0] + N[1] * N[1] + N[2] * N[2]);
                if(origin >= 0) {
                    const double* o = &P[3 * size_t(origin)];
                    double W[3] = {o[0] - a[0], o[1] - a[1], o[2] - a[2]};
                    double vol =
                        std::fabs(N[0] * W[0] + N[1] * W[1] + N[2] * W[2]) / 6.0;
                    m += vol;
                    for(int k = 0; k < 3; ++k) {
                        mg[k] += vol * (o[k] + a[k] + b[k] + c[k]) / 4.0;
                    }
                    if(energy) {
                        const double* q4[4] = {o, a, b, c};
                        double e = 0.0;
                        for(int i = 0; i < 4; ++i) {
                            for(int j = i; j < 4; ++j) {
                                for(int k = 0; k < 3; ++k) {
                                    e += (q4[i][k] - seed[k]) * (q4[j][k] - seed[k]);
                                }
                            }
                        }
                        *energy += vol / 10.0 * e;
                    }
                    ++nI;
                }
            }
            if(origin < 0) {
                origin = v2t[v];
            }
            F.push_back(std::make_pair(gid[v], area));
        }
        std::sort(F.begin(), F.end());
        for(size_t f = 0; f < F.size() && f < fcap; ++f) {
            if(ids) {
                ids[f] = F[f].first;
            }
            if(areas) {
                areas[f] = F[f].second;
            }
        }
        return uint32_t(F.size());
    }
};

/* delaunay/periodic.cpp:44-86 */
const int kTranslation[27][3] = {
    {0, 0, 0},   {-1, -1, -1}, {-1, -1, 0}, {-1, -1, 1}, {-1, 0, -1}, {-1, 0, 0},
    {-1, 0, 1},  {-1, 1, -1},  {-1, 1, 0},  {-1, 1, 1},  {0, -1, -1}, {0, -1, 0},
    {0, -1, 1},  {0, 0, -1},   {0, 0, 1},   {0, 1, -1},  {0, 1, 0},   {0, 1, 1},
    {1, -1, -1}, {1, -1, 0},   {1, -1, 1},  {1, 0, -1},  {1, 0, 0},   {1, 0, 1},
    {1, 1, -1},  {1, 1, 0},    {1, 1, 1}};

int n_threads() {
#ifdef _OPENMP
    return g_threads > 0 ? g_threads : omp_get_max_threads();
#else
    return 1;
#endif
}

}  // namespace

extern "C" {

/* convex domain (b200vor_set_domain_planes): cells are clipped by these planes before any neighbour
 * (ConvexCell::clip_by_plane, convex_cell.cpp:467-516); a face on plane k gets the id VO_DOMAIN_ID + k */
static std::vector<double> g_domain;
void vo_set_domain_planes(uint32_t n, const double* planes) {
    g_domain.assign(planes, planes + 4 * size_t(n));
}
void vo_set_num_threads(int nthreads) { g_threads = nthreads; }
void vo_set_energy_output(double* energy) { g_energy = energy; }
int vo_num_threads(void) { return n_threads(); }

int vo_knn(uint32_t n, const double* xyz, uint32_t k, uint32_t* nbr, double* d2) {
    return vo_knn_points(n, xyz, n, xyz, k, nbr, d2);
}

int vo_knn_points(uint32_t n, const double* xyz, uint32_t nq, const double* q,
                  uint32_t k, uint32_t* nbr, double* d2) {
    if(k > n) {
        return 1;
    }
    KdIndex T;
    T.build(n, xyz);
#pragma omp parallel for schedule(dynamic, 256) num_threads(n_threads())
    for(int64_t i = 0; i < int64_t(nq); ++i) {
        T.query(q + 3 * size_t(i), k, nbr + size_t(i) * k, d2 + size_t(i) * k);
    }
    return 0;
}

int vo_cells(uint32_t n, const double* xyz, const double* w, int periodic,
             double period, const double* box, uint32_t k0, uint32_t kcap,
             double* m, double* mg, uint32_t* nfaces, uint32_t* face_nbr,
             double* face_area, uint32_t fcap, uint32_t* kstop, uint32_t* flags,
             double* stats) {
    /* periodic: search structure over the 27 translated copies; index of a copy
     * = real + n*instance (periodic.h:95-99). */
    std::vector<double> all;
    const double* pts = xyz;
    uint32_t ntot = n;
    if(periodic) {
        all.resize(size_t(27) * n * 3);
        for(int inst = 0; inst < 27; ++inst) {
            for(uint32_t i = 0; i < n; ++i) {
                for(int c = 0; c < 3; ++c) {
                    all[(size_t(inst) * n + i) * 3 + c] =
                        xyz[3 * size_t(i) + c] + period * kTranslation[inst][c];
                }
            }
        }
        pts = all.data();
        ntot = 27 * n;
    }
    if(ntot < 2) {
        return 1;
    }
    KdIndex T;
    T.build(ntot, pts);
    double wmax = 0.0;
    if(w) {
        wmax = w[0];
        for(uint32_t i = 1; i < n; ++i) {
            wmax = std::max(wmax, w[i]);
        }
    }
    kcap = std::min(kcap, ntot - 1);
    double sP = 0, sC = 0, sV = 0, sI = 0, sU = 0, sR = 0;
#pragma omp parallel num_threads(n_threads())
    {
        DualCell C;
        std::vector<uint32_t> nb;
        std::vector<double> d2;
        double lP = 0, lC = 0, lV = 0, lI = 0, lU = 0, lR = 0;
#pragma omp for schedule(dynamic, 64)
        for(int64_t ii = 0; ii < int64_t(n); ++ii) {
            uint32_t i = uint32_t(ii);
            const double* pi = xyz + 3 * size_t(i);
            double wi = w ? w[i] : 0.0;
            double pi2 = pi[0] * pi[0] + pi[1] * pi[1] + pi[2] * pi[2];
            double bx[6];
            if(periodic) {
                /* helper box = one period centred on the seed; its planes must
                 * all be clipped away for the cell to be valid */
                for(int c = 0; c < 3; ++c) {
                    bx[c] = pi[c] - 0.5 * period;
                    bx[3 + c] = pi[c] + 0.5 * period;
                }
            } else {
                std::memcpy(bx, box, sizeof(bx));
            }
            C.init_box(bx);
            if(!periodic) {
                for(size_t kd = 0; kd < g_domain.size() / 4; ++kd) {
                    Plane4 D = {g_domain[4 * kd], g_domain[4 * kd + 1], g_domain[4 * kd + 2], g_domain[4 * kd + 3]};
                    C.clip(D, 0xFFFFFF00u + uint32_t(kd));
                }
            }
            C.nP = C.nC = C.nV = 0;
            uint32_t k = std::min(std::max(k0, 1u), kcap), jj = 1;
            bool certified = false;
            uint32_t fl = 0;
            for(;;) {
                nb.resize(k + 1);
                d2.resize(k + 1);
                T.query(pi, k + 1, nb.data(), d2.data());
                lR += 1;
                for(; jj <= k; ++jj) {
                    /* generic_RVD.h:2303-2322, weighted variant in DESIGN.md */
                    double R2 = C.squared_radius(pi);
                    double bound = 4.1 * R2;
                    if(w) {
                        double s = std::sqrt(R2) + std::sqrt(R2 + (wmax - wi));
                        bound = 1.025 * s * s;
                    }
                    if(d2[jj] > bound) {
                        certified = true;
                        break;
                    }
                    uint32_t j = nb[jj];
                    if(d2[jj] == 0.0 && !w) {
                        continue;   /* duplicate point: no bisector (delaunay_nn.cpp:122-135) */
                    }
                    const double* pj = pts + 3 * size_t(j);
                    double wj = w ? w[j % n] : 0.0;
                    double pj2 = pj[0] * pj[0] + pj[1] * pj[1] + pj[2] * pj[2];
                    Plane4 E = {2.0 * (pi[0] - pj[0]), 2.0 * (pi[1] - pj[1]),
                                2.0 * (pi[2] - pj[2]), (wi - pi2) - (wj - pj2)};
                    C.clip(E, j);
                }
                if(certified || k >= kcap) {
                    break;
                }
                k = (k > 8) ? k + k / 8 : k + 1;   /* generic_RVD.h:2334-2338 */
                k = std::min(k, kcap);
            }
            if(!certified && kcap + 1 >= ntot) {
                certified = true;
            }
            if(!certified) {
                fl |= VO_FLAG_UNCERTIFIED;
                lU += 1;
            }
            if(C.tie) {
                fl |= VO_FLAG_TIE;
            }
            if(C.live_count() == 0) {
                fl |= VO_FLAG_EMPTY;
            }
            uint32_t ids[256];
            double areas[256];
            uint64_t nI = 0;
            double mm, g[3];
            double en = 0.0;
            uint32_t nf = C.integrate(mm, g, 256, ids, areas, nI, pi, g_energy ? &en : nullptr);
            if(g_energy) {
                g_energy[i] = (fl & VO_FLAG_EMPTY) ? 0.0 : en;
            }
            for(uint32_t f = 0; f < std::min(nf, 256u); ++f) {
                if(periodic && ids[f] >= VO_BOX_ID) {
                    fl |= VO_FLAG_BOXFACE;
                }
            }
            m[i] = mm;
            mg[3 * size_t(i)] = g[0];
            mg[3 * size_t(i) + 1] = g[1];
            mg[3 * size_t(i) + 2] = g[2];
            if(nfaces) {
                nfaces[i] = nf;
            }
            for(uint32_t f = 0; f < std::min(nf, fcap); ++f) {
                if(face_nbr) {
                    face_nbr[size_t(i) * fcap + f] = ids[f];
                }
                if(face_area) {
                    face_area[size_t(i) * fcap + f] = areas[f];
                }
            }
            if(kstop) {
                kstop[i] = certified ? jj : VO_UNCERTIFIED;
            }
            if(flags) {
                flags[i] = fl;
            }
            lP += double(C.nP);
            lC += double(C.nC);
            lV += double(C.nV);
            lI += double(nI);
        }
#pragma omp critical
        {
            sP += lP; sC += lC; sV += lV; sI += lI; sU += lU; sR += lR;
        }
    }
    if(stats) {
        stats[0] = sP; stats[1] = sC; stats[2] = sV;
        stats[3] = sI; stats[4] = sU; stats[5] = sR;
    }
    return 0;
}

int vo_lloyd_update(uint32_t n, const double* xyz, const double* m, const double* mg,
                    int periodic, double period, double* xyz_out) {
    for(uint32_t i = 0; i < n; ++i) {
        for(int c = 0; c < 3; ++c) {
            double x = xyz[3 * size_t(i) + c];
            if(m[i] > 1e-30) {                       /* CVT.cpp:154-163 */
                x = (1.0 / m[i]) * mg[3 * size_t(i) + c];   /* s = 1/m; p = s*mg, CVT.cpp:156-159 */
            }
            if(periodic) {                           /* demo_Delaunay3d/main.cpp:443-451 */
                if(x < 0.0) {
                    x += period;
                }
                if(x > period) {
                    x -= period;
                }
            }
            xyz_out[3 * size_t(i) + c] = x;
        }
    }
    return 0;
}

}  // extern "C"
