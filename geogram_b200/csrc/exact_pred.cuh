// exact_pred.cuh -- exact sign of the plane / cell-vertex test by floating-point expansions.
//
// The reference decides triangle_is_in_conflict exactly by default (use_exact_predicates_,
// voronoi/convex_cell.cpp:137, :905-911: PCK::det_4d(p1, p2, p3, eqn) >= 0, filter
// numerics/predicates/det4d.h:4-140, exact branch numerics/predicates.cpp:2073-2110 on
// numerics/multi_precision.*).  Here the same decision is taken when the fp64 filter of the cell
// kernels cannot certify a sign: the value of plane P at the vertex v = P0 ^ P1 ^ P2 is
//     s = N / W,   W = det3[n0, n1, n2],
//     N = d det3[n0, n1, n2] - d0 det3[n, n1, n2] - d1 det3[n, n2, n0] - d2 det3[n, n0, n1]
// (the 4x4 determinant of the four plane equations developed along its last column), and both
// N and W are evaluated EXACTLY as non-overlapping expansions (Shewchuk 1997: two-sum,
// two-product by FMA, grow / scale with zero elimination), so sign(s) = sign(N) sign(W) is the
// sign of the exact rational value of the fp64 inputs.  __host__ __device__: the host emulation
// of the thread-per-cell kernel runs the same code.
#pragma once

#include <math.h>
#include <stdint.h>

#include "grid_types.h"

namespace b200vor {
namespace exact {

// Individually rounded operations.  On the device the compiler must NOT contract a product with
// the sum it feeds (fma(a, b, c) is a different number than round(round(a b) + c), and the error
// terms below are only exact for the latter): the _rn intrinsics are never contracted.
B200_HD inline double add_rn(double a, double b) {
#if defined(__CUDA_ARCH__)
    return __dadd_rn(a, b);
#else
    return a + b;
#endif
}
B200_HD inline double sub_rn(double a, double b) {
#if defined(__CUDA_ARCH__)
    return __dsub_rn(a, b);
#else
    return a - b;
#endif
}
B200_HD inline double mul_rn(double a, double b) {
#if defined(__CUDA_ARCH__)
    return __dmul_rn(a, b);
#else
    return a * b;
#endif
}

B200_HD inline void two_sum(double a, double b, double& x, double& y) {
    x = add_rn(a, b);
    const double bv = sub_rn(x, a);
    const double av = sub_rn(x, bv);
    y = add_rn(sub_rn(a, av), sub_rn(b, bv));
}

B200_HD inline void two_product(double a, double b, double& x, double& y) {
    x = mul_rn(a, b);
    y = fma(a, b, -x);
}

// h = b * e, zero components dropped; returns the length (<= 2 elen)
B200_HD inline int scale_expansion(int elen, const double* e, double b, double* h) {
    if(elen == 0) {
        return 0;
    }
    int hlen = 0;
    double Q, hh;
    two_product(e[0], b, Q, hh);
    if(hh != 0.0) {
        h[hlen++] = hh;
    }
    for(int i = 1; i < elen; ++i) {
        double p1, p0, sum;
        two_product(e[i], b, p1, p0);
        two_sum(Q, p0, sum, hh);
        if(hh != 0.0) {
            h[hlen++] = hh;
        }
        // fast two-sum (|p1| >= |sum| holds for non-overlapping e)
        two_sum(p1, sum, Q, hh);
        if(hh != 0.0) {
            h[hlen++] = hh;
        }
    }
    if(Q != 0.0) {
        h[hlen++] = Q;
    }
    return hlen;
}

// h <- h + b in place (Shewchuk's grow-expansion with zero elimination): exact for any
// non-overlapping h sorted by increasing magnitude; returns the new length (<= hlen + 1)
B200_HD inline int grow_inplace(int hlen, double* h, double b) {
    double Q = b;
    int o = 0;
    for(int i = 0; i < hlen; ++i) {
        double s, err;
        two_sum(Q, h[i], s, err);
        if(err != 0.0) {
            h[o++] = err;   // o <= i: the write never overtakes the read
        }
        Q = s;
    }
    if(Q != 0.0) {
        h[o++] = Q;
    }
    return o;
}

// Scratch of one exact sign evaluation.  ONE object owned by the caller: with separate local
// arrays in the (inlined) helpers, nvcc 12.9 overlaid `term` of add_triple_product on the live
// accumulator of plane_vertex_sign (stack colouring of the merged allocas; the device returned wrong
// signs while every helper called on its own was right) -- see DESIGN.md section 9.
struct Scratch {
    double W[24], D[24], sc[48], acc[200];
    double pe[2], term[4];
};

// acc <- acc + a b c exactly (one of the six terms of a 3x3 determinant); returns the new length
B200_HD inline int add_triple_product(Scratch& S, int alen, double* acc, double a, double b, double c) {
    double hi, lo;
    two_product(a, b, hi, lo);
    int plen = 0;
    if(lo != 0.0) {
        S.pe[plen++] = lo;
    }
    if(hi != 0.0) {
        S.pe[plen++] = hi;
    }
    const int tlen = scale_expansion(plen, S.pe, c, S.term);
    for(int i = 0; i < tlen; ++i) {
        alen = grow_inplace(alen, acc, S.term[i]);
    }
    return alen;
}

// exact det3 of three fp64 rows as an expansion (<= 24 components), written out term by term
B200_HD inline int det3_expansion(Scratch& S, const double* r0, const double* r1, const double* r2, double* acc /*[24]*/) {
    int alen = 0;
    alen = add_triple_product(S, alen, acc, r0[0], r1[1], r2[2]);
    alen = add_triple_product(S, alen, acc, r0[1], r1[2], r2[0]);
    alen = add_triple_product(S, alen, acc, r0[2], r1[0], r2[1]);
    alen = add_triple_product(S, alen, acc, r0[0], r1[2], -r2[1]);
    alen = add_triple_product(S, alen, acc, r0[1], r1[0], -r2[2]);
    alen = add_triple_product(S, alen, acc, r0[2], r1[1], -r2[0]);
    return alen;
}

B200_HD inline int expansion_sign(int len, const double* e) {
    // the last non-zero component carries the sign (components are non-overlapping, increasing)
    for(int i = len - 1; i >= 0; --i) {
        if(e[i] > 0.0) {
            return 1;
        }
        if(e[i] < 0.0) {
            return -1;
        }
    }
    return 0;
}

// S.acc <- S.acc - d det3[r0, r1, r2] exactly; returns the new length of S.acc
B200_HD inline int sub_scaled_det3(Scratch& S, int alen, const double* r0, const double* r1, const double* r2, double d) {
    const int dlen = det3_expansion(S, r0, r1, r2, S.D);
    const int slen = scale_expansion(dlen, S.D, -d, S.sc);
    for(int i = 0; i < slen; ++i) {
        alen = grow_inplace(alen, S.acc, S.sc[i]);
    }
    return alen;
}

// Exact sign of plane (a, b, c, d) at the intersection of the planes P0, P1, P2 (each 4 doubles:
// normal, offset; the planes must meet in one point).  Returns -1, 0, +1; 2 if the three planes do
// not meet in a point (W == 0 exactly).
#if defined(__CUDACC__)
#define B200_EXACT_NOINLINE __noinline__
#else
#define B200_EXACT_NOINLINE
#endif
// (never inlined on the device: its scratch must not be coloured together with the caller's locals)
B200_EXACT_NOINLINE B200_HD static int plane_vertex_sign(const double* P0, const double* P1, const double* P2, const double* P) {
    Scratch S;
    const int wlen = det3_expansion(S, P0, P1, P2, S.W);
    const int sw = expansion_sign(wlen, S.W);
    if(sw == 0) {
        return 2;
    }
    // N = d W - d0 det3[n, n1, n2] - d1 det3[n, n2, n0] - d2 det3[n, n0, n1]
    int alen = scale_expansion(wlen, S.W, P[3], S.acc);
    alen = sub_scaled_det3(S, alen, P, P1, P2, P0[3]);
    alen = sub_scaled_det3(S, alen, P, P2, P0, P1[3]);
    alen = sub_scaled_det3(S, alen, P, P0, P1, P2[3]);
    return expansion_sign(alen, S.acc) * sw;
}

}  // namespace exact
}  // namespace b200vor
