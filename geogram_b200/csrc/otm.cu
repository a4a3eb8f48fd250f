// otm.cu -- semi-discrete optimal transport on the device: damped Newton on the Laguerre weights.
//
// SURVEY.md 8(f) rank 1 / 8(a) a20.  The reference's solver lives in BrunoLevy/exploragram
// (OptimalTransportMap3d; NOT under /root/reference, unpinned -- call sites
// src/examples/exploragram/compute_OTM/main.cpp:284-297), so this file restates the PUBLISHED
// algorithm and is parity-unpinned against exploragram itself:
//   * B. Levy, "A numerical algorithm for L2 semi-discrete optimal transport in 3D" (M2AN 2015):
//     objective, gradient g_i = vol(Lag_i) - nu_i, Hessian H_ij = A_ij / (2 |p_i - p_j|);
//   * Kitagawa, Merigot, Thibert, "Convergence of a Newton algorithm for semi-discrete optimal
//     transport" (JEMS 2019), Algorithm 1: damped Newton, step halved until every cell keeps at
//     least eps0 = 1/2 min(min_i vol_i(w0), min_i nu_i) and the error decreases by (1 - alpha/2).
// What IS pinned against the reference are the inputs of every Newton step: Laguerre cell
// volumes and facet areas A_ij (tests/test_gpu_parity.py against the exact Delaunay oracle).
//
// All vectors stay in HBM.  Per Newton iteration: one cell pass with face lists (fused.cu), one
// coefficient kernel, a Jacobi-preconditioned CG whose SpMV reads the face lists as an ELL
// matrix (16 lanes per row, coalesced 64-byte segments), and one cell pass per damping trial.
// Every reduction is a fixed-partition block tree + ordered host sum: bitwise reproducible.
#include "common.cuh"

#include <math.h>
#include <string.h>

#include <algorithm>
#include <vector>

namespace b200vor {

int build_grid(Ctx* c);
int otm_cells(Ctx* c);   // capi.cu: cell pass with face lists on the current weights

namespace {

constexpr int kNB = 256;       // reduction blocks
constexpr int kNT = 256;       // threads per block
constexpr uint32_t kFcap = 64;
static_assert(kNB == kNT, "the scalar kernels reduce the kNB partials with one block of kNT threads");

// translation (in periods) of periodic instance `inst`: inverse of translation_to_instance
__device__ inline void instance_translation(uint32_t inst, int& tx, int& ty, int& tz) {
    // reorder_instances: 0 -> 13, 1..13 -> 0..12, 14..26 unchanged (periodic.cpp:80-84)
    int i = inst == 0 ? 13 : (inst <= 13 ? int(inst) - 1 : int(inst));
    tz = i % 3 - 1;
    ty = (i / 3) % 3 - 1;
    tx = i / 9 - 1;
}

// c_if = A_if / (2 |p_i - p_j|) for every face f of every cell i, column j_f = real index of the
// neighbour (box faces: coefficient 0); diag_i = sum_f c_if.  16 lanes per row.
// Rows and columns are numbered in GRID order (row i = seed perm[i]; column = inv[neighbour]):
// neighbours are then close in memory and the x[col] gathers of the SpMV stay in L1/L2.
__global__ void otm_coef_kernel(uint32_t n, const double* __restrict__ xyz, int periodic, double period,
                                const uint32_t* __restrict__ perm, const uint32_t* __restrict__ inv,
                                const uint32_t* __restrict__ nfaces, const uint32_t* __restrict__ face_nbr,
                                const double* __restrict__ face_area, uint32_t* __restrict__ col,
                                double* __restrict__ coef, double* __restrict__ diag,
                                uint32_t* __restrict__ nfg) {
    const uint32_t lane = threadIdx.x & 15u;
    const uint32_t row = (blockIdx.x * blockDim.x + threadIdx.x) >> 4;
    const bool alive = row < n;
    const uint32_t r = alive ? perm[row] : perm[0];
    const uint32_t nf = alive ? min(nfaces[r], kFcap) : 0u;
    const double px = xyz[3 * size_t(r)], py = xyz[3 * size_t(r) + 1], pz = xyz[3 * size_t(r) + 2];
    double d = 0.0;
    for(uint32_t f = lane; f < kFcap; f += 16u) {
        uint32_t j = 0;
        double cf = 0.0;
        if(f < nf) {
            const uint32_t id = face_nbr[size_t(r) * kFcap + f];
            if(id < B200VOR_DOMAIN_ID) {   // (not a face on the clipping box or on a plane of the convex domain)
                j = id % n;
                const uint32_t inst = id / n;
                int tx = 0, ty = 0, tz = 0;
                if(periodic) {
                    instance_translation(inst, tx, ty, tz);
                }
                const double qx = xyz[3 * size_t(j)] + tx * period - px;
                const double qy = xyz[3 * size_t(j) + 1] + ty * period - py;
                const double qz = xyz[3 * size_t(j) + 2] + tz * period - pz;
                const double dist = sqrt(qx * qx + qy * qy + qz * qz);
                cf = dist > 0.0 ? face_area[size_t(r) * kFcap + f] / (2.0 * dist) : 0.0;
                j = inv[j];
            } else {
                j = row;
            }
        } else {
            j = row;
        }
        if(alive) {
            col[size_t(row) * kFcap + f] = j;
            coef[size_t(row) * kFcap + f] = cf;
        }
        d += cf;
    }
    for(int o = 8; o > 0; o >>= 1) {
        d += __shfl_xor_sync(0xffffffffu, d, o, 16);
    }
    if(alive && lane == 0) {
        diag[row] = d;
        nfg[row] = nf;
    }
}

__global__ void invert_perm_kernel(uint32_t n, const uint32_t* __restrict__ perm, uint32_t* __restrict__ inv) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if(i < n) {
        inv[perm[i]] = i;
    }
}

// out[i] = in[perm[i]] (user order -> grid order)
__global__ void gather_kernel(uint32_t n, const uint32_t* __restrict__ perm, const double* __restrict__ in,
                              double* __restrict__ out) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if(i < n) {
        out[i] = in[perm[i]];
    }
}

// out[perm[i]] = a[perm[i]] + s * d[i] (d in grid order, a and out in user order)
__global__ void axpy_scatter_kernel(uint32_t n, const uint32_t* __restrict__ perm, const double* __restrict__ a,
                                    double s, const double* __restrict__ d, double* __restrict__ out) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if(i < n) {
        const uint32_t r = perm[i];
        out[r] = a[r] + s * d[i];
    }
}

// y = L x, L_ii = sum_f c_if, L_ij = -c_if: y_i = sum_f c_if (x_i - x_jf).  16 lanes per row.
__global__ void otm_spmv_kernel(uint32_t n, const uint32_t* __restrict__ nfaces, const uint32_t* __restrict__ col,
                                const double* __restrict__ coef, const double* __restrict__ x,
                                double* __restrict__ y) {
    const uint32_t lane = threadIdx.x & 15u;
    const uint32_t row = (blockIdx.x * blockDim.x + threadIdx.x) >> 4;
    const bool alive = row < n;
    const uint32_t r = alive ? row : 0u;
    const uint32_t nf = alive ? min(nfaces[r], kFcap) : 0u;
    const double xi = x[r];
    double acc = 0.0;
    for(uint32_t f = lane; f < nf; f += 16u) {
        acc += coef[size_t(r) * kFcap + f] * (xi - x[col[size_t(r) * kFcap + f]]);
    }
    for(int o = 8; o > 0; o >>= 1) {
        acc += __shfl_xor_sync(0xffffffffu, acc, o, 16);
    }
    if(alive && lane == 0) {
        y[r] = acc;
    }
}

// fixed-partition block reductions: block b owns the contiguous slice [b*per, (b+1)*per)
__device__ inline double block_sum(double v) {
    __shared__ double sm[kNT];
    sm[threadIdx.x] = v;
    __syncthreads();
    for(uint32_t o = kNT / 2; o > 0; o >>= 1) {
        if(threadIdx.x < o) {
            sm[threadIdx.x] += sm[threadIdx.x + o];
        }
        __syncthreads();
    }
    double r = sm[0];
    __syncthreads();
    return r;
}

__global__ void dot_kernel(uint32_t n, const double* __restrict__ a, const double* __restrict__ b,
                           double* __restrict__ part) {
    const uint32_t per = (n + gridDim.x - 1) / gridDim.x;
    const uint32_t lo = blockIdx.x * per, hi = min(n, lo + per);
    double acc = 0.0;
    for(uint32_t i = lo + threadIdx.x; i < hi; i += kNT) {
        acc += a[i] * b[i];
    }
    acc = block_sum(acc);
    if(threadIdx.x == 0) {
        part[blockIdx.x] = acc;
    }
}

// The CG scalars live on the device (no host round trip per iteration):
// sc[0] = r.z, sc[1] = alpha, sc[2] = beta, sc[3] = r.r, sc[4] = stop threshold on r.r,
// sc[5] = 1 once converged (or the direction fell into the null space), sc[6] = iterations applied
enum { kRZ = 0, kAlpha, kBeta, kRR, kStop, kDone, kIters, kNScalars };

// x += alpha p ; r -= alpha q ; z = r / diag ; partials of r.z and r.r
__global__ void cg_update_kernel(uint32_t n, const double* __restrict__ sc, const double* __restrict__ p,
                                 const double* __restrict__ q, const double* __restrict__ diag, double* __restrict__ x,
                                 double* __restrict__ r, double* __restrict__ z, double* __restrict__ part) {
    const double alpha = sc[kAlpha];
    const uint32_t per = (n + gridDim.x - 1) / gridDim.x;
    const uint32_t lo = blockIdx.x * per, hi = min(n, lo + per);
    double rz = 0.0, rr = 0.0;
    for(uint32_t i = lo + threadIdx.x; i < hi; i += kNT) {
        x[i] += alpha * p[i];
        const double ri = r[i] - alpha * q[i];
        r[i] = ri;
        const double di = diag[i];
        const double zi = di > 0.0 ? ri / di : 0.0;
        z[i] = zi;
        rz += ri * zi;
        rr += ri * ri;
    }
    rz = block_sum(rz);
    rr = block_sum(rr);
    if(threadIdx.x == 0) {
        part[blockIdx.x] = rz;
        part[gridDim.x + blockIdx.x] = rr;
    }
}

// alpha = r.z / p.q from the fixed-tree sum of the partials of p.q (one block of kNB threads)
__global__ void cg_alpha_kernel(const double* __restrict__ part, double* __restrict__ sc) {
    const double pq = block_sum(part[threadIdx.x]);
    if(threadIdx.x == 0) {
        const bool stop = sc[kDone] != 0.0 || !(sc[kRR] > sc[kStop]) || !(pq > 0.0);
        sc[kDone] = stop ? 1.0 : 0.0;
        sc[kAlpha] = stop ? 0.0 : sc[kRZ] / pq;
    }
}

// beta = (r.z)_new / (r.z)_old from the fixed-tree sums of the partials of the update kernel
__global__ void cg_beta_kernel(const double* __restrict__ part, double* __restrict__ sc) {
    const double rz_new = block_sum(part[threadIdx.x]);
    const double rr = block_sum(part[kNB + threadIdx.x]);
    if(threadIdx.x == 0 && sc[kDone] == 0.0) {
        sc[kBeta] = rz_new / sc[kRZ];
        sc[kRZ] = rz_new;
        sc[kRR] = rr;
        sc[kIters] += 1.0;
    }
}

// p = z + beta p
__global__ void cg_dir_kernel(uint32_t n, const double* __restrict__ sc, const double* __restrict__ z,
                              double* __restrict__ p) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if(i < n && sc[kDone] == 0.0) {
        p[i] = z[i] + sc[kBeta] * p[i];
    }
}

// b = nu - vol (nu == NULL: uniform total / n) ; partials: sum (vol-nu)^2, max |vol-nu|/nu, min vol, sum vol
__global__ void otm_residual_kernel(uint32_t n, const double* __restrict__ vol, const double* __restrict__ nu,
                                    double nu_uniform, double* __restrict__ b, double* __restrict__ part) {
    __shared__ double smx[kNT], smn[kNT];
    const uint32_t per = (n + gridDim.x - 1) / gridDim.x;
    const uint32_t lo = blockIdx.x * per, hi = min(n, lo + per);
    double s2 = 0.0, sv = 0.0, mx = 0.0, mn = 1e300;
    for(uint32_t i = lo + threadIdx.x; i < hi; i += kNT) {
        const double v = vol[i], t = nu ? nu[i] : nu_uniform;
        const double g = t - v;
        if(b) {
            b[i] = g;
        }
        s2 += g * g;
        sv += v;
        mx = fmax(mx, fabs(g) / t);
        mn = fmin(mn, v);
    }
    s2 = block_sum(s2);
    sv = block_sum(sv);
    smx[threadIdx.x] = mx;
    smn[threadIdx.x] = mn;
    __syncthreads();
    for(uint32_t o = kNT / 2; o > 0; o >>= 1) {
        if(threadIdx.x < o) {
            smx[threadIdx.x] = fmax(smx[threadIdx.x], smx[threadIdx.x + o]);
            smn[threadIdx.x] = fmin(smn[threadIdx.x], smn[threadIdx.x + o]);
        }
        __syncthreads();
    }
    if(threadIdx.x == 0) {
        part[blockIdx.x] = s2;
        part[gridDim.x + blockIdx.x] = sv;
        part[2 * gridDim.x + blockIdx.x] = smx[0];
        part[3 * gridDim.x + blockIdx.x] = smn[0];
    }
}

__global__ void jacobi_kernel(uint32_t n, const double* __restrict__ r, const double* __restrict__ diag,
                              double* __restrict__ z, double* __restrict__ p) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if(i < n) {
        const double di = diag[i];
        const double zi = di > 0.0 ? r[i] / di : 0.0;
        z[i] = zi;
        p[i] = zi;
    }
}

struct Resid {
    double norm2, sumvol, maxrel, minvol;
    uint32_t bad = 0;   // cells that are not OK / EMPTY, or rows wider than the ELL matrix
};

struct Work {
    DevBuf<double> nu, b, x, r, z, p, q, diag, coef, w0, w1, part;
    DevBuf<uint32_t> col, inv, nfg;
    DevBuf<double> bg;
    ~Work() { release(); }
    void release() {
        nu.release(); b.release(); x.release(); r.release(); z.release(); p.release(); q.release();
        diag.release(); coef.release(); w0.release(); w1.release(); part.release(); col.release();
        inv.release(); nfg.release(); bg.release();
    }
};

int host_sum(Ctx* c, const double* part, int count, double* out) {
    std::vector<double> h(count);
    B200_CUDA_CHECK(cudaMemcpyAsync(h.data(), part, sizeof(double) * count, cudaMemcpyDeviceToHost, c->stream));
    B200_CUDA_CHECK(cudaStreamSynchronize(c->stream));
    memcpy(out, h.data(), sizeof(double) * count);
    return B200VOR_OK;
}

// cells whose measure or Hessian row cannot be trusted: anything but OK / EMPTY, or more faces than
// the ELL rows hold (their m is "best effort", otm_coef_kernel would truncate the row)
__global__ void otm_check_kernel(uint32_t n, const uint8_t* __restrict__ status, const uint32_t* __restrict__ nfaces,
                                 uint32_t* __restrict__ bad) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if(i < n) {
        const uint32_t st = status[i] & ~uint32_t(B200VOR_CELL_EMPTY);
        if(st != 0u || nfaces[i] > kFcap) {
            atomicAdd(bad, 1u);
        }
    }
}

int residual(Ctx* c, Work& W, const double* nu, double nu_uniform, double* b, Resid* R) {
    int rc;
    if((rc = c->counter.reserve(4))) {
        return rc;
    }
    B200_CUDA_CHECK(cudaMemsetAsync(c->counter.p, 0, sizeof(uint32_t), c->stream));
    otm_check_kernel<<<(c->n + kNT - 1) / kNT, kNT, 0, c->stream>>>(c->n, c->status.p, c->nfaces.p, c->counter.p);
    otm_residual_kernel<<<kNB, kNT, 0, c->stream>>>(c->n, c->m.p, nu, nu_uniform, b, W.part.p);
    c->stats.gpu_launches += 2;
    uint32_t hbad = 0;
    B200_CUDA_CHECK(cudaMemcpyAsync(&hbad, c->counter.p, sizeof(uint32_t), cudaMemcpyDeviceToHost, c->stream));
    double h[4 * kNB];
    rc = host_sum(c, W.part.p, 4 * kNB, h);
    if(rc) {
        return rc;
    }
    R->bad = hbad;
    R->norm2 = 0.0; R->sumvol = 0.0; R->maxrel = 0.0; R->minvol = 1e300;
    for(int k = 0; k < kNB; ++k) {
        R->norm2 += h[k];
        R->sumvol += h[kNB + k];
        R->maxrel = fmax(R->maxrel, h[2 * kNB + k]);
        R->minvol = fmin(R->minvol, h[3 * kNB + k]);
    }
    R->norm2 = sqrt(R->norm2);
    return B200VOR_OK;
}

int dot(Ctx* c, Work& W, const double* a, const double* b, double* out) {
    dot_kernel<<<kNB, kNT, 0, c->stream>>>(c->n, a, b, W.part.p);
    c->stats.gpu_launches++;
    double h[kNB];
    int rc = host_sum(c, W.part.p, kNB, h);
    if(rc) {
        return rc;
    }
    double s = 0.0;
    for(int k = 0; k < kNB; ++k) {
        s += h[k];
    }
    *out = s;
    return B200VOR_OK;
}

// weights -> context, grid rebuilt (the weight travels with the seed in `sorted`), cells + faces
int eval_cells(Ctx* c, const double* w_dev, double wmax) {
    int rc;
    if((rc = c->wgt.reserve(c->n))) {
        return rc;
    }
    B200_CUDA_CHECK(cudaMemcpyAsync(c->wgt.p, w_dev, size_t(c->n) * sizeof(double), cudaMemcpyDeviceToDevice,
                                    c->stream));
    c->has_w = true;
    c->wmax = wmax;
    if((rc = build_grid(c))) {
        return rc;
    }
    c->grid_dirty = false;
    return otm_cells(c);
}

__global__ void max_part_kernel(uint32_t n, const double* __restrict__ w, double* __restrict__ part) {
    __shared__ double sm[kNT];
    const uint32_t per = (n + gridDim.x - 1) / gridDim.x;
    const uint32_t lo = blockIdx.x * per, hi = min(n, lo + per);
    double mx = -1e300;
    for(uint32_t i = lo + threadIdx.x; i < hi; i += kNT) {
        mx = fmax(mx, w[i]);
    }
    sm[threadIdx.x] = mx;
    __syncthreads();
    for(uint32_t o = kNT / 2; o > 0; o >>= 1) {
        if(threadIdx.x < o) {
            sm[threadIdx.x] = fmax(sm[threadIdx.x], sm[threadIdx.x + o]);
        }
        __syncthreads();
    }
    if(threadIdx.x == 0) {
        part[blockIdx.x] = sm[0];
    }
}

int vec_max(Ctx* c, Work& W, const double* w, double* out) {
    max_part_kernel<<<kNB, kNT, 0, c->stream>>>(c->n, w, W.part.p);
    c->stats.gpu_launches++;
    double h[kNB];
    int rc = host_sum(c, W.part.p, kNB, h);
    if(rc) {
        return rc;
    }
    double m = h[0];
    for(int k = 1; k < kNB; ++k) {
        m = fmax(m, h[k]);
    }
    *out = m;
    return B200VOR_OK;
}

}  // namespace

int otm_newton(Ctx* c, const double* nu_in, double* w_io, const b200vor_otm_params* prm, b200vor_otm_result* res,
               int on_device) {
    b200vor_otm_params P = {50, 1000, 0.0, 0.01, 0.0};
    // cg_tol = 0: INEXACT Newton -- the linear system of a step is solved to a tenth of the current
    // relative volume error (never looser than 1e-2, never tighter than epsilon / 100): the early steps,
    // far from the solution, do not pay for 8 digits they cannot use (Eisenstat & Walker 1996)
    bool adaptive_tol = true;
    if(prm) {
        P = *prm;
        if(P.max_newton == 0) P.max_newton = 50;
        if(P.max_cg == 0) P.max_cg = 1000;
        adaptive_tol = !(P.cg_tol > 0.0);
        if(!(P.epsilon > 0.0)) P.epsilon = 0.01;
        if(!(P.total_mass > 0.0)) P.total_mass = 0.0;
    }
    const uint32_t n = c->n;
    const uint32_t nblk = (n + kNT - 1) / kNT;
    const uint32_t rblk = uint32_t((size_t(n) * 16 + kNT - 1) / kNT);
    Work W;
    int rc = 0;
    b200vor_otm_result R;
    memset(&R, 0, sizeof(R));
    auto fail = [&](int code) { return code; };   // W releases its buffers when it goes out of scope
    if((rc = W.b.reserve(n)) || (rc = W.x.reserve(n)) || (rc = W.r.reserve(n)) || (rc = W.z.reserve(n)) ||
       (rc = W.p.reserve(n)) || (rc = W.q.reserve(n)) || (rc = W.diag.reserve(n)) || (rc = W.w0.reserve(n)) ||
       (rc = W.w1.reserve(n)) || (rc = W.part.reserve(4 * kNB + 16)) || (rc = W.coef.reserve(size_t(n) * kFcap)) ||
       (rc = W.col.reserve(size_t(n) * kFcap)) || (rc = W.inv.reserve(n)) || (rc = W.nfg.reserve(n)) ||
       (rc = W.bg.reserve(n))) {
        return fail(rc);
    }
    cudaStream_t st = c->stream;
    const double* nu = nullptr;
    if(nu_in) {
        if(on_device) {
            nu = nu_in;
        } else {
            if((rc = W.nu.reserve(n))) {
                return fail(rc);
            }
            if(cudaMemcpyAsync(W.nu.p, nu_in, size_t(n) * sizeof(double), cudaMemcpyHostToDevice, st) != cudaSuccess) {
                set_error("b200vor_otm_newton: copy of nu failed");
                return fail(B200VOR_ERR_CUDA);
            }
            nu = W.nu.p;
        }
    }
    if(cudaMemcpyAsync(W.w0.p, w_io, size_t(n) * sizeof(double),
                       on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, st) != cudaSuccess) {
        set_error("b200vor_otm_newton: copy of the weights failed");
        return fail(B200VOR_ERR_CUDA);
    }
    if(!nu_in && c->ndom > 0 && !(P.total_mass > 0.0)) {
        set_error("b200vor_otm_newton: on a convex domain (b200vor_set_domain_planes) give nu or total_mass (= the measure of the domain)");
        return fail(B200VOR_ERR_ARG);
    }
    if(!c->piece_off.empty()) {
        set_error("b200vor_otm_newton: not available on a domain made of several pieces (no facet areas there)");
        return fail(B200VOR_ERR_ARG);
    }
    // measure of the domain (uniform target = |domain| / n)
    double domain = c->periodic ? c->period * c->period * c->period
                                : (c->box[3] - c->box[0]) * (c->box[4] - c->box[1]) * (c->box[5] - c->box[2]);
    const double nu_uniform = (P.total_mass > 0.0 ? P.total_mass : domain) / double(n);
    double wmax = 0.0;
    if((rc = vec_max(c, W, W.w0.p, &wmax)) || (rc = eval_cells(c, W.w0.p, wmax))) {
        return fail(rc);
    }
    R.cell_evals = 1;
    Resid cur;
    if((rc = residual(c, W, nu, nu_uniform, W.b.p, &cur))) {
        return fail(rc);
    }
    if(cur.bad) {
        // never iterate on volumes that are only best effort (flagged cells) or on truncated rows
        set_error("b200vor_otm_newton: %u cell(s) of the start weights are flagged (not OK / EMPTY) or have more "
                  "than %u faces", cur.bad, kFcap);
        return fail(B200VOR_ERR_CAPACITY);
    }
    // smallest target measure
    double numin = nu_uniform;
    if(nu) {
        // min nu = -max(-nu): reuse the residual kernel with vol := nu (min of "vol")
        otm_residual_kernel<<<kNB, kNT, 0, st>>>(n, nu, nullptr, 1.0, nullptr, W.part.p);
        c->stats.gpu_launches++;
        double h[4 * kNB];
        if((rc = host_sum(c, W.part.p, 4 * kNB, h))) {
            return fail(rc);
        }
        numin = 1e300;
        for(int k = 0; k < kNB; ++k) {
            numin = fmin(numin, h[3 * kNB + k]);
        }
    }
    // KMT: every cell keeps at least eps0
    const double eps0 = 0.5 * fmin(cur.minvol, numin);
    R.initial_error = cur.norm2;
    if(!(eps0 > 0.0)) {
        set_error("b200vor_otm_newton: an initial cell is empty (min volume %g): the damped Newton method needs "
                  "weights whose Laguerre cells are all non-empty (KMT 2019, hypothesis of Algorithm 1)",
                  cur.minvol);
        return fail(B200VOR_ERR_ARG);
    }
    uint32_t it = 0;
    for(; it < P.max_newton && cur.maxrel > P.epsilon; ++it) {
        // ---- Hessian of this iterate from the face lists ----
        // (the grid order depends on the points only: identical for every weight vector)
        invert_perm_kernel<<<nblk, kNT, 0, st>>>(n, c->perm.p, W.inv.p);
        otm_coef_kernel<<<rblk, kNT, 0, st>>>(n, c->xyz.p, c->periodic, c->period, c->perm.p, W.inv.p, c->nfaces.p,
                                              c->face_nbr.p, c->face_area.p, W.col.p, W.coef.p, W.diag.p, W.nfg.p);
        c->stats.gpu_launches += 2;
        // ---- L delta = nu - vol by Jacobi-preconditioned CG in grid order, delta_0 = 0 ----
        B200_CUDA_CHECK(cudaMemsetAsync(W.x.p, 0, size_t(n) * sizeof(double), st));
        gather_kernel<<<nblk, kNT, 0, st>>>(n, c->perm.p, W.b.p, W.r.p);
        c->stats.gpu_launches++;
        jacobi_kernel<<<nblk, kNT, 0, st>>>(n, W.r.p, W.diag.p, W.z.p, W.p.p);
        c->stats.gpu_launches++;
        double rz = 0.0, bb = 0.0;
        if((rc = dot(c, W, W.r.p, W.z.p, &rz)) || (rc = dot(c, W, W.r.p, W.r.p, &bb))) {
            return fail(rc);
        }
        // device-resident scalars; the host looks at them every 16 iterations only
        const double tol = adaptive_tol ? fmin(1e-2, fmax(0.1 * fmin(cur.maxrel, 1.0), 0.01 * P.epsilon)) : P.cg_tol;
        double hsc[kNScalars] = {rz, 0.0, 0.0, bb, tol * tol * bb, 0.0, 0.0};
        double* sc = W.part.p + 4 * kNB;
        B200_CUDA_CHECK(cudaMemcpyAsync(sc, hsc, sizeof(hsc), cudaMemcpyHostToDevice, st));
        uint32_t k = 0;
        for(uint32_t launched = 0; launched < P.max_cg;) {
            const uint32_t burst = std::min<uint32_t>(16u, P.max_cg - launched);
            for(uint32_t t = 0; t < burst; ++t) {
                otm_spmv_kernel<<<rblk, kNT, 0, st>>>(n, W.nfg.p, W.col.p, W.coef.p, W.p.p, W.q.p);
                dot_kernel<<<kNB, kNT, 0, st>>>(n, W.p.p, W.q.p, W.part.p);
                cg_alpha_kernel<<<1, kNB, 0, st>>>(W.part.p, sc);
                cg_update_kernel<<<kNB, kNT, 0, st>>>(n, sc, W.p.p, W.q.p, W.diag.p, W.x.p, W.r.p, W.z.p, W.part.p);
                cg_beta_kernel<<<1, kNB, 0, st>>>(W.part.p, sc);
                cg_dir_kernel<<<nblk, kNT, 0, st>>>(n, sc, W.z.p, W.p.p);
            }
            c->stats.gpu_launches += 6ull * burst;
            launched += burst;
            B200_CUDA_CHECK(cudaMemcpyAsync(hsc, sc, sizeof(hsc), cudaMemcpyDeviceToHost, st));
            B200_CUDA_CHECK(cudaStreamSynchronize(st));
            k = uint32_t(hsc[kIters]);
            if(hsc[kDone] != 0.0 || !(hsc[kRR] > hsc[kStop])) {
                break;
            }
        }
        R.cg_iters += k;
        // ---- KMT damping ----
        double alpha = 1.0;
        Resid trial = cur;
        bool accepted = false;
        for(int halv = 0; halv < 16; ++halv) {   // alpha >= 1.5e-5: below that the iterate sits on the rounding floor
            axpy_scatter_kernel<<<nblk, kNT, 0, st>>>(n, c->perm.p, W.w0.p, alpha, W.x.p, W.w1.p);
            c->stats.gpu_launches++;
            if((rc = vec_max(c, W, W.w1.p, &wmax)) || (rc = eval_cells(c, W.w1.p, wmax))) {
                return fail(rc);
            }
            R.cell_evals++;
            if((rc = residual(c, W, nu, nu_uniform, W.q.p, &trial))) {
                return fail(rc);
            }
            if(trial.bad == 0 && trial.minvol >= eps0 && trial.norm2 <= (1.0 - 0.5 * alpha) * cur.norm2) {
                accepted = true;
                break;
            }
            alpha *= 0.5;
        }
        if(!accepted) {
            // no admissible step: restore the cells of the last accepted weights and stop
            if((rc = vec_max(c, W, W.w0.p, &wmax)) || (rc = eval_cells(c, W.w0.p, wmax))) {
                return fail(rc);
            }
            R.cell_evals++;
            break;
        }
        std::swap(W.w0.p, W.w1.p);
        std::swap(W.w0.cap, W.w1.cap);
        std::swap(W.b.p, W.q.p);
        std::swap(W.b.cap, W.q.cap);
        cur = trial;
        R.last_alpha = alpha;
    }
    R.newton_iters = it;
    R.final_error = cur.norm2;
    R.max_rel_error = cur.maxrel;
    R.min_volume = cur.minvol;
    R.sum_volume = cur.sumvol;
    R.converged = cur.maxrel <= P.epsilon ? 1 : 0;
    if(cudaMemcpyAsync(w_io, W.w0.p, size_t(n) * sizeof(double),
                       on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, st) != cudaSuccess) {
        set_error("b200vor_otm_newton: copy of the result failed");
        return fail(B200VOR_ERR_CUDA);
    }
    B200_CUDA_CHECK(cudaStreamSynchronize(st));
    if(res) {
        *res = R;
    }
    return B200VOR_OK;
}

}  // namespace b200vor
