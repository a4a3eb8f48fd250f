// capi.cu -- the C ABI of include/b200vor.h: context, buffers, pass orchestration.
#include "common.cuh"

#include <math.h>
#include <stdarg.h>
#include <string.h>
#include <vector>

namespace b200vor {

static thread_local char g_error[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_error, sizeof(g_error), fmt, ap);
    va_end(ap);
}

namespace {

// which kernel takes the first pass of mode 0 (measured: DESIGN.md section 9)
constexpr bool kTpcFirstByDefault = false;

__global__ void max_kernel(const double* __restrict__ w, uint32_t n, double* out) {
    __shared__ double sm[8];
    double mx = -1e300;
    for(uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        mx = fmax(mx, w[i]);
    }
    for(int o = 16; o > 0; o >>= 1) {
        mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    }
    if((threadIdx.x & 31) == 0) {
        sm[threadIdx.x >> 5] = mx;
    }
    __syncthreads();
    if(threadIdx.x == 0) {
        for(int k = 1; k < 8; ++k) {
            mx = fmax(mx, sm[k]);
        }
        out[blockIdx.x] = mx;
    }
}

__global__ void gather3_kernel(const double* __restrict__ src, uint32_t stride, uint32_t n,
                               double* __restrict__ dst) {
    uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if(i < n) {
        dst[3 * size_t(i)] = src[size_t(i) * stride];
        dst[3 * size_t(i) + 1] = src[size_t(i) * stride + 1];
        dst[3 * size_t(i) + 2] = src[size_t(i) * stride + 2];
    }
}

// g_i = 2 (m_i p_i - mg_i): the per-seed sum of the gradient terms of
// ComputeCVTFuncGradVolumetric (voronoi/RVD.cpp:861-869)
__global__ void cvt_grad_kernel(uint32_t n, const double* __restrict__ xyz, const double* __restrict__ m,
                                const double* __restrict__ mg, double* __restrict__ g) {
    uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    if(k < 3 * n) {
        g[k] = 2.0 * (m[k / 3] * xyz[k] - mg[k]);
    }
}

// fixed-order sum: block b adds its contiguous slice with a fixed tree; the host adds the
// partials in block order
__global__ void sum_kernel(const double* __restrict__ v, uint32_t n, double* __restrict__ part) {
    __shared__ double sm[256];
    const uint32_t per = (n + gridDim.x - 1) / gridDim.x;
    const uint32_t b = blockIdx.x * per, e = min(n, b + per);
    double acc = 0.0;
    for(uint32_t i = b + threadIdx.x; i < e; i += blockDim.x) {
        acc += v[i];
    }
    sm[threadIdx.x] = acc;
    __syncthreads();
    for(uint32_t o = blockDim.x / 2; o > 0; o >>= 1) {
        if(threadIdx.x < o) {
            sm[threadIdx.x] += sm[threadIdx.x + o];
        }
        __syncthreads();
    }
    if(threadIdx.x == 0) {
        part[blockIdx.x] = sm[0];
    }
}

struct Timer {
    // device-side timing of a sequence of launches on the context's stream; the events come from a
    // pool owned by the context (created once, never destroyed between calls)
    std::vector<cudaEvent_t> ev;
    std::vector<int> kind;
    Ctx* c;
    explicit Timer(Ctx* ctx) : c(ctx) {}
    ~Timer() {
        for(cudaEvent_t e : ev) {
            c->event_pool.push_back(e);
        }
    }
    void mark(int k) {
        cudaEvent_t e;
        if(!c->event_pool.empty()) {
            e = c->event_pool.back();
            c->event_pool.pop_back();
        } else {
            cudaEventCreate(&e);
        }
        cudaEventRecord(e, c->stream);
        ev.push_back(e);
        kind.push_back(k);
    }
    // sums the intervals that START with a mark of kind k
    float sum(int k) {
        float total = 0.f;
        for(size_t i = 0; i + 1 < ev.size(); ++i) {
            if(kind[i] == k) {
                float ms = 0.f;
                cudaEventElapsedTime(&ms, ev[i], ev[i + 1]);
                total += ms;
            }
        }
        return total;
    }
};

int copy_in(void* dst, const void* src, size_t bytes, int on_device, cudaStream_t st) {
    B200_CUDA_CHECK(cudaMemcpyAsync(dst, src, bytes,
                                    on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, st));
    return B200VOR_OK;
}

int copy_out(void* dst, const void* src, size_t bytes, int on_device, cudaStream_t st) {
    B200_CUDA_CHECK(cudaMemcpyAsync(dst, src, bytes,
                                    on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyDeviceToHost, st));
    return B200VOR_OK;
}

int cells_pipeline_one(Ctx* c, const b200vor_cell_params* params, bool collect_events) {
    b200vor_cell_params p = {48, 384, 0, 0, 0};
    if(params) {
        p = *params;
    }
    if(p.k0 == 0) {
        p.k0 = 48;
    }
    if(p.kmax == 0) {
        p.kmax = 384;
    }
    if(p.kmax > 569) {
        p.kmax = 569;   // row = k + 1 entries, kNN kernel capacity 570
    }
    if(c->n == 0) {
        set_error("b200vor_compute_cells: no points");
        return B200VOR_ERR_ARG;
    }
    if(c->ndom > 0 && p.mode == 2) {
        set_error("b200vor_compute_cells: the kNN-table pipeline (mode 2) knows no domain planes");
        return B200VOR_ERR_ARG;
    }
    const uint32_t n = c->n;
    const uint64_t ntotal = c->periodic ? 27ull * n : n;
    // row length: k neighbours + the seed itself at rank 0
    auto row_len = [&](uint32_t k) { return uint32_t(std::min<uint64_t>(uint64_t(k) + 1, ntotal)); };
    int rc;
    if((rc = c->m.reserve(n)) || (rc = c->mg.reserve(3 * size_t(n))) ||
       (rc = c->status.reserve(n))) {
        return rc;
    }
    if(p.want_faces) {
        c->fcap = 64;
        if((rc = c->nfaces.reserve(n)) || (rc = c->face_nbr.reserve(size_t(n) * c->fcap)) ||
           (rc = c->face_area.reserve(size_t(n) * c->fcap))) {
            return rc;
        }
    }
    if(p.want_faces == 2) {
        if((rc = c->nverts.reserve(n)) || (rc = c->vert_xyz.reserve(size_t(n) * c->vcap * 3)) ||
           (rc = c->vert_faces.reserve(size_t(n) * c->vcap))) {
            return rc;
        }
        // cells finished by the kNN-table pipeline have no vertex list: nverts = 0xFFFFFFFF
        B200_CUDA_CHECK(cudaMemsetAsync(c->nverts.p, 0xFF, size_t(n) * sizeof(uint32_t), c->stream));
    }
    if(p.want_energy && (rc = c->energy.reserve(n))) {
        return rc;
    }
    if(collect_events) {
        if((rc = c->evcount.reserve(2 * size_t(n)))) {
            return rc;
        }
    }
    Timer T(c);
    uint32_t k = std::min(p.k0, p.kmax);
    uint32_t nq = n;
    const uint32_t* qlist = nullptr;
    if(c->q0 > 0 || c->q1 < n) {
        // sharded run: only the cells of the seeds with user index in [q0, q1)
        if((rc = c->qlist.reserve(n)) || (rc = c->qlist2.reserve(n)) ||
           (rc = select_range(c, c->qlist2.p, &nq))) {
            return rc;
        }
        qlist = c->qlist2.p;
        B200_CUDA_CHECK(cudaMemsetAsync(c->status.p, 0, n, c->stream));
        if(collect_events) {
            B200_CUDA_CHECK(cudaMemsetAsync(c->evcount.p, 0, 2 * size_t(n) * sizeof(uint32_t), c->stream));
        }
    }
    c->stats.passes = 0;
    memset(c->stats.n_pass, 0, sizeof(c->stats.n_pass));
    memset(c->stats.k_pass, 0, sizeof(c->stats.k_pass));
    int pass0 = 0;
    c->stats.tpc_cells = 0;
    const uint32_t nq0 = nq;
    const uint32_t* const qlist0 = qlist;
    // two-level grid on a box domain: far-reaching cells are swept by the large-cell instance (fused_kernel.inc)
    const bool far_sweep = c->grid.two_level && !c->grid.periodic && p.mode != 1 && p.mode != 2;
    c->stats.exact_cells = 0;
    if((p.mode == 4 || (p.mode == 0 && kTpcFirstByDefault)) && p.want_faces <= 1 && !p.want_energy &&
       tpc_applicable(c)) {
        // thread-per-cell kernel (tpc.cu): one launch; what it flags goes to the team kernel
        T.mark(2);
        if((rc = run_cells_tpc(c, nq, qlist, int(p.want_faces), collect_events ? 1 : 0, 0))) {
            return rc;
        }
        T.mark(0);
        c->stats.n_pass[0] = nq;
        c->stats.k_pass[0] = 0;
        c->stats.passes = 1;
        c->stats.tpc_cells = nq;
        pass0 = 1;
        if((rc = c->qlist.reserve(nq)) || (rc = c->qlist2.reserve(nq))) {
            return rc;
        }
        uint32_t* out = (qlist == c->qlist.p) ? c->qlist2.p : c->qlist.p;
        uint32_t nfail = 0;
        if((rc = compact_failed(c, qlist, nq, out, nullptr, &nfail,
                                B200VOR_CELL_UNCERTIFIED | B200VOR_CELL_OVERFLOW))) {
            return rc;
        }
        qlist = out;
        nq = nfail;
    }
    if(p.mode != 2 && nq > 0) {
        // team-kernel grid walk: one launch, no neighbour table (fused.cu)
        T.mark(2);
        c->fused_bail_far = p.mode != 1;   // the large-cell instance follows
        if((rc = run_cells_fused(c, nq, qlist, int(p.want_faces), collect_events ? 1 : 0, 0.f,
                                 int(p.want_energy), 0))) {
            return rc;
        }
        T.mark(0);
        c->stats.n_pass[pass0] = nq;
        c->stats.k_pass[pass0] = 0;
        c->stats.passes = ++pass0;
        if(p.mode == 1) {
            nq = 0;
        } else {
            // the cells it could not finish inside its fixed capacities are re-done by the
            // large-cell instance of the same kernel (one warp per cell, 128 vertices / 64
            // planes); what that leaves goes through the kNN-table pipeline below
            for(int big = 1; big >= 0 && nq > 0; --big) {
                if((rc = c->qlist.reserve(nq)) || (rc = c->qlist2.reserve(nq))) {
                    return rc;
                }
                uint32_t* out = (qlist == c->qlist.p) ? c->qlist2.p : c->qlist.p;
                uint32_t nfail = 0;
                if((rc = compact_failed(c, qlist, nq, out, nullptr, &nfail,
                                        B200VOR_CELL_UNCERTIFIED | B200VOR_CELL_OVERFLOW))) {
                    return rc;
                }
                qlist = out;
                nq = nfail;
                if(big && nq > 0) {
                    T.mark(2);
                    c->fused_sweep_mode = 0;
                    if((rc = run_cells_fused(c, nq, qlist, int(p.want_faces), collect_events ? 1 : 0, 0.f,
                                             int(p.want_energy), 1))) {
                        return rc;
                    }
                    T.mark(0);
                    c->stats.n_pass[pass0] = nq;
                    c->stats.k_pass[pass0] = 0;
                    c->stats.passes = ++pass0;
                }
            }
            // two-level grid: the few far-reaching cells whose sweep ran out of capacity (clipping out of
            // distance order piles up transient faces) are swept again in much finer distance bands, and
            // what even that leaves walks its shells to the end
            for(int sm = 1; sm <= 2 && nq > 0 && far_sweep; ++sm) {
                T.mark(2);
                c->fused_sweep_mode = sm;
                rc = run_cells_fused(c, nq, qlist, int(p.want_faces), collect_events ? 1 : 0, 0.f, int(p.want_energy), 1);
                c->fused_sweep_mode = 0;
                if(rc) {
                    return rc;
                }
                T.mark(0);
                if(pass0 < 8) {
                    c->stats.n_pass[pass0] = nq;
                    c->stats.k_pass[pass0] = 0;
                    c->stats.passes = ++pass0;
                }
                if((rc = c->qlist.reserve(nq)) || (rc = c->qlist2.reserve(nq))) {
                    return rc;
                }
                uint32_t* out = (qlist == c->qlist.p) ? c->qlist2.p : c->qlist.p;
                uint32_t nfail = 0;
                if((rc = compact_failed(c, qlist, nq, out, nullptr, &nfail,
                                        B200VOR_CELL_UNCERTIFIED | B200VOR_CELL_OVERFLOW))) {
                    return rc;
                }
                qlist = out;
                nq = nfail;
            }
        }
    }
    // (the kNN-table fallback knows nothing of open faces: on a slab-local grid the flags stay)
    // (nor can a table of <= kmax neighbours certify the far-reaching cells of clustered seeds that
    //  the large-cell instance gave up on: their flags stay too)
    for(int pass = pass0; pass < 8 && nq > 0 && !c->open && !far_sweep && c->ndom == 0; ++pass) {
        uint32_t rl = row_len(k);
        if((rc = c->knn.reserve(size_t(nq) * rl)) || (rc = c->knn_flag.reserve(nq))) {
            return rc;
        }
        T.mark(1);
        KnnJob kj = {nq, qlist, nullptr, rl, c->knn.p, nullptr, c->knn_flag.p, 0};
        if((rc = run_knn(c, kj))) {
            return rc;
        }
        T.mark(2);
        CellJob cj = {nq, qlist, nullptr, 2, rl, c->knn.p, c->knn_flag.p, int(p.want_faces),
                      collect_events ? 1 : 0, int(p.mode), uint64_t(rl) >= ntotal ? 1 : 0,
                      int(p.want_energy)};
        if((rc = run_cells(c, cj))) {
            return rc;
        }
        // cells that outgrew the 64-triangle register table: same rows, 128-triangle kernel
        {
            if((rc = c->qlist3.reserve(nq)) || (rc = c->qrows.reserve(nq))) {
                return rc;
            }
            uint32_t novf = 0;
            if((rc = compact_failed(c, qlist, nq, c->qlist3.p, c->qrows.p, &novf,
                                    B200VOR_CELL_OVERFLOW))) {
                return rc;
            }
            if(novf) {
                CellJob big = cj;
                big.nq = novf;
                big.qlist = c->qlist3.p;
                big.rowmap = c->qrows.p;
                big.slots = 4;
                if((rc = run_cells(c, big))) {
                    return rc;
                }
            }
        }
        T.mark(0);
        c->stats.n_pass[pass] = nq;
        c->stats.k_pass[pass] = k;
        c->stats.passes = pass + 1;
        if(uint64_t(rl) >= ntotal || k >= p.kmax) {
            break;
        }
        // re-launch only the cells that did not reach their security radius
        if((rc = c->qlist.reserve(nq)) || (rc = c->qlist2.reserve(nq))) {
            return rc;
        }
        uint32_t* out = (qlist == c->qlist.p) ? c->qlist2.p : c->qlist.p;
        uint32_t nfail = 0;
        if((rc = compact_failed(c, qlist, nq, out, nullptr, &nfail, B200VOR_CELL_UNCERTIFIED))) {
            return rc;
        }
        qlist = out;
        nq = nfail;
        k = std::min(2 * k, p.kmax);
    }
    // ---- exact pass: cells a kernel flagged UNCERTAIN (a sign test inside the fp64 error bound:
    //      degenerate input) are re-done by the large thread-per-cell instance, which decides such
    //      ties with expansion arithmetic and the reference's default rule (convex_cell.cpp:905-911)
    if(p.mode != 1 && p.mode != 2 && p.want_faces <= 1 && !p.want_energy && tpc_applicable(c)) {
        if((rc = c->qlist3.reserve(nq0))) {
            return rc;
        }
        uint32_t nunc = 0;
        if((rc = compact_failed(c, qlist0, nq0, c->qlist3.p, nullptr, &nunc, B200VOR_CELL_UNCERTAIN))) {
            return rc;
        }
        if(nunc > 0) {
            T.mark(2);
            if((rc = run_cells_tpc(c, nunc, c->qlist3.p, int(p.want_faces), collect_events ? 1 : 0, 1))) {
                return rc;
            }
            T.mark(0);
            c->stats.exact_cells = nunc;
            if(c->stats.passes < 8) {
                c->stats.n_pass[c->stats.passes] = nunc;
                c->stats.k_pass[c->stats.passes] = 0;
                c->stats.passes++;
            }
        }
    }
    T.mark(0);
    B200_CUDA_CHECK(cudaStreamSynchronize(c->stream));
    c->stats.ms_knn += T.sum(1);
    c->stats.ms_clip += T.sum(2);
    c->cells_valid = true;
    c->faces_valid = p.want_faces != 0;
    c->energy_valid = p.want_energy != 0;
    c->verts_valid = p.want_faces == 2;
    return B200VOR_OK;
}

// acc += the cells of one piece; status: flags of any piece, EMPTY only if empty in every piece
__global__ void piece_accumulate_kernel(uint32_t n, int first, int want_e, const double* __restrict__ m,
                                        const double* __restrict__ mg, const double* __restrict__ e,
                                        const uint8_t* __restrict__ st, double* __restrict__ am,
                                        double* __restrict__ amg, double* __restrict__ ae, uint8_t* __restrict__ ast) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if(i >= n) {
        return;
    }
    const uint8_t s = st[i];
    const uint8_t keep = uint8_t(s & ~uint8_t(B200VOR_CELL_EMPTY));
    if(first) {
        am[i] = m[i];
        amg[3 * size_t(i)] = mg[3 * size_t(i)];
        amg[3 * size_t(i) + 1] = mg[3 * size_t(i) + 1];
        amg[3 * size_t(i) + 2] = mg[3 * size_t(i) + 2];
        if(want_e) {
            ae[i] = e[i];
        }
        ast[i] = s;
    } else {
        am[i] += m[i];
        amg[3 * size_t(i)] += mg[3 * size_t(i)];
        amg[3 * size_t(i) + 1] += mg[3 * size_t(i) + 1];
        amg[3 * size_t(i) + 2] += mg[3 * size_t(i) + 2];
        if(want_e) {
            ae[i] += e[i];
        }
        const uint8_t a = ast[i];
        ast[i] = uint8_t((a & ~uint8_t(B200VOR_CELL_EMPTY)) | keep | (a & s & uint8_t(B200VOR_CELL_EMPTY)));
    }
}

// The cell pass.  A domain made of several convex pieces (b200vor_set_domain_pieces / _tets: the
// tetrahedra of a mesh, generic_RVD.h:1464-1597): every cell is integrated over each piece in turn
// (piece planes = the first candidates of the cell, fused_clip.inc) and the integrals are added up.
int cells_pipeline(Ctx* c, const b200vor_cell_params* params, bool collect_events) {
    const size_t np = c->piece_off.empty() ? 0 : c->piece_off.size() - 1;
    if(np == 0) {
        return cells_pipeline_one(c, params, collect_events);
    }
    b200vor_cell_params p = {0, 0, 0, 0, 0};
    if(params) {
        p = *params;
    }
    if(p.want_faces) {
        set_error("b200vor_compute_cells: no face lists on a domain made of several pieces");
        return B200VOR_ERR_ARG;
    }
    const uint32_t n = c->n;
    int rc;
    if((rc = c->acc_m.reserve(n)) || (rc = c->acc_mg.reserve(3 * size_t(n))) || (rc = c->acc_st.reserve(n)) ||
       (p.want_energy && (rc = c->acc_e.reserve(n)))) {
        return rc;
    }
    float ms_clip = 0.f, ms_knn = 0.f;
    for(size_t k = 0; k < np; ++k) {
        const uint32_t a = c->piece_off[k], b = c->piece_off[k + 1];
        if((rc = c->dom_dev.reserve(4 * size_t(b - a)))) {
            return rc;
        }
        B200_CUDA_CHECK(cudaMemcpyAsync(c->dom_dev.p, &c->piece_planes[4 * size_t(a)], 4 * size_t(b - a) * sizeof(double),
                                        cudaMemcpyHostToDevice, c->stream));
        c->ndom = int(b - a);
        c->stats.ms_clip = c->stats.ms_knn = 0.f;
        rc = cells_pipeline_one(c, &p, collect_events && k + 1 == np);
        if(rc) {
            c->ndom = 0;
            return rc;
        }
        ms_clip += c->stats.ms_clip;
        ms_knn += c->stats.ms_knn;
        piece_accumulate_kernel<<<(n + 255) / 256, 256, 0, c->stream>>>(
            n, k == 0 ? 1 : 0, p.want_energy ? 1 : 0, c->m.p, c->mg.p, p.want_energy ? c->energy.p : nullptr, c->status.p,
            c->acc_m.p, c->acc_mg.p, p.want_energy ? c->acc_e.p : nullptr, c->acc_st.p);
        c->stats.gpu_launches++;
        B200_CUDA_CHECK(cudaGetLastError());
    }
    c->ndom = 0;
    B200_CUDA_CHECK(cudaMemcpyAsync(c->m.p, c->acc_m.p, size_t(n) * sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
    B200_CUDA_CHECK(cudaMemcpyAsync(c->mg.p, c->acc_mg.p, 3 * size_t(n) * sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
    B200_CUDA_CHECK(cudaMemcpyAsync(c->status.p, c->acc_st.p, size_t(n), cudaMemcpyDeviceToDevice, c->stream));
    if(p.want_energy) {
        B200_CUDA_CHECK(cudaMemcpyAsync(c->energy.p, c->acc_e.p, size_t(n) * sizeof(double), cudaMemcpyDeviceToDevice, c->stream));
    }
    B200_CUDA_CHECK(cudaStreamSynchronize(c->stream));
    c->stats.ms_clip = ms_clip;
    c->stats.ms_knn = ms_knn;
    return B200VOR_OK;
}

}  // namespace

// cell pass with face lists on the context's current weights (otm.cu)
int otm_cells(Ctx* c) {
    b200vor_cell_params p = {0, 0, 1, 0, 0};
    int rc = cells_pipeline(c, &p, false);
    return rc;
}
int otm_newton(Ctx* c, const double* nu, double* w, const b200vor_otm_params* prm, b200vor_otm_result* res,
               int on_device);

namespace {

int ensure_grid(Ctx* c) {
    Timer T(c);
    T.mark(1);
    int rc = build_grid(c);
    if(rc) {
        return rc;
    }
    T.mark(0);
    B200_CUDA_CHECK(cudaStreamSynchronize(c->stream));
    c->stats.ms_grid += T.sum(1);
    c->grid_dirty = false;
    return B200VOR_OK;
}

}  // namespace

int ctx_build_grid_timed(Ctx* c) { return ensure_grid(c); }
int ctx_cells(Ctx* c, const b200vor_cell_params* params, bool collect_events) {
    return cells_pipeline(c, params, collect_events);
}
}  // namespace b200vor

using namespace b200vor;

extern "C" {

int b200vor_version(void) { return B200VOR_VERSION; }

const char* b200vor_last_error(void) { return g_error; }

int b200vor_create(b200vor_ctx** out, int device) {
    if(!out) {
        set_error("b200vor_create: null output pointer");
        return B200VOR_ERR_ARG;
    }
    *out = nullptr;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if(e != cudaSuccess || count == 0) {
        set_error("b200vor_create: no CUDA device (%s); this library has no CPU fallback",
                  e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
        return B200VOR_ERR_CUDA;
    }
    if(device < 0 || device >= count) {
        set_error("b200vor_create: device %d out of range (0..%d)", device, count - 1);
        return B200VOR_ERR_ARG;
    }
    B200_CUDA_CHECK(cudaSetDevice(device));
    b200vor_ctx* h = new b200vor_ctx();
    h->c.device = device;
    {
        cudaError_t se = cudaStreamCreateWithFlags(&h->c.stream, cudaStreamNonBlocking);
        if(se != cudaSuccess) {
            set_error("b200vor_create: cudaStreamCreateWithFlags -> %s", cudaGetErrorString(se));
            delete h;
            return B200VOR_ERR_CUDA;
        }
    }
    h->c.own_stream = true;
    *out = h;
    return B200VOR_OK;
}

void b200vor_destroy(b200vor_ctx* h) {
    if(!h) {
        return;
    }
    Ctx& c = h->c;
    cudaSetDevice(c.device);
    cudaStreamSynchronize(c.stream);
    for(cudaEvent_t e : c.event_pool) {
        cudaEventDestroy(e);
    }
    c.event_pool.clear();
    slab_release(&c);
    c.xyz.release(); c.wgt.release(); c.cellid.release(); c.rank.release();
    c.sub.release(); c.vstart.release(); c.vcount.release(); c.sbcount.release(); c.dom_dev.release(); c.acc_m.release(); c.acc_mg.release(); c.acc_e.release(); c.acc_st.release(); c.wloc.release(); c.wloc_tmp.release();
    c.cell_start.release(); c.perm.release(); c.scan_tmp.release(); c.sorted.release();
    c.knn.release(); c.knn_flag.release(); c.qlist.release(); c.qlist2.release(); c.qlist3.release(); c.qrows.release();
    c.counter.release(); c.m.release(); c.mg.release(); c.status.release();
    c.energy.release(); c.nverts.release(); c.vert_xyz.release(); c.vert_faces.release(); c.nfaces.release(); c.face_nbr.release(); c.face_area.release(); c.evcount.release();
    c.sums.release(); c.out_nbr.release(); c.out_d2.release(); c.qpts.release();
    if(c.own_stream) {
        cudaStreamDestroy(c.stream);
    }
    delete h;
}

int b200vor_set_domain_box(b200vor_ctx* h, const double* box) {
    if(!h || !box) {
        set_error("b200vor_set_domain_box: null argument");
        return B200VOR_ERR_ARG;
    }
    for(int a = 0; a < 3; ++a) {
        if(!(box[3 + a] > box[a])) {
            set_error("b200vor_set_domain_box: empty box");
            return B200VOR_ERR_ARG;
        }
    }
    slab_deactivate(&h->c);   // (a slab run captures the domain again at its next set_points)
    memcpy(h->c.box, box, 6 * sizeof(double));
    h->c.periodic = 0;
    h->c.cells_valid = h->c.faces_valid = false;
    return B200VOR_OK;
}

int b200vor_set_domain_planes(b200vor_ctx* h, uint32_t n, const double* planes) {
    if(!h || n > 32 || (n > 0 && !planes)) {
        set_error("b200vor_set_domain_planes: invalid argument (at most 32 planes)");
        return B200VOR_ERR_ARG;
    }
    Ctx* c = &h->c;
    if(n > 0 && c->periodic) {
        set_error("b200vor_set_domain_planes: box domains only");
        return B200VOR_ERR_ARG;
    }
    B200_CUDA_CHECK(cudaSetDevice(c->device));
    for(uint32_t k = 0; k < n; ++k) {
        const double* p = planes + 4 * size_t(k);
        if(!(p[0] * p[0] + p[1] * p[1] + p[2] * p[2] > 0.0)) {
            set_error("b200vor_set_domain_planes: plane %u has no normal", k);
            return B200VOR_ERR_ARG;
        }
    }
    int rc;
    if(n > 0) {
        if((rc = c->dom_dev.reserve(4 * size_t(n)))) {
            return rc;
        }
        B200_CUDA_CHECK(cudaMemcpyAsync(c->dom_dev.p, planes, 4 * size_t(n) * sizeof(double), cudaMemcpyHostToDevice, c->stream));
        B200_CUDA_CHECK(cudaStreamSynchronize(c->stream));
    }
    c->ndom = int(n);
    c->piece_off.clear();
    c->piece_planes.clear();
    c->cells_valid = c->faces_valid = false;
    return B200VOR_OK;
}

int b200vor_set_domain_pieces(b200vor_ctx* h, uint32_t npieces, const uint32_t* nplanes, const double* planes) {
    if(!h || npieces > 4096 || (npieces > 0 && (!nplanes || !planes))) {
        set_error("b200vor_set_domain_pieces: invalid argument (at most 4096 pieces)");
        return B200VOR_ERR_ARG;
    }
    Ctx* c = &h->c;
    if(npieces > 0 && c->periodic) {
        set_error("b200vor_set_domain_pieces: box domains only");
        return B200VOR_ERR_ARG;
    }
    std::vector<uint32_t> off(1, 0u);
    for(uint32_t k = 0; k < npieces; ++k) {
        if(nplanes[k] == 0 || nplanes[k] > 32) {
            set_error("b200vor_set_domain_pieces: piece %u has %u planes (1..32)", k, nplanes[k]);
            return B200VOR_ERR_ARG;
        }
        off.push_back(off.back() + nplanes[k]);
    }
    c->ndom = 0;
    if(npieces == 0) {
        c->piece_off.clear();
        c->piece_planes.clear();
    } else {
        c->piece_off = off;
        c->piece_planes.assign(planes, planes + 4 * size_t(off.back()));
    }
    c->cells_valid = c->faces_valid = false;
    return B200VOR_OK;
}

int b200vor_set_domain_tets(b200vor_ctx* h, uint32_t nv, const double* v, uint32_t nt, const uint32_t* tets) {
    if(!h || !v || !tets || nt == 0 || nv < 4) {
        set_error("b200vor_set_domain_tets: invalid argument");
        return B200VOR_ERR_ARG;
    }
    std::vector<uint32_t> np(nt, 4u);
    std::vector<double> pl(16 * size_t(nt));
    for(uint32_t t = 0; t < nt; ++t) {
        const uint32_t* q = tets + 4 * size_t(t);
        for(int f = 0; f < 4; ++f) {
            if(q[f] >= nv) {
                set_error("b200vor_set_domain_tets: vertex index out of range in tetrahedron %u", t);
                return B200VOR_ERR_ARG;
            }
        }
        for(int f = 0; f < 4; ++f) {
            // face opposite to vertex f, normal turned towards vertex f
            const double* a = v + 3 * size_t(q[(f + 1) & 3]);
            const double* b = v + 3 * size_t(q[(f + 2) & 3]);
            const double* cc = v + 3 * size_t(q[(f + 3) & 3]);
            const double* o = v + 3 * size_t(q[f]);
            const double ux = b[0] - a[0], uy = b[1] - a[1], uz = b[2] - a[2];
            const double wx = cc[0] - a[0], wy = cc[1] - a[1], wz = cc[2] - a[2];
            double nx = uy * wz - uz * wy, ny = uz * wx - ux * wz, nz = ux * wy - uy * wx;
            double d = -(nx * a[0] + ny * a[1] + nz * a[2]);
            const double side = nx * o[0] + ny * o[1] + nz * o[2] + d;
            if(side == 0.0) {
                set_error("b200vor_set_domain_tets: tetrahedron %u is flat", t);
                return B200VOR_ERR_ARG;
            }
            if(side < 0.0) {
                nx = -nx; ny = -ny; nz = -nz; d = -d;
            }
            double* out = &pl[16 * size_t(t) + 4 * f];
            out[0] = nx; out[1] = ny; out[2] = nz; out[3] = d;
        }
    }
    // the clipping box only has to contain the mesh; it is set a little larger than the mesh so that
    // no plane of a tetrahedron passes exactly through a corner or an edge of the box (exact ties)
    double lo[3] = {1e300, 1e300, 1e300}, hi[3] = {-1e300, -1e300, -1e300};
    for(uint32_t t = 0; t < nt; ++t) {
        for(int f = 0; f < 4; ++f) {
            const double* q = v + 3 * size_t(tets[4 * size_t(t) + f]);
            for(int a = 0; a < 3; ++a) {
                lo[a] = fmin(lo[a], q[a]);
                hi[a] = fmax(hi[a], q[a]);
            }
        }
    }
    const double ext = fmax(hi[0] - lo[0], fmax(hi[1] - lo[1], hi[2] - lo[2]));
    int rc = b200vor_set_domain_pieces(h, nt, np.data(), pl.data());
    if(rc) {
        return rc;
    }
    const double margin[3] = {0.0371, 0.0413, 0.0457};   // (unequal: no symmetric coincidence either)
    for(int a = 0; a < 3; ++a) {
        h->c.box[a] = lo[a] - margin[a] * ext;
        h->c.box[3 + a] = hi[a] + margin[a] * ext;
    }
    h->c.periodic = 0;
    return B200VOR_OK;
}

int b200vor_set_domain_periodic(b200vor_ctx* h, double period) {
    if(!h || !(period > 0.0)) {
        set_error("b200vor_set_domain_periodic: invalid argument");
        return B200VOR_ERR_ARG;
    }
    if(uint64_t(h->c.n) * 27ull >= 0xFFFFFFF0ull) {
        // the points were set first: the same check as in b200vor_set_points
        set_error("b200vor_set_domain_periodic: n too large for 32-bit periodic ids");
        return B200VOR_ERR_ARG;
    }
    if(h->c.ndom > 0 || !h->c.piece_off.empty()) {
        set_error("b200vor_set_domain_periodic: remove the domain planes / pieces first (b200vor_set_domain_planes(ctx, 0, NULL))");
        return B200VOR_ERR_ARG;
    }
    slab_deactivate(&h->c);
    h->c.periodic = 1;
    h->c.period = period;
    h->c.cells_valid = h->c.faces_valid = false;
    return B200VOR_OK;
}

int b200vor_set_stream(b200vor_ctx* h, void* cuda_stream) {
    if(!h) {
        set_error("b200vor_set_stream: null context");
        return B200VOR_ERR_ARG;
    }
    Ctx& c = h->c;
    cudaSetDevice(c.device);
    cudaStreamSynchronize(c.stream);
    if(c.own_stream) {
        cudaStreamDestroy(c.stream);
    }
    c.stream = static_cast<cudaStream_t>(cuda_stream);
    c.own_stream = false;
    return B200VOR_OK;
}

int b200vor_set_query_range(b200vor_ctx* h, uint32_t first, uint32_t last) {
    if(!h || first > last) {
        set_error("b200vor_set_query_range: invalid range");
        return B200VOR_ERR_ARG;
    }
    h->c.q0 = first;
    h->c.q1 = last;
    h->c.cells_valid = h->c.faces_valid = false;
    return B200VOR_OK;
}

int b200vor_set_points_per_cell(b200vor_ctx* h, double ppc) {
    if(!h || !(ppc > 0.01)) {
        set_error("b200vor_set_points_per_cell: invalid argument");
        return B200VOR_ERR_ARG;
    }
    h->c.ppc = ppc;
    return B200VOR_OK;
}

int b200vor_set_points(b200vor_ctx* h, uint32_t n, const double* xyz, uint32_t stride,
                       const double* w, int on_device) {
    if(!h || !xyz || n == 0 || stride < 3) {
        set_error("b200vor_set_points: invalid argument (n=%u stride=%u)", n, stride);
        return B200VOR_ERR_ARG;
    }
    if(uint64_t(n) * 27ull >= 0xFFFFFFF0ull && h->c.periodic) {
        set_error("b200vor_set_points: n too large for 32-bit periodic ids");
        return B200VOR_ERR_ARG;
    }
    Ctx* c = &h->c;
    B200_CUDA_CHECK(cudaSetDevice(c->device));
    slab_deactivate(c);
    int rc;
    if((rc = c->xyz.reserve(3 * size_t(n)))) {
        return rc;
    }
    c->n = n;
    c->cells_valid = c->faces_valid = false;
    if(stride == 3) {
        if((rc = copy_in(c->xyz.p, xyz, 3 * size_t(n) * sizeof(double), on_device, c->stream))) {
            return rc;
        }
    } else if(on_device) {
        gather3_kernel<<<(n + 255) / 256, 256, 0, c->stream>>>(xyz, stride, n, c->xyz.p);
        c->stats.gpu_launches++;
    } else {
        B200_CUDA_CHECK(cudaMemcpy2DAsync(c->xyz.p, 3 * sizeof(double), xyz,
                                          stride * sizeof(double), 3 * sizeof(double), n,
                                          cudaMemcpyHostToDevice, c->stream));
    }
    c->has_w = (w != nullptr);
    c->wmax = 0.0;
    if(w) {
        if((rc = c->wgt.reserve(n)) ||
           (rc = copy_in(c->wgt.p, w, size_t(n) * sizeof(double), on_device, c->stream))) {
            return rc;
        }
        const int nb = 128;
        if((rc = c->sums.reserve(256))) {
            return rc;
        }
        double* part = reinterpret_cast<double*>(c->sums.p);
        max_kernel<<<nb, 256, 0, c->stream>>>(c->wgt.p, n, part);
        c->stats.gpu_launches++;
        double hmax[nb];
        B200_CUDA_CHECK(cudaMemcpyAsync(hmax, part, sizeof(hmax), cudaMemcpyDeviceToHost, c->stream));
        B200_CUDA_CHECK(cudaStreamSynchronize(c->stream));
        c->wmax = hmax[0];
        for(int b = 1; b < nb; ++b) {
            c->wmax = fmax(c->wmax, hmax[b]);
        }
    }
    c->stats.ms_grid = c->stats.ms_knn = c->stats.ms_clip = 0.f;
    return ensure_grid(c);
}

int b200vor_knn(b200vor_ctx* h, uint32_t k, uint32_t* nbr, double* d2, int on_device) {
    if(h && slab_active(&h->c)) {
        set_error("b200vor_knn: not available on a slab context (b200vor_slab_set_points)");
        return B200VOR_ERR_ARG;
    }
    if(!h || !nbr || k == 0) {
        set_error("b200vor_knn: invalid argument");
        return B200VOR_ERR_ARG;
    }
    Ctx* c = &h->c;
    if(c->n == 0 || k > c->n) {
        set_error("b200vor_knn: k=%u exceeds the number of points (%u)", k, c->n);
        return B200VOR_ERR_ARG;
    }
    B200_CUDA_CHECK(cudaSetDevice(c->device));
    int rc;
    uint32_t* dn = nbr;
    double* dd = d2;
    size_t cnt = size_t(c->n) * k;
    if(!on_device) {
        if((rc = c->out_nbr.reserve(cnt)) || (d2 && (rc = c->out_d2.reserve(cnt)))) {
            return rc;
        }
        dn = c->out_nbr.p;
        dd = d2 ? c->out_d2.p : nullptr;
    }
    // the public kNN is not offered on a periodic domain (the reference's kd-tree has none)
    if(c->periodic) {
        set_error("b200vor_knn: not available on a periodic domain");
        return B200VOR_ERR_ARG;
    }
    if((c->grid_dirty || c->grid.periodic != c->periodic) && (rc = ensure_grid(c))) {
        return rc;   // the domain changed after set_points: the grid is rebuilt for the current one
    }
    Timer T(c);
    T.mark(1);
    KnnJob kj = {c->n, nullptr, nullptr, k, dn, dd, nullptr, 1};
    if((rc = run_knn(c, kj))) {
        return rc;
    }
    T.mark(0);
    if(!on_device) {
        if((rc = copy_out(nbr, dn, cnt * sizeof(uint32_t), 0, c->stream)) ||
           (d2 && (rc = copy_out(d2, dd, cnt * sizeof(double), 0, c->stream)))) {
            return rc;
        }
    }
    B200_CUDA_CHECK(cudaStreamSynchronize(c->stream));
    c->stats.ms_knn = T.sum(1);
    return B200VOR_OK;
}

int b200vor_knn_points(b200vor_ctx* h, uint32_t nq, const double* q, uint32_t k,
                       uint32_t* nbr, double* d2, int on_device) {
    if(!h || !nbr || !q || k == 0 || nq == 0) {
        set_error("b200vor_knn_points: invalid argument");
        return B200VOR_ERR_ARG;
    }
    Ctx* c = &h->c;
    if(c->n == 0 || k > c->n) {
        set_error("b200vor_knn_points: k=%u exceeds the number of points (%u)", k, c->n);
        return B200VOR_ERR_ARG;
    }
    if(c->periodic) {
        set_error("b200vor_knn_points: not available on a periodic domain");
        return B200VOR_ERR_ARG;
    }
    B200_CUDA_CHECK(cudaSetDevice(c->device));
    int rc;
    if((c->grid_dirty || c->grid.periodic != c->periodic) && (rc = ensure_grid(c))) {
        return rc;
    }
    uint32_t* dn = nbr;
    double* dd = d2;
    const double* dq = q;
    size_t cnt = size_t(nq) * k;
    if(!on_device) {
        if((rc = c->out_nbr.reserve(cnt)) || (d2 && (rc = c->out_d2.reserve(cnt))) ||
           (rc = c->qpts.reserve(3 * size_t(nq))) ||
           (rc = copy_in(c->qpts.p, q, 3 * size_t(nq) * sizeof(double), 0, c->stream))) {
            return rc;
        }
        dn = c->out_nbr.p;
        dd = d2 ? c->out_d2.p : nullptr;
        dq = c->qpts.p;
    }
    KnnJob kj = {nq, nullptr, dq, k, dn, dd, nullptr, 1};
    if((rc = run_knn(c, kj))) {
        return rc;
    }
    if(!on_device) {
        if((rc = copy_out(nbr, dn, cnt * sizeof(uint32_t), 0, c->stream)) ||
           (d2 && (rc = copy_out(d2, dd, cnt * sizeof(double), 0, c->stream)))) {
            return rc;
        }
    }
    B200_CUDA_CHECK(cudaStreamSynchronize(c->stream));
    return B200VOR_OK;
}

int b200vor_compute_cells(b200vor_ctx* h, const b200vor_cell_params* params) {
    if(h && slab_active(&h->c)) {
        set_error("b200vor_compute_cells: the context holds a slab (b200vor_slab_set_points): use the b200vor_slab_* calls");
        return B200VOR_ERR_ARG;
    }
    if(!h) {
        set_error("b200vor_compute_cells: null context");
        return B200VOR_ERR_ARG;
    }
    Ctx* c = &h->c;
    B200_CUDA_CHECK(cudaSetDevice(c->device));
    if(c->grid_dirty || c->grid.periodic != c->periodic ||
       (c->periodic && c->grid.period != c->period)) {
        c->stats.ms_grid = 0.f;
        int rc = ensure_grid(c);   // points moved (Lloyd) or the domain changed after set_points
        if(rc) {
            return rc;
        }
    }
    c->stats.ms_knn = c->stats.ms_clip = 0.f;
    int rc = cells_pipeline(c, params, true);
    if(rc) {
        return rc;
    }
    c->stats.ms_total = c->stats.ms_grid + c->stats.ms_knn + c->stats.ms_clip;
    return reduce_events(c);
}

int b200vor_get_cells(b200vor_ctx* h, double* m, double* mg, uint8_t* status, int on_device) {
    if(!h || !h->c.cells_valid) {
        set_error("b200vor_get_cells: no cells computed");
        return B200VOR_ERR_ARG;
    }
    Ctx* c = &h->c;
    B200_CUDA_CHECK(cudaSetDevice(c->device));
    int rc;
    if((m && (rc = copy_out(m, c->m.p, size_t(out_rows(c)) * sizeof(double), on_device, c->stream))) ||
       (mg && (rc = copy_out(mg, c->mg.p, 3 * size_t(out_rows(c)) * sizeof(double), on_device, c->stream))) ||
       (status && (rc = copy_out(status, c->status.p, size_t(out_rows(c)), on_device, c->stream)))) {
        return rc;
    }
    B200_CUDA_CHECK(cudaStreamSynchronize(c->stream));
    return B200VOR_OK;
}

int b200vor_get_cvt(b200vor_ctx* h, double* f_cell, double* g, double* f_total, int on_device) {
    if(!h || !h->c.cells_valid || !h->c.energy_valid) {
        set_error("b200vor_get_cvt: cells were not computed with want_energy=1");
        return B200VOR_ERR_ARG;
    }
    Ctx* c = &h->c;
    B200_CUDA_CHECK(cudaSetDevice(c->device));
    int rc;
    const uint32_t n = c->n;
    if(f_cell && (rc = copy_out(f_cell, c->energy.p, size_t(n) * sizeof(double), on_device, c->stream))) {
        return rc;
    }
    if(g) {
        double* dg = g;
        if(!on_device) {
            if((rc = c->out_d2.reserve(3 * size_t(n)))) {
                return rc;
            }
            dg = c->out_d2.p;
        }
        cvt_grad_kernel<<<(3 * n + 255) / 256, 256, 0, c->stream>>>(n, c->xyz.p, c->m.p, c->mg.p, dg);
        c->stats.gpu_launches++;
        B200_CUDA_CHECK(cudaGetLastError());
        if(!on_device && (rc = copy_out(g, dg, 3 * size_t(n) * sizeof(double), 0, c->stream))) {
            return rc;
        }
    }
    if(f_total) {
        const int nb = 128;
        if((rc = c->sums.reserve(256))) {
            return rc;
        }
        double* part = reinterpret_cast<double*>(c->sums.p);
        sum_kernel<<<nb, 256, 0, c->stream>>>(c->energy.p, n, part);
        c->stats.gpu_launches++;
        B200_CUDA_CHECK(cudaGetLastError());
        double hp[nb];
        B200_CUDA_CHECK(cudaMemcpyAsync(hp, part, sizeof(hp), cudaMemcpyDeviceToHost, c->stream));
        B200_CUDA_CHECK(cudaStreamSynchronize(c->stream));
        double t = 0.0;
        for(int b = 0; b < nb; ++b) {
            t += hp[b];
        }
        *f_total = t;
    }
    B200_CUDA_CHECK(cudaStreamSynchronize(c->stream));
    return B200VOR_OK;
}

int b200vor_get_vertices(b200vor_ctx* h, uint32_t* nverts, double* vert_xyz, uint32_t* vert_faces,
                         uint32_t vcap, int on_device) {
    if(!h || !h->c.cells_valid || !h->c.verts_valid) {
        set_error("b200vor_get_vertices: cells were not computed with want_faces=2");
        return B200VOR_ERR_ARG;
    }
    Ctx* c = &h->c;
    if(vcap != c->vcap) {
        set_error("b200vor_get_vertices: vcap must be %u", c->vcap);
        return B200VOR_ERR_ARG;
    }
    B200_CUDA_CHECK(cudaSetDevice(c->device));
    int rc;
    size_t cnt = size_t(c->n) * vcap;
    if((nverts && (rc = copy_out(nverts, c->nverts.p, size_t(c->n) * sizeof(uint32_t), on_device, c->stream))) ||
       (vert_xyz && (rc = copy_out(vert_xyz, c->vert_xyz.p, cnt * 3 * sizeof(double), on_device, c->stream))) ||
       (vert_faces && (rc = copy_out(vert_faces, c->vert_faces.p, cnt * sizeof(uint32_t), on_device, c->stream)))) {
        return rc;
    }
    B200_CUDA_CHECK(cudaStreamSynchronize(c->stream));
    return B200VOR_OK;
}

int b200vor_otm_newton(b200vor_ctx* h, const double* nu, double* w, const b200vor_otm_params* params,
                       b200vor_otm_result* result, int on_device) {
    if(h && slab_active(&h->c)) {
        set_error("b200vor_otm_newton: not available on a slab context (b200vor_slab_set_points)");
        return B200VOR_ERR_ARG;
    }
    if(!h || !w || h->c.n == 0) {
        set_error("b200vor_otm_newton: needs a context with points and a weight array");
        return B200VOR_ERR_ARG;
    }
    Ctx* c = &h->c;
    if(c->q0 > 0 || c->q1 < c->n) {
        set_error("b200vor_otm_newton: not available on a sharded context (b200vor_set_query_range)");
        return B200VOR_ERR_ARG;
    }
    B200_CUDA_CHECK(cudaSetDevice(c->device));
    c->stats.ms_grid = c->stats.ms_knn = c->stats.ms_clip = 0.f;
    int rc = otm_newton(c, nu, w, params, result, on_device);
    c->stats.ms_total = c->stats.ms_grid + c->stats.ms_knn + c->stats.ms_clip;
    return rc;
}

int b200vor_get_faces(b200vor_ctx* h, uint32_t* nfaces, uint32_t* face_nbr, double* face_area,
                      uint32_t fcap, int on_device) {
    if(!h || !h->c.faces_valid) {
        set_error("b200vor_get_faces: cells were not computed with want_faces=1");
        return B200VOR_ERR_ARG;
    }
    Ctx* c = &h->c;
    if(fcap != c->fcap) {
        set_error("b200vor_get_faces: fcap must be %u", c->fcap);
        return B200VOR_ERR_ARG;
    }
    B200_CUDA_CHECK(cudaSetDevice(c->device));
    int rc;
    size_t cnt = size_t(out_rows(c)) * fcap;
    if((nfaces && (rc = copy_out(nfaces, c->nfaces.p, size_t(out_rows(c)) * sizeof(uint32_t), on_device, c->stream))) ||
       (face_nbr && (rc = copy_out(face_nbr, c->face_nbr.p, cnt * sizeof(uint32_t), on_device, c->stream))) ||
       (face_area && (rc = copy_out(face_area, c->face_area.p, cnt * sizeof(double), on_device, c->stream)))) {
        return rc;
    }
    B200_CUDA_CHECK(cudaStreamSynchronize(c->stream));
    return B200VOR_OK;
}

int b200vor_lloyd(b200vor_ctx* h, uint32_t iters, const b200vor_cell_params* params,
                  double* xyz_out, int on_device) {
    if(h && slab_active(&h->c)) {
        set_error("b200vor_lloyd: the context holds a slab (b200vor_slab_set_points): use the b200vor_slab_* calls");
        return B200VOR_ERR_ARG;
    }
    if(!h || h->c.n == 0) {
        set_error("b200vor_lloyd: no points");
        return B200VOR_ERR_ARG;
    }
    Ctx* c = &h->c;
    B200_CUDA_CHECK(cudaSetDevice(c->device));
    b200vor_cell_params p = {48, 384, 0, 0, 0};
    if(params) {
        p = *params;
    }
    p.want_faces = 0;
    p.want_energy = 0;
    int rc;
    c->stats.ms_knn = c->stats.ms_clip = 0.f;
    if(c->grid_dirty) {
        c->stats.ms_grid = 0.f;   // a clean grid keeps the time of the build that made it
    }
    for(uint32_t it = 0; it < iters; ++it) {
        if(c->grid_dirty && (rc = ensure_grid(c))) {
            return rc;
        }
        if((rc = cells_pipeline(c, &p, false)) || (rc = lloyd_update(c))) {
            return rc;
        }
        c->grid_dirty = true;
    }
    if(iters > 0) {
        c->cells_valid = c->faces_valid = false;   // the grid is rebuilt lazily at the next use
    }
    if(xyz_out) {
        if((rc = copy_out(xyz_out, c->xyz.p, 3 * size_t(c->n) * sizeof(double), on_device, c->stream))) {
            return rc;
        }
    }
    B200_CUDA_CHECK(cudaStreamSynchronize(c->stream));
    c->stats.ms_total = c->stats.ms_grid + c->stats.ms_knn + c->stats.ms_clip;
    return B200VOR_OK;
}

int b200vor_get_spatial_order(b200vor_ctx* h, uint32_t* order, int on_device) {
    if(!h || !order || h->c.n == 0) {
        set_error("b200vor_get_spatial_order: needs a context with points and an output array");
        return B200VOR_ERR_ARG;
    }
    Ctx* c = &h->c;
    B200_CUDA_CHECK(cudaSetDevice(c->device));
    int rc;
    if(c->grid_dirty && (rc = ensure_grid(c))) {
        return rc;
    }
    if((rc = copy_out(order, c->perm.p, size_t(c->n) * sizeof(uint32_t), on_device, c->stream))) {
        return rc;
    }
    B200_CUDA_CHECK(cudaStreamSynchronize(c->stream));
    return B200VOR_OK;
}

int b200vor_get_hilbert_order(b200vor_ctx* h, uint32_t* order, int on_device) {
    if(!h || !order || h->c.n == 0) {
        set_error("b200vor_get_hilbert_order: needs a context with points and an output array");
        return B200VOR_ERR_ARG;
    }
    Ctx* c = &h->c;
    B200_CUDA_CHECK(cudaSetDevice(c->device));
    int rc;
    if(c->grid_dirty && (rc = ensure_grid(c))) {
        return rc;
    }
    uint32_t* dst = order;
    if(!on_device) {
        if((rc = c->out_nbr.reserve(c->n))) {
            return rc;
        }
        dst = c->out_nbr.p;
    }
    rc = hilbert_order(c, dst);
    if(rc == 1) {
        // a grid too elongated for the curve enumeration: the raster order of the bins
        B200_CUDA_CHECK(cudaMemcpyAsync(dst, c->perm.p, size_t(c->n) * sizeof(uint32_t), cudaMemcpyDeviceToDevice, c->stream));
    } else if(rc) {
        return rc;
    }
    if(!on_device && (rc = copy_out(order, dst, size_t(c->n) * sizeof(uint32_t), 0, c->stream))) {
        return rc;
    }
    B200_CUDA_CHECK(cudaStreamSynchronize(c->stream));
    return B200VOR_OK;
}

int b200vor_selftest_exact_sign(b200vor_ctx* h, uint32_t n, const double* planes, int* signs) {
    if(!h || !planes || !signs || n == 0) {
        set_error("b200vor_selftest_exact_sign: invalid argument");
        return B200VOR_ERR_ARG;
    }
    B200_CUDA_CHECK(cudaSetDevice(h->c.device));
    return exact_sign_selftest(&h->c, n, planes, signs);
}

int b200vor_microbench(b200vor_ctx* h, b200vor_peaks* peaks) {
    if(!h || !peaks) {
        set_error("b200vor_microbench: null argument");
        return B200VOR_ERR_ARG;
    }
    B200_CUDA_CHECK(cudaSetDevice(h->c.device));
    memset(peaks, 0, sizeof(*peaks));
    return microbench(&h->c, peaks);
}

int b200vor_get_stats(b200vor_ctx* h, b200vor_stats* stats) {
    if(!h || !stats) {
        set_error("b200vor_get_stats: null argument");
        return B200VOR_ERR_ARG;
    }
    *stats = h->c.stats;
    return B200VOR_OK;
}

}  // extern "C"
