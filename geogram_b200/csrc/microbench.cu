// microbench.cu -- measured pipe peaks of the GPU the context runs on: the denominators of the
// flop / issue rooflines bench.py reports for the cell kernels (MEASURED_PEAKS.json holds only
// the HBM copy bandwidth and the bf16 tensor throughput; the cell path is FP64 / FP32 / issue
// bound, so its peaks are measured here, in the same process, with CUDA events).
#include "common.cuh"

namespace b200vor {
namespace {

// `chains` independent FMA chains per thread (enough to cover the pipe latency), `iters` rounds
template <class T, int CH> __global__ void __launch_bounds__(256) fma_peak_kernel(T* out, int iters, T a, T b) {
    T acc[CH];
#pragma unroll
    for(int k = 0; k < CH; ++k) {
        acc[k] = T(threadIdx.x + k);
    }
    for(int it = 0; it < iters; ++it) {
#pragma unroll
        for(int k = 0; k < CH; ++k) {
            acc[k] = fma(acc[k], a, b);
        }
    }
    T s = T(0);
#pragma unroll
    for(int k = 0; k < CH; ++k) {
        s += acc[k];
    }
    if(s == T(-1.2345)) {   // never true: keeps the chains alive
        out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    }
}

template <class T> int measure(Ctx* c, double* tflops) {
    constexpr int CH = 8;
    int sms = 0;
    B200_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, c->device));
    const int blocks = sms * 8, threads = 256, iters = 4096;
    int rc;
    if((rc = c->out_d2.reserve(size_t(blocks) * threads))) {
        return rc;
    }
    cudaEvent_t e0, e1;
    B200_CUDA_CHECK(cudaEventCreate(&e0));
    B200_CUDA_CHECK(cudaEventCreate(&e1));
    double best = 0.0;
    for(int rep = 0; rep < 5; ++rep) {
        B200_CUDA_CHECK(cudaEventRecord(e0, c->stream));
        fma_peak_kernel<T, CH><<<blocks, threads, 0, c->stream>>>(reinterpret_cast<T*>(c->out_d2.p), iters,
                                                                  T(1.0000001), T(1e-9));
        c->stats.gpu_launches++;
        B200_CUDA_CHECK(cudaEventRecord(e1, c->stream));
        B200_CUDA_CHECK(cudaEventSynchronize(e1));
        float ms = 0.f;
        B200_CUDA_CHECK(cudaEventElapsedTime(&ms, e0, e1));
        const double fl = 2.0 * double(blocks) * threads * double(iters) * CH;
        if(rep > 0) {   // first launch = warm-up
            best = fmax(best, fl / (double(ms) * 1e-3) / 1e12);
        }
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    *tflops = best;
    return B200VOR_OK;
}

}  // namespace

int microbench(Ctx* c, b200vor_peaks* out) {
    int rc;
    if((rc = measure<double>(c, &out->fp64_tflops)) || (rc = measure<float>(c, &out->fp32_tflops))) {
        return rc;
    }
    int sms = 0, khz = 0;
    B200_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, c->device));
    B200_CUDA_CHECK(cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, c->device));
    out->sm_count = uint32_t(sms);
    out->sm_clock_mhz = double(khz) * 1e-3;
    // one warp instruction per scheduler per cycle, 4 schedulers per SM
    out->issue_ginst_s = 4.0 * sms * double(khz) * 1e3 / 1e9;
    return B200VOR_OK;
}

}  // namespace b200vor
