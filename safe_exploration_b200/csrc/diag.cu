// Diagnostics: measured FP64 tensor-pipe (DMMA) rate, the roofline tri_sumsq is bounded by.
#include "segp_internal.cuh"

namespace segp {

__global__ void __launch_bounds__(256) dmma_peak_kernel(int iters, double* out) {
    double acc[32][2];
#pragma unroll
    for (int i = 0; i < 32; ++i) acc[i][0] = acc[i][1] = 0.0;
    double a = 1.0 + threadIdx.x * 1e-9, b = 1.0 - threadIdx.x * 1e-9;
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 2; ++r)
#pragma unroll
            for (int i = 0; i < 32; ++i)
                asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                             : "+d"(acc[i][0]), "+d"(acc[i][1])
                             : "d"(a), "d"(b));
    }
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < 32; ++i) s += acc[i][0] + acc[i][1];
    if (s == 123.456) out[0] = s;   // keep the accumulators alive
}

}  // namespace segp

extern "C" int segp_dmma_peak(int device, int iters, double* tflops) {
    using namespace segp;
    if (tflops == nullptr || iters < 1) {
        set_error("segp_dmma_peak: bad argument");
        return SEGP_ERR_INVALID;
    }
    int prev = 0;
    SEGP_CUDA_CHECK(cudaGetDevice(&prev));
    SEGP_CUDA_CHECK(cudaSetDevice(device));
    cudaDeviceProp prop;
    SEGP_CUDA_CHECK(cudaGetDeviceProperties(&prop, device));
    double* d_out = nullptr;
    SEGP_CUDA_CHECK(cudaMalloc(reinterpret_cast<void**>(&d_out), sizeof(double)));
    cudaEvent_t e0, e1;
    SEGP_CUDA_CHECK(cudaEventCreate(&e0));
    SEGP_CUDA_CHECK(cudaEventCreate(&e1));
    const int blocks = prop.multiProcessorCount * 4;
    dmma_peak_kernel<<<blocks, 256>>>(iters / 4 + 1, d_out);   // warm-up
    SEGP_CUDA_CHECK(cudaEventRecord(e0));
    dmma_peak_kernel<<<blocks, 256>>>(iters, d_out);
    SEGP_CUDA_CHECK(cudaEventRecord(e1));
    SEGP_CUDA_CHECK(cudaEventSynchronize(e1));
    float ms = 0.f;
    SEGP_CUDA_CHECK(cudaEventElapsedTime(&ms, e0, e1));
    // per warp and iteration: 64 DMMA x (8*8*4 MAC) x 2 flop
    const double flop = (double)blocks * 8.0 * iters * 64.0 * 512.0;
    *tflops = flop / (ms * 1e-3) / 1e12;
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(d_out);
    cudaSetDevice(prev);
    return SEGP_OK;
}
