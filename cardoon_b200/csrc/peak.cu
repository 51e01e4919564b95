// FP64 FMA throughput microbenchmark: the roofline denominator of the
// device-evaluation kernels (MEASURED_PEAKS.json has no FP64 figure).
#include "common.cuh"

__global__ void __launch_bounds__(256) fp64_peak_kernel(double* out, int iters) {
    double a0 = threadIdx.x * 1e-3, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3;
    double a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
    const double b = 1.0000001, c = 1e-9;
    for (int i = 0; i < iters; ++i) {
        a0 = fma(a0, b, c); a1 = fma(a1, b, c); a2 = fma(a2, b, c); a3 = fma(a3, b, c);
        a4 = fma(a4, b, c); a5 = fma(a5, b, c); a6 = fma(a6, b, c); a7 = fma(a7, b, c);
    }
    out[(long)blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
}

extern "C" int cdn_fp64_peak(cdn_ctx* ctx, double* flops_per_s) {
    if (!ctx || !flops_per_s) return CDN_ERR_ARG;
    CDN_USE_DEVICE(ctx);
    const int blocks = ctx->sm_count * 8, tpb = 256, iters = 4096;
    double* buf = nullptr;
    CDN_CUDA(ctx, cudaMalloc(&buf, sizeof(double) * blocks * tpb));
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    fp64_peak_kernel<<<blocks, tpb>>>(buf, iters);     // warm-up
    double best = 0.0;
    for (int rep = 0; rep < 5; ++rep) {
        cudaEventRecord(e0);
        fp64_peak_kernel<<<blocks, tpb>>>(buf, iters);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        const double fl = 2.0 * 8.0 * iters * (double)blocks * tpb / (ms * 1e-3);
        if (fl > best) best = fl;
    }
    cudaEventDestroy(e0); cudaEventDestroy(e1);
    cudaError_t e = cudaGetLastError();
    cudaFree(buf);
    if (e != cudaSuccess) return cdn_fail(ctx, CDN_ERR_CUDA, cudaGetErrorString(e));
    *flops_per_s = best;
    return CDN_OK;
}
