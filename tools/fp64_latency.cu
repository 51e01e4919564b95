// Microbenchmark (run on the GPU box): latency of dependent FP64 operations as
// one warp sees them, and how much independent work per warp hides it.  Numbers
// feed the latency models in DESIGN.md (cluster kernel, Monte-Carlo tile kernel).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/fp64_latency tools/fp64_latency.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int ILP>
__global__ void dfma_chain(double* out, long long* cyc, int n, double a, double b) {
    double x[ILP];
#pragma unroll
    for (int j = 0; j < ILP; ++j) x[j] = threadIdx.x * 1e-3 + j;
    const long long t0 = clock64();
    for (int i = 0; i < n; ++i) {
#pragma unroll
        for (int j = 0; j < ILP; ++j) x[j] = fma(x[j], a, b);
    }
    const long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int j = 0; j < ILP; ++j) s += x[j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

__global__ void rcp_chain(double* out, long long* cyc, int n) {
    double x = 1.5 + threadIdx.x * 1e-3;
    const long long t0 = clock64();
    for (int i = 0; i < n; ++i) x = 1.0 / x + 0.25;
    const long long t1 = clock64();
    out[threadIdx.x] = x;
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
}

__global__ void lds_chain(int* out, long long* cyc, int n) {
    __shared__ int idx[1024];
    for (int k = threadIdx.x; k < 1024; k += blockDim.x) idx[k] = (k * 17 + 5) & 1023;
    __syncthreads();
    int p = threadIdx.x;
    const long long t0 = clock64();
    for (int i = 0; i < n; ++i) p = idx[p];
    const long long t1 = clock64();
    out[threadIdx.x] = p;
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
}

template <int ILP>
static void run(int warps, int blocks, double* out, long long* cyc) {
    const int n = 4096;
    dfma_chain<ILP><<<blocks, 32 * warps>>>(out, cyc, n, 0.999, 1e-3);
    cudaDeviceSynchronize();
    dfma_chain<ILP><<<blocks, 32 * warps>>>(out, cyc, n, 0.999, 1e-3);
    cudaDeviceSynchronize();
    long long c;
    cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    printf("DFMA  warps/CTA %2d  ILP %2d : %6.2f cycles per dependent step, %6.2f cycles per warp instruction\n",
           warps, ILP, (double)c / n, (double)c / n / ILP);
}

int main() {
    double* out; long long* cyc; int* iout;
    cudaMalloc(&out, 1 << 22); cudaMalloc(&cyc, 1 << 16); cudaMalloc(&iout, 1 << 16);
    for (int w : {1, 2, 4, 8, 16}) { run<1>(w, 1, out, cyc); }
    run<2>(1, 1, out, cyc); run<4>(1, 1, out, cyc); run<8>(1, 1, out, cyc); run<16>(1, 1, out, cyc);
    run<4>(4, 1, out, cyc); run<4>(8, 1, out, cyc); run<8>(8, 1, out, cyc);
    long long c;
    rcp_chain<<<1, 32>>>(out, cyc, 2048); cudaDeviceSynchronize();
    rcp_chain<<<1, 32>>>(out, cyc, 2048); cudaDeviceSynchronize();
    cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    printf("1.0/x + c (IEEE division + DADD), one warp: %6.2f cycles per step\n", (double)c / 2048);
    lds_chain<<<1, 32>>>(iout, cyc, 4096); cudaDeviceSynchronize();
    lds_chain<<<1, 32>>>(iout, cyc, 4096); cudaDeviceSynchronize();
    cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    printf("dependent LDS (pointer chase), one warp: %6.2f cycles per load\n", (double)c / 4096);
    return 0;
}
