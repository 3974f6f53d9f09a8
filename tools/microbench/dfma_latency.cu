// DFMA dependent-issue latency and per-SMSP throughput on the device under test.
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a dfma_latency.cu -o dfma_latency && ./dfma_latency
#include <cstdio>
#include <cuda_runtime.h>

template <int CHAINS>
__global__ void k(double *out, long long *cyc, int iters, double a, double b)
{
    double x[CHAINS];
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) x[c] = threadIdx.x * 1e-3 + c;
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 8; ++u)
#pragma unroll
            for (int c = 0; c < CHAINS; ++c) x[c] = fma(x[c], a, b);
    }
    long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int c = 0; c < CHAINS; ++c) s += x[c];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

template <int CHAINS>
void run(int warps)
{
    double *out; long long *cyc, h;
    cudaMalloc(&out, sizeof(double) * 4096); cudaMalloc(&cyc, 8);
    const int iters = 4096;
    k<CHAINS><<<1, 32 * warps>>>(out, cyc, iters, 0.999999, 1e-9);
    k<CHAINS><<<1, 32 * warps>>>(out, cyc, iters, 0.999999, 1e-9);
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    const double per = (double)h / (iters * 8.0 * CHAINS);
    printf("warps/CTA %2d (%.2f per SMSP) chains %d: %.2f cycles per DFMA per warp, %.2f DFMA/cycle/SMSP\n", warps,
           warps / 4.0, CHAINS, per, (warps / 4.0) / per);
    cudaFree(out); cudaFree(cyc);
}

int main()
{
    for (int w : {1, 4, 8, 12, 16, 32}) { run<1>(w); run<2>(w); run<3>(w); run<4>(w); run<8>(w); }
    return 0;
}
