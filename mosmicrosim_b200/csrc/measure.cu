// measure.cu -- roofline denominators measured on the device the benchmark
// runs on: FP64 FMA throughput (not in MEASURED_PEAKS.json, so bench.py takes it
// live) and a device-to-device copy as a cross-check of the HBM figure.
#include <algorithm>
#include <cstdint>
#include <cstdio>

#include <cuda_runtime.h>

#include "../../include/mmsgpu.h"

namespace {

constexpr int kChains = 8;

__global__ void __launch_bounds__(256) dfma_kernel(double *out, int iters, double a, double b)
{
    double x[kChains];
#pragma unroll
    for (int k = 0; k < kChains; ++k) x[k] = (double)(threadIdx.x + k) * 1e-3;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int k = 0; k < kChains; ++k) x[k] = fma(x[k], a, b);
    }
    double s = 0.0;
#pragma unroll
    for (int k = 0; k < kChains; ++k) s += x[k];
    if (s == 123.456) out[blockIdx.x * blockDim.x + threadIdx.x] = s;   // keep the chains alive
}

void *g_flush[64] = {nullptr};
constexpr size_t kFlushBytes = (size_t)256 << 20;   // > 126 MB L2

}  // namespace

extern "C" {

int mms_measure_fp64_peak(int32_t device, double *tflops)
{
    if (!tflops) return MMS_ERR_INVALID_ARG;
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n < 1) return MMS_ERR_NO_DEVICE;
    if (device < 0 || device >= n) return MMS_ERR_INVALID_ARG;
    if (cudaSetDevice(device) != cudaSuccess) return MMS_ERR_CUDA;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return MMS_ERR_CUDA;
    const int blocks = prop.multiProcessorCount * 8, threads = 256, iters = 1 << 16;
    double *out = nullptr;
    if (cudaMalloc(&out, sizeof(double) * blocks * threads) != cudaSuccess) return MMS_ERR_ALLOC;
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    double best = 0.0;
    for (int rep = 0; rep < 4; ++rep) {          // rep 0 is warm-up
        cudaEventRecord(e0);
        dfma_kernel<<<blocks, threads>>>(out, iters, 0.999999, 1e-9);
        cudaEventRecord(e1);
        if (cudaEventSynchronize(e1) != cudaSuccess) { cudaFree(out); return MMS_ERR_CUDA; }
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        const double flop = 2.0 * kChains * (double)iters * blocks * threads;
        if (rep > 0) best = std::max(best, flop / (ms * 1e-3) / 1e12);
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(out);
    *tflops = best;
    return MMS_OK;
}

int mms_measure_hbm_copy(int32_t device, double *gbs)
{
    if (!gbs) return MMS_ERR_INVALID_ARG;
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n < 1) return MMS_ERR_NO_DEVICE;
    if (device < 0 || device >= n) return MMS_ERR_INVALID_ARG;
    if (cudaSetDevice(device) != cudaSuccess) return MMS_ERR_CUDA;
    const size_t bytes = (size_t)1 << 30;
    void *a = nullptr, *b = nullptr;
    if (cudaMalloc(&a, bytes) != cudaSuccess) return MMS_ERR_ALLOC;
    if (cudaMalloc(&b, bytes) != cudaSuccess) { cudaFree(a); return MMS_ERR_ALLOC; }
    cudaMemset(a, 1, bytes);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    double best = 0.0;
    for (int rep = 0; rep < 6; ++rep) {
        cudaEventRecord(e0);
        cudaMemcpyAsync(b, a, bytes, cudaMemcpyDeviceToDevice);
        cudaEventRecord(e1);
        if (cudaEventSynchronize(e1) != cudaSuccess) { cudaFree(a); cudaFree(b); return MMS_ERR_CUDA; }
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        if (rep > 0) best = std::max(best, 2.0 * bytes / (ms * 1e-3) / 1e9);
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(a);
    cudaFree(b);
    *gbs = best;
    return MMS_OK;
}

// SM count and the maximum SM clock of a device: the nominal FP64 peak is n_sm x 64 FMA/clk x 2 x clock
int mms_device_info(int32_t device, int32_t *n_sm, int32_t *sm_clock_khz_max, uint64_t *memory_bytes)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n < 1) return MMS_ERR_NO_DEVICE;
    if (device < 0 || device >= n) return MMS_ERR_INVALID_ARG;
    int v = 0;
    if (n_sm) { if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, device) != cudaSuccess) return MMS_ERR_CUDA; *n_sm = v; }
    if (sm_clock_khz_max) { if (cudaDeviceGetAttribute(&v, cudaDevAttrClockRate, device) != cudaSuccess) return MMS_ERR_CUDA; *sm_clock_khz_max = v; }
    if (memory_bytes) {
        cudaDeviceProp prop;
        if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return MMS_ERR_CUDA;
        *memory_bytes = (uint64_t)prop.totalGlobalMem;
    }
    return MMS_OK;
}

int mms_flush_l2(int32_t device)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n < 1) return MMS_ERR_NO_DEVICE;
    if (device < 0 || device >= n || device >= 64) return MMS_ERR_INVALID_ARG;
    if (cudaSetDevice(device) != cudaSuccess) return MMS_ERR_CUDA;
    if (!g_flush[device] && cudaMalloc(&g_flush[device], kFlushBytes) != cudaSuccess) return MMS_ERR_ALLOC;
    if (cudaMemset(g_flush[device], 0x5a, kFlushBytes) != cudaSuccess) return MMS_ERR_CUDA;
    if (cudaDeviceSynchronize() != cudaSuccess) return MMS_ERR_CUDA;
    return MMS_OK;
}

}  // extern "C"
