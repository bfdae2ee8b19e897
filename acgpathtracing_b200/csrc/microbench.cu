// microbench.cu -- device microbenchmarks that supply the roofline denominators MEASURED_PEAKS.json does not
// carry (FP32 FFMA peak, L2 streaming bandwidth) plus an HBM copy for cross-checking its hbm_gbs.
// Timed with CUDA events, best of several repetitions, after a warm-up launch.
#include <cuda_runtime.h>
#include <stdint.h>

#include "pt_scene.h"

namespace ptb {

#define MB_TRY(expr)                        \
    do {                                    \
        cudaError_t e__ = (expr);           \
        if (e__ != cudaSuccess) return e__; \
    } while (0)

// 8 independent FFMA chains per thread; 2 flop per FFMA.
__global__ void __launch_bounds__(1024) ffma_kernel(float* out, int iters, float a, float b)
{
    float x0 = threadIdx.x, x1 = x0 + 1.f, x2 = x0 + 2.f, x3 = x0 + 3.f, x4 = x0 + 4.f, x5 = x0 + 5.f, x6 = x0 + 6.f, x7 = x0 + 7.f;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            x0 = __fmaf_rn(x0, a, b);
            x1 = __fmaf_rn(x1, a, b);
            x2 = __fmaf_rn(x2, a, b);
            x3 = __fmaf_rn(x3, a, b);
            x4 = __fmaf_rn(x4, a, b);
            x5 = __fmaf_rn(x5, a, b);
            x6 = __fmaf_rn(x6, a, b);
            x7 = __fmaf_rn(x7, a, b);
        }
    }
    const float s = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
    if (s == 123.456f) out[0] = s;   // never true in practice; keeps the chains alive
}

cudaError_t bench_fp32_peak(double* tflops)
{
    int dev = 0, sms = 0;
    MB_TRY(cudaGetDevice(&dev));
    MB_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    float* d_out = nullptr;
    MB_TRY(cudaMalloc((void**)&d_out, 4));
    cudaEvent_t e0, e1;
    MB_TRY(cudaEventCreate(&e0));
    MB_TRY(cudaEventCreate(&e1));
    const int iters = 4096, blocks = sms * 2 * 4, threads = 1024;
    double best = 0.0;
    for (int rep = 0; rep < 6; ++rep) {
        cudaEventRecord(e0);
        ffma_kernel<<<blocks, threads>>>(d_out, iters, 0.999f, 0.001f);
        cudaEventRecord(e1);
        MB_TRY(cudaEventSynchronize(e1));
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        const double flops = 2.0 * 64.0 * iters * (double)blocks * threads;
        const double tf = flops / (ms * 1e-3) / 1e12;
        if (rep > 0 && tf > best) best = tf;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(d_out);
    *tflops = best;
    return cudaGetLastError();
}

// grid-stride streaming read, 16 B per thread per step, of `n4` float4; `passes` sweeps
__global__ void __launch_bounds__(512) stream_read_kernel(const float4* __restrict__ src, size_t n4, int passes, float* out)
{
    float acc = 0.f;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (int p = 0; p < passes; ++p) {
        for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
            float4 v;
            asm volatile("ld.global.cg.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(src + i));
            acc += v.x + v.y + v.z + v.w;
        }
    }
    if (acc == 123.456f) out[0] = acc;
}

cudaError_t bench_l2_bandwidth(double* gbps)
{
    int dev = 0, sms = 0, l2 = 0;
    MB_TRY(cudaGetDevice(&dev));
    MB_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    MB_TRY(cudaDeviceGetAttribute(&l2, cudaDevAttrL2CacheSize, dev));
    size_t bytes = (size_t)l2 / 4;   // comfortably L2-resident (both dies)
    if (bytes < (8u << 20)) bytes = 8u << 20;
    bytes &= ~(size_t)4095;
    float4* d_src = nullptr;
    float* d_out = nullptr;
    MB_TRY(cudaMalloc((void**)&d_src, bytes));
    MB_TRY(cudaMalloc((void**)&d_out, 4));
    MB_TRY(cudaMemset(d_src, 0, bytes));
    cudaEvent_t e0, e1;
    MB_TRY(cudaEventCreate(&e0));
    MB_TRY(cudaEventCreate(&e1));
    const int passes = 40;
    double best = 0.0;
    for (int rep = 0; rep < 6; ++rep) {
        cudaEventRecord(e0);
        stream_read_kernel<<<sms * 4, 512>>>(d_src, bytes / 16, passes, d_out);
        cudaEventRecord(e1);
        MB_TRY(cudaEventSynchronize(e1));
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        const double gb = (double)bytes * passes / (ms * 1e-3) / 1e9;
        if (rep > 0 && gb > best) best = gb;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(d_src);
    cudaFree(d_out);
    *gbps = best;
    return cudaGetLastError();
}

__global__ void __launch_bounds__(512) stream_copy_kernel(const float4* __restrict__ src, float4* __restrict__ dst, size_t n4)
{
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) dst[i] = src[i];
}

cudaError_t bench_hbm_bandwidth(double* gbps)
{
    int dev = 0, sms = 0;
    MB_TRY(cudaGetDevice(&dev));
    MB_TRY(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const size_t bytes = (size_t)1 << 30;   // 1 GiB each way, far larger than L2
    float4 *d_src = nullptr, *d_dst = nullptr;
    MB_TRY(cudaMalloc((void**)&d_src, bytes));
    MB_TRY(cudaMalloc((void**)&d_dst, bytes));
    MB_TRY(cudaMemset(d_src, 1, bytes));
    cudaEvent_t e0, e1;
    MB_TRY(cudaEventCreate(&e0));
    MB_TRY(cudaEventCreate(&e1));
    double best = 0.0;
    for (int rep = 0; rep < 6; ++rep) {
        cudaEventRecord(e0);
        stream_copy_kernel<<<sms * 8, 512>>>(d_src, d_dst, bytes / 16);
        cudaEventRecord(e1);
        MB_TRY(cudaEventSynchronize(e1));
        float ms = 0.f;
        cudaEventElapsedTime(&ms, e0, e1);
        const double gb = 2.0 * (double)bytes / (ms * 1e-3) / 1e9;
        if (rep > 0 && gb > best) best = gb;
    }
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(d_src);
    cudaFree(d_dst);
    *gbps = best;
    return cudaGetLastError();
}

}  // namespace ptb
