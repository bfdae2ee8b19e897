// fetchbench.cu -- how should a warp fetch 32 DIFFERENT 64-byte (or 128-byte) records?  (experiment behind the
// node-fetch design of wavefront.cu; build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fetchbench fetchbench.cu)
//   A  every lane loads its own record with LDG.128 x4 (x8)            -- 32 lines per instruction
//   B  4 (8) adjacent lanes load one record, 16 B each; the warp transposes through shared memory
//   A' as A with LDG.E.256 x2 (x4): 32 bytes per lane and instruction (sm_100+)
//   C  every lane loads its own record with tex1Dfetch<float4> x4 (x8) -- texture path
// Each record's first word decides (with a hash) the next record: a dependent chain as in a BVH traversal.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

__device__ __forceinline__ uint32_t mix(uint32_t a, uint32_t b) { a ^= b * 0x9e3779b9u; a ^= a >> 15; a *= 0x85ebca6bu; a ^= a >> 13; return a; }

template <int Q>   // quads per record (4 = 64 B, 8 = 128 B)
__global__ void __launch_bounds__(768, 1) k_direct(const float4* __restrict__ rec, uint32_t n_rec, int iters, uint32_t* out)
{
    uint32_t idx = mix(blockIdx.x * blockDim.x + threadIdx.x, 12345u) % n_rec;
    float acc = 0.f;
    for (int it = 0; it < iters; ++it) {
        float4 q[Q];
#pragma unroll
        for (int k = 0; k < Q; ++k) q[k] = __ldg(rec + (size_t)idx * Q + k);
#pragma unroll
        for (int k = 0; k < Q; ++k) acc += q[k].y + q[k].z + q[k].w;
        idx = mix(idx, __float_as_uint(q[0].x)) % n_rec;
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = idx + (uint32_t)acc;
}

struct alignas(32) F8 { float4 a, b; };
__device__ __forceinline__ F8 ldg256(const void* p)      // LDG.E.256 (sm_100+): one 32-byte sector per lane and instruction
{
    F8 r;
    asm volatile("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=f"(r.a.x), "=f"(r.a.y), "=f"(r.a.z), "=f"(r.a.w), "=f"(r.b.x), "=f"(r.b.y), "=f"(r.b.z), "=f"(r.b.w)
                 : "l"(p));
    return r;
}

template <int Q>
__global__ void __launch_bounds__(768, 1) k_direct256(const float4* __restrict__ rec, uint32_t n_rec, int iters, uint32_t* out)
{
    uint32_t idx = mix(blockIdx.x * blockDim.x + threadIdx.x, 12345u) % n_rec;
    float acc = 0.f;
    for (int it = 0; it < iters; ++it) {
        F8 q[Q / 2];
#pragma unroll
        for (int k = 0; k < Q / 2; ++k) q[k] = ldg256(rec + (size_t)idx * Q + 2 * k);
#pragma unroll
        for (int k = 0; k < Q / 2; ++k) acc += q[k].a.y + q[k].a.z + q[k].a.w + q[k].b.y + q[k].b.z + q[k].b.w;
        idx = mix(idx, __float_as_uint(q[0].a.x)) % n_rec;
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = idx + (uint32_t)acc;
}

template <int Q>
__global__ void __launch_bounds__(768, 1) k_tex(cudaTextureObject_t tex, uint32_t n_rec, int iters, uint32_t* out)
{
    uint32_t idx = mix(blockIdx.x * blockDim.x + threadIdx.x, 12345u) % n_rec;
    float acc = 0.f;
    for (int it = 0; it < iters; ++it) {
        float4 q[Q];
#pragma unroll
        for (int k = 0; k < Q; ++k) q[k] = tex1Dfetch<float4>(tex, (int)(idx * Q + k));
#pragma unroll
        for (int k = 0; k < Q; ++k) acc += q[k].y + q[k].z + q[k].w;
        idx = mix(idx, __float_as_uint(q[0].x)) % n_rec;
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = idx + (uint32_t)acc;
}

// cooperative: Q lanes per record; instruction k fetches records (32/Q)*k .. +32/Q-1 of the warp
template <int Q>
__global__ void __launch_bounds__(768, 1) k_coop(const float4* __restrict__ rec, uint32_t n_rec, int iters, uint32_t* out)
{
    extern __shared__ float4 stage[];                      // per warp: Q * 32 float4
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    float4* st = stage + warp * (Q * 32);
    uint32_t idx = mix(blockIdx.x * blockDim.x + threadIdx.x, 12345u) % n_rec;
    float acc = 0.f;
    constexpr int RPI = 32 / Q;                            // records per instruction
    const int sub = lane % Q, grp = lane / Q;
    for (int it = 0; it < iters; ++it) {
        float4 q[Q];
#pragma unroll
        for (int k = 0; k < Q; ++k) {
            const int r = RPI * k + grp;                   // record (= owner lane) this lane helps to fetch
            const uint32_t ridx = __shfl_sync(0xffffffffu, idx, r);
            q[k] = __ldg(rec + (size_t)ridx * Q + sub);
        }
#pragma unroll
        for (int k = 0; k < Q; ++k) {
            const int r = RPI * k + grp;
            st[sub * 32 + ((r + 2 * sub) & 31)] = q[k];    // swizzle: conflict-free stores and loads
        }
        __syncwarp();
#pragma unroll
        for (int k = 0; k < Q; ++k) q[k] = st[k * 32 + ((lane + 2 * k) & 31)];
        __syncwarp();
#pragma unroll
        for (int k = 0; k < Q; ++k) acc += q[k].y + q[k].z + q[k].w;
        idx = mix(idx, __float_as_uint(q[0].x)) % n_rec;
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = idx + (uint32_t)acc;
}

template <class F>
static float time_ms(F launch)
{
    cudaEvent_t a, b;
    CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
    launch();
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(a));
    launch();
    CK(cudaEventRecord(b));
    CK(cudaEventSynchronize(b));
    float ms; CK(cudaEventElapsedTime(&ms, a, b));
    return ms;
}

template <int Q>
static void run(const char* label, size_t bytes, int block, int iters)
{
    const uint32_t n_rec = (uint32_t)(bytes / (16 * Q));
    float4* rec; uint32_t* out;
    CK(cudaMalloc(&rec, (size_t)n_rec * Q * 16));
    CK(cudaMalloc(&out, 148 * 768 * 4));
    float4* h = (float4*)malloc((size_t)n_rec * Q * 16);
    for (size_t i = 0; i < (size_t)n_rec * Q; ++i) h[i] = make_float4((float)(i * 2654435761u % 1000003u), 1.f, 2.f, 3.f);
    CK(cudaMemcpy(rec, h, (size_t)n_rec * Q * 16, cudaMemcpyHostToDevice));
    free(h);
    cudaResourceDesc rd = {}; rd.resType = cudaResourceTypeLinear; rd.res.linear.devPtr = rec;
    rd.res.linear.desc = cudaCreateChannelDesc<float4>(); rd.res.linear.sizeInBytes = (size_t)n_rec * Q * 16;
    cudaTextureDesc td = {}; td.readMode = cudaReadModeElementType;
    cudaTextureObject_t tex = 0;
    const bool have_tex = (size_t)n_rec * Q <= (1u << 27) && cudaCreateTextureObject(&tex, &rd, &td, nullptr) == cudaSuccess;
    const size_t pad = (Q == 8 ? 100 : 150) * 1024;                         // occupy shared memory like the render kernel does (small L1)
    CK(cudaFuncSetAttribute(k_direct<Q>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pad));
    CK(cudaFuncSetAttribute(k_direct256<Q>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pad));
    CK(cudaFuncSetAttribute(k_tex<Q>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pad));
    CK(cudaFuncSetAttribute(k_coop<Q>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)(pad + 24 * Q * 32 * 16)));
    const double fetches = 148.0 * block * iters;
    float a = time_ms([&] { k_direct<Q><<<148, block, pad>>>(rec, n_rec, iters, out); });
    float d = time_ms([&] { k_direct256<Q><<<148, block, pad>>>(rec, n_rec, iters, out); });
    float b = time_ms([&] { k_coop<Q><<<148, block, pad + (block / 32) * Q * 32 * 16>>>(rec, n_rec, iters, out); });
    float c = have_tex ? time_ms([&] { k_tex<Q><<<148, block, pad>>>(tex, n_rec, iters, out); }) : 0.f;
    CK(cudaGetLastError());
    printf("%-22s rec %3d B  set %8.1f MB  block %3d | direct %7.2f G rec/s | LDG.256 %7.2f | coop+smem %7.2f | tex %7.2f\n", label, 16 * Q,
           bytes / 1e6, block, fetches / a / 1e6, fetches / d / 1e6, fetches / b / 1e6, have_tex ? fetches / c / 1e6 : 0.0);
    if (have_tex) cudaDestroyTextureObject(tex);
    cudaFree(rec); cudaFree(out);
}

int main()
{
    for (int block : {768, 384}) {
        run<4>("L1-resident", 24 << 10, block, 4000);
        run<4>("L2-resident (332 KB)", 332 << 10, block, 4000);
        run<4>("L2-resident (32 MB)", 32 << 20, block, 2000);
        run<4>("DRAM (512 MB)", 512 << 20, block, 1000);
        run<8>("L1-resident", 24 << 10, block, 4000);
        run<8>("L2-resident (332 KB)", 332 << 10, block, 4000);
        run<8>("L2-resident (32 MB)", 32 << 20, block, 2000);
        run<8>("DRAM (512 MB)", 512 << 20, block, 1000);
    }
    return 0;
}
