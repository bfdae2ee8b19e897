// lbvh.cu -- deterministic LBVH build on the GPU; replaces optixAccelBuild + optixAccelCompact
// (PathTracerMain.cpp:311-398), whose algorithm is closed.
//
// Pipeline (DESIGN.md "LBVH spec"; CPU statement of the same arithmetic: oracle/oracle_geom.h):
//   tri_setup   per triangle: AABB, (v0,e1,e2) record, geometric normal; scene bounds by atomic min/max
//               on order-preserving integer images of the floats (order-independent => deterministic)
//   morton      30-bit Morton code of the AABB centre, normalised by the scene bounds
//   radix sort  4 passes x 8 bits, LSD, STABLE (histogram -> exclusive scan -> ranked scatter with
//               warp match/ballot ranks), values = primitive ids.  Input is in primitive order, so a stable
//               sort by Morton code equals sorting the unique 64-bit keys morton<<32|prim: bit-exact.
//   karras      binary radix tree over the sorted unique keys (Karras 2012) -- a pure function of the keys
//   refit       bottom-up AABB union with per-node arrival counters (fmin/fmax are order-independent)
//   pack        64-byte two-child nodes + 48-byte triangle records in leaf order
#include <cuda_runtime.h>
#include <math.h>
#include <float.h>
#include <stdint.h>

#include <algorithm>
#include <chrono>
#include <vector>

#include "pt_device.cuh"
#include "pt_scene.h"

namespace ptb {

using namespace ptd;

// ---- order-preserving float <-> uint image ---------------------------------------------------------------
__device__ __forceinline__ uint32_t float_to_ordered(float f)
{
    uint32_t u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float ordered_to_float(uint32_t u)
{
    return __uint_as_float((u & 0x80000000u) ? (u & 0x7fffffffu) : ~u);
}

// bounds[0..2] = min (ordered uint), bounds[3..5] = max
__global__ void init_bounds_kernel(uint32_t* bounds)
{
    if (threadIdx.x < 3) bounds[threadIdx.x] = 0xffffffffu;
    else if (threadIdx.x < 6) bounds[threadIdx.x] = 0u;
}

__global__ void __launch_bounds__(256) tri_setup_kernel(const float4* __restrict__ verts, const uint32_t* __restrict__ indices,
                                                        const uint32_t* __restrict__ mat_ids, uint32_t n_tris,
                                                        float4* __restrict__ tri_in,      // 3 per primitive
                                                        float4* __restrict__ tri_bb,      // 2 per primitive (lo, hi)
                                                        float4* __restrict__ tri_shade,   // (N_0, material)
                                                        uint32_t* __restrict__ bounds)
{
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    float3 lo = f3(FLT_MAX), hi = f3(-FLT_MAX);
    if (k < n_tris) {
        const float4 a = verts[indices[3 * k + 0]];
        const float4 b = verts[indices[3 * k + 1]];
        const float4 c = verts[indices[3 * k + 2]];
        const float3 v0 = f3(a.x, a.y, a.z), v1 = f3(b.x, b.y, b.z), v2 = f3(c.x, c.y, c.z);
        const float3 e1 = v1 - v0, e2 = v2 - v0;
        lo = f3(fminf(fminf(v0.x, v1.x), v2.x), fminf(fminf(v0.y, v1.y), v2.y), fminf(fminf(v0.z, v1.z), v2.z));
        hi = f3(fmaxf(fmaxf(v0.x, v1.x), v2.x), fmaxf(fmaxf(v0.y, v1.y), v2.y), fmaxf(fmaxf(v0.z, v1.z), v2.z));
        const uint32_t mat = mat_ids[k];
        tri_in[3 * k + 0] = make_float4(v0.x, v0.y, v0.z, __uint_as_float(k));
        tri_in[3 * k + 1] = make_float4(e1.x, e1.y, e1.z, __uint_as_float(mat));
        tri_in[3 * k + 2] = make_float4(e2.x, e2.y, e2.z, 0.0f);
        tri_bb[2 * k + 0] = make_float4(lo.x, lo.y, lo.z, 0.0f);
        tri_bb[2 * k + 1] = make_float4(hi.x, hi.y, hi.z, 0.0f);
        // pathTracerPrograms.cu:891  N_0 = normalize(cross(v1 - v0, v2 - v0))
        const float3 n0 = normalize(cross(e1, e2));
        tri_shade[k] = make_float4(n0.x, n0.y, n0.z, __uint_as_float(mat));
    }
    // block reduction of the bounds (min/max: any order gives the same bits)
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        lo.x = fminf(lo.x, __shfl_xor_sync(0xffffffffu, lo.x, off));
        lo.y = fminf(lo.y, __shfl_xor_sync(0xffffffffu, lo.y, off));
        lo.z = fminf(lo.z, __shfl_xor_sync(0xffffffffu, lo.z, off));
        hi.x = fmaxf(hi.x, __shfl_xor_sync(0xffffffffu, hi.x, off));
        hi.y = fmaxf(hi.y, __shfl_xor_sync(0xffffffffu, hi.y, off));
        hi.z = fmaxf(hi.z, __shfl_xor_sync(0xffffffffu, hi.z, off));
    }
    if ((threadIdx.x & 31) == 0) {
        atomicMin(&bounds[0], float_to_ordered(lo.x));
        atomicMin(&bounds[1], float_to_ordered(lo.y));
        atomicMin(&bounds[2], float_to_ordered(lo.z));
        atomicMax(&bounds[3], float_to_ordered(hi.x));
        atomicMax(&bounds[4], float_to_ordered(hi.y));
        atomicMax(&bounds[5], float_to_ordered(hi.z));
    }
}

__device__ __forceinline__ uint32_t expand_bits10(uint32_t v)
{
    v &= 0x3ffu;
    v = (v | (v << 16)) & 0x030000ffu;
    v = (v | (v << 8)) & 0x0300f00fu;
    v = (v | (v << 4)) & 0x030c30c3u;
    v = (v | (v << 2)) & 0x09249249u;
    return v;
}

__global__ void __launch_bounds__(256) morton_kernel(const float4* __restrict__ tri_bb, uint32_t n_tris,
                                                     const uint32_t* __restrict__ bounds, uint32_t* __restrict__ keys,
                                                     uint32_t* __restrict__ vals, float* __restrict__ bounds_f)
{
    const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
    const float3 slo = f3(ordered_to_float(bounds[0]), ordered_to_float(bounds[1]), ordered_to_float(bounds[2]));
    const float3 shi = f3(ordered_to_float(bounds[3]), ordered_to_float(bounds[4]), ordered_to_float(bounds[5]));
    if (k == 0) {
        bounds_f[0] = slo.x; bounds_f[1] = slo.y; bounds_f[2] = slo.z;
        bounds_f[3] = shi.x; bounds_f[4] = shi.y; bounds_f[5] = shi.z;
    }
    if (k >= n_tris) return;
    const float3 e = shi - slo;
    const float3 inv = f3(e.x > 0.0f ? 1.0f / e.x : 0.0f, e.y > 0.0f ? 1.0f / e.y : 0.0f, e.z > 0.0f ? 1.0f / e.z : 0.0f);
    const float4 lo = tri_bb[2 * k + 0];
    const float4 hi = tri_bb[2 * k + 1];
    const float cx = (lo.x + hi.x) * 0.5f, cy = (lo.y + hi.y) * 0.5f, cz = (lo.z + hi.z) * 0.5f;
    const float nx = (cx - slo.x) * inv.x, ny = (cy - slo.y) * inv.y, nz = (cz - slo.z) * inv.z;
    const uint32_t qx = (uint32_t)fminf(fmaxf(nx * 1024.0f, 0.0f), 1023.0f);
    const uint32_t qy = (uint32_t)fminf(fmaxf(ny * 1024.0f, 0.0f), 1023.0f);
    const uint32_t qz = (uint32_t)fminf(fmaxf(nz * 1024.0f, 0.0f), 1023.0f);
    keys[k] = (expand_bits10(qx) << 2) | (expand_bits10(qy) << 1) | expand_bits10(qz);
    vals[k] = k;
}

// ---- stable LSD radix sort, 8-bit digits ---------------------------------------------------------------------
constexpr int kSortThreads = 256;
constexpr int kSortItems = 8;                               // per thread
constexpr int kSortTile = kSortThreads * kSortItems;        // 2048 keys per block
constexpr int kSortWarps = kSortThreads / 32;

// block_hist[digit * n_blocks + block]
__global__ void __launch_bounds__(kSortThreads) sort_hist_kernel(const uint32_t* __restrict__ keys, uint32_t n, int shift,
                                                                 uint32_t* __restrict__ block_hist, uint32_t n_blocks)
{
    __shared__ uint32_t hist[256];
    hist[threadIdx.x] = 0;
    __syncthreads();
    const uint32_t base = blockIdx.x * kSortTile;
#pragma unroll
    for (int j = 0; j < kSortItems; ++j) {
        const uint32_t i = base + j * kSortThreads + threadIdx.x;
        if (i < n) atomicAdd(&hist[(keys[i] >> shift) & 255u], 1u);
    }
    __syncthreads();
    block_hist[threadIdx.x * n_blocks + blockIdx.x] = hist[threadIdx.x];
}

// exclusive scan of `count` uint32 in place, single block of 1024 threads
__global__ void __launch_bounds__(1024) scan_kernel(uint32_t* __restrict__ data, uint32_t count)
{
    __shared__ uint32_t warp_sums[32];
    __shared__ uint32_t carry;
    if (threadIdx.x == 0) carry = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (uint32_t base = 0; base < count; base += 1024) {
        const uint32_t i = base + threadIdx.x;
        const uint32_t v = i < count ? data[i] : 0u;
        uint32_t x = v;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            const uint32_t y = __shfl_up_sync(0xffffffffu, x, off);
            if (lane >= off) x += y;
        }
        if (lane == 31) warp_sums[warp] = x;
        __syncthreads();
        if (warp == 0) {
            uint32_t s = warp_sums[lane];
#pragma unroll
            for (int off = 1; off < 32; off <<= 1) {
                const uint32_t y = __shfl_up_sync(0xffffffffu, s, off);
                if (lane >= off) s += y;
            }
            warp_sums[lane] = s;   // inclusive
        }
        __syncthreads();
        const uint32_t warp_off = warp > 0 ? warp_sums[warp - 1] : 0u;
        const uint32_t incl = x + warp_off + carry;
        if (i < count) data[i] = incl - v;
        __syncthreads();
        if (threadIdx.x == 1023) carry = incl;
        __syncthreads();
    }
}

// ---- multi-block exclusive scan ------------------------------------------------------------------------------------
// scan_kernel above walks the whole array with one CTA: fine for a few thousand entries, 2 ms for the 2 M flags of a
// PLOC round.  For larger arrays: every CTA scans a tile of 4096 entries in place and writes the tile's total, the
// totals are scanned (recursively), and a third kernel adds each tile's offset.  Integer sums: deterministic.
constexpr uint32_t kScanTile = 4096;   // 1024 threads x 4 entries

__global__ void __launch_bounds__(1024) scan_tiles_kernel(uint32_t* __restrict__ data, uint32_t count, uint32_t* __restrict__ tile_sums)
{
    __shared__ uint32_t warp_sums[32];
    const uint32_t base = blockIdx.x * kScanTile + threadIdx.x * 4u;
    uint32_t v[4], sum = 0u;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        v[k] = base + k < count ? data[base + k] : 0u;
        sum += v[k];
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t x = sum;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        const uint32_t y = __shfl_up_sync(0xffffffffu, x, off);
        if (lane >= off) x += y;
    }
    if (lane == 31) warp_sums[warp] = x;
    __syncthreads();
    if (warp == 0) {
        uint32_t t = warp_sums[lane];
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            const uint32_t y = __shfl_up_sync(0xffffffffu, t, off);
            if (lane >= off) t += y;
        }
        warp_sums[lane] = t;   // inclusive
    }
    __syncthreads();
    uint32_t run = x - sum + (warp > 0 ? warp_sums[warp - 1] : 0u);   // exclusive prefix of this thread's four entries
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        if (base + k < count) data[base + k] = run;
        run += v[k];
    }
    if (threadIdx.x == 1023) tile_sums[blockIdx.x] = warp_sums[31];
}

__global__ void __launch_bounds__(1024) scan_add_kernel(uint32_t* __restrict__ data, uint32_t count, const uint32_t* __restrict__ tile_offsets)
{
    const uint32_t off = tile_offsets[blockIdx.x];
    const uint32_t base = blockIdx.x * kScanTile + threadIdx.x * 4u;
#pragma unroll
    for (int k = 0; k < 4; ++k)
        if (base + k < count) data[base + k] += off;
}

// scratch: at least scan_scratch_entries(count) entries
static inline size_t scan_scratch_entries(size_t count)
{
    size_t total = 0;
    while (count > kScanTile) {
        count = (count + kScanTile - 1) / kScanTile;
        total += count;
    }
    return total + 1;
}

static void scan_exclusive(uint32_t* data, uint32_t count, uint32_t* scratch, cudaStream_t stream)
{
    if (count <= kScanTile) {
        scan_kernel<<<1, 1024, 0, stream>>>(data, count);
        return;
    }
    const uint32_t tiles = (count + kScanTile - 1) / kScanTile;
    scan_tiles_kernel<<<tiles, 1024, 0, stream>>>(data, count, scratch);
    scan_exclusive(scratch, tiles, scratch + tiles, stream);
    scan_add_kernel<<<tiles, 1024, 0, stream>>>(data, count, scratch);
}

// Ranked, stable scatter.  Warp w of a block owns the contiguous chunk [w*256, (w+1)*256) of the block's
// tile and walks it in rounds of 32 consecutive keys, so (block, warp, round, lane) is the input order.
__global__ void __launch_bounds__(kSortThreads) sort_scatter_kernel(const uint32_t* __restrict__ keys_in,
                                                                    const uint32_t* __restrict__ vals_in,
                                                                    uint32_t* __restrict__ keys_out,
                                                                    uint32_t* __restrict__ vals_out, uint32_t n, int shift,
                                                                    const uint32_t* __restrict__ block_base, uint32_t n_blocks)
{
    __shared__ uint32_t warp_count[kSortWarps][256];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int d = threadIdx.x; d < kSortWarps * 256; d += kSortThreads) (&warp_count[0][0])[d] = 0;
    __syncthreads();
    const uint32_t chunk = blockIdx.x * kSortTile + warp * (32 * kSortItems);
    uint32_t key[kSortItems], val[kSortItems], rank[kSortItems];
    const uint32_t lt_mask = (1u << lane) - 1u;
#pragma unroll
    for (int j = 0; j < kSortItems; ++j) {
        const uint32_t i = chunk + j * 32 + lane;
        const bool valid = i < n;
        key[j] = valid ? keys_in[i] : 0xffffffffu;
        val[j] = valid ? vals_in[i] : 0u;
        const uint32_t digit = (key[j] >> shift) & 255u;
        // peers: valid lanes with the same digit (invalid lanes form their own group via bit 8)
        const uint32_t peers = __match_any_sync(0xffffffffu, valid ? digit : 256u);
        const uint32_t before = __popc(peers & lt_mask);
        uint32_t base = 0;
        if (valid) base = warp_count[warp][digit];
        __syncwarp();
        if (valid && before == 0) warp_count[warp][digit] = base + __popc(peers);
        __syncwarp();
        rank[j] = base + before;
    }
    __syncthreads();
    // exclusive prefix over warps for each digit (thread d handles digit d), then add the global base
    {
        const int d = threadIdx.x;
        uint32_t run = block_base[d * n_blocks + blockIdx.x];
#pragma unroll
        for (int w = 0; w < kSortWarps; ++w) {
            const uint32_t c = warp_count[w][d];
            warp_count[w][d] = run;
            run += c;
        }
    }
    __syncthreads();
#pragma unroll
    for (int j = 0; j < kSortItems; ++j) {
        const uint32_t i = chunk + j * 32 + lane;
        if (i < n) {
            const uint32_t digit = (key[j] >> shift) & 255u;
            const uint32_t pos = warp_count[warp][digit] + rank[j];
            keys_out[pos] = key[j];
            vals_out[pos] = val[j];
        }
    }
}

// ---- Karras 2012 radix tree -------------------------------------------------------------------------------------
__device__ __forceinline__ int delta64(const uint32_t* __restrict__ mort, const uint32_t* __restrict__ prim, int n, int i, int j)
{
    if (j < 0 || j >= n) return -1;
    const unsigned long long a = ((unsigned long long)mort[i] << 32) | prim[i];
    const unsigned long long b = ((unsigned long long)mort[j] << 32) | prim[j];
    return __clzll((long long)(a ^ b));
}

__global__ void __launch_bounds__(256) karras_kernel(const uint32_t* __restrict__ mort, const uint32_t* __restrict__ prim, int n,
                                                     int* __restrict__ left, int* __restrict__ right,
                                                     int* __restrict__ node_parent, int* __restrict__ leaf_parent)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n - 1) return;
    const int d = (delta64(mort, prim, n, i, i + 1) - delta64(mort, prim, n, i, i - 1)) >= 0 ? 1 : -1;
    const int dmin = delta64(mort, prim, n, i, i - d);
    int lmax = 2;
    while (delta64(mort, prim, n, i, i + lmax * d) > dmin) lmax *= 2;
    int l = 0;
    for (int t = lmax / 2; t >= 1; t /= 2)
        if (delta64(mort, prim, n, i, i + (l + t) * d) > dmin) l += t;
    const int j = i + l * d;
    const int dnode = delta64(mort, prim, n, i, j);
    int s = 0;
    int t = l;
    do {
        t = (t + 1) >> 1;
        if (delta64(mort, prim, n, i, i + (s + t) * d) > dnode) s += t;
    } while (t > 1);
    const int gamma = i + s * d + min(d, 0);
    const int lo = min(i, j), hi = max(i, j);
    const int lref = (lo == gamma) ? ~gamma : gamma;
    const int rref = (hi == gamma + 1) ? ~(gamma + 1) : gamma + 1;
    left[i] = lref;
    right[i] = rref;
    if (lref >= 0) node_parent[lref] = i; else leaf_parent[~lref] = i;
    if (rref >= 0) node_parent[rref] = i; else leaf_parent[~rref] = i;
    if (i == 0) node_parent[0] = -1;
}

// ---- bottom-up refit ----------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) refit_kernel(int n, const int* __restrict__ left, const int* __restrict__ right,
                                                    const int* __restrict__ node_parent, const int* __restrict__ leaf_parent,
                                                    const uint32_t* __restrict__ sorted_prim, const float4* __restrict__ tri_bb,
                                                    float4* node_lo, float4* node_hi, uint32_t* __restrict__ visit)
{
    const int leaf = blockIdx.x * blockDim.x + threadIdx.x;
    if (leaf >= n) return;
    int cur = leaf_parent[leaf];
    while (cur >= 0) {
        // the second arrival at a node sees both children complete
        __threadfence();
        if (atomicAdd(&visit[cur], 1u) == 0u) return;
        __threadfence();
        float4 lo[2], hi[2];
        const int c[2] = {left[cur], right[cur]};
#pragma unroll
        for (int s = 0; s < 2; ++s) {
            if (c[s] < 0) {
                const uint32_t p = sorted_prim[~c[s]];
                lo[s] = tri_bb[2 * p + 0];
                hi[s] = tri_bb[2 * p + 1];
            } else {
                lo[s] = __ldcg(&node_lo[c[s]]);
                hi[s] = __ldcg(&node_hi[c[s]]);
            }
        }
        __stcg(&node_lo[cur], make_float4(fminf(lo[0].x, lo[1].x), fminf(lo[0].y, lo[1].y), fminf(lo[0].z, lo[1].z), 0.0f));
        __stcg(&node_hi[cur], make_float4(fmaxf(hi[0].x, hi[1].x), fmaxf(hi[0].y, hi[1].y), fmaxf(hi[0].z, hi[1].z), 0.0f));
        cur = node_parent[cur];
    }
}

// ---- pack -----------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) pack_nodes_kernel(int n_internal, const int* __restrict__ left, const int* __restrict__ right,
                                                         const uint32_t* __restrict__ sorted_prim, const float4* __restrict__ tri_bb,
                                                         const float4* __restrict__ node_lo, const float4* __restrict__ node_hi,
                                                         float4* __restrict__ nodes)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_internal) return;
    float4 lo[2], hi[2];
    const int c[2] = {left[i], right[i]};
#pragma unroll
    for (int s = 0; s < 2; ++s) {
        if (c[s] < 0) {
            const uint32_t p = sorted_prim[~c[s]];
            lo[s] = tri_bb[2 * p + 0];
            hi[s] = tri_bb[2 * p + 1];
        } else {
            lo[s] = node_lo[c[s]];
            hi[s] = node_hi[c[s]];
        }
    }
    nodes[4 * i + 0] = make_float4(lo[0].x, lo[0].y, lo[0].z, __int_as_float(c[0]));
    nodes[4 * i + 1] = make_float4(hi[0].x, hi[0].y, hi[0].z, __int_as_float(c[1]));
    nodes[4 * i + 2] = make_float4(lo[1].x, lo[1].y, lo[1].z, 0.0f);
    nodes[4 * i + 3] = make_float4(hi[1].x, hi[1].y, hi[1].z, 0.0f);
}

__global__ void __launch_bounds__(256) pack_tris_kernel(uint32_t n, const uint32_t* __restrict__ sorted_prim,
                                                        const float4* __restrict__ tri_in, float4* __restrict__ tris)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t p = sorted_prim[i];
    tris[3 * i + 0] = tri_in[3 * p + 0];
    tris[3 * i + 1] = tri_in[3 * p + 1];
    tris[3 * i + 2] = tri_in[3 * p + 2];
}

// single triangle: one node whose two children are the same leaf
__global__ void single_node_kernel(const float4* __restrict__ tri_bb, float4* __restrict__ nodes, int* node_parent)
{
    const float4 lo = tri_bb[0], hi = tri_bb[1];
    nodes[0] = make_float4(lo.x, lo.y, lo.z, __int_as_float(~0));
    nodes[1] = make_float4(hi.x, hi.y, hi.z, __int_as_float(~0));
    nodes[2] = make_float4(lo.x, lo.y, lo.z, 0.0f);
    nodes[3] = make_float4(hi.x, hi.y, hi.z, 0.0f);
    node_parent[0] = -1;
}

// Tree depth in internal nodes (root alone = 1): the traversal stack never holds more than this many entries.
__global__ void __launch_bounds__(256) depth_kernel(int n, const int* __restrict__ node_parent,
                                                    const int* __restrict__ leaf_parent, uint32_t* __restrict__ depth_out)
{
    const int leaf = blockIdx.x * blockDim.x + threadIdx.x;
    uint32_t d = 0;
    if (leaf < n) {
        int cur = leaf_parent[leaf];
        while (cur >= 0) {
            ++d;
            cur = node_parent[cur];
        }
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) d = max(d, __shfl_xor_sync(0xffffffffu, d, off));
    if ((threadIdx.x & 31) == 0 && d) atomicMax(depth_out, d);
}

// ---- PLOC: a better tree than Morton order for scenes too large for the host SAH sweep (sah.cu) ------------------------
// Parallel locally-ordered clustering (Meister & Bittner 2018) over the LBVH's leaves in Morton order.  Round: every
// cluster i looks at its neighbours j in [i - R, i + R] and picks the one whose union box has the smallest half-area
// (float, ex*ey + ey*ez + ez*ex; first strict minimum for ascending j); mutual choices merge into a node that takes the
// place of the lower index; the survivors are compacted in order.  Rounds repeat until one cluster is left.  Everything
// is a pure function of the leaf boxes and their order => deterministic; oracle/oracle_geom.h restates it serially and
// the node records are compared bit for bit.  Same 64-byte node record and the same leaves as the LBVH, so no pixel can
// change; node visits per ray drop by 22-27 % on the tessellated scene (fewer dependent fetches: C4 is latency-bound).
// Node ids count down from n - 2 in creation order, so the root -- created last -- is node 0 like in the LBVH.
constexpr int kPlocRadius = 8;

__device__ __forceinline__ float ploc_half_area(float4 lo_a, float4 hi_a, float4 lo_b, float4 hi_b)
{
    const float ex = fmaxf(hi_a.x, hi_b.x) - fminf(lo_a.x, lo_b.x);
    const float ey = fmaxf(hi_a.y, hi_b.y) - fminf(lo_a.y, lo_b.y);
    const float ez = fmaxf(hi_a.z, hi_b.z) - fminf(lo_a.z, lo_b.z);
    return ex * ey + ey * ez + ez * ex;
}

__global__ void __launch_bounds__(256) ploc_init_kernel(uint32_t n, const float4* __restrict__ tri_bb,
                                                        const uint32_t* __restrict__ sorted_prim, float4* __restrict__ lo,
                                                        float4* __restrict__ hi, int* __restrict__ ref)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const uint32_t p = sorted_prim[i];
    lo[i] = tri_bb[2 * p + 0];
    hi[i] = tri_bb[2 * p + 1];
    ref[i] = ~(int)i;
}

// The round kernels take the cluster count c and the next node id from device memory (state[0], state[1]): the host
// enqueues several rounds back to back with grids sized for the last count it knows and reads the state only every
// kPlocRoundsPerSync rounds -- a read-back per round made the build time hostage to the host's sync latency (10 ms on one
// box, 300 ms on another).  Rounds enqueued after the last merge find c == 1 and do nothing.
constexpr int kPlocRoundsPerSync = 8;

__global__ void __launch_bounds__(256) ploc_nearest_kernel(const int* __restrict__ state, const float4* __restrict__ lo,
                                                           const float4* __restrict__ hi, int* __restrict__ nn)
{
    const int c = state[0];
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (c <= 1 || i >= c) return;
    const float4 lo_i = lo[i], hi_i = hi[i];
    const int first = max(0, i - kPlocRadius), last = min(c - 1, i + kPlocRadius);
    int best_j = i > 0 ? i - 1 : i + 1;                  // never left undecided, whatever the boxes hold
    float best = __int_as_float(0x7f800000);
    for (int j = first; j <= last; ++j) {
        if (j == i) continue;
        const float a = ploc_half_area(lo_i, hi_i, lo[j], hi[j]);
        if (a < best) {
            best = a;
            best_j = j;
        }
    }
    nn[i] = best_j;
}

// keep[i]: cluster i survives the round (at its compacted position); make[i]: i is the lower index of a merging pair
__global__ void __launch_bounds__(256) ploc_flags_kernel(const int* __restrict__ state, int c_upper, const int* __restrict__ nn,
                                                         uint32_t* __restrict__ keep, uint32_t* __restrict__ make)
{
    const int c = state[0];
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i > c_upper) return;
    uint32_t k = 0u, m = 0u;
    if (c > 1 && i < c) {
        const int j = nn[i];
        const bool mutual = nn[j] == i;
        k = (mutual && i > j) ? 0u : 1u;
        m = (mutual && i < j) ? 1u : 0u;
    }
    keep[i] = k;      // entries c .. c_upper stay 0: after the exclusive scans entry c holds the totals
    make[i] = m;
}

// one thread: next round's cluster count and node id
__global__ void ploc_advance_kernel(int* __restrict__ state, const uint32_t* __restrict__ keep, const uint32_t* __restrict__ make)
{
    const int c = state[0];
    if (c <= 1) return;
    const uint32_t kept = keep[c], made = make[c];
    if (made == 0u || kept + made != (uint32_t)c) state[3] = 1;      // cannot happen: every round merges at least one pair
    state[0] = made == 0u ? 1 : (int)kept;
    state[1] -= (int)made;
    state[2] += 1;
}

__global__ void __launch_bounds__(256) ploc_emit_kernel(const int* __restrict__ state, const int* __restrict__ nn,
                                                        const uint32_t* __restrict__ pos, const uint32_t* __restrict__ rank,
                                                        const float4* __restrict__ lo_in, const float4* __restrict__ hi_in,
                                                        const int* __restrict__ ref_in, float4* __restrict__ lo_out,
                                                        float4* __restrict__ hi_out, int* __restrict__ ref_out,
                                                        float4* __restrict__ nodes, int* __restrict__ node_parent,
                                                        int* __restrict__ leaf_parent)
{
    const int c = state[0], next_id = state[1];
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (c <= 1 || i >= c) return;
    const int j = nn[i];
    const bool mutual = nn[j] == i;
    if (mutual && i > j) return;                         // absorbed by cluster j
    const uint32_t p = pos[i];
    const float4 lo_i = lo_in[i], hi_i = hi_in[i];
    const int ref_i = ref_in[i];
    if (!mutual) {
        lo_out[p] = lo_i;
        hi_out[p] = hi_i;
        ref_out[p] = ref_i;
        return;
    }
    const float4 lo_j = lo_in[j], hi_j = hi_in[j];
    const int ref_j = ref_in[j];
    const int id = next_id - (int)rank[i];
    nodes[4 * (size_t)id + 0] = make_float4(lo_i.x, lo_i.y, lo_i.z, __int_as_float(ref_i));
    nodes[4 * (size_t)id + 1] = make_float4(hi_i.x, hi_i.y, hi_i.z, __int_as_float(ref_j));
    nodes[4 * (size_t)id + 2] = make_float4(lo_j.x, lo_j.y, lo_j.z, 0.0f);
    nodes[4 * (size_t)id + 3] = make_float4(hi_j.x, hi_j.y, hi_j.z, 0.0f);
    if (ref_i >= 0) node_parent[ref_i] = id; else leaf_parent[~ref_i] = id;
    if (ref_j >= 0) node_parent[ref_j] = id; else leaf_parent[~ref_j] = id;
    lo_out[p] = make_float4(fminf(lo_i.x, lo_j.x), fminf(lo_i.y, lo_j.y), fminf(lo_i.z, lo_j.z), 0.0f);
    hi_out[p] = make_float4(fmaxf(hi_i.x, hi_j.x), fmaxf(hi_i.y, hi_j.y), fmaxf(hi_i.z, hi_j.z), 0.0f);
    ref_out[p] = id;
}

// ---- quantised nodes: 32 bytes per node for scenes that are read from global memory ------------------------------------
// Twelve 16-bit plane positions on the grid  lo + q * step  of the scene's bounding box and the two child references (layout:
// pt_device.cuh FetchGlobal::qnode4).  lo planes are rounded down, hi planes up, and the rounding is CHECKED with the very
// expression the traversal evaluates (one fused multiply-add), so a quantised box always contains the exact one: the
// traversal reaches a superset of the leaves and the hits -- which come from the exact triangle records -- cannot change.
// plane q of the grid: (2^23 + q) * step + base in one fused multiply-add -- the traversal's expression (pt_device.cuh qnode4)
__host__ __device__ __forceinline__ float quant_plane(uint32_t q, float base, float step)
{
#ifdef __CUDA_ARCH__
    return __fmaf_rn(__uint_as_float(0x4b000000u | q), step, base);
#else
    return fmaf(8388608.0f + (float)q, step, base);
#endif
}
__device__ __forceinline__ uint32_t quant_down(float x, float base, float step)
{
    float t = floorf((x - quant_plane(0u, base, step)) / step);
    t = fminf(fmaxf(t, 0.0f), 65535.0f);
    uint32_t q = (uint32_t)t;
    while (q > 0u && quant_plane(q, base, step) > x) --q;
    while (q < 65535u && quant_plane(q + 1u, base, step) <= x) ++q;
    return q;
}
__device__ __forceinline__ uint32_t quant_up(float x, float base, float step)
{
    float t = ceilf((x - quant_plane(0u, base, step)) / step);
    t = fminf(fmaxf(t, 0.0f), 65535.0f);
    uint32_t q = (uint32_t)t;
    while (q < 65535u && quant_plane(q, base, step) < x) ++q;
    while (q > 0u && quant_plane(q - 1u, base, step) >= x) --q;
    return q;
}

__global__ void __launch_bounds__(256) quantize_nodes_kernel(int n_internal, const float4* __restrict__ nodes, float3 lo, float3 step,
                                                             uint4* __restrict__ qnodes, uint32_t* __restrict__ bad)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_internal) return;
    const float4 q0 = nodes[4 * (size_t)i + 0], q1 = nodes[4 * (size_t)i + 1], q2 = nodes[4 * (size_t)i + 2], q3 = nodes[4 * (size_t)i + 3];
    const uint32_t l0x = quant_down(q0.x, lo.x, step.x), l0y = quant_down(q0.y, lo.y, step.y), l0z = quant_down(q0.z, lo.z, step.z);
    const uint32_t h0x = quant_up(q1.x, lo.x, step.x), h0y = quant_up(q1.y, lo.y, step.y), h0z = quant_up(q1.z, lo.z, step.z);
    const uint32_t l1x = quant_down(q2.x, lo.x, step.x), l1y = quant_down(q2.y, lo.y, step.y), l1z = quant_down(q2.z, lo.z, step.z);
    const uint32_t h1x = quant_up(q3.x, lo.x, step.x), h1y = quant_up(q3.y, lo.y, step.y), h1z = quant_up(q3.z, lo.z, step.z);
    // the guarantee, re-checked: quantised lo <= exact lo and quantised hi >= exact hi on every plane
    const bool ok = quant_plane(l0x, lo.x, step.x) <= q0.x && quant_plane(l0y, lo.y, step.y) <= q0.y && quant_plane(l0z, lo.z, step.z) <= q0.z &&
                    quant_plane(h0x, lo.x, step.x) >= q1.x && quant_plane(h0y, lo.y, step.y) >= q1.y && quant_plane(h0z, lo.z, step.z) >= q1.z &&
                    quant_plane(l1x, lo.x, step.x) <= q2.x && quant_plane(l1y, lo.y, step.y) <= q2.y && quant_plane(l1z, lo.z, step.z) <= q2.z &&
                    quant_plane(h1x, lo.x, step.x) >= q3.x && quant_plane(h1y, lo.y, step.y) >= q3.y && quant_plane(h1z, lo.z, step.z) >= q3.z;
    if (!ok) atomicAdd(bad, 1u);
    qnodes[2 * (size_t)i + 0] = make_uint4(l0x | (l0y << 16), l0z | (h0x << 16), h0y | (h0z << 16), l1x | (l1y << 16));
    qnodes[2 * (size_t)i + 1] = make_uint4(l1z | (h1x << 16), h1y | (h1z << 16), __float_as_uint(q0.w), __float_as_uint(q1.w));
}

#define PTB_CUDA_TRY(expr)                  \
    do {                                    \
        cudaError_t e__ = (expr);           \
        if (e__ != cudaSuccess) return e__; \
    } while (0)

static inline unsigned div_up(unsigned a, unsigned b) { return (a + b - 1) / b; }

// Builds every device array of `s` from the uploaded inputs.  All work is enqueued on `stream`.
size_t lbvh_scan_scratch_entries(size_t n_tris)
{
    const size_t n_blocks = (n_tris + kSortTile - 1) / kSortTile;
    return scan_scratch_entries(std::max<size_t>(256 * n_blocks, n_tris + 1));
}

cudaError_t lbvh_build(Scene& s, cudaStream_t stream)
{
    const uint32_t n = s.n_tris;
    if (n == 0) return cudaSuccess;
    const unsigned g = div_up(n, 256);
    init_bounds_kernel<<<1, 32, 0, stream>>>(s.d_bounds_ordered);
    tri_setup_kernel<<<g, 256, 0, stream>>>(s.d_verts, s.d_indices, s.d_mat_ids, n, s.d_tri_in, s.d_tri_bb, s.d_tri_shade,
                                            s.d_bounds_ordered);
    morton_kernel<<<g, 256, 0, stream>>>(s.d_tri_bb, n, s.d_bounds_ordered, s.d_keys[0], s.d_vals[0], s.d_bounds);
    PTB_CUDA_TRY(cudaMemcpyAsync(s.d_morton, s.d_keys[0], sizeof(uint32_t) * n, cudaMemcpyDeviceToDevice, stream));
    const uint32_t n_blocks = div_up(n, kSortTile);
    int src = 0;
    for (int pass = 0; pass < 4; ++pass) {
        const int shift = 8 * pass;
        sort_hist_kernel<<<n_blocks, kSortThreads, 0, stream>>>(s.d_keys[src], n, shift, s.d_block_hist, n_blocks);
        scan_exclusive(s.d_block_hist, 256u * n_blocks, s.d_scan_scratch, stream);
        sort_scatter_kernel<<<n_blocks, kSortThreads, 0, stream>>>(s.d_keys[src], s.d_vals[src], s.d_keys[src ^ 1],
                                                                  s.d_vals[src ^ 1], n, shift, s.d_block_hist, n_blocks);
        src ^= 1;
    }
    // after 4 passes the sorted data is back in buffer 0
    s.sorted_buf = src;
    const uint32_t* mort = s.d_keys[src];
    const uint32_t* prim = s.d_vals[src];
    PTB_CUDA_TRY(cudaMemsetAsync(s.d_work_counter + 1, 0, sizeof(uint32_t), stream));
    if (n == 1) {
        single_node_kernel<<<1, 1, 0, stream>>>(s.d_tri_bb, s.d_nodes, s.d_node_parent);
        PTB_CUDA_TRY(cudaMemsetAsync(s.d_work_counter + 1, 1, 1, stream));   // depth 1
    } else {
        karras_kernel<<<div_up(n - 1, 256), 256, 0, stream>>>(mort, prim, (int)n, s.d_left, s.d_right, s.d_node_parent,
                                                             s.d_leaf_parent);
        PTB_CUDA_TRY(cudaMemsetAsync(s.d_visit, 0, sizeof(uint32_t) * (n - 1), stream));
        refit_kernel<<<g, 256, 0, stream>>>((int)n, s.d_left, s.d_right, s.d_node_parent, s.d_leaf_parent, prim, s.d_tri_bb,
                                            s.d_node_lo, s.d_node_hi, s.d_visit);
        pack_nodes_kernel<<<div_up(n - 1, 256), 256, 0, stream>>>((int)n - 1, s.d_left, s.d_right, prim, s.d_tri_bb, s.d_node_lo,
                                                                 s.d_node_hi, s.d_nodes);
        depth_kernel<<<g, 256, 0, stream>>>((int)n, s.d_node_parent, s.d_leaf_parent, s.d_work_counter + 1);
    }
    pack_tris_kernel<<<g, 256, 0, stream>>>(n, prim, s.d_tri_in, s.d_tris);
    return cudaGetLastError();
}

// Builds s.d_nodes_ploc (n - 1 nodes) and s.ploc_depth from the sorted leaves; synchronous (reads two counters per round).
// Leaves s.d_nodes_ploc null when the scene has fewer than 2 triangles or the rounds do not converge.
cudaError_t ploc_build(Scene& s, cudaStream_t stream)
{
    const uint32_t n = s.n_tris;
    if (n < 2) return cudaSuccess;
    float4 *lo[2] = {nullptr, nullptr}, *hi[2] = {nullptr, nullptr}, *nodes = nullptr;
    int *ref[2] = {nullptr, nullptr}, *nn = nullptr, *node_parent = nullptr, *leaf_parent = nullptr;
    uint32_t *keep = nullptr, *make = nullptr;
    int* state = nullptr;                                // [0] clusters left  [1] next node id  [2] rounds done  [3] error
    cudaError_t e = cudaSuccess;
    auto A = [&](cudaError_t r) {
        if (e == cudaSuccess) e = r;
    };
    // one arena for the scratch arrays (a dozen separate cudaMalloc / cudaFree pairs cost more than the build)
    char* arena = nullptr;
    size_t arena_bytes = 0;
    auto carve = [&](size_t bytes) {
        const size_t off = arena_bytes;
        arena_bytes += (bytes + 255) & ~(size_t)255;
        return off;
    };
    const size_t o_lo0 = carve(sizeof(float4) * n), o_lo1 = carve(sizeof(float4) * n), o_hi0 = carve(sizeof(float4) * n),
                 o_hi1 = carve(sizeof(float4) * n), o_ref0 = carve(sizeof(int) * n), o_ref1 = carve(sizeof(int) * n),
                 o_nn = carve(sizeof(int) * n), o_state = carve(sizeof(int) * 4), o_keep = carve(sizeof(uint32_t) * ((size_t)n + 1)),
                 o_make = carve(sizeof(uint32_t) * ((size_t)n + 1)), o_np = carve(sizeof(int) * n), o_lp = carve(sizeof(int) * n);
    A(cudaMalloc((void**)&arena, arena_bytes));
    A(cudaMalloc((void**)&nodes, sizeof(float4) * 4 * ((size_t)n - 1)));
    if (e == cudaSuccess) {
        lo[0] = (float4*)(arena + o_lo0), lo[1] = (float4*)(arena + o_lo1), hi[0] = (float4*)(arena + o_hi0), hi[1] = (float4*)(arena + o_hi1);
        ref[0] = (int*)(arena + o_ref0), ref[1] = (int*)(arena + o_ref1), nn = (int*)(arena + o_nn), state = (int*)(arena + o_state);
        keep = (uint32_t*)(arena + o_keep), make = (uint32_t*)(arena + o_make);
        node_parent = (int*)(arena + o_np), leaf_parent = (int*)(arena + o_lp);
    }
    bool ok = e == cudaSuccess;
    uint32_t depth = 0;
    if (ok) {
        ploc_init_kernel<<<div_up(n, 256), 256, 0, stream>>>(n, s.d_tri_bb, s.d_vals[s.sorted_buf], lo[0], hi[0], ref[0]);
        int h_state[4] = {(int)n, (int)n - 2, 0, 0};
        A(cudaMemcpyAsync(state, h_state, sizeof(h_state), cudaMemcpyHostToDevice, stream));
        int c_upper = (int)n, cur = 0;
        for (int batch = 0; c_upper > 1 && ok; ++batch) {
            if (batch > 100000) {
                ok = false;
                break;
            }
            for (int r = 0; r < kPlocRoundsPerSync; ++r) {
                ploc_nearest_kernel<<<div_up(c_upper, 256), 256, 0, stream>>>(state, lo[cur], hi[cur], nn);
                ploc_flags_kernel<<<div_up(c_upper + 1, 256), 256, 0, stream>>>(state, c_upper, nn, keep, make);
                scan_exclusive(keep, (uint32_t)c_upper + 1u, s.d_scan_scratch, stream);
                scan_exclusive(make, (uint32_t)c_upper + 1u, s.d_scan_scratch, stream);
                ploc_emit_kernel<<<div_up(c_upper, 256), 256, 0, stream>>>(state, nn, keep, make, lo[cur], hi[cur], ref[cur], lo[cur ^ 1],
                                                                          hi[cur ^ 1], ref[cur ^ 1], nodes, node_parent, leaf_parent);
                ploc_advance_kernel<<<1, 1, 0, stream>>>(state, keep, make);
                cur ^= 1;
            }
            A(cudaMemcpyAsync(h_state, state, sizeof(h_state), cudaMemcpyDeviceToHost, stream));
            A(cudaStreamSynchronize(stream));
            if (e != cudaSuccess || h_state[3] != 0) {
                ok = false;
                break;
            }
            c_upper = h_state[0];
        }
        const int next_id = h_state[1];
        if (ok && next_id != -1) ok = false;
        if (ok) {
            const int root_parent = -1;
            A(cudaMemcpyAsync(node_parent, &root_parent, sizeof(int), cudaMemcpyHostToDevice, stream));
            A(cudaMemsetAsync(s.d_work_counter + 2, 0, sizeof(uint32_t), stream));
            depth_kernel<<<div_up(n, 256), 256, 0, stream>>>((int)n, node_parent, leaf_parent, s.d_work_counter + 2);
            A(cudaMemcpyAsync(&depth, s.d_work_counter + 2, sizeof(uint32_t), cudaMemcpyDeviceToHost, stream));
            A(cudaStreamSynchronize(stream));
            A(cudaGetLastError());
            ok = e == cudaSuccess;
        }
    }
    if (arena) cudaFree(arena);
    if (ok) {
        s.d_nodes_ploc = nodes;
        s.ploc_depth = depth;
    } else if (nodes) {
        cudaFree(nodes);
    }
    return e;
}

// Quantised copy of the tree the kernels walk (scene_nodes), rebuilt when that choice changes.  Leaves s.d_qnodes null
// when a box cannot be contained (never expected: the grid spans the scene's bounding box).
cudaError_t quantize_nodes(Scene& s, cudaStream_t stream)
{
    const int tree = scene_tree(s);
    if (s.n_tris < 2) return cudaSuccess;
    if (s.d_qnodes && s.qnodes_tree == tree) return cudaSuccess;
    if (!s.d_qnodes) PTB_CUDA_TRY(cudaMalloc((void**)&s.d_qnodes, sizeof(uint4) * 2 * (size_t)s.n_nodes));
    s.qnodes_tree = 0;
    for (int a = 0; a < 3; ++a) {
        // the grid: 65536 planes, a few dozen steps of margin on either side; `base` = origin - 2^23 * step is rounded to a
        // float (its ulp is about one step), so both ends are checked with the traversal's own expression and nudged
        const float lo = s.h_bounds[a], hi = s.h_bounds[3 + a];
        float step = (hi - lo) / 65500.0f;
        if (!(step > 1e-30f)) step = 1e-30f;
        float base = (float)((double)lo - 16.0 * (double)step - 8388608.0 * (double)step);
        for (int it = 0; it < 256; ++it) {
            if (quant_plane(0u, base, step) > lo) base = nextafterf(base, -INFINITY);
            else if (quant_plane(65535u, base, step) < hi) step = nextafterf(step, INFINITY);
            else break;
        }
        s.q_lo[a] = base;
        s.q_step[a] = step;
    }
    PTB_CUDA_TRY(cudaMemsetAsync(s.d_work_counter + 3, 0, sizeof(uint32_t), stream));
    quantize_nodes_kernel<<<div_up(s.n_nodes, 256), 256, 0, stream>>>((int)s.n_nodes, scene_nodes(s), make_float3(s.q_lo[0], s.q_lo[1], s.q_lo[2]),
                                                                      make_float3(s.q_step[0], s.q_step[1], s.q_step[2]),
                                                                      reinterpret_cast<uint4*>(s.d_qnodes), s.d_work_counter + 3);
    uint32_t bad = 0;
    PTB_CUDA_TRY(cudaMemcpyAsync(&bad, s.d_work_counter + 3, sizeof(uint32_t), cudaMemcpyDeviceToHost, stream));
    PTB_CUDA_TRY(cudaStreamSynchronize(stream));
    if (bad == 0) s.qnodes_tree = tree;
    return cudaGetLastError();
}


// ---- 4-wide copy of the quantised tree (pt_scene_set_traversal 4) --------------------------------------------------------
// Every wide node takes a binary node's two children and twice replaces the internal child with the largest surface area (first
// one on ties) by that child's own two children, so one 64-byte fetch resolves up to two binary levels along the rays' most
// likely way down.  Child boxes are the 16-bit planes of the binary quantised nodes, copied: the containment guarantee of
// quantize_nodes carries over.  Nodes are numbered breadth first (root = 0: the levels every ray walks share cache lines), which
// also makes the build a loop over levels: count the internal children of the level's nodes, scan, write the records and the next
// level's list.  Record layout: pt_device.cuh FetchGlobal::wnode16.
struct WideChild {
    uint32_t w[3];     // lo | hi << 16 per axis
    int ref;
    double area;
};
__device__ __forceinline__ WideChild wide_child(const uint32_t* __restrict__ q, int node, int k, double sx, double sy, double sz)
{
    const uint32_t* r = q + 8 * (size_t)node;
    const uint32_t w0 = r[3 * k], w1 = r[3 * k + 1], w2 = r[3 * k + 2];   // lo.x | lo.y << 16, lo.z | hi.x << 16, hi.y | hi.z << 16
    const uint32_t lx = w0 & 0xffffu, ly = w0 >> 16, lz = w1 & 0xffffu, hx = w1 >> 16, hy = w2 & 0xffffu, hz = w2 >> 16;
    WideChild c;
    c.w[0] = lx | (hx << 16), c.w[1] = ly | (hy << 16), c.w[2] = lz | (hz << 16);
    c.ref = (int)r[6 + k];
    const double dx = sx * ((double)hx - (double)lx), dy = sy * ((double)hy - (double)ly), dz = sz * ((double)hz - (double)lz);
    c.area = dx * dy + dy * dz + dz * dx;
    return c;
}
__device__ __forceinline__ int wide_children(const uint32_t* __restrict__ q, int node, double sx, double sy, double sz, WideChild ch[4])
{
    int n = 2;
    ch[0] = wide_child(q, node, 0, sx, sy, sz), ch[1] = wide_child(q, node, 1, sx, sy, sz);
    while (n < 4) {
        int best = -1;
        for (int k = 0; k < n; ++k)
            if (ch[k].ref >= 0 && (best < 0 || ch[k].area > ch[best].area)) best = k;
        if (best < 0) break;
        const int b = ch[best].ref;
        ch[best] = wide_child(q, b, 0, sx, sy, sz);
        ch[n++] = wide_child(q, b, 1, sx, sy, sz);
    }
    return n;
}

// state: [0] first wide index of the level  [1] nodes in the level  [2] levels so far
__global__ void wide_init_kernel(int* __restrict__ state, int* __restrict__ level_nodes)
{
    state[0] = 0, state[1] = 1, state[2] = 1;
    level_nodes[0] = 0;
}
__global__ void __launch_bounds__(256) wide_count_kernel(uint32_t bound, const int* __restrict__ state, const int* __restrict__ level_nodes,
                                                         const uint32_t* __restrict__ q, double sx, double sy, double sz,
                                                         uint32_t* __restrict__ counts)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i > bound) return;
    uint32_t c = 0;
    if (i < (uint32_t)state[1]) {
        WideChild ch[4];
        const int n = wide_children(q, level_nodes[i], sx, sy, sz, ch);
        for (int k = 0; k < n; ++k) c += ch[k].ref >= 0 ? 1u : 0u;
    }
    counts[i] = c;                                       // counts[bound] = 0: after the scan, the size of the next level
}
__global__ void __launch_bounds__(256) wide_emit_kernel(uint32_t bound, const int* __restrict__ state, const int* __restrict__ level_nodes,
                                                        const uint32_t* __restrict__ q, double sx, double sy, double sz,
                                                        const uint32_t* __restrict__ offsets, int* __restrict__ next_nodes,
                                                        uint4* __restrict__ out)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= bound || i >= (uint32_t)state[1]) return;
    WideChild ch[4];
    const int n = wide_children(q, level_nodes[i], sx, sy, sz, ch);
    uint32_t rec[16];
    uint32_t at = offsets[i];
    const uint32_t next_first = (uint32_t)(state[0] + state[1]);
    for (int k = 0; k < 4; ++k) {
        if (k >= n) {                                    // empty slot: child 0 again behind an inside-out box
            rec[3 * k] = rec[3 * k + 1] = rec[3 * k + 2] = 0x0000ffffu;
            continue;
        }
        rec[3 * k] = ch[k].w[0], rec[3 * k + 1] = ch[k].w[1], rec[3 * k + 2] = ch[k].w[2];
        if (ch[k].ref >= 0) {
            next_nodes[at] = ch[k].ref;
            rec[12 + k] = next_first + at++;
        } else {
            rec[12 + k] = (uint32_t)ch[k].ref;
        }
    }
    for (int k = n; k < 4; ++k) rec[12 + k] = rec[12];
    uint4* o = out + 4 * ((size_t)state[0] + i);
    for (int k = 0; k < 4; ++k) o[k] = make_uint4(rec[4 * k], rec[4 * k + 1], rec[4 * k + 2], rec[4 * k + 3]);
}
__global__ void wide_advance_kernel(uint32_t bound, int* __restrict__ state, const uint32_t* __restrict__ offsets)
{
    const int next = (int)offsets[bound];
    state[0] += state[1];
    state[1] = next;
    if (next > 0) state[2] += 1;
}

// Leaves s.d_wnodes null when the quantised nodes are not available.  Synchronous (reads the node count back).
cudaError_t build_wide_nodes(Scene& s, cudaStream_t stream)
{
    const int tree = scene_tree(s);
    if (s.n_tris < 2) return cudaSuccess;
    if (s.d_wnodes && s.wnodes_tree == tree) return cudaSuccess;
    PTB_CUDA_TRY(quantize_nodes(s, stream));
    if (!s.d_qnodes || s.qnodes_tree != tree) return cudaSuccess;
    s.wnodes_tree = 0;
    if (s.d_wnodes) cudaFree(s.d_wnodes);
    s.d_wnodes = nullptr;
    const auto t_begin = std::chrono::steady_clock::now();
    const uint32_t nn = s.n_nodes;
    char* arena = nullptr;                               // two level lists, the counts, the state, all nn records
    const size_t o_list0 = 0, o_list1 = o_list0 + sizeof(int) * (size_t)nn, o_counts = o_list1 + sizeof(int) * (size_t)nn,
                 o_state = o_counts + sizeof(uint32_t) * ((size_t)nn + 1), o_out = (o_state + 16 + 255) & ~(size_t)255;
    PTB_CUDA_TRY(cudaMalloc((void**)&arena, o_out + 64 * (size_t)nn));
    int* lists[2] = {(int*)(arena + o_list0), (int*)(arena + o_list1)};
    uint32_t* counts = (uint32_t*)(arena + o_counts);
    int* state = (int*)(arena + o_state);
    uint4* out = (uint4*)(arena + o_out);
    const uint32_t* q = reinterpret_cast<const uint32_t*>(s.d_qnodes);
    const double sx = s.q_step[0], sy = s.q_step[1], sz = s.q_step[2];
    wide_init_kernel<<<1, 1, 0, stream>>>(state, lists[0]);
    const uint32_t levels = scene_depth(s) + 1u;         // a wide level never spans less than one binary level
    uint64_t width = 1;
    for (uint32_t level = 0; level < levels; ++level) {
        const uint32_t bound = (uint32_t)std::min<uint64_t>(width, nn);
        const int cur = (int)(level & 1u);
        wide_count_kernel<<<div_up(bound + 1, 256), 256, 0, stream>>>(bound, state, lists[cur], q, sx, sy, sz, counts);
        scan_exclusive(counts, bound + 1u, s.d_scan_scratch, stream);
        wide_emit_kernel<<<div_up(bound, 256), 256, 0, stream>>>(bound, state, lists[cur], q, sx, sy, sz, counts, lists[cur ^ 1], out);
        wide_advance_kernel<<<1, 1, 0, stream>>>(bound, state, counts);
        width = std::min<uint64_t>(width * 4, nn);
    }
    int h_state[3] = {0, 0, 0};
    cudaError_t e = cudaMemcpyAsync(h_state, state, sizeof(h_state), cudaMemcpyDeviceToHost, stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
    if (e == cudaSuccess) e = cudaGetLastError();
    if (e == cudaSuccess && h_state[1] == 0 && h_state[0] >= 1 && (uint32_t)h_state[0] <= nn) {   // every level consumed
        e = cudaMalloc((void**)&s.d_wnodes, 64 * (size_t)h_state[0]);
        if (e == cudaSuccess) e = cudaMemcpyAsync(s.d_wnodes, out, 64 * (size_t)h_state[0], cudaMemcpyDeviceToDevice, stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
        if (e == cudaSuccess) {
            s.n_wnodes = (uint32_t)h_state[0];
            s.wide_depth = (uint32_t)h_state[2];
            s.wnodes_tree = tree;
        } else if (s.d_wnodes) {
            cudaFree(s.d_wnodes);
            s.d_wnodes = nullptr;
        }
    }
    cudaFree(arena);
    s.wide_ms = std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - t_begin).count();
    return e;
}

}  // namespace ptb
