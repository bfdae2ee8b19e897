// render.cu -- the render kernel: what optixLaunch(width, height, 1) runs in the reference
// (__raygen__rg / __closesthit__diffuse__ch / __miss__ms, pathTracerPrograms.cu:707-1031), re-designed for
// sm_100a without OptiX.
//
// render_persistent_kernel: a persistent grid (SMs x occupancy CTAs) of warps; one LANE = one pixel = one path
// in flight, kept entirely in registers.
//   * Per-pixel order is the reference's: samples in sequence, segments in path order, so the float sums and the
//     RNG chaining are bit-identical; when a path ends the lane regenerates the pixel's next sample in
//     place, and when the pixel has used all its samples the lane writes it and takes the next pixel index from a
//     global counter (8x4-pixel tiles, tile-major), so lanes never wait for the slowest pixel of a tile.
//   * A lane is always in one of four stages: NODE (internal-node slab tests), LEAF (triangle test), SHADE
//     (closest-hit shading + BSDF sample + NEE set-up) or END (miss / shadow result, accumulate, Russian roulette,
//     next ray or next sample or next pixel).  Radiance rays and shadow rays share the same traversal code.
//     Each round the warp votes (ballot + popc) and runs the stage with the most waiting lanes, so every stage
//     body is executed once per round for many lanes instead of once per lane-specific control path.
//   * The whole scene (64-B BVH nodes, 48-B triangles, normals, materials) is staged into shared memory when it
//     fits (Cornell box: 4 KB), else read through the read-only path from L1/L2.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <vector>

#include "pt_scene.h"
#include "pt_shade.cuh"

namespace ptb {

using namespace ptd;

constexpr int kBlock = 256;
constexpr int kMinBlocksPerSm = 2;
constexpr uint32_t kFull = 0xffffffffu;
constexpr size_t kSmemSceneLimit = 40 * 1024;   // stage the scene in shared memory up to this size

struct ThreadCounters {
    uint32_t n_samples, n_radiance, n_shadow, n_node, n_tri, n_hit0, n_hit1, n_hit2, n_miss;
};

__device__ __forceinline__ void flush_counter(unsigned long long* dst, uint32_t v)
{
    unsigned long long x = v;
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) x += __shfl_xor_sync(kFull, x, off);
    if ((threadIdx.x & 31) == 0 && x) atomicAdd(dst, x);
}

__device__ __forceinline__ float safe_divide(float a, float b) { return b == 0.0f ? 0.0f : a / b; }   // :265-284

// Scene bytes when staged: nodes | tris | tri_shade | materials (all float4)
__host__ __device__ inline size_t staged_scene_float4s(uint32_t n_nodes, uint32_t n_tris, uint32_t n_mats)
{
    return 4 * (size_t)n_nodes + 3 * (size_t)n_tris + (size_t)n_tris + 3 * (size_t)n_mats;
}

template <bool SMEM> struct FetchSelect { typedef FetchGlobal type; };
template <> struct FetchSelect<true> { typedef FetchShared type; };

enum : int { PH_RAD = 0, PH_SHADOW = 1, PH_END = 2, PH_IDLE = 3 };

// Work item j of this rank -> pixel: owned tile j / 32 (image tile rank + k * n_ranks), pixel j % 32 of the 8x4 tile.
// Items that fall outside a ragged image edge are skipped by taking the next index.
__device__ __forceinline__ bool take_pixel(const LaunchConst& lc, uint32_t n_work, uint32_t j, uint32_t& ix, uint32_t& iy)
{
    for (;;) {
        if (j >= n_work) return false;
        const uint32_t tile = lc.tile_order ? __ldg(lc.tile_order + (j >> 5)) : owned_tile(j >> 5, lc.rank, lc.span, lc.n_ranks);
        const uint32_t tx = tile % lc.tiles_x, ty = tile / lc.tiles_x;
        ix = tx * 8u + (j & 7u);
        iy = ty * 4u + ((j >> 3) & 3u);
        if (ix < lc.width && iy < lc.height) return true;
        j = atomicAdd(lc.work_counter, 1u);
    }
}

template <bool COUNT, bool PROBE, bool SMEM>
__global__ void __launch_bounds__(kBlock, kMinBlocksPerSm) render_persistent_kernel(const __grid_constant__ LaunchConst lc)
{
    extern __shared__ float4 smem4[];
    typename FetchSelect<SMEM>::type f;
    f.nodes = lc.bvh.nodes;
    f.tris = lc.bvh.tris;
    f.shade = lc.tri_shade;
    f.mats = reinterpret_cast<const float4*>(lc.mats);
    if (SMEM) {
        const uint32_t n_nodes4 = 4 * lc.n_nodes, n_tris4 = 3 * lc.bvh.n_tris, n_sh = lc.bvh.n_tris, n_m4 = 3 * lc.n_mats;
        float4* s_nodes = smem4;
        float4* s_tris = s_nodes + n_nodes4;
        float4* s_shade = s_tris + n_tris4;
        float4* s_mats = s_shade + n_sh;
        for (uint32_t i = threadIdx.x; i < n_nodes4; i += kBlock) s_nodes[i] = __ldg(f.nodes + i);
        for (uint32_t i = threadIdx.x; i < n_tris4; i += kBlock) s_tris[i] = __ldg(f.tris + i);
        for (uint32_t i = threadIdx.x; i < n_sh; i += kBlock) s_shade[i] = __ldg(f.shade + i);
        for (uint32_t i = threadIdx.x; i < n_m4; i += kBlock) s_mats[i] = __ldg(f.mats + i);
        __syncthreads();
        f.nodes = s_nodes;
        f.tris = s_tris;
        f.shade = s_shade;
        f.mats = s_mats;
    }

    const float l_area = length(cross(lc.l_v1, lc.l_v2));   // :1021
    const int lane = threadIdx.x & 31;
    const bool have_tris = lc.bvh.n_tris != 0;
    const uint32_t n_work = lc.n_owned_tiles * 32u;
    ThreadCounters ctr = {};

    // ---- lane state ------------------------------------------------------------------------------------------
    uint32_t ix = 0, iy = 0, pixel_seed = 0, left = 0;   // pixel, its seed (advanced by jitter draws only), samples left
    float3 result = f3(0.0f);                            // sum over this launch's samples (:724-780)
    Path p;                                              // next ray / attenuation / path seed / depth
    p.o = p.d = p.att = f3(0.0f);
    p.seed = 0;
    p.depth = 0;
    Trav t;                                              // traversal in flight (radiance or shadow ray)
    t.r.o = t.r.d = t.r.inv = t.r.ood = f3(0.0f);
    t.tfar = 0.0f;
    t.prim = kNoPrim;
    t.sp = 0;
    t.cur = kTravDone;
    int stack_mem[kStackSize];
    const StackLocal stack{stack_mem};
    float3 rad = f3(0.0f), add = f3(0.0f);               // radiance of the segment; NEE term added if unoccluded
    bool done = false;
    int phase = PH_IDLE;

    {   // first pixel: 32 consecutive work items per warp (one 8x4 tile)
        uint32_t j = 0;
        if (lane == 0) j = atomicAdd(lc.work_counter, 32u);
        j = __shfl_sync(kFull, j, 0) + (uint32_t)lane;
        if (take_pixel(lc, n_work, j, ix, iy)) {
            const uint32_t image_index = iy * lc.width + ix;
            pixel_seed = tea4(image_index, lc.frame);                                          // :721
            left = lc.spp;
            generate_camera_ray(lc, ix, iy, pixel_seed, p);
            if (COUNT) ctr.n_samples++, ctr.n_radiance++;
            if (PROBE) {
                if (lc.dbg_seed_after_jitter) lc.dbg_seed_after_jitter[image_index] = pixel_seed;
                if (lc.dbg_primary_dir) {
                    lc.dbg_primary_dir[3 * image_index + 0] = p.d.x;
                    lc.dbg_primary_dir[3 * image_index + 1] = p.d.y;
                    lc.dbg_primary_dir[3 * image_index + 2] = p.d.z;
                }
            }
            trav_begin(t, p.o, p.d, 1e16f, have_tris);                                         // tmax :755
            phase = PH_RAD;
        }
    }

    const int w_node = lc.sched[0], w_leaf = lc.sched[1], w_shade = lc.sched[2], w_end = lc.sched[3];
    const int node_min = lc.sched[4];

    for (;;) {
        // ---- vote -------------------------------------------------------------------------------------------
        const int n_node = __popc(__ballot_sync(kFull, t.cur >= 0));
        int stage = 0;
        if (n_node < node_min) {
            const bool fin = t.cur == kTravDone;
            const bool sh = fin && phase == PH_RAD && t.prim != kNoPrim;
            const bool en = fin && !sh && phase != PH_IDLE;
            const int n_leaf = __popc(__ballot_sync(kFull, !fin && t.cur < 0));
            const int n_sh = __popc(__ballot_sync(kFull, sh));
            const int n_en = __popc(__ballot_sync(kFull, en));
            if ((n_node | n_leaf | n_sh | n_en) == 0) break;
            const int s0 = n_node * w_node, s1 = n_leaf * w_leaf, s2 = n_sh * w_shade, s3 = n_en * w_end;
            if (s0 >= s1 && s0 >= s2 && s0 >= s3 && n_node) {
                stage = 0;
            } else if (s1 >= s2 && s1 >= s3 && n_leaf) {
                stage = 1;
            } else if (s2 >= s3 && n_sh) {
                stage = 2;
            } else if (n_en) {
                stage = 3;
            } else {
                stage = n_node ? 0 : (n_leaf ? 1 : 2);
            }
        }

        if (stage == 0) {
            // ---- NODE: two child-box slab tests, descend / push / pop ----------------------------------------
            if (t.cur >= 0) {
                if (COUNT) ctr.n_node++;
                trav_node_step(t, stack, f, 0.01f);                              // tmin :754
            }
        } else if (stage == 1) {
            // ---- LEAF: one triangle test; a shadow ray ends on its first hit (traceOcclusion :651-684) -------
            if (t.cur < 0 && t.cur != kTravDone) {
                if (COUNT) ctr.n_tri++;
                trav_leaf_step(t, stack, f, phase == PH_SHADOW, 0.01f, 1e16f);
            }
        } else if (stage == 2) {
            // ---- SHADE: __closesthit__diffuse__ch :866-1031 for lanes whose radiance ray found a hit --------
            if (t.cur == kTravDone && phase == PH_RAD && t.prim != kNoPrim) {
                if (PROBE && left == lc.spp && p.depth == 0) {
                    const uint32_t image_index = iy * lc.width + ix;
                    if (lc.dbg_primary_prim) lc.dbg_primary_prim[image_index] = (int32_t)t.prim;
                    if (lc.dbg_primary_t) lc.dbg_primary_t[image_index] = t.tfar;
                }
                ShadowQuery sq;
                int hit_type, mat_id;
                float3 emission;
                shade_hit(lc, f, l_area, t.prim, t.tfar, p, emission, rad, done, sq, hit_type, mat_id);
                result = result + emission;                                                    // :760
                if (COUNT) {
                    ctr.n_hit0 += hit_type == 0;
                    ctr.n_hit1 += hit_type == 1;
                    ctr.n_hit2 += hit_type == 2;
                }
                if (sq.want) {
                    // traceOcclusion :651-684: "anything in (0.01, Ldist - 0.01)" 
                    if (COUNT) ctr.n_shadow++;
                    add = sq.add;
                    trav_begin(t, sq.org, sq.dir, sq.tmax, true);
                    phase = PH_SHADOW;
                } else {
                    phase = PH_END;
                }
            }
        } else {
            // ---- END: finish the segment (__miss__ms :833-847, shadow result, :760-778), start the next ray -----
            if (t.cur == kTravDone && phase != PH_IDLE && !(phase == PH_RAD && t.prim != kNoPrim)) {
                if (phase == PH_RAD) {                  // miss
                    if (PROBE && left == lc.spp && p.depth == 0) {
                        const uint32_t image_index = iy * lc.width + ix;
                        if (lc.dbg_primary_prim) lc.dbg_primary_prim[image_index] = -1;
                        if (lc.dbg_primary_t) lc.dbg_primary_t[image_index] = 0.0f;
                    }
                    rad = lc.background;
                    done = true;
                    result = result + f3(0.0f);                                                // emissionColor = 0, :760
                    if (COUNT) ctr.n_miss++;
                } else if (phase == PH_SHADOW) {
                    if (t.prim == kNoPrim) rad = rad + add;                                    // unoccluded, :1019-1024
                }
                result = result + rad * p.att;                                                 // :761
                const float pr = dot(p.att, f3(0.30f, 0.59f, 0.11f));                          // :763
                const bool rr = rnd(p.seed) > pr;                                              // :765
                bool alive = true;
                if (done || rr || (uint32_t)p.depth >= lc.max_depth) {                         // :767
                    if (--left == 0u) {
                        const uint32_t image_index = iy * lc.width + ix;
                        write_pixel(lc, image_index, result);
                        if (PROBE) {
                            if (lc.dbg_final_path_seed) lc.dbg_final_path_seed[image_index] = p.seed;
                            if (lc.dbg_final_pixel_seed) lc.dbg_final_pixel_seed[image_index] = pixel_seed;
                        }
                        alive = take_pixel(lc, n_work, atomicAdd(lc.work_counter, 1u), ix, iy);
                        if (alive) {
                            pixel_seed = tea4(iy * lc.width + ix, lc.frame);                   // :721
                            result = f3(0.0f);
                            left = lc.spp;
                        }
                    }
                    if (alive) {
                        generate_camera_ray(lc, ix, iy, pixel_seed, p);                        // next sample, in place
                        if (COUNT) ctr.n_samples++;
                        if (PROBE && left == lc.spp) {
                            const uint32_t image_index = iy * lc.width + ix;
                            if (lc.dbg_seed_after_jitter) lc.dbg_seed_after_jitter[image_index] = pixel_seed;
                            if (lc.dbg_primary_dir) {
                                lc.dbg_primary_dir[3 * image_index + 0] = p.d.x;
                                lc.dbg_primary_dir[3 * image_index + 1] = p.d.y;
                                lc.dbg_primary_dir[3 * image_index + 2] = p.d.z;
                            }
                        }
                    }
                } else {
                    p.att = f3(safe_divide(p.att.x, pr), safe_divide(p.att.y, pr), safe_divide(p.att.z, pr));   // :773
                    ++p.depth;
                }
                if (alive) {
                    if (COUNT) ctr.n_radiance++;
                    trav_begin(t, p.o, p.d, 1e16f, have_tris);                                 // traceRadiance :577-635
                    phase = PH_RAD;
                } else {
                    phase = PH_IDLE;
                }
            }
        }
    }

    if (COUNT && lc.counters) {
        flush_counter(lc.counters + 0, ctr.n_samples);
        flush_counter(lc.counters + 1, ctr.n_radiance);
        flush_counter(lc.counters + 2, ctr.n_shadow);
        flush_counter(lc.counters + 3, ctr.n_node);
        flush_counter(lc.counters + 4, ctr.n_tri);
        flush_counter(lc.counters + 5, ctr.n_hit0);
        flush_counter(lc.counters + 6, ctr.n_hit1);
        flush_counter(lc.counters + 7, ctr.n_hit2);
        flush_counter(lc.counters + 8, ctr.n_miss);
    }
}

// ---- batch ray queries (parity probe) --------------------------------------------------------------------------
__global__ void __launch_bounds__(256) trace_rays_kernel(BvhView bvh, const float* __restrict__ org, const float* __restrict__ dir,
                                                         const float* __restrict__ tmin, const float* __restrict__ tmax,
                                                         uint64_t n, int32_t* prim_out, float* t_out, uint8_t* any_out,
                                                         unsigned long long* counters2)
{
    const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    TraceCount tc = {0, 0};
    FetchGlobal f;
    f.nodes = bvh.nodes;
    f.tris = bvh.tris;
    f.shade = nullptr;
    f.mats = nullptr;
    if (i < n) {
        const float3 o = f3(org[3 * i], org[3 * i + 1], org[3 * i + 2]);
        const float3 d = f3(dir[3 * i], dir[3 * i + 1], dir[3 * i + 2]);
        if (prim_out || t_out) {
            float t;
            uint32_t prim;
            const bool hit = trace_closest<true>(f, bvh.n_tris != 0, o, d, tmin[i], tmax[i], t, prim, tc);
            if (prim_out) prim_out[i] = hit ? (int32_t)prim : -1;
            if (t_out) t_out[i] = hit ? t : 0.0f;
        }
        if (any_out) any_out[i] = trace_any<true>(f, bvh.n_tris != 0, o, d, tmin[i], tmax[i], tc) ? 1 : 0;
    }
    if (counters2) {
        flush_counter(counters2 + 0, tc.n_node);
        flush_counter(counters2 + 1, tc.n_tri);
    }
}

cudaError_t trace_rays_device(Scene& s, const float* d_org, const float* d_dir, const float* d_tmin, const float* d_tmax,
                              uint64_t n, int32_t* d_prim, float* d_t, uint8_t* d_any, unsigned long long* d_counters2,
                              cudaStream_t stream)
{
    if (n == 0) return cudaSuccess;
    BvhView bvh{scene_nodes(s), s.d_tris, s.n_tris};
    const unsigned grid = (unsigned)((n + 255) / 256);
    trace_rays_kernel<<<grid, 256, 0, stream>>>(bvh, d_org, d_dir, d_tmin, d_tmax, n, d_prim, d_t, d_any, d_counters2);
    return cudaGetLastError();
}

// ---- pixel partition: tiles of 8x4 pixels, row-major; tile L has class L % n_ranks; a launch owns `span` consecutive
// classes starting at `rank` (span = 1: tile k of rank r is image tile r + k * n_ranks) ----------------------------------
uint32_t partition_owned_tiles(uint32_t width, uint32_t height, uint32_t rank, uint32_t n_ranks, uint32_t span)
{
    const uint32_t tiles = ((width + 7) / 8) * ((height + 3) / 4);
    uint32_t owned = 0;
    for (uint32_t c = rank; c < rank + span && c < n_ranks; ++c) owned += c < tiles ? (tiles - c + n_ranks - 1) / n_ranks : 0;
    return owned;
}
uint32_t partition_max_owned_tiles(uint32_t width, uint32_t height, uint32_t n_ranks)
{
    const uint32_t tiles = ((width + 7) / 8) * ((height + 3) / 4);
    return (tiles + n_ranks - 1) / n_ranks;
}

__global__ void __launch_bounds__(256) partition_pack_kernel(const float4* __restrict__ accum_full, const uchar4* __restrict__ frame_full,
                                                             float4* __restrict__ accum_compact, uchar4* __restrict__ frame_compact,
                                                             uint32_t width, uint32_t height, uint32_t tiles_x, uint32_t rank,
                                                             uint32_t n_ranks, uint32_t capacity_tiles, uint32_t owned)
{
    const uint32_t slot = blockIdx.x * blockDim.x + threadIdx.x;   // k * 32 + lane
    if (slot >= capacity_tiles * 32u) return;
    const uint32_t k = slot >> 5, lane = slot & 31u;
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
    uchar4 c = make_uchar4(0, 0, 0, 0);
    if (k < owned) {
        const uint32_t tile = rank + k * n_ranks;
        const uint32_t ix = (tile % tiles_x) * 8u + (lane & 7u), iy = (tile / tiles_x) * 4u + (lane >> 3);
        if (ix < width && iy < height) {
            a = accum_full[iy * width + ix];
            if (frame_full) c = frame_full[iy * width + ix];
        }
    }
    accum_compact[slot] = a;
    if (frame_compact) frame_compact[slot] = c;
}

__global__ void __launch_bounds__(256) partition_unpack_kernel(const float4* __restrict__ accum_compact_all,
                                                               const uchar4* __restrict__ frame_compact_all,
                                                               float4* __restrict__ accum_full, uchar4* __restrict__ frame_full,
                                                               uint32_t width, uint32_t height, uint32_t tiles_x, uint32_t n_tiles,
                                                               uint32_t n_ranks, uint32_t capacity_tiles)
{
    const uint32_t idx = blockIdx.x * blockDim.x + threadIdx.x;    // tile * 32 + lane
    if (idx >= n_tiles * 32u) return;
    const uint32_t tile = idx >> 5, lane = idx & 31u;
    const uint32_t ix = (tile % tiles_x) * 8u + (lane & 7u), iy = (tile / tiles_x) * 4u + (lane >> 3);
    if (ix >= width || iy >= height) return;
    const uint32_t r = tile % n_ranks, k = tile / n_ranks;
    const size_t slot = ((size_t)r * capacity_tiles + k) * 32u + lane;
    accum_full[iy * width + ix] = accum_compact_all[slot];
    if (frame_full && frame_compact_all) frame_full[iy * width + ix] = frame_compact_all[slot];
}

cudaError_t partition_pack(const void* accum_full, const void* frame_full, void* accum_compact, void* frame_compact,
                           uint32_t width, uint32_t height, uint32_t rank, uint32_t n_ranks, cudaStream_t stream)
{
    const uint32_t tiles_x = (width + 7) / 8;
    const uint32_t cap = partition_max_owned_tiles(width, height, n_ranks);
    const uint32_t owned = partition_owned_tiles(width, height, rank, n_ranks, 1);
    const unsigned grid = (cap * 32u + 255u) / 256u;
    partition_pack_kernel<<<grid, 256, 0, stream>>>((const float4*)accum_full, (const uchar4*)frame_full, (float4*)accum_compact,
                                                    (uchar4*)frame_compact, width, height, tiles_x, rank, n_ranks, cap, owned);
    return cudaGetLastError();
}

cudaError_t partition_unpack(const void* accum_compact_all, const void* frame_compact_all, void* accum_full,
                             void* frame_full, uint32_t width, uint32_t height, uint32_t n_ranks, cudaStream_t stream)
{
    const uint32_t tiles_x = (width + 7) / 8;
    const uint32_t n_tiles = tiles_x * ((height + 3) / 4);
    const uint32_t cap = partition_max_owned_tiles(width, height, n_ranks);
    const unsigned grid = (n_tiles * 32u + 255u) / 256u;
    partition_unpack_kernel<<<grid, 256, 0, stream>>>((const float4*)accum_compact_all, (const uchar4*)frame_compact_all,
                                                      (float4*)accum_full, (uchar4*)frame_full, width, height, tiles_x, n_tiles,
                                                      n_ranks, cap);
    return cudaGetLastError();
}

// ---- launch ---------------------------------------------------------------------------------------------------------
static inline float3 to3(const pt_float3& v) { return make_float3(v.x, v.y, v.z); }

// Work order.  A pixel whose camera rays cannot reach the scene costs one node test per sample; one that looks into
// the scene costs a whole path.  Paths take pixels from a counter, so the order decides the tail of the launch: with
// the expensive tiles first the cheap ones fill the gaps at the end (longest-processing-time-first).  A tile counts as
// expensive when the camera ray through its centre or one of its corners hits the scene's bounding box (host, double
// precision; only the ORDER depends on it, never a pixel value).  Returns the device array or null (natural order).
static const uint32_t* tile_order_update(Scene& s, const pt_params& p, uint32_t rank, uint32_t n_ranks, uint32_t span,
                                         uint32_t owned, cudaStream_t stream)
{
    if (!s.tile_order_enabled || s.n_tris == 0 || owned == 0) return nullptr;
    const float key[17] = {(float)span, (float)p.width, (float)p.height, (float)rank, (float)n_ranks, p.cameraEye.x, p.cameraEye.y,
                           p.cameraEye.z, p.cameraU.x, p.cameraU.y, p.cameraU.z, p.cameraV.x, p.cameraV.y, p.cameraV.z,
                           p.cameraW.x, p.cameraW.y, p.cameraW.z};
    // a few orders are kept (a multi-GPU image alternates between whole frames and class ranges): a hit costs nothing and,
    // unlike the upload below (pageable source: the copy waits for the stream), does not stall the host
    TileOrder* slot = &s.tile_orders[0];
    for (TileOrder& t : s.tile_orders) {
        if (t.d && t.count == owned && memcmp(key, t.key, sizeof(key)) == 0) {
            t.stamp = ++s.tile_order_clock;
            return t.d;
        }
        if (t.stamp < slot->stamp) slot = &t;
    }
    const uint32_t tiles_x = (p.width + 7) / 8;
    const double lo[3] = {s.h_bounds[0], s.h_bounds[1], s.h_bounds[2]}, hi[3] = {s.h_bounds[3], s.h_bounds[4], s.h_bounds[5]};
    const double o[3] = {p.cameraEye.x, p.cameraEye.y, p.cameraEye.z};
    auto hits = [&](double px, double py) {
        const double dx = 2.0 * (px / (double)p.width) - 1.0, dy = 2.0 * (py / (double)p.height) - 1.0;   // :732-737
        const double d[3] = {dx * p.cameraU.x + dy * p.cameraV.x + p.cameraW.x, dx * p.cameraU.y + dy * p.cameraV.y + p.cameraW.y,
                             dx * p.cameraU.z + dy * p.cameraV.z + p.cameraW.z};
        double tn = 0.0, tf = 1e300;
        for (int a = 0; a < 3; ++a) {
            if (d[a] == 0.0) {
                if (o[a] < lo[a] || o[a] > hi[a]) return false;
                continue;
            }
            double t0 = (lo[a] - o[a]) / d[a], t1 = (hi[a] - o[a]) / d[a];
            if (t0 > t1) std::swap(t0, t1);
            tn = std::max(tn, t0);
            tf = std::min(tf, t1);
        }
        return tn <= tf;
    };
    // the slot's own pinned staging array: the upload is then truly asynchronous (a pageable source would make the copy wait
    // for everything queued on the stream -- with several devices driven by one thread that serialises them)
    if (slot->capacity < owned) {
        if (slot->d) cudaFree(slot->d);
        if (slot->h) cudaFreeHost(slot->h);
        slot->d = nullptr;
        slot->h = nullptr;
        slot->capacity = 0;
        slot->count = 0;
        if (cudaMalloc((void**)&slot->d, sizeof(uint32_t) * owned) != cudaSuccess) return nullptr;
        if (cudaMallocHost((void**)&slot->h, sizeof(uint32_t) * owned) != cudaSuccess) {
            cudaFree(slot->d);
            slot->d = nullptr;
            return nullptr;
        }
        slot->capacity = owned;
    }
    if (!slot->copied && cudaEventCreateWithFlags(&slot->copied, cudaEventDisableTiming) != cudaSuccess) return nullptr;
    if (slot->count) cudaEventSynchronize(slot->copied);   // the previous upload from this staging array has been read
    slot->count = 0;
    uint32_t* order = slot->h;
    uint32_t front = 0;
    std::vector<uint32_t> cheap;
    cheap.reserve(owned);
    for (uint32_t k = 0; k < owned; ++k) {
        const uint32_t tile = owned_tile(k, rank, span, n_ranks);
        const double x0 = (tile % tiles_x) * 8.0, y0 = (tile / tiles_x) * 4.0;
        const bool expensive = hits(x0 + 4.0, y0 + 2.0) || hits(x0, y0) || hits(x0 + 8.0, y0) || hits(x0, y0 + 4.0) ||
                               hits(x0 + 8.0, y0 + 4.0);
        if (expensive) order[front++] = tile;
        else cheap.push_back(tile);
    }
    for (uint32_t t : cheap) order[front++] = t;
    if (cudaMemcpyAsync(slot->d, order, sizeof(uint32_t) * owned, cudaMemcpyHostToDevice, stream) != cudaSuccess) return nullptr;
    cudaEventRecord(slot->copied, stream);
    slot->count = owned;
    slot->stamp = ++s.tile_order_clock;
    memcpy(slot->key, key, sizeof(key));
    return slot->d;
}

template <bool COUNT, bool PROBE, bool SMEM>
static cudaError_t launch_persistent(Scene& s, const LaunchConst& lc, size_t smem_bytes, cudaStream_t stream)
{
    auto kernel = render_persistent_kernel<COUNT, PROBE, SMEM>;
    cudaError_t e;
    if (smem_bytes > 0) {
        e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes);
        if (e != cudaSuccess) return e;
    }
    int per_sm = 0;
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, kBlock, smem_bytes);
    if (e != cudaSuccess) return e;
    if (per_sm < 1) per_sm = 1;
    int dev = 0, sms = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    // persistent grid: every resident warp pulls tiles until the counter runs out
    const uint32_t warps_needed = lc.n_owned_tiles;
    uint32_t grid = (uint32_t)(sms * per_sm);
    const uint32_t grid_needed = (warps_needed + (kBlock / 32) - 1) / (kBlock / 32);
    if (grid > grid_needed) grid = grid_needed;
    if (grid == 0) grid = 1;
    e = cudaMemsetAsync(s.d_work_counter, 0, sizeof(uint32_t), stream);
    if (e != cudaSuccess) return e;
    kernel<<<grid, kBlock, smem_bytes, stream>>>(lc);
    s.last_grid = grid;
    s.last_block = kBlock;
    s.last_smem = (uint32_t)smem_bytes;
    s.last_kernels = 1;
    return cudaGetLastError();
}

// ---- frame fold: the running average of pathTracerPrograms.cu:803-811 applied to per-launch means, in frame order ----
struct FoldArgs {
    const float4* means[kMaxFoldFrames];
    uint32_t n_frames, first_frame;
};

__global__ void __launch_bounds__(256) fold_frames_kernel(float4* __restrict__ accum, uchar4* __restrict__ frame_out,
                                                          const __grid_constant__ FoldArgs a, uint32_t n_pixels)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_pixels) return;
    float3 acc;
    uint32_t k = 0;
    if (a.first_frame == 0u) {                       // frame 0 overwrites (:803: only frames > 0 read the buffer)
        const float4 x = a.means[0][i];
        acc = f3(x.x, x.y, x.z);
        k = 1;
    } else {
        const float4 p = accum[i];
        acc = f3(p.x, p.y, p.z);
    }
    for (; k < a.n_frames; ++k) {
        const float w = 1.0f / (float)((int)(a.first_frame + k) + 1);
        const float4 x4 = a.means[k][i];
        const float3 x = f3(x4.x, x4.y, x4.z);
        acc = acc + w * (x - acc);                   // lerp(prev, accum, a), vec_math.h:500 -- same order as write_pixel
    }
    accum[i] = make_float4(acc.x, acc.y, acc.z, 1.0f);
    if (frame_out) frame_out[i] = make_color(acc);
}

cudaError_t fold_frames(void* accum, void* frame_out, const void* const* means, uint32_t first_frame, uint32_t n_frames,
                        uint32_t width, uint32_t height, cudaStream_t stream)
{
    FoldArgs a = {};
    for (uint32_t k = 0; k < n_frames; ++k) a.means[k] = (const float4*)means[k];
    a.n_frames = n_frames;
    a.first_frame = first_frame;
    const uint32_t n = width * height;
    fold_frames_kernel<<<(n + 255u) / 256u, 256, 0, stream>>>((float4*)accum, (uchar4*)frame_out, a, n);
    return cudaGetLastError();
}

cudaError_t render_launch(Scene& s, const pt_params& params, uint32_t rank, uint32_t n_ranks, cudaStream_t stream,
                          const LaunchConst* probe_overrides, bool raw_mean, uint32_t span)
{
    if (span == 0 || rank + span > n_ranks) return cudaErrorInvalidValue;
    LaunchConst lc = {};
    lc.raw_mean = raw_mean ? 1 : 0;
    lc.frame = params.currentFrameIdx;
    lc.width = params.width;
    lc.height = params.height;
    lc.spp = params.samplesPerPixel;
    lc.max_depth = params.maxDepth;
    lc.eye = to3(params.cameraEye);
    lc.U = to3(params.cameraU);
    lc.V = to3(params.cameraV);
    lc.W = to3(params.cameraW);
    lc.l_corner = to3(params.areaLight.corner);
    lc.l_v1 = to3(params.areaLight.v1);
    lc.l_v2 = to3(params.areaLight.v2);
    lc.l_normal = to3(params.areaLight.normal);
    lc.l_emission = to3(params.areaLight.emission);
    lc.use_nee = params.useDirectLighting ? 1 : 0;
    lc.use_is = params.useImportanceSampling ? 1 : 0;
    lc.background = make_float3(s.background[0], s.background[1], s.background[2]);
    lc.accum = (float4*)params.accumulationBuffer;
    lc.frame_buffer = raw_mean ? nullptr : (uchar4*)params.frameBuffer;
    lc.bvh = BvhView{scene_nodes(s), s.d_tris, s.n_tris};
    lc.n_nodes = s.n_nodes;
    lc.n_mats = s.n_mats;
    lc.tri_shade = s.d_tri_shade;
    lc.mats = s.d_mats;
    lc.tiles_x = (params.width + 7) / 8;
    lc.tiles_y = (params.height + 3) / 4;
    lc.rank = rank;
    lc.n_ranks = n_ranks;
    lc.span = span;
    lc.n_owned_tiles = partition_owned_tiles(params.width, params.height, rank, n_ranks, span);
    lc.tile_order = tile_order_update(s, params, rank, n_ranks, span, lc.n_owned_tiles, stream);
    lc.counters = s.count_enabled ? s.d_counters : nullptr;
    lc.work_counter = s.d_work_counter;
    lc.error_flag = s.d_counters + 15;
    for (int i = 0; i < 5; ++i) lc.sched[i] = s.sched[i];
    bool probe = false;
    if (probe_overrides) {
        lc.dbg_seed_after_jitter = probe_overrides->dbg_seed_after_jitter;
        lc.dbg_primary_dir = probe_overrides->dbg_primary_dir;
        lc.dbg_primary_prim = probe_overrides->dbg_primary_prim;
        lc.dbg_primary_t = probe_overrides->dbg_primary_t;
        lc.dbg_final_path_seed = probe_overrides->dbg_final_path_seed;
        lc.dbg_final_pixel_seed = probe_overrides->dbg_final_pixel_seed;
        probe = true;
    }
    if (lc.n_owned_tiles == 0) return cudaSuccess;

    // pipeline 2 (wavefront.cu) whenever it can hold the scene and the launch, unless pipeline 1 was asked for.  Its path
    // record packs the depth into 8 bits and the pixel into 16 + 16 bits: deeper paths (maxDepth > 255 never terminates a
    // path that Russian roulette keeps alive) and larger images run on pipeline 1, which has neither limit.
    const bool wf_fits = params.maxDepth <= 255u && params.width <= 65535u && params.height <= 65535u;
    if (s.pipeline != 1 && wf_fits && wavefront_supported(s)) return wavefront_launch(s, lc, probe, stream);
    if (s.pipeline == 2) return cudaErrorInvalidConfiguration;
    s.last_trav = 0;

    const size_t staged = staged_scene_float4s(s.n_nodes, s.n_tris, s.n_mats) * sizeof(float4);
    const bool smem = s.n_tris > 0 && staged <= kSmemSceneLimit;
    const bool count = s.count_enabled;
    if (probe) {
        return smem ? launch_persistent<false, true, true>(s, lc, staged, stream)
                    : launch_persistent<false, true, false>(s, lc, 0, stream);
    }
    if (count) {
        return smem ? launch_persistent<true, false, true>(s, lc, staged, stream)
                    : launch_persistent<true, false, false>(s, lc, 0, stream);
    }
    return smem ? launch_persistent<false, false, true>(s, lc, staged, stream)
                : launch_persistent<false, false, false>(s, lc, 0, stream);
}

}  // namespace ptb
