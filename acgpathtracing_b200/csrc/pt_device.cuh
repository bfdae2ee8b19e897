// pt_device.cuh -- device-side building blocks shared by the LBVH build and the render kernels:
// float3 helpers with the reference's rounding order, the reference's RNG, and the intersection spec.
//
// The library is compiled with -fmad=false: the compiler never fuses a*b+c, so plain C++ float
// expressions below round exactly like the reference source does under IEEE rules (and like the CPU
// oracle compiled with -ffp-contract=off).  Where a fused multiply-add is part of OUR spec (ray/box,
// ray/triangle -- arithmetic the reference leaves to the OptiX driver) it is written as __fmaf_rn.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "pt_detmath.h"

namespace ptd {

// ---- float3 helpers: sutil/vec_math.h semantics (operation order matters) -------------------------------
__device__ __forceinline__ float3 f3(float x, float y, float z) { return make_float3(x, y, z); }
__device__ __forceinline__ float3 f3(float s) { return make_float3(s, s, s); }
__device__ __forceinline__ float3 operator+(float3 a, float3 b) { return f3(a.x + b.x, a.y + b.y, a.z + b.z); }
__device__ __forceinline__ float3 operator-(float3 a, float3 b) { return f3(a.x - b.x, a.y - b.y, a.z - b.z); }
__device__ __forceinline__ float3 operator-(float3 a) { return f3(-a.x, -a.y, -a.z); }
__device__ __forceinline__ float3 operator*(float3 a, float3 b) { return f3(a.x * b.x, a.y * b.y, a.z * b.z); }
__device__ __forceinline__ float3 operator*(float3 a, float s) { return f3(a.x * s, a.y * s, a.z * s); }
__device__ __forceinline__ float3 operator*(float s, float3 a) { return f3(s * a.x, s * a.y, s * a.z); }
__device__ __forceinline__ float3 operator/(float3 a, float3 b) { return f3(a.x / b.x, a.y / b.y, a.z / b.z); }   // vec_math.h:479
__device__ __forceinline__ float3 div_by_scalar(float3 a, float s)      // vec_math.h:483-487: multiply by reciprocal
{
    float inv = 1.0f / s;
    return a * inv;
}
__device__ __forceinline__ float dot(float3 a, float3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }            // :527
__device__ __forceinline__ float3 cross(float3 a, float3 b)                                                        // :533
{
    return f3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}
__device__ __forceinline__ float length(float3 v) { return sqrtf(dot(v, v)); }                                     // :539
__device__ __forceinline__ float3 normalize(float3 v)                                                              // :545
{
    float invLen = 1.0f / sqrtf(dot(v, v));
    return v * invLen;
}
__device__ __forceinline__ float3 reflect(float3 i, float3 n) { return i - (2.0f * n) * dot(n, i); }               // :558
__device__ __forceinline__ float clampf(float f, float a, float b) { return fmaxf(a, fminf(f, b)); }               // :114

// ---- RNG: cuda/random.h ------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t tea4(uint32_t val0, uint32_t val1)      // random.h:32-46, N = 4
{
    uint32_t v0 = val0, v1 = val1, s0 = 0;
#pragma unroll
    for (int n = 0; n < 4; ++n) {
        s0 += 0x9e3779b9u;
        v0 += ((v1 << 4) + 0xa341316cu) ^ (v1 + s0) ^ ((v1 >> 5) + 0xc8013ea4u);
        v1 += ((v0 << 4) + 0xad90777du) ^ (v0 + s0) ^ ((v0 >> 5) + 0x7e95761eu);
    }
    return v0;
}
__device__ __forceinline__ float rnd(uint32_t& prev)                        // random.h:49-67
{
    prev = 1664525u * prev + 1013904223u;
    // (float)(prev & 0xffffff) / 2^24: both exact, so a multiply by 2^-24 is the same value
    return (float)(prev & 0x00ffffffu) * 5.9604644775390625e-8f;
}

// ---- pixel partition: the k-th tile of a launch that owns classes first .. first + span - 1 of n (tile L has class L % n) ----
__host__ __device__ __forceinline__ uint32_t owned_tile(uint32_t k, uint32_t first, uint32_t span, uint32_t n)
{
    return span == 1u ? first + k * n : (k / span) * n + first + k % span;
}

// ---- acceleration-structure records (DESIGN.md "LBVH spec") ---------------------------------------------------
// Node, 64 B = 4 x float4:  q0 = (lo0.xyz, child0)  q1 = (hi0.xyz, child1)  q2 = (lo1.xyz, 0)  q3 = (hi1.xyz, 0)
//   child >= 0: internal node index; child < 0: leaf, ~child = slot in the leaf-ordered triangle array.
// Triangle, 48 B = 3 x float4: t0 = (v0.xyz, primIdx)  t1 = (e1.xyz, materialId)  t2 = (e2.xyz, 0),  e1 = v1-v0, e2 = v2-v0.
struct BvhView {
    const float4* nodes;   // 4 per node
    const float4* tris;    // 3 per leaf
    uint32_t n_tris;
};

// Where the scene records are read from.  Every kernel takes one of these "fetch" policies:
//   FetchGlobal      global memory through the read-only path (L1/L2), one copy
//   FetchShared      one copy staged in shared memory (plain LDS)
//   FetchSharedBanked `1 << SHIFT` interleaved copies in shared memory: element e of copy k lives at float4 index
//                    (e << SHIFT) + k and lane l reads copy l & (copies - 1).  With 8 copies the 8 lanes of every
//                    quarter-warp hit 8 different 16-byte bank groups whatever elements they ask for, so a
//                    divergent LDS.128 costs exactly its 4 data wavefronts.
// 256-bit read-only load (LDG.E.256, sm_100+): one whole 32-byte sector per lane and instruction.  A divergent fetch of
// 64-byte records costs half the L1 requests of LDG.128 x4 (tools/fetchbench.cu: +45-50 % records/s from L2, +22 % from
// DRAM, never slower).  `p` must be 32-byte aligned.
__device__ __forceinline__ void ldg256(const float4* p, float4& a, float4& b)
{
    asm("ld.global.nc.v8.f32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=f"(a.x), "=f"(a.y), "=f"(a.z), "=f"(a.w), "=f"(b.x), "=f"(b.y), "=f"(b.z), "=f"(b.w)
                 : "l"(p));
}

struct FetchGlobal {
    const float4 *nodes, *tris, *shade, *mats;
    // quantised nodes (lbvh.cu quantize_nodes_kernel), 32 bytes = one LDG.E.256 per node: twelve 16-bit plane positions on
    // the scene's bounding-box grid (w0 = lo0.x | lo0.y << 16, w1 = lo0.z | hi0.x << 16, w2 = hi0.y | hi0.z << 16, w3..w5 the
    // same for child 1) and the two child references (w6, w7).  A plane is (2^23 + q) * step + base in ONE fused multiply-add
    // -- 2^23 + q is the float with bit pattern 0x4b000000 | q, so no integer conversion is needed -- and that is the
    // expression the build checks its rounding with: lo planes never above, hi planes never below the exact box.
    const float4* qnodes;
    float3 q_lo, q_step;                                  // base = grid origin - 2^23 * step, and the step, per axis
    bool prefetch = true;                                 // pull deferred children towards L2; the wavefront kernel sets it from
                                                          // WfTune::prefetch (off: measured), the fallback kernels keep it
    __device__ __forceinline__ float qf_lo(uint32_t w) const { return __uint_as_float(__byte_perm(w, 0x4b000000u, 0x7410)); }
    __device__ __forceinline__ float qf_hi(uint32_t w) const { return __uint_as_float(__byte_perm(w, 0x4b000000u, 0x7432)); }
    __device__ __forceinline__ void qnode4(int i, float4& q0, float4& q1, float4& q2, float4& q3) const
    {
        float4 a, b;
        ldg256(qnodes + 2 * (size_t)i, a, b);
        const uint32_t w0 = __float_as_uint(a.x), w1 = __float_as_uint(a.y), w2 = __float_as_uint(a.z), w3 = __float_as_uint(a.w);
        const uint32_t w4 = __float_as_uint(b.x), w5 = __float_as_uint(b.y);
        q0 = make_float4(__fmaf_rn(qf_lo(w0), q_step.x, q_lo.x), __fmaf_rn(qf_hi(w0), q_step.y, q_lo.y), __fmaf_rn(qf_lo(w1), q_step.z, q_lo.z), b.z);
        q1 = make_float4(__fmaf_rn(qf_hi(w1), q_step.x, q_lo.x), __fmaf_rn(qf_lo(w2), q_step.y, q_lo.y), __fmaf_rn(qf_hi(w2), q_step.z, q_lo.z), b.w);
        q2 = make_float4(__fmaf_rn(qf_lo(w3), q_step.x, q_lo.x), __fmaf_rn(qf_hi(w3), q_step.y, q_lo.y), __fmaf_rn(qf_lo(w4), q_step.z, q_lo.z), 0.0f);
        q3 = make_float4(__fmaf_rn(qf_hi(w4), q_step.x, q_lo.x), __fmaf_rn(qf_lo(w5), q_step.y, q_lo.y), __fmaf_rn(qf_hi(w5), q_step.z, q_lo.z), 0.0f);
    }
    __device__ __forceinline__ void prefetch_qchild(int c) const
    {
        if (!prefetch) return;
        const void* p = c >= 0 ? (const void*)(qnodes + 2 * (size_t)c) : (const void*)(tris + 3 * (~c));
        asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
    }
    // 4-wide quantised nodes (lbvh.cu build_wide_nodes), 64 bytes = two LDG.E.256 per node, read through `qnodes`: four child
    // boxes on the same grid, one word per child and axis (lo | hi << 16; words 3k .. 3k+2 = x, y, z of child k), and four child
    // references (words 12..15).  The planes are those of the binary quantised nodes; an empty slot repeats child 0 behind an
    // inside-out box (lo = 65535, hi = 0 on every axis: never entered before it is left -- and harmless if it ever were).
    __device__ __forceinline__ void wnode16(int i, float4& a, float4& b, float4& c, float4& d) const
    {
        ldg256(qnodes + 4 * (size_t)i, a, b);
        ldg256(qnodes + 4 * (size_t)i + 2, c, d);
    }
    __device__ __forceinline__ void prefetch_wchild(int c) const
    {
        const void* p = c >= 0 ? (const void*)(qnodes + 4 * (size_t)c) : (const void*)(tris + 3 * (~c));
        asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
    }
    __device__ __forceinline__ void node4(int i, float4& q0, float4& q1, float4& q2, float4& q3) const
    {
        ldg256(nodes + 4 * (size_t)i, q0, q1);
        ldg256(nodes + 4 * (size_t)i + 2, q2, q3);
    }
    // 48-byte triangle record in two loads instead of three: records start 32-byte aligned in even slots and at 16 mod 32
    // in odd ones, so one half of the record is always a whole aligned sector pair
    __device__ __forceinline__ void tri3(int i, float4& t0, float4& t1, float4& t2) const
    {
        const float4* p = tris + 3 * (size_t)i;
        if (i & 1) {
            t0 = __ldg(p);
            ldg256(p + 1, t1, t2);
        } else {
            ldg256(p, t0, t1);
            t2 = __ldg(p + 2);
        }
    }
    // a deferred child will be fetched after the subtree of its sibling: start pulling its record towards L2/L1 now
    __device__ __forceinline__ void prefetch_child(int c) const
    {
        if (!prefetch) return;
        const void* p = c >= 0 ? (const void*)(nodes + 4 * c) : (const void*)(tris + 3 * (~c));
        asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
    }
    __device__ __forceinline__ float4 node(int i, int q) const { return __ldg(nodes + 4 * i + q); }
    __device__ __forceinline__ float4 tri(int i, int q) const { return __ldg(tris + 3 * i + q); }
    __device__ __forceinline__ float4 tri_shade(uint32_t prim) const { return __ldg(shade + prim); }
    __device__ __forceinline__ float4 mat(int m, int q) const { return __ldg(mats + 3 * m + q); }
};
struct FetchShared {
    const float4 *nodes, *tris, *shade, *mats;
    __device__ __forceinline__ void prefetch_child(int) const {}
    __device__ __forceinline__ void tri3(int i, float4& t0, float4& t1, float4& t2) const { t0 = tri(i, 0), t1 = tri(i, 1), t2 = tri(i, 2); }
    __device__ __forceinline__ void node4(int i, float4& q0, float4& q1, float4& q2, float4& q3) const
    {
        q0 = node(i, 0), q1 = node(i, 1), q2 = node(i, 2), q3 = node(i, 3);
    }
    __device__ __forceinline__ float4 node(int i, int q) const { return nodes[4 * i + q]; }
    __device__ __forceinline__ float4 tri(int i, int q) const { return tris[3 * i + q]; }
    __device__ __forceinline__ float4 tri_shade(uint32_t prim) const { return shade[prim]; }
    __device__ __forceinline__ float4 mat(int m, int q) const { return mats[3 * m + q]; }
};
template <int SHIFT>
struct FetchSharedBanked {
    const float4 *nodes, *tris, *shade, *mats;   // each already offset by this lane's copy index
    __device__ __forceinline__ void prefetch_child(int) const {}
    __device__ __forceinline__ void tri3(int i, float4& t0, float4& t1, float4& t2) const { t0 = tri(i, 0), t1 = tri(i, 1), t2 = tri(i, 2); }
    __device__ __forceinline__ void node4(int i, float4& q0, float4& q1, float4& q2, float4& q3) const
    {
        q0 = node(i, 0), q1 = node(i, 1), q2 = node(i, 2), q3 = node(i, 3);
    }
    __device__ __forceinline__ float4 node(int i, int q) const { return nodes[((4 * i) << SHIFT) + (q << SHIFT)]; }
    __device__ __forceinline__ float4 tri(int i, int q) const { return tris[((3 * i) << SHIFT) + (q << SHIFT)]; }
    __device__ __forceinline__ float4 tri_shade(uint32_t prim) const { return shade[prim << SHIFT]; }
    __device__ __forceinline__ float4 mat(int m, int q) const { return mats[((3 * m) << SHIFT) + (q << SHIFT)]; }
};

// Where a traversal keeps its stack of deferred children.
struct StackLocal {      // per-thread array (local memory)
    int* a;
    __device__ __forceinline__ void put(int i, int v) const { a[i] = v; }
    __device__ __forceinline__ int get(int i) const { return a[i]; }
    __device__ __forceinline__ void put3(int sp, int n, int c3, int c2, int c1) const
    {
        if (n > 3) put(sp, c3);
        if (n > 2) put(sp + n - 3, c2);
        if (n > 1) put(sp + n - 2, c1);
    }
};
struct StackStrided {    // level-major words, `stride` paths apart (shared memory; bank = path id mod 32)
    int* base;
    int stride;
    __device__ __forceinline__ void put(int i, int v) const { base[i * stride] = v; }
    __device__ __forceinline__ int get(int i) const { return base[i * stride]; }
    __device__ __forceinline__ void put3(int sp, int n, int c3, int c2, int c1) const
    {
        if (n > 3) put(sp, c3);
        if (n > 2) put(sp + n - 3, c2);
        if (n > 1) put(sp + n - 2, c1);
    }
};

// The first `cap` levels in shared memory as above, deeper levels in a global-memory array (level-major, `ostride`
// paths apart).  For scenes read from global memory: a traversal rarely defers more than a dozen children, but the
// stack has to be sized for the tree's depth (36 for the 2 M-triangle PLOC tree) -- kept whole in shared memory it takes
// half of a path's record and with it a third of the paths a CTA can hold.
struct StackSpill {
    int* base;
    int stride;
    int cap;
    int* over;
    int ostride;
    __device__ __forceinline__ void put(int i, int v) const
    {
        if (i < cap) base[i * stride] = v;
        else over[(size_t)(i - cap) * ostride] = v;
    }
    __device__ __forceinline__ int get(int i) const { return i < cap ? base[i * stride] : over[(size_t)(i - cap) * ostride]; }
    // up to three pushes of one wide-node step (n - 1 of them, farthest child first): predicated shared-memory stores while
    // all three levels are in shared memory
    __device__ __forceinline__ void put3(int sp, int n, int c3, int c2, int c1) const
    {
        if (sp + 3 <= cap) {
            int* p = base + sp * stride;
            if (n > 3) p[0] = c3;
            if (n > 2) p[(n - 3) * stride] = c2;
            if (n > 1) p[(n - 2) * stride] = c1;
        } else {
            if (n > 3) put(sp, c3);
            if (n > 2) put(sp + n - 3, c2);
            if (n > 1) put(sp + n - 2, c1);
        }
    }
};

struct Ray {
    float3 o, d;
    float3 inv, ood;   // 1/d (|d| clamped away from 0) and o*inv
};

__device__ __forceinline__ float safe_inv(float d)
{
    float dd = fabsf(d) > 1e-20f ? d : copysignf(1e-20f, d);
    return 1.0f / dd;
}

__device__ __forceinline__ Ray make_ray(float3 o, float3 d)
{
    Ray r;
    r.o = o;
    r.d = d;
    r.inv = f3(safe_inv(d.x), safe_inv(d.y), safe_inv(d.z));
    r.ood = f3(o.x * r.inv.x, o.y * r.inv.y, o.z * r.inv.z);
    return r;
}

// Slab test against one child box; far bound padded by 2 ulp.  Returns hit, writes entry distance.
__device__ __forceinline__ bool ray_box(const Ray& r, float3 lo, float3 hi, float tmin, float tfar, float& tnear)
{
    float t0x = __fmaf_rn(lo.x, r.inv.x, -r.ood.x), t1x = __fmaf_rn(hi.x, r.inv.x, -r.ood.x);
    float t0y = __fmaf_rn(lo.y, r.inv.y, -r.ood.y), t1y = __fmaf_rn(hi.y, r.inv.y, -r.ood.y);
    float t0z = __fmaf_rn(lo.z, r.inv.z, -r.ood.z), t1z = __fmaf_rn(hi.z, r.inv.z, -r.ood.z);
    float tn = fmaxf(fmaxf(fminf(t0x, t1x), fminf(t0y, t1y)), fmaxf(fminf(t0z, t1z), tmin));
    float tf = fminf(fminf(fmaxf(t0x, t1x), fmaxf(t0y, t1y)), fminf(fmaxf(t0z, t1z), tfar));
    tnear = tn;
    return tn <= tf * 1.0000004f;
}

__device__ __forceinline__ float fms(float a, float b, float c, float d) { return __fmaf_rn(a, b, -(c * d)); }
__device__ __forceinline__ float dot_fma(float3 a, float3 b) { return __fmaf_rn(a.x, b.x, __fmaf_rn(a.y, b.y, a.z * b.z)); }

// Moeller-Trumbore on a (v0, e1, e2) record, both sides, tmin < t < tmax.
__device__ __forceinline__ bool ray_tri(const Ray& r, float3 v0, float3 e1, float3 e2, float tmin, float tmax, float& t_out)
{
    float3 p = f3(fms(r.d.y, e2.z, r.d.z, e2.y), fms(r.d.z, e2.x, r.d.x, e2.z), fms(r.d.x, e2.y, r.d.y, e2.x));
    float det = dot_fma(e1, p);
    if (det == 0.0f) return false;
    float inv = 1.0f / det;
    float3 tv = f3(r.o.x - v0.x, r.o.y - v0.y, r.o.z - v0.z);
    float u = dot_fma(tv, p) * inv;
    if (!(u >= 0.0f && u <= 1.0f)) return false;
    float3 q = f3(fms(tv.y, e1.z, tv.z, e1.y), fms(tv.z, e1.x, tv.x, e1.z), fms(tv.x, e1.y, tv.y, e1.x));
    float v = dot_fma(r.d, q) * inv;
    if (!(v >= 0.0f && u + v <= 1.0f)) return false;
    float t = dot_fma(e2, q) * inv;
    if (!(t > tmin && t < tmax)) return false;
    t_out = t;
    return true;
}

struct TraceCount {
    uint32_t n_node, n_tri;
};

constexpr int kStackSize = 64;   // >= depth of a radix tree over 62-bit unique keys
constexpr int kTravDone = (int)0x80000000;        // `cur` value of a finished traversal (never a node or ~leaf)
constexpr uint32_t kNoPrim = 0xffffffffu;

// One traversal in flight.  `cur` >= 0: internal node to test next; `cur` < 0: leaf ~cur to test next;
// kTravDone: finished.  Closest hit: `tfar` shrinks to the best t and `prim` is the best primitive under the
// total order (t, primIdx).  Any hit: `tfar` stays the ray's tmax and `prim` becomes 0 on the first hit.
struct Trav {
    Ray r;
    float tfar;
    uint32_t prim;
    int cur, sp;
};

__device__ __forceinline__ void trav_begin(Trav& t, float3 o, float3 d, float tmax, bool have_tris)
{
    t.r = make_ray(o, d);
    t.tfar = tmax;
    t.prim = kNoPrim;
    t.sp = 0;
    t.cur = have_tris ? 0 : kTravDone;
}

template <class STK>
__device__ __forceinline__ void trav_pop(Trav& t, const STK& stack)
{
    if (t.sp > 0) {
        t.cur = stack.get(--t.sp);
    } else {
        t.cur = kTravDone;
    }
}

// Internal-node step.  Logical visiting order (mirrored by the oracle, which matters only for the work counters):
// nearer hit child first, child 1 only when strictly nearer.  Uses r.inv / r.ood only.
template <bool QUANT = false, class FETCH, class STK>
__device__ __forceinline__ void trav_node_apply(Trav& t, const STK& stack, const FETCH& f, float tmin, float4 q0, float4 q1,
                                                float4 q2, float4 q3)
{
    float tn0, tn1;
    const bool h0 = ray_box(t.r, f3(q0.x, q0.y, q0.z), f3(q1.x, q1.y, q1.z), tmin, t.tfar, tn0);
    const bool h1 = ray_box(t.r, f3(q2.x, q2.y, q2.z), f3(q3.x, q3.y, q3.z), tmin, t.tfar, tn1);
    const int c0 = __float_as_int(q0.w), c1 = __float_as_int(q1.w);
    // selects instead of a branch ladder: the warp runs this once for all lanes.  The stack top is read
    // unconditionally (one word) so that "no child hit -> pop" needs no divergent branch either.
    const bool both = h0 && h1;
    const bool swap = both && tn1 < tn0;
    const int far_child = swap ? c0 : c1;
    const int near_child = (h0 && !swap) ? c0 : c1;
    const bool none = !(h0 || h1);
    const bool can_pop = t.sp > 0;
    const int top = stack.get(can_pop ? t.sp - 1 : 0);
    if (both) {
        stack.put(t.sp, far_child);
        if constexpr (QUANT) f.prefetch_qchild(far_child);
        else f.prefetch_child(far_child);
    }
    t.cur = none ? (can_pop ? top : kTravDone) : near_child;
    t.sp += both ? 1 : ((none && can_pop) ? -1 : 0);
}

// Node step on the quantised node array (scenes in global memory): one 32-byte fetch per node instead of 64 bytes.
template <class FETCH, class STK>
__device__ __forceinline__ void trav_qnode_step(Trav& t, const STK& stack, const FETCH& f, float tmin)
{
    float4 q0, q1, q2, q3;
    f.qnode4(t.cur, q0, q1, q2, q3);
    trav_node_apply<true>(t, stack, f, tmin, q0, q1, q2, q3);
}

// Node step on the 4-wide quantised nodes: one fetch resolves up to two levels of the binary tree.  Every axis word holds both
// planes of a child; which half is the entry plane follows from the ray's direction sign (selectors `sel`: byte-permute codes
// of the entry half per axis, the exit half is the other one), so the slab test needs no min/max per axis -- the distances are
// the same numbers ray_box() computes.  The hit children are visited nearest first (entry distance with the low two mantissa
// bits replaced by the slot number); the order only matters for the amount of work.
struct WideSel {
    uint32_t x, y, z;
};
__device__ __forceinline__ WideSel wide_selectors(float3 inv)
{
    return WideSel{inv.x < 0.0f ? 0x7432u : 0x7410u, inv.y < 0.0f ? 0x7432u : 0x7410u, inv.z < 0.0f ? 0x7432u : 0x7410u};
}
template <class FETCH>
__device__ __forceinline__ uint32_t wide_child_key(const Trav& t, const FETCH& f, WideSel sel, float wx, float wy, float wz, float tmin)
{
    const uint32_t ux = __float_as_uint(wx), uy = __float_as_uint(wy), uz = __float_as_uint(wz);
    const float nx = __fmaf_rn(__uint_as_float(__byte_perm(ux, 0x4b000000u, sel.x)), f.q_step.x, f.q_lo.x);
    const float ny = __fmaf_rn(__uint_as_float(__byte_perm(uy, 0x4b000000u, sel.y)), f.q_step.y, f.q_lo.y);
    const float nz = __fmaf_rn(__uint_as_float(__byte_perm(uz, 0x4b000000u, sel.z)), f.q_step.z, f.q_lo.z);
    const float fx = __fmaf_rn(__uint_as_float(__byte_perm(ux, 0x4b000000u, sel.x ^ 0x0022u)), f.q_step.x, f.q_lo.x);
    const float fy = __fmaf_rn(__uint_as_float(__byte_perm(uy, 0x4b000000u, sel.y ^ 0x0022u)), f.q_step.y, f.q_lo.y);
    const float fz = __fmaf_rn(__uint_as_float(__byte_perm(uz, 0x4b000000u, sel.z ^ 0x0022u)), f.q_step.z, f.q_lo.z);
    const float tn = fmaxf(fmaxf(__fmaf_rn(nx, t.r.inv.x, -t.r.ood.x), __fmaf_rn(ny, t.r.inv.y, -t.r.ood.y)),
                           fmaxf(__fmaf_rn(nz, t.r.inv.z, -t.r.ood.z), tmin));
    const float tf = fminf(fminf(__fmaf_rn(fx, t.r.inv.x, -t.r.ood.x), __fmaf_rn(fy, t.r.inv.y, -t.r.ood.y)),
                           fminf(__fmaf_rn(fz, t.r.inv.z, -t.r.ood.z), t.tfar));
    return tn <= tf * 1.0000004f ? __float_as_uint(tn) : 0x7f800000u;    // tn >= tmin > 0: the bit patterns order like the floats
}

__device__ __forceinline__ void wide_order(uint32_t& ka, int& ra, uint32_t& kb, int& rb)
{
    const bool sw = kb < ka;
    const uint32_t k = sw ? kb : ka;
    const int r = sw ? rb : ra;
    kb = sw ? ka : kb, rb = sw ? ra : rb;
    ka = k, ra = r;
}

template <class FETCH, class STK>
__device__ __forceinline__ void trav_wnode_step(Trav& t, const STK& stack, const FETCH& f, float tmin, WideSel sel)
{
    float4 a, b, c, d;
    f.wnode16(t.cur, a, b, c, d);
    int r0 = __float_as_int(d.x), r1 = __float_as_int(d.y), r2 = __float_as_int(d.z), r3 = __float_as_int(d.w);
    constexpr uint32_t kMiss = 0x7f800000u;
    uint32_t k0 = wide_child_key(t, f, sel, a.x, a.y, a.z, tmin);
    uint32_t k1 = wide_child_key(t, f, sel, a.w, b.x, b.y, tmin);
    uint32_t k2 = wide_child_key(t, f, sel, b.z, b.w, c.x, tmin);
    uint32_t k3 = wide_child_key(t, f, sel, c.y, c.z, c.w, tmin);
    // sorting network on (entry distance, child): misses last
    wide_order(k0, r0, k1, r1);
    wide_order(k2, r2, k3, r3);
    wide_order(k0, r0, k2, r2);
    wide_order(k1, r1, k3, r3);
    wide_order(k1, r1, k2, r2);
    // selects and predicated stores instead of a branch ladder (the warp runs this once for all lanes): the stack top is read
    // unconditionally so that "no child hit -> next deferred child" needs no divergent branch
    const int sp = t.sp;
    const int top = stack.get(sp > 0 ? sp - 1 : 0);
    const int n = (k0 < kMiss) + (k1 < kMiss) + (k2 < kMiss) + (k3 < kMiss);
    stack.put3(sp, n, r3, r2, r1);                        // farthest first
    if (f.prefetch) {
        if (n > 3) f.prefetch_wchild(r3);
        if (n > 2) f.prefetch_wchild(r2);
        if (n > 1) f.prefetch_wchild(r1);
    }
    t.cur = n > 0 ? r0 : (sp > 0 ? top : kTravDone);
    t.sp = n > 0 ? sp + n - 1 : (sp > 0 ? sp - 1 : 0);
}

template <class FETCH, class STK>
__device__ __forceinline__ void trav_node_step(Trav& t, const STK& stack, const FETCH& f, float tmin)
{
    float4 q0, q1, q2, q3;
    f.node4(t.cur, q0, q1, q2, q3);
    trav_node_apply(t, stack, f, tmin, q0, q1, q2, q3);
}

// Leaf step.  any = false: closest hit in (tmin, tmax_closest) under (t, primIdx); any = true: the first hit in
// (tmin, t.tfar) ends the traversal.  Uses r.o / r.d only.
template <class FETCH, class STK>
__device__ __forceinline__ void trav_leaf_step(Trav& t, const STK& stack, const FETCH& f, bool any, float tmin,
                                               float tmax_closest)
{
    const int leaf = ~t.cur;
    float4 t0, t1, t2;
    f.tri3(leaf, t0, t1, t2);
    float th;
    const bool hit = ray_tri(t.r, f3(t0.x, t0.y, t0.z), f3(t1.x, t1.y, t1.z), f3(t2.x, t2.y, t2.z), tmin,
                             any ? t.tfar : tmax_closest, th);
    if (hit) {
        if (any) {
            t.prim = 0u;
            t.cur = kTravDone;
            return;
        }
        const uint32_t prim = __float_as_uint(t0.w);
        if (th < t.tfar || (th == t.tfar && prim < t.prim)) {
            t.tfar = th;
            t.prim = prim;
        }
    }
    trav_pop(t, stack);
}

// Whole-ray wrappers (batch ray queries; the render kernels schedule the steps themselves).
template <bool COUNT, class FETCH>
__device__ __forceinline__ bool trace_closest(const FETCH& f, bool have_tris, float3 o, float3 d, float tmin, float tmax,
                                              float& t_hit, uint32_t& prim_hit, TraceCount& tc)
{
    Trav t;
    int stack_mem[kStackSize];
    const StackLocal stack{stack_mem};
    trav_begin(t, o, d, tmax, have_tris);
    while (t.cur != kTravDone) {
        if (t.cur >= 0) {
            if (COUNT) tc.n_node++;
            trav_node_step(t, stack, f, tmin);
        } else {
            if (COUNT) tc.n_tri++;
            trav_leaf_step(t, stack, f, false, tmin, tmax);
        }
    }
    t_hit = t.tfar;
    prim_hit = t.prim;
    return t.prim != kNoPrim;
}

// Any hit in (tmin, tmax): boolean, order-free.
template <bool COUNT, class FETCH>
__device__ __forceinline__ bool trace_any(const FETCH& f, bool have_tris, float3 o, float3 d, float tmin, float tmax,
                                          TraceCount& tc)
{
    Trav t;
    int stack_mem[kStackSize];
    const StackLocal stack{stack_mem};
    trav_begin(t, o, d, tmax, have_tris);
    while (t.cur != kTravDone) {
        if (t.cur >= 0) {
            if (COUNT) tc.n_node++;
            trav_node_step(t, stack, f, tmin);
        } else {
            if (COUNT) tc.n_tri++;
            trav_leaf_step(t, stack, f, true, tmin, tmax);
        }
    }
    return t.prim != kNoPrim;
}

}  // namespace ptd
