// oracle/shim/optix.h -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// A host stand-in for the slice of the OptiX 8 *device* API that the reference's
// PathTracer_Optix/pathTracerPrograms.cu uses, so that the reference's own raygen /
// closest-hit / miss programs can be #included and executed UNMODIFIED on the CPU
// (see oracle/ref_shim.cpp).  Nothing from the reference is copied here: this file only
// declares the API names the reference calls (pathTracerPrograms.cu:118-123,186-210,
// 600-622,660-674,716,836-838,869-894) and maps them onto a thread_local context.
//
// What OptiX does that is NOT in the reference tree (BVH build, traversal, ray/triangle
// test) is supplied by the oracle's own intersection routine through `shim_closest_hit`
// / `shim_any_hit` function pointers (set by ref_shim.cpp).
#pragma once

#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <vector_types.h>
#include <vector_functions.h>

// ---- CUDA keywords the device file uses ---------------------------------------------------
#ifndef __CUDACC__
#undef __device__
#undef __host__
#undef __global__
#undef __constant__
#undef __forceinline__
#undef __inline__
#define __device__
#define __host__
#define __global__
#define __constant__
#define __forceinline__ inline
#define __inline__ inline
#endif

static inline float __uint_as_float(unsigned int u) { float f; std::memcpy(&f, &u, 4); return f; }
static inline unsigned int __float_as_uint(float f) { unsigned int u; std::memcpy(&u, &f, 4); return u; }

// Device code resolves the unqualified abs(float) at pathTracerPrograms.cu:470 to the float
// overload; plain g++ would pick int abs(int) from <cstdlib>.
static inline float abs(float x) { return std::fabs(x); }

// ---- OptiX types / enums referenced by pathTracer.h and the programs ----------------------
typedef unsigned long long OptixTraversableHandle;
typedef unsigned int OptixVisibilityMask;

enum OptixPayloadTypeID {
    OPTIX_PAYLOAD_TYPE_DEFAULT = 0,
    OPTIX_PAYLOAD_TYPE_ID_0 = (1 << 0u),
    OPTIX_PAYLOAD_TYPE_ID_1 = (1 << 1u),
};

enum OptixPayloadSemantics {
    OPTIX_PAYLOAD_SEMANTICS_TRACE_CALLER_NONE = 0,
    OPTIX_PAYLOAD_SEMANTICS_TRACE_CALLER_READ = 1u << 0,
    OPTIX_PAYLOAD_SEMANTICS_TRACE_CALLER_WRITE = 2u << 0,
    OPTIX_PAYLOAD_SEMANTICS_TRACE_CALLER_READ_WRITE = 3u << 0,
    OPTIX_PAYLOAD_SEMANTICS_CH_NONE = 0,
    OPTIX_PAYLOAD_SEMANTICS_CH_READ = 1u << 2,
    OPTIX_PAYLOAD_SEMANTICS_CH_WRITE = 2u << 2,
    OPTIX_PAYLOAD_SEMANTICS_CH_READ_WRITE = 3u << 2,
    OPTIX_PAYLOAD_SEMANTICS_MS_NONE = 0,
    OPTIX_PAYLOAD_SEMANTICS_MS_READ = 1u << 4,
    OPTIX_PAYLOAD_SEMANTICS_MS_WRITE = 2u << 4,
    OPTIX_PAYLOAD_SEMANTICS_MS_READ_WRITE = 3u << 4,
};

enum OptixRayFlags {
    OPTIX_RAY_FLAG_NONE = 0u,
    OPTIX_RAY_FLAG_DISABLE_ANYHIT = 1u << 0,
    OPTIX_RAY_FLAG_TERMINATE_ON_FIRST_HIT = 1u << 2,
};

// ---- per-thread execution context ----------------------------------------------------------
struct ShimHit {
    bool hit;
    unsigned int prim;
    float t;
};

struct ShimContext {
    unsigned int payload[32];
    uint3 launch_index;
    // state of the radiance ray whose closest-hit / miss program is executing
    float3 ray_origin, ray_direction;
    ShimHit radiance_hit;
    // "hit object" of the most recent optixTraverse (what optixHitObjectIsHit() reports)
    ShimHit last_traverse;
    const void* sbt_data;  // record of the currently executing program
    // counters (segments, shadow rays) for the Grays/s metric
    unsigned long long n_radiance, n_shadow;
};

extern thread_local ShimContext g_shim;

// Supplied by the harness (ref_shim.cpp): intersection + SBT lookup + the two programs.
ShimHit shim_closest_hit(float3 o, float3 d, float tmin, float tmax);
bool shim_any_hit(float3 o, float3 d, float tmin, float tmax);
const void* shim_hit_record(unsigned int prim);
const void* shim_miss_record();
void shim_run_closesthit();
void shim_run_miss();

// ---- device API ---------------------------------------------------------------------------
static inline uint3 optixGetLaunchIndex() { return g_shim.launch_index; }
static inline unsigned int optixGetPrimitiveIndex() { return g_shim.radiance_hit.prim; }
static inline float3 optixGetWorldRayOrigin() { return g_shim.ray_origin; }
static inline float3 optixGetWorldRayDirection() { return g_shim.ray_direction; }
static inline float optixGetRayTmax() { return g_shim.radiance_hit.t; }  // in CH: the hit distance
static inline unsigned long long optixGetSbtDataPointer() { return (unsigned long long)g_shim.sbt_data; }
static inline bool optixHitObjectIsHit() { return g_shim.last_traverse.hit; }
static inline void optixReorder() {}
static inline void optixSetPayloadTypes(unsigned int) {}

#define SHIM_PAYLOAD_ACCESSOR(N)                                                     \
    static inline unsigned int optixGetPayload_##N() { return g_shim.payload[N]; }   \
    static inline void optixSetPayload_##N(unsigned int v) { g_shim.payload[N] = v; }
SHIM_PAYLOAD_ACCESSOR(0) SHIM_PAYLOAD_ACCESSOR(1) SHIM_PAYLOAD_ACCESSOR(2) SHIM_PAYLOAD_ACCESSOR(3)
SHIM_PAYLOAD_ACCESSOR(4) SHIM_PAYLOAD_ACCESSOR(5) SHIM_PAYLOAD_ACCESSOR(6) SHIM_PAYLOAD_ACCESSOR(7)
SHIM_PAYLOAD_ACCESSOR(8) SHIM_PAYLOAD_ACCESSOR(9) SHIM_PAYLOAD_ACCESSOR(10) SHIM_PAYLOAD_ACCESSOR(11)
SHIM_PAYLOAD_ACCESSOR(12) SHIM_PAYLOAD_ACCESSOR(13) SHIM_PAYLOAD_ACCESSOR(14) SHIM_PAYLOAD_ACCESSOR(15)
SHIM_PAYLOAD_ACCESSOR(16) SHIM_PAYLOAD_ACCESSOR(17) SHIM_PAYLOAD_ACCESSOR(18)
#undef SHIM_PAYLOAD_ACCESSOR

// Radiance traversal: 19 payload registers travel with the ray (pathTracerPrograms.cu:600-613).
// Registers the caller left uninitialised (u5..u17) are copied as they are, exactly like
// registers on the device.
#if defined(__GNUC__)
#pragma GCC diagnostic push
#pragma GCC diagnostic ignored "-Wmaybe-uninitialized"
#pragma GCC diagnostic ignored "-Wuninitialized"
#endif
static inline void optixTraverse(OptixPayloadTypeID, OptixTraversableHandle, float3 o, float3 d, float tmin,
                                 float tmax, float /*rayTime*/, OptixVisibilityMask, unsigned int /*flags*/,
                                 unsigned int, unsigned int, unsigned int,
                                 unsigned int& u0, unsigned int& u1, unsigned int& u2, unsigned int& u3,
                                 unsigned int& u4, unsigned int& u5, unsigned int& u6, unsigned int& u7,
                                 unsigned int& u8, unsigned int& u9, unsigned int& u10, unsigned int& u11,
                                 unsigned int& u12, unsigned int& u13, unsigned int& u14, unsigned int& u15,
                                 unsigned int& u16, unsigned int& u17, unsigned int& u18)
{
    unsigned int* p = g_shim.payload;
    p[0] = u0; p[1] = u1; p[2] = u2; p[3] = u3; p[4] = u4; p[18] = u18;
    (void)u5; (void)u6; (void)u7; (void)u8; (void)u9; (void)u10; (void)u11; (void)u12; (void)u13;
    (void)u14; (void)u15; (void)u16; (void)u17;
    g_shim.ray_origin = o;
    g_shim.ray_direction = d;
    g_shim.radiance_hit = shim_closest_hit(o, d, tmin, tmax);
    g_shim.last_traverse = g_shim.radiance_hit;
    g_shim.n_radiance++;
}
#if defined(__GNUC__)
#pragma GCC diagnostic pop
#endif

// Probe (occlusion) traversal: no payload (pathTracerPrograms.cu:660-671).
static inline void optixTraverse(OptixTraversableHandle, float3 o, float3 d, float tmin, float tmax,
                                 float /*rayTime*/, OptixVisibilityMask, unsigned int /*flags*/, unsigned int,
                                 unsigned int, unsigned int)
{
    g_shim.last_traverse.hit = shim_any_hit(o, d, tmin, tmax);
    g_shim.n_shadow++;
}

// Invoke closest-hit or miss for the radiance hit object (pathTracerPrograms.cu:620-623).
static inline void optixInvoke(OptixPayloadTypeID, unsigned int& u0, unsigned int& u1, unsigned int& u2,
                               unsigned int& u3, unsigned int& u4, unsigned int& u5, unsigned int& u6,
                               unsigned int& u7, unsigned int& u8, unsigned int& u9, unsigned int& u10,
                               unsigned int& u11, unsigned int& u12, unsigned int& u13, unsigned int& u14,
                               unsigned int& u15, unsigned int& u16, unsigned int& u17, unsigned int& u18)
{
    if (g_shim.radiance_hit.hit) {
        g_shim.sbt_data = shim_hit_record(g_shim.radiance_hit.prim);
        shim_run_closesthit();
    } else {
        g_shim.sbt_data = shim_miss_record();
        shim_run_miss();
    }
    const unsigned int* p = g_shim.payload;
    u0 = p[0]; u1 = p[1]; u2 = p[2]; u3 = p[3]; u4 = p[4]; u5 = p[5]; u6 = p[6]; u7 = p[7]; u8 = p[8];
    u9 = p[9]; u10 = p[10]; u11 = p[11]; u12 = p[12]; u13 = p[13]; u14 = p[14]; u15 = p[15]; u16 = p[16];
    u17 = p[17]; u18 = p[18];
}
