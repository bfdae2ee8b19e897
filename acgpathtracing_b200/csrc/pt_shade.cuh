// pt_shade.cuh -- the reference's closest-hit / miss / raygen arithmetic as device functions.
//
// Follows /root/reference/PathTracer_Optix/pathTracerPrograms.cu operation by operation (line numbers
// cited) so that every float rounds as in the reference source under IEEE rules; the only substitutions
// are sinf/cosf/acosf/powf -> ptdet:: (pt_detmath.h), which makes the results reproducible on the CPU
// oracle bit for bit.  Structure differs from the reference: no payload registers, no SBT -- a path is a
// handful of registers and the per-triangle geometric normal is precomputed at scene build.
#pragma once
#include "pt_device.cuh"

namespace ptd {

struct alignas(16) DevMaterial {   // 48 B = 3 x float4; from pt_material (TinyObjWrapper.h:33-40)
    float3 diffuse;
    float ior;
    float3 emission;
    int bsdf;          // 0 diffuse, 1 metallic, 2 refraction
    int emissive;      // length(emission) > 0.0f (pathTracerPrograms.cu:992), evaluated at scene build
    int pad[3];
};

struct LaunchConst {   // everything a launch needs, passed by value as a kernel parameter
    // pt_params (pathTracer.h:85-108)
    uint32_t frame, width, height, spp, max_depth;
    float3 eye, U, V, W;
    float3 l_corner, l_v1, l_v2, l_normal, l_emission;
    int use_nee, use_is;
    int raw_mean;                 // 1: write this launch's own mean (no fold into the running average, no frame buffer)
    float3 background;            // MissData.backgroundColor, PathTracerMain.cpp:568
    float4* accum;
    uchar4* frame_buffer;         // may be null
    // scene (global memory; kernels may stage it into shared memory)
    BvhView bvh;
    const float4* qnodes;         // quantised nodes of the tree in use (lbvh.cu quantize_nodes_kernel), 2 float4 each, or null
    float3 q_lo, q_step;
    uint32_t n_nodes, n_mats;
    const float4* tri_shade;      // per primitive: (N_0.xyz, material id)
    const DevMaterial* mats;      // 3 x float4 per material
    // pixel partition (tiles of 8x4 pixels; tile L belongs to class L % n_ranks; this launch owns the `span` classes
    // rank .. rank + span - 1, i.e. its k-th tile is (k / span) * n_ranks + rank + k % span)
    uint32_t tiles_x, tiles_y, rank, n_ranks, span, n_owned_tiles;
    // work order: image tile of this launch's k-th work tile (expensive tiles first, render.cu tile_order_update);
    // null = natural order
    const uint32_t* tile_order;
    // optional outputs
    unsigned long long* counters;   // pt_counters layout, or null
    uint32_t* work_counter;         // dynamic pixel fetch
    unsigned long long* error_flag; // set when a kernel watchdog fired (never expected)
    // stage scheduling of the render kernel: weights of the NODE / LEAF / SHADE / END votes and the lane count at
    // or above which the NODE stage runs without a full vote
    int sched[5];
    // probes (null on the hot path)
    uint32_t* dbg_seed_after_jitter;
    float* dbg_primary_dir;
    int32_t* dbg_primary_prim;
    float* dbg_primary_t;
    uint32_t* dbg_final_path_seed;
    uint32_t* dbg_final_pixel_seed;
};

constexpr float kPi = 3.14159265358979323846f;        // M_PIf, sutil/vec_math.h:43
constexpr float kTwoPi = 2.0f * 3.14159265358979323846f;

// pathTracerPrograms.cu:54-85.  inverse_transform(p) = p.x*tangent + p.y*binormal + p.z*normal
__device__ __forceinline__ float3 onb_to_world(float3 n, float3 p)
{
    float3 b;
    if (fabsf(n.x) > fabsf(n.z)) {
        b = f3(-n.y, n.x, 0.0f);
    } else {
        b = f3(0.0f, -n.z, n.y);
    }
    b = normalize(b);
    const float3 t = cross(b, n);
    return p.x * t + p.y * b + p.z * n;
}

// :341-353
__device__ __forceinline__ float3 cosine_sample_hemisphere(float eta1, float eta2)
{
    const float theta = ptdet::acosf_(sqrtf(eta1));
    const float phi = kTwoPi * eta2;
    float st, ct, sp, cp;
    ptdet::sincosf_(theta, st, ct);
    ptdet::sincosf_(phi, sp, cp);
    return f3(st * cp, st * sp, ct);
}

// :368-380
__device__ __forceinline__ float3 uniform_sample_hemisphere(float u1, float u2)
{
    const float phi = kTwoPi * u2;
    float sp, cp;
    ptdet::sincosf_(phi, sp, cp);
    const float r = sqrtf(1.0f - u1 * u1);
    return f3(cp * r, sp * r, u1);
}

// :455-476 (roughness is the literal 0.2 at the call site, :880)
__device__ __forceinline__ float3 sample_ggx(float u1, float u2, float roughness, float3 N)
{
    const float phi = kTwoPi * u1;
    const float cosTheta = sqrtf((1.0f - u2) / (1.0f + (roughness * roughness - 1.0f) * u2));
    const float sinTheta = sqrtf(1.0f - cosTheta * cosTheta);
    float sp, cp;
    ptdet::sincosf_(phi, sp, cp);
    const float3 H = f3(sinTheta * cp, sinTheta * sp, cosTheta);
    // :470 abs(N.z) < 0.999 -- float abs, double comparison
    const float3 up = ((double)fabsf(N.z) < 0.999) ? f3(0.0f, 0.0f, 1.0f) : f3(1.0f, 0.0f, 0.0f);
    const float3 tangent = normalize(cross(up, N));
    const float3 bitangent = cross(N, tangent);
    const float3 sampleDir = H.x * tangent + H.y * bitangent + H.z * N;
    return normalize(sampleDir);
}

// :494-510
__device__ __forceinline__ float3 fresnel_conductor(float cosTheta, float3 eta, float3 k)
{
    const float3 eta2 = eta * eta;
    const float3 k2 = k * k;
    const float3 c2 = f3(cosTheta * cosTheta);
    const float3 t1 = eta2 - k2 - c2;
    const float3 a2plusb2 = f3(sqrtf(t1.x * t1.x + 4.0f * eta2.x * k2.x), sqrtf(t1.y * t1.y + 4.0f * eta2.y * k2.y),
                               sqrtf(t1.z * t1.z + 4.0f * eta2.z * k2.z));
    const float3 t2 = a2plusb2 + c2;
    const float3 tec = (2.0f * eta) * cosTheta;
    const float3 Rs = (t2 - tec + c2) / (t2 + tec + c2);
    const float3 Rp = (Rs * (t2 - tec + f3(1.0f))) / (t2 + tec + f3(1.0f));
    return (Rs + Rp) * 0.5f;
}

// :534-559
__device__ __forceinline__ float fr_dielectric(float cosThetaI, float etaI, float etaT)
{
    cosThetaI = clampf(cosThetaI, -1.0f, 1.0f);
    if (!(cosThetaI > 0.0f)) {
        const float temp = etaI;
        etaI = etaT;
        etaT = temp;
        cosThetaI = fabsf(cosThetaI);
    }
    const float sinThetaI = sqrtf(fmaxf(0.0f, 1.0f - cosThetaI * cosThetaI));
    const float sinThetaT = etaI / etaT * sinThetaI;
    if (sinThetaT >= 1.0f) return 1.0f;
    const float cosThetaT = sqrtf(fmaxf(0.0f, 1.0f - sinThetaT * sinThetaT));
    const float rParl = ((etaT * cosThetaI) - (etaI * cosThetaT)) / ((etaT * cosThetaI) + (etaI * cosThetaT));
    const float rPerp = ((etaI * cosThetaI) - (etaT * cosThetaT)) / ((etaI * cosThetaI) + (etaT * cosThetaT));
    return (rParl * rParl + rPerp * rPerp) / 2.0f;
}

// cuda/helpers.h:107-137
__device__ __forceinline__ bool refract(float3& r, float3 i, float3 n, float ior)
{
    float3 nn = n;
    float negNdotV = dot(i, nn);
    float eta;
    if (negNdotV > 0.0f) {
        eta = ior;
        nn = -n;
        negNdotV = -negNdotV;
    } else {
        eta = 1.f / ior;
    }
    const float k = 1.f - eta * eta * (1.f - negNdotV * negNdotV);
    if (k < 0.0f) {
        r = f3(0.f);
        return false;
    }
    r = normalize(eta * i - (eta * negNdotV + sqrtf(k)) * nn);
    return true;
}

// cuda/helpers.h:35-62
__device__ __forceinline__ unsigned char quantize8(float x)
{
    x = clampf(x, 0.0f, 1.0f);
    const unsigned int q = (unsigned int)(x * 256.0f);
    return (unsigned char)(q < 255u ? q : 255u);
}
__device__ __forceinline__ float to_srgb(float c)
{
    const float invGamma = 1.0f / 2.4f;
    const float powed = ptdet::powf_(c, invGamma);
    return c < 0.0031308f ? 12.92f * c : 1.055f * powed - 0.055f;
}
// one channel; out of line so that the three channels share one copy of the ~220-instruction powf (it runs once per pixel and
// launch, and the render kernel's hot code is larger than the instruction cache as it is)
static __device__ __noinline__ unsigned int srgb8(float c) { return quantize8(to_srgb(clampf(c, 0.0f, 1.0f))); }
__device__ __forceinline__ uchar4 make_color(float3 c)
{
    return make_uchar4((unsigned char)srgb8(c.x), (unsigned char)srgb8(c.y), (unsigned char)srgb8(c.z), 255u);
}

// Per-path registers (RadiancePayloadRayData, pathTracer.h:19-32, minus what never leaves a stage).
struct Path {
    float3 o, d;        // current ray
    float3 att;         // attenuation
    uint32_t seed;      // prd.randomSeed
    int depth;
};

struct ShadowQuery {    // NEE probe ray produced by shading (pathTracerPrograms.cu:1003-1026)
    bool want;
    float3 dir;         // L; origin is the shading point = next ray origin for diffuse, recomputed for metal
    float3 org;
    float tmax;         // Ldist - 0.01
    float3 add;         // light.emission * weight, added to radiance when unoccluded
};

// __closesthit__diffuse__ch :866-1031 up to (not including) the occlusion trace.
// In: path (ray that hit, attenuation, seed, depth), hit (prim, t).  Out: path advanced to the next
// ray/attenuation/seed; emission, radiance, done; the shadow query.  `hit_type` receives the BSDF type.
// `f` says where the per-triangle shading records and the material table live (pt_device.cuh); `l_area` is
// length(cross(light.v1, light.v2)) (:1021), constant per launch.  `mat_id` receives the material index.
template <class FETCH>
__device__ __forceinline__ void shade_hit(const LaunchConst& lc, const FETCH& f, float l_area,
                                          uint32_t prim, float t_hit, Path& p, float3& emission, float3& radiance,
                                          bool& done, ShadowQuery& sq, int& hit_type, int& mat_id)
{
    const float4 ts = f.tri_shade(prim);
    DevMaterial m;
    {
        mat_id = __float_as_int(ts.w);
        const float4 m0 = f.mat(mat_id, 0), m1 = f.mat(mat_id, 1), m2 = f.mat(mat_id, 2);
        m.diffuse = f3(m0.x, m0.y, m0.z);
        m.ior = m0.w;
        m.emission = f3(m1.x, m1.y, m1.z);
        m.bsdf = __float_as_int(m1.w);
        m.emissive = __float_as_int(m2.x);
    }
    const float3 ray_dir = p.d;
    const float3 N_0 = f3(ts.x, ts.y, ts.z);                                           // :891 (precomputed)
    const float3 N = N_0 * copysignf(1.0f, dot(-ray_dir, N_0));                        // :893 faceforward
    const float3 P = p.o + t_hit * ray_dir;                                            // :894
    emission = (p.depth == 0) ? m.emission : f3(0.0f);                                 // :898-901
    uint32_t seed = p.seed;
    hit_type = m.bsdf;

    if (m.bsdf == 0) {                                                                 // :907-930
        const float z1 = rnd(seed);
        const float z2 = rnd(seed);
        const float3 w = lc.use_is ? cosine_sample_hemisphere(z1, z2) : uniform_sample_hemisphere(z1, z2);
        p.d = onb_to_world(N, w);
        p.o = P;
        p.att = p.att * m.diffuse;
    } else if (m.bsdf == 1) {                                                          // :931-953
        const float z1 = rnd(seed);
        const float z2 = rnd(seed);
        const float3 mn = sample_ggx(z1, z2, 0.2f, N);
        const float3 R = reflect(ray_dir, mn);
        p.d = R;
        p.o = P + R * 1e-4f;
        const float3 eta = f3(1.45f, 0.7f, 1.55f);
        const float3 k = f3(3.0f, 2.2f, 3.5f);
        const float cosTheta = fmaxf(dot(mn, -ray_dir), 0.0f);
        const float3 F = fresnel_conductor(cosTheta, eta, k);
        p.att = p.att * (F * m.diffuse);
    } else {                                                                           // :954-982
        const float3 inc = normalize(ray_dir);
        const float cos_theta = dot(normalize(-ray_dir), N_0);
        const float F = fr_dielectric(cos_theta, 1.0f, m.ior);
        float3 dir;
        if (rnd(seed) < F) {
            dir = reflect(inc, N_0);
        } else {
            float3 refr;
            dir = refract(refr, inc, N_0, m.ior) ? refr : reflect(inc, N_0);
        }
        p.d = dir;
        p.o = P + dir * 1e-3f;
        p.att = p.att * m.diffuse;
    }

    const float z1 = rnd(seed);                                                        // :985-987
    const float z2 = rnd(seed);
    p.seed = seed;

    if (m.emissive) {                                                                  // :992-1000
        radiance = m.emission;
        done = true;
    } else {
        radiance = f3(0.0f);
        done = false;
    }

    sq.want = false;
    if (lc.use_nee && m.bsdf != 2) {                                                   // :1003
        const float3 light_pos = lc.l_corner + lc.l_v1 * z1 + lc.l_v2 * z2;            // :1007
        const float3 toL = light_pos - P;
        const float Ldist = length(toL);
        const float3 L = normalize(toL);
        const float nDl = dot(N, L);
        const float LnDl = -dot(lc.l_normal, L);
        if (nDl > 0.0f && LnDl > 0.0f) {                                               // :1015
            sq.want = true;
            sq.org = P;
            sq.dir = L;
            sq.tmax = Ldist - 0.01f;                                                   // :1017
            const float weight = nDl * LnDl * l_area / (kPi * Ldist * Ldist);       // :1022
            sq.add = lc.l_emission * weight;
        }
    }
}

// Camera ray of one sample (__raygen__rg :730-745).  Draws the two jitter values from the PIXEL seed
// (x first: nvcc's evaluation order of make_float2(rnd(seed), rnd(seed))), then the path gets a
// COPY of the pixel seed.
__device__ __forceinline__ void generate_camera_ray(const LaunchConst& lc, uint32_t ix, uint32_t iy, uint32_t& pixel_seed, Path& p)
{
    const float jx = rnd(pixel_seed);
    const float jy = rnd(pixel_seed);
    const float dx = 2.0f * (((float)ix + jx) / (float)(int)lc.width) - 1.0f;           // :732-735
    const float dy = 2.0f * (((float)iy + jy) / (float)(int)lc.height) - 1.0f;
    p.d = normalize(dx * lc.U + dy * lc.V + lc.W);                                      // :737
    p.o = lc.eye;
    p.att = f3(1.0f);                                                                   // :742-745
    p.seed = pixel_seed;
    p.depth = 0;
}

// Fold `result` (sum over this launch's samples) into the running average and write both buffers
// (__raygen__rg :784-814).
__device__ __forceinline__ void write_pixel(const LaunchConst& lc, uint32_t image_index, float3 result)
{
    float3 accum = div_by_scalar(result, (float)lc.spp);                                // :784
    if (!lc.raw_mean && (int)lc.frame > 0) {                                            // :803-810
        const float a = 1.0f / (float)((int)lc.frame + 1);
        const float4 prev4 = lc.accum[image_index];
        const float3 prev = f3(prev4.x, prev4.y, prev4.z);
        accum = prev + a * (accum - prev);                                              // lerp, vec_math.h:500
    }
    lc.accum[image_index] = make_float4(accum.x, accum.y, accum.z, 1.0f);               // :811
    if (lc.frame_buffer) lc.frame_buffer[image_index] = make_color(accum);              // :814
}

}  // namespace ptd
