// oracle/pt_oracle.cpp -- TEST INFRASTRUCTURE (CPU oracle), NOT PRODUCT CODE.
//
// Scalar host restatement ("port") of the reference's device programs
//   /root/reference/PathTracer_Optix/pathTracerPrograms.cu   (__raygen__rg :707, __miss__ms :833,
//   __closesthit__diffuse__ch :866 and helpers), cuda/random.h, cuda/helpers.h, sutil/Camera.cpp.
// Each function cites the reference lines it follows.  Only tests/, __graft_entry__.smoke() and
// bench.py's cpu_baseline / --impl reference legs may load this library; the product
// (acgpathtracing_b200/) never does.
//
// PINNING.  The reference has no tests or golden vectors.  This port is pinned
// against the reference's OWN source executed on the host through oracle/shim/optix.h
// (oracle/ref_shim.cpp -> oracle/_ref/libptref.so): tests/test_oracle_vs_ref.py requires
// bit-identical accumulation buffers, framebuffers and final RNG states, and the committed
// fixtures under tests/golden/ were produced by that library (tests/golden/make_golden.py).
// What OptiX itself computes (BVH, traversal, ray/triangle test) is not in the reference tree:
// that part is specified by us (oracle_geom.h) -> "parity unpinned" at the hit-distance level.
//
// Two math modes:
//   mode 0 "libm": sinf/cosf/acosf/powf from glibc -- the semantics of the reference source
//                  compiled without fast-math; bit-identical to the shim-over-reference build.
//   mode 1 "det" : the same transcription with those four functions taken from
//                  acgpathtracing_b200/csrc/pt_detmath.h (portable add/mul/fma polynomials), the
//                  arithmetic the GPU kernels are specified to use, so CPU == GPU bit-for-bit.
// Compile with -ffp-contract=off.
#include <atomic>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <thread>
#include <vector>

#include "../acgpathtracing_b200/csrc/pt_detmath.h"
#include "oracle_geom.h"
#include "oracle_vec.h"

using namespace ora;

// ---- launch contract (byte-identical to pathTracer.h:77-108; offsets checked in tests) ----------
struct OrAreaLight {
    float corner[3], v1[3], v2[3], normal[3], emission[3];
};
struct OrParams {
    uint32_t currentFrameIdx;
    float* accumulationBuffer;   // float4*
    uint8_t* frameBuffer;        // uchar4*
    uint32_t width, height, samplesPerPixel, maxDepth;
    float cameraEye[3], cameraU[3], cameraV[3], cameraW[3];
    OrAreaLight areaLight;
    uint64_t handle;
    bool useDirectLighting;
    bool useImportanceSampling;
};
static_assert(sizeof(OrParams) == 168, "PathTraceParams must be 168 bytes");
static_assert(offsetof(OrParams, accumulationBuffer) == 8 && offsetof(OrParams, width) == 24 &&
                  offsetof(OrParams, cameraEye) == 40 && offsetof(OrParams, areaLight) == 88 &&
                  offsetof(OrParams, handle) == 152 && offsetof(OrParams, useDirectLighting) == 160,
              "PathTraceParams layout");

// TinyObjWrapper.h:27-40
enum { BSDF_DIFFUSE = 0, BSDF_METALLIC = 1, BSDF_REFRACTION = 2 };
struct OrMaterial {
    float diffuse[3], emission[3], roughness, metallic, ior;
    int32_t bsdfType;
};
static_assert(sizeof(OrMaterial) == 40, "Material must be 40 bytes");

struct OrCounters {
    unsigned long long n_samples, n_radiance, n_shadow, n_node, n_tri, n_hit[3], n_miss;
};

// ---- RNG: cuda/random.h ---------------------------------------------------------------------------
// random.h:32-46
static inline uint32_t tea4(uint32_t val0, uint32_t val1)
{
    uint32_t v0 = val0, v1 = val1, s0 = 0;
    for (int n = 0; n < 4; ++n) {
        s0 += 0x9e3779b9u;
        v0 += ((v1 << 4) + 0xa341316cu) ^ (v1 + s0) ^ ((v1 >> 5) + 0xc8013ea4u);
        v1 += ((v0 << 4) + 0xad90777du) ^ (v0 + s0) ^ ((v0 >> 5) + 0x7e95761eu);
    }
    return v0;
}
// random.h:49-55
static inline uint32_t lcg(uint32_t& prev)
{
    prev = 1664525u * prev + 1013904223u;
    return prev & 0x00ffffffu;
}
// random.h:64-67
static inline float rnd(uint32_t& prev) { return (float)lcg(prev) / (float)0x01000000; }

// ---- math modes -----------------------------------------------------------------------------------
struct MathLibm {
    static float sin(float x) { return sinf(x); }
    static float cos(float x) { return cosf(x); }
    static float acos(float x) { return acosf(x); }
    static float pow(float x, float y) { return powf(x, y); }
};
struct MathDet {
    static float sin(float x) { return ptdet::sinf_(x); }
    static float cos(float x) { return ptdet::cosf_(x); }
    static float acos(float x) { return ptdet::acosf_(x); }
    static float pow(float x, float y) { return ptdet::powf_(x, y); }
};

static const float kPi = 3.14159265358979323846f;  // M_PIf, vec_math.h:43-45

// ---- helpers of pathTracerPrograms.cu -------------------------------------------------------------
// :54-85 OrthonormalBasis
struct Onb {
    V3 tangent, binormal, normal;
    explicit Onb(V3 n)
    {
        normal = n;
        if (fabs(n.x) > fabs(n.z)) {
            binormal = v3(-n.y, n.x, 0.0f);
        } else {
            binormal = v3(0.0f, -n.z, n.y);
        }
        binormal = normalize(binormal);
        tangent = cross(binormal, normal);
    }
    V3 inverse_transform(V3 p) const { return p.x * tangent + p.y * binormal + p.z * normal; }
};

// :265-284
static inline float safeDivide(float a, float b) { return b == 0.0f ? 0.0f : a / b; }
static inline V3 safeDivide(V3 a, float b) { return v3(safeDivide(a.x, b), safeDivide(a.y, b), safeDivide(a.z, b)); }

// :341-353
template <class M>
static inline V3 cosine_sample_hemisphere(float eta1, float eta2)
{
    const float theta = M::acos(sqrtf(eta1));
    const float phi = 2.0f * kPi * eta2;
    V3 p;
    p.x = M::sin(theta) * M::cos(phi);
    p.y = M::sin(theta) * M::sin(phi);
    p.z = M::cos(theta);
    return p;
}

// :368-380 (the acosf at :374 is dead code)
template <class M>
static inline V3 uniform_sample_hemisphere(float u1, float u2)
{
    const float phi = 2.0f * kPi * u2;
    V3 wi;
    wi.x = M::cos(phi) * sqrtf(1 - u1 * u1);
    wi.y = M::sin(phi) * sqrtf(1 - u1 * u1);
    wi.z = u1;
    return wi;
}

// :455-476
template <class M>
static inline V3 sampleGGX(float u1, float u2, float roughness, V3 N)
{
    float phi = 2.0f * kPi * u1;
    float cosTheta = sqrtf((1.0f - u2) / (1.0f + (roughness * roughness - 1.0f) * u2));
    float sinTheta = sqrtf(1.0f - cosTheta * cosTheta);
    V3 H;
    H.x = sinTheta * M::cos(phi);
    H.y = sinTheta * M::sin(phi);
    H.z = cosTheta;
    // :470 -- float abs on the device, compared against a double literal (Q8)
    V3 up = (double)fabsf(N.z) < 0.999 ? v3(0, 0, 1) : v3(1, 0, 0);
    V3 tangent = normalize(cross(up, N));
    V3 bitangent = cross(N, tangent);
    V3 sampleDir = H.x * tangent + H.y * bitangent + H.z * N;
    return normalize(sampleDir);
}

// :494-510
static inline V3 fresnelSchlickConductor(float cosTheta, V3 eta, V3 k)
{
    V3 eta2 = eta * eta;
    V3 k2 = k * k;
    V3 c2 = v3(cosTheta * cosTheta);
    V3 t1 = eta2 - k2 - c2;
    V3 a2plusb2 = v3(sqrtf(t1.x * t1.x + 4 * eta2.x * k2.x), sqrtf(t1.y * t1.y + 4 * eta2.y * k2.y),
                     sqrtf(t1.z * t1.z + 4 * eta2.z * k2.z));
    V3 t2 = a2plusb2 + c2;
    V3 Rs = (t2 - 2 * eta * cosTheta + c2) / (t2 + 2 * eta * cosTheta + c2);
    V3 Rp = Rs * (t2 - 2 * eta * cosTheta + v3(1)) / (t2 + 2 * eta * cosTheta + v3(1));
    return (Rs + Rp) * 0.5f;
}

// :534-559
static inline float FrDielectric(float cosThetaI, float etaI, float etaT)
{
    cosThetaI = clampf(cosThetaI, -1.0f, 1.0f);
    bool entering = cosThetaI > 0.0f;
    if (!entering) {
        float temp = etaI;
        etaI = etaT;
        etaT = temp;
        cosThetaI = fabs(cosThetaI);
    }
    float sinThetaI = sqrtf(fmaxf(0.0f, 1.0f - cosThetaI * cosThetaI));
    float sinThetaT = etaI / etaT * sinThetaI;
    if (sinThetaT >= 1.0f) return 1.0f;
    float cosThetaT = sqrtf(fmaxf(0.0f, 1.0f - sinThetaT * sinThetaT));
    float rParl = ((etaT * cosThetaI) - (etaI * cosThetaT)) / ((etaT * cosThetaI) + (etaI * cosThetaT));
    float rPerp = ((etaI * cosThetaI) - (etaT * cosThetaT)) / ((etaI * cosThetaI) + (etaT * cosThetaT));
    return (rParl * rParl + rPerp * rPerp) / 2.0f;
}

// cuda/helpers.h:107-137
static inline bool refract(V3& r, V3 i, V3 n, float ior)
{
    V3 nn = n;
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
        r = v3(0.f);
        return false;
    }
    r = normalize(eta * i - (eta * negNdotV + sqrtf(k)) * nn);
    return true;
}

// cuda/helpers.h:35-62
template <class M>
static inline void make_color(V3 c, uint8_t out[4])
{
    c = clamp(c, 0.0f, 1.0f);
    float invGamma = 1.0f / 2.4f;
    V3 powed = v3(M::pow(c.x, invGamma), M::pow(c.y, invGamma), M::pow(c.z, invGamma));
    V3 s = v3(c.x < 0.0031308f ? 12.92f * c.x : 1.055f * powed.x - 0.055f,
              c.y < 0.0031308f ? 12.92f * c.y : 1.055f * powed.y - 0.055f,
              c.z < 0.0031308f ? 12.92f * c.z : 1.055f * powed.z - 0.055f);
    const float comp[3] = {s.x, s.y, s.z};
    for (int a = 0; a < 3; ++a) {
        float x = clampf(comp[a], 0.0f, 1.0f);
        unsigned int q = (unsigned int)(x * (float)256);
        out[a] = (uint8_t)(q < 255u ? q : 255u);
    }
    out[3] = 255u;
}

// ---- scene ----------------------------------------------------------------------------------------
struct OrScene {
    std::vector<float> verts;        // xyzw
    std::vector<uint32_t> indices;   // 3 per triangle
    std::vector<uint32_t> mat_ids;
    std::vector<OrMaterial> mats;
    float background[4] = {0, 0, 0, 0};   // MissData.backgroundColor, PathTracerMain.cpp:568
    // 0: jitter x takes the first draw (nvcc device code) -- the semantics the GPU path follows.
    // 1: jitter y takes the first draw (what g++ makes of pathTracerPrograms.cu:730 in the host build of
    //    the reference); used only to compare this port with oracle/_ref bit for bit.
    int jitter_order = 0;
    // optional [w*h] output: per pixel a hash of every discrete decision of the launch's paths -- the primitive each segment
    // hit (or the miss), the dielectric's reflect / refract / total-internal-reflection branch, whether a shadow ray was cast
    // and what it found, and how the segment ended (done / Russian roulette / depth).  Two renders whose arithmetic differs in
    // the last place (deterministic vs glibc transcendentals) agree to ~1e-6 on a pixel exactly when these hashes agree.
    uint32_t* decision_probe = nullptr;
    Lbvh bvh;
};

struct PathState {   // RadiancePayloadRayData, pathTracer.h:19-32
    V3 attenuation;
    uint32_t randomSeed;
    int depth;
    V3 emissionColor, radiance, origin, direction;
    int done;
    uint32_t decisions;   // test aid: hash of the discrete decisions the path took (see OrScene::decision_probe)
};

// FNV-1a step: the hash of a pixel's decisions only has to tell "same sequence" from "different sequence"
static inline void decide(PathState& prd, uint32_t what) { prd.decisions = (prd.decisions ^ what) * 16777619u; }

template <bool USE_BVH>
static inline Hit trace_closest(const OrScene& sc, V3 o, V3 d, float tmin, float tmax, TraceCounters* tc)
{
    const float oo[3] = {o.x, o.y, o.z}, dd[3] = {d.x, d.y, d.z};
    if (USE_BVH) return bvh_closest(sc.bvh, oo, dd, tmin, tmax, tc);
    return brute_closest(sc.bvh.tris_input, oo, dd, tmin, tmax);
}
template <bool USE_BVH>
static inline bool trace_any(const OrScene& sc, V3 o, V3 d, float tmin, float tmax, TraceCounters* tc)
{
    const float oo[3] = {o.x, o.y, o.z}, dd[3] = {d.x, d.y, d.z};
    if (USE_BVH) return bvh_any(sc.bvh, oo, dd, tmin, tmax, tc);
    return brute_any(sc.bvh.tris_input, oo, dd, tmin, tmax);
}

// __miss__ms, pathTracerPrograms.cu:833-847
static inline void miss_program(const OrScene& sc, PathState& prd)
{
    prd.radiance = v3(sc.background[0], sc.background[1], sc.background[2]);
    prd.emissionColor = v3(0.f);
    prd.done = true;
}

// __closesthit__diffuse__ch, pathTracerPrograms.cu:866-1031
template <class M, bool USE_BVH>
static inline void closesthit_program(const OrScene& sc, const OrParams& params, uint32_t prim_idx, float t_hit,
                                      V3 ray_org, V3 ray_dir, PathState& prd, OrCounters& ctr, TraceCounters* tc)
{
    const OrMaterial& rt = sc.mats[sc.mat_ids[prim_idx]];                     // SBT record = material (:871)
    const uint32_t* idx = &sc.indices[3 * (size_t)prim_idx];                  // :877
    const bool useDirectLighting = params.useDirectLighting;
    const float roughness = 0.2;                                              // :880 (literal; Q6)
    const float IOR = rt.ior;
    const int bsdfType = rt.bsdfType;
    const bool useImportanceSampling = params.useImportanceSampling;
    const float* p0 = &sc.verts[4 * (size_t)idx[0]];
    const float* p1 = &sc.verts[4 * (size_t)idx[1]];
    const float* p2 = &sc.verts[4 * (size_t)idx[2]];
    const V3 v0 = v3(p0[0], p0[1], p0[2]), v1 = v3(p1[0], p1[1], p1[2]), v2 = v3(p2[0], p2[1], p2[2]);   // :886-888

    const V3 N_0 = normalize(cross(v1 - v0, v2 - v0));                        // :891
    const V3 N = faceforward(N_0, -ray_dir, N_0);                             // :893
    const V3 P = ray_org + t_hit * ray_dir;                                   // :894

    const V3 matEmission = v3(rt.emission[0], rt.emission[1], rt.emission[2]);
    const V3 matDiffuse = v3(rt.diffuse[0], rt.diffuse[1], rt.diffuse[2]);
    prd.emissionColor = (prd.depth == 0) ? matEmission : v3(0.0f);            // :898-901

    uint32_t seed = prd.randomSeed;                                           // :903
    if (bsdfType >= 0 && bsdfType < 3) ctr.n_hit[bsdfType]++;

    switch (bsdfType) {                                                       // :905
        case BSDF_DIFFUSE: {
            const float z1 = rnd(seed);
            const float z2 = rnd(seed);
            Onb onb(N);
            V3 w_in = useImportanceSampling ? cosine_sample_hemisphere<M>(z1, z2) : uniform_sample_hemisphere<M>(z1, z2);
            w_in = onb.inverse_transform(w_in);
            prd.direction = w_in;
            prd.origin = P;
            prd.attenuation *= matDiffuse;
            break;
        }
        case BSDF_METALLIC: {
            const float z1 = rnd(seed);
            const float z2 = rnd(seed);
            V3 microfacetNormal = sampleGGX<M>(z1, z2, roughness, N);
            V3 R = reflect(ray_dir, microfacetNormal);
            prd.direction = R;
            prd.origin = P + R * 1e-4f;
            V3 eta = v3(1.45, 0.7, 1.55);                                      // :943 (double literals -> float)
            V3 k = v3(3.0, 2.2, 3.5);                                          // :944
            float cosTheta = fmaxf(dot(microfacetNormal, -ray_dir), 0.0f);
            V3 F = fresnelSchlickConductor(cosTheta, eta, k);
            V3 color = F * matDiffuse;
            prd.attenuation *= color;
            break;
        }
        case BSDF_REFRACTION: {
            V3 incidentRayDir = normalize(ray_dir);
            float cos_theta = dot(normalize(-ray_dir), N_0);                  // :959
            float F = FrDielectric(cos_theta, 1.0f, IOR);
            if (rnd(seed) < F) {
                prd.direction = reflect(incidentRayDir, N_0);
                decide(prd, 0x101u);
            } else {
                V3 refractedDir;
                bool didRefract = refract(refractedDir, incidentRayDir, N_0, IOR);
                prd.direction = didRefract ? refractedDir : reflect(incidentRayDir, N_0);
                decide(prd, didRefract ? 0x102u : 0x103u);
            }
            prd.origin = P + prd.direction * 1e-3f;
            prd.attenuation *= matDiffuse;
            break;
        }
        default: break;
    }

    const float z1 = rnd(seed);                                               // :985
    const float z2 = rnd(seed);                                               // :986
    prd.randomSeed = seed;                                                    // :987

    if (length(matEmission) > 0.0f) {                                         // :992
        prd.radiance = matEmission;
        prd.done = true;
    } else {
        prd.radiance = v3(0.0f);
        prd.done = false;
    }

    if (useDirectLighting && bsdfType != BSDF_REFRACTION) {                   // :1003
        const OrAreaLight& lt = params.areaLight;
        const V3 lc = v3(lt.corner[0], lt.corner[1], lt.corner[2]);
        const V3 lv1 = v3(lt.v1[0], lt.v1[1], lt.v1[2]), lv2 = v3(lt.v2[0], lt.v2[1], lt.v2[2]);
        const V3 ln = v3(lt.normal[0], lt.normal[1], lt.normal[2]);
        const V3 le = v3(lt.emission[0], lt.emission[1], lt.emission[2]);
        const V3 light_pos = lc + lv1 * z1 + lv2 * z2;                        // :1007
        const float Ldist = length(light_pos - P);
        const V3 L = normalize(light_pos - P);
        const float nDl = dot(N, L);
        const float LnDl = -dot(ln, L);
        if (nDl > 0.0f && LnDl > 0.0f) {                                      // :1015
            ctr.n_shadow++;
            // traceOcclusion :651-684 reduces to "anything hit" 
            const bool occluded = trace_any<USE_BVH>(sc, P, L, 0.01f, Ldist - 0.01f, tc);
            decide(prd, occluded ? 0x201u : 0x202u);
            if (!occluded) {
                const float A = length(cross(lv1, lv2));
                float weight = nDl * LnDl * A / (kPi * Ldist * Ldist);        // :1022
                prd.radiance += le * weight;
            }
        }
    }
}

struct PixelProbe {   // optional per-pixel debug outputs (first sample of the launch)
    int32_t* primary_prim;   // -1 on miss
    float* primary_t;
    uint32_t* seed_after_jitter;   // pixel seed after the first sample's two jitter draws
    uint32_t* final_path_seed;     // prd.randomSeed when the last sample of the launch ended
    uint32_t* final_pixel_seed;    // pixel seed after all jitter draws
    float* primary_dir;            // 3 floats: first camera ray direction
};

// __raygen__rg, pathTracerPrograms.cu:707-816
template <class M, bool USE_BVH>
static void raygen_pixel(const OrScene& sc, const OrParams& params, uint32_t ix, uint32_t iy, float* accum,
                         uint8_t* fb, OrCounters& ctr, TraceCounters* tc, const PixelProbe* probe)
{
    const int w = params.width;
    const int h = params.height;
    const V3 eye = v3(params.cameraEye[0], params.cameraEye[1], params.cameraEye[2]);
    const V3 U = v3(params.cameraU[0], params.cameraU[1], params.cameraU[2]);
    const V3 V = v3(params.cameraV[0], params.cameraV[1], params.cameraV[2]);
    const V3 W = v3(params.cameraW[0], params.cameraW[1], params.cameraW[2]);
    const int subframe_index = params.currentFrameIdx;
    const unsigned int maxDepth = params.maxDepth;
    const unsigned int image_index = iy * params.width + ix;

    unsigned int seed = tea4(iy * w + ix, subframe_index);                     // :721
    V3 result = v3(0.0f);
    int i = params.samplesPerPixel;
    PathState prd;
    std::memset(&prd, 0, sizeof(prd));
    bool first = true;
    do {
        // :730 -- nvcc evaluates the two draws left to right: x first, then y
        float jx, jy;
        if (sc.jitter_order == 0) {
            jx = rnd(seed);
            jy = rnd(seed);
        } else {
            jy = rnd(seed);
            jx = rnd(seed);
        }
        // :732-735
        const float dx = 2.0f * ((static_cast<float>(ix) + jx) / static_cast<float>(w)) - 1.0f;
        const float dy = 2.0f * ((static_cast<float>(iy) + jy) / static_cast<float>(h)) - 1.0f;
        V3 ray_direction = normalize(dx * U + dy * V + W);                     // :737
        V3 ray_origin = eye;
        prd.attenuation = v3(1.f);                                             // :742-745
        prd.randomSeed = seed;
        prd.depth = 0;
        ctr.n_samples++;
        if (first && probe) {
            if (probe->seed_after_jitter) probe->seed_after_jitter[image_index] = seed;
            if (probe->primary_dir) {
                probe->primary_dir[3 * image_index + 0] = ray_direction.x;
                probe->primary_dir[3 * image_index + 1] = ray_direction.y;
                probe->primary_dir[3 * image_index + 2] = ray_direction.z;
            }
        }
        for (;;) {
            // traceRadiance :577-635, tmin 0.01 tmax 1e16 (:754-755)
            ctr.n_radiance++;
            Hit hit = trace_closest<USE_BVH>(sc, ray_origin, ray_direction, 0.01f, 1e16f, tc);
            if (first && probe) {
                if (probe->primary_prim) probe->primary_prim[image_index] = hit.hit ? (int32_t)hit.prim : -1;
                if (probe->primary_t) probe->primary_t[image_index] = hit.hit ? hit.t : 0.0f;
                first = false;
            }
            first = false;
            decide(prd, hit.hit ? hit.prim : 0xffffffffu);
            if (hit.hit) {
                closesthit_program<M, USE_BVH>(sc, params, hit.prim, hit.t, ray_origin, ray_direction, prd, ctr, tc);
            } else {
                ctr.n_miss++;
                miss_program(sc, prd);
            }
            result += prd.emissionColor;                                       // :760
            result += prd.radiance * prd.attenuation;                          // :761
            const float p = dot(prd.attenuation, v3(0.30f, 0.59f, 0.11f));     // :763
            bool russianRoulette = rnd(prd.randomSeed) > p;                    // :765
            const bool done = prd.done || russianRoulette || (unsigned int)prd.depth >= maxDepth;   // :767
            decide(prd, 0x300u | (prd.done ? 1u : 0u) | (russianRoulette ? 2u : 0u) | ((unsigned int)prd.depth >= maxDepth ? 4u : 0u));
            if (done) break;
            prd.attenuation = safeDivide(prd.attenuation, p);                  // :773
            ray_origin = prd.origin;
            ray_direction = prd.direction;
            ++prd.depth;
        }
    } while (--i);

    V3 accum_color = result / static_cast<float>(params.samplesPerPixel);     // :784
    if (subframe_index > 0) {                                                  // :803-810
        const float a = 1.0f / static_cast<float>(subframe_index + 1);
        const V3 prev = v3(accum[4 * (size_t)image_index + 0], accum[4 * (size_t)image_index + 1],
                           accum[4 * (size_t)image_index + 2]);
        accum_color = lerp(prev, accum_color, a);
    }
    accum[4 * (size_t)image_index + 0] = accum_color.x;                       // :811
    accum[4 * (size_t)image_index + 1] = accum_color.y;
    accum[4 * (size_t)image_index + 2] = accum_color.z;
    accum[4 * (size_t)image_index + 3] = 1.0f;
    if (fb) make_color<M>(accum_color, fb + 4 * (size_t)image_index);        // :814
    if (sc.decision_probe) sc.decision_probe[image_index] = prd.decisions;
    if (probe) {
        if (probe->final_path_seed) probe->final_path_seed[image_index] = prd.randomSeed;
        if (probe->final_pixel_seed) probe->final_pixel_seed[image_index] = seed;
    }
}

template <class M, bool USE_BVH>
static void render_rows(const OrScene& sc, const OrParams& params, float* accum, uint8_t* fb, int nthreads,
                        OrCounters* out_ctr, const PixelProbe* probe, uint32_t y0, uint32_t y1, uint32_t x0, uint32_t x1)
{
    if (nthreads < 1) nthreads = 1;
    std::atomic<uint32_t> next_row(y0);
    std::vector<OrCounters> ctrs(nthreads);
    std::vector<std::thread> pool;
    auto work = [&](int tid) {
        OrCounters c;
        std::memset(&c, 0, sizeof(c));
        TraceCounters tc;
        for (;;) {
            uint32_t y = next_row.fetch_add(1);
            if (y >= y1) break;
            for (uint32_t x = x0; x < x1; ++x) raygen_pixel<M, USE_BVH>(sc, params, x, y, accum, fb, c, &tc, probe);
        }
        c.n_node = tc.n_node;
        c.n_tri = tc.n_tri;
        ctrs[tid] = c;
    };
    for (int t = 1; t < nthreads; ++t) pool.emplace_back(work, t);
    work(0);
    for (auto& th : pool) th.join();
    if (out_ctr) {
        std::memset(out_ctr, 0, sizeof(*out_ctr));
        for (const OrCounters& c : ctrs) {
            out_ctr->n_samples += c.n_samples;
            out_ctr->n_radiance += c.n_radiance;
            out_ctr->n_shadow += c.n_shadow;
            out_ctr->n_node += c.n_node;
            out_ctr->n_tri += c.n_tri;
            out_ctr->n_miss += c.n_miss;
            for (int a = 0; a < 3; ++a) out_ctr->n_hit[a] += c.n_hit[a];
        }
    }
}

// =====================================================================================================
// C API (ctypes)
// =====================================================================================================
extern "C" {

uint32_t ptor_tea4(uint32_t a, uint32_t b) { return tea4(a, b); }
uint32_t ptor_lcg(uint32_t* s) { return lcg(*s); }
float ptor_rnd(uint32_t* s) { return rnd(*s); }

unsigned ptor_sizeof_params() { return (unsigned)sizeof(OrParams); }
unsigned ptor_sizeof_material() { return (unsigned)sizeof(OrMaterial); }

#define OR_DISPATCH_MATH(mode, CALL_LIBM, CALL_DET) \
    do {                                            \
        if ((mode) == 0) { CALL_LIBM; } else { CALL_DET; } \
    } while (0)

void ptor_cosine_sample_hemisphere(int mode, float u1, float u2, float* out)
{
    V3 p;
    OR_DISPATCH_MATH(mode, p = cosine_sample_hemisphere<MathLibm>(u1, u2), p = cosine_sample_hemisphere<MathDet>(u1, u2));
    out[0] = p.x; out[1] = p.y; out[2] = p.z;
}
void ptor_uniform_sample_hemisphere(int mode, float u1, float u2, float* out)
{
    V3 p;
    OR_DISPATCH_MATH(mode, p = uniform_sample_hemisphere<MathLibm>(u1, u2), p = uniform_sample_hemisphere<MathDet>(u1, u2));
    out[0] = p.x; out[1] = p.y; out[2] = p.z;
}
void ptor_sample_ggx(int mode, float u1, float u2, float roughness, const float* n, float* out)
{
    V3 N = v3(n[0], n[1], n[2]), p;
    OR_DISPATCH_MATH(mode, p = sampleGGX<MathLibm>(u1, u2, roughness, N), p = sampleGGX<MathDet>(u1, u2, roughness, N));
    out[0] = p.x; out[1] = p.y; out[2] = p.z;
}
void ptor_fresnel_conductor(float cosTheta, const float* eta, const float* k, float* out)
{
    V3 r = fresnelSchlickConductor(cosTheta, v3(eta[0], eta[1], eta[2]), v3(k[0], k[1], k[2]));
    out[0] = r.x; out[1] = r.y; out[2] = r.z;
}
float ptor_fr_dielectric(float c, float etaI, float etaT) { return FrDielectric(c, etaI, etaT); }
int ptor_refract(const float* i, const float* n, float ior, float* out)
{
    V3 r;
    bool ok = refract(r, v3(i[0], i[1], i[2]), v3(n[0], n[1], n[2]), ior);
    out[0] = r.x; out[1] = r.y; out[2] = r.z;
    return ok ? 1 : 0;
}
void ptor_onb(const float* n, float* tangent, float* binormal)
{
    Onb o(v3(n[0], n[1], n[2]));
    tangent[0] = o.tangent.x; tangent[1] = o.tangent.y; tangent[2] = o.tangent.z;
    binormal[0] = o.binormal.x; binormal[1] = o.binormal.y; binormal[2] = o.binormal.z;
}
void ptor_make_color(int mode, const float* c, uint8_t* out)
{
    OR_DISPATCH_MATH(mode, make_color<MathLibm>(v3(c[0], c[1], c[2]), out), make_color<MathDet>(v3(c[0], c[1], c[2]), out));
}
// det-math probes (for accuracy tests against glibc)
void ptor_detmath(int fn, const float* x, float* y, uint64_t n, float yexp)
{
    for (uint64_t i = 0; i < n; ++i) {
        switch (fn) {
            case 0: y[i] = ptdet::sinf_(x[i]); break;
            case 1: y[i] = ptdet::cosf_(x[i]); break;
            case 2: y[i] = ptdet::acosf_(x[i]); break;
            case 3: y[i] = ptdet::powf_(x[i], yexp); break;
            case 10: y[i] = sinf(x[i]); break;
            case 11: y[i] = cosf(x[i]); break;
            case 12: y[i] = acosf(x[i]); break;
            case 13: y[i] = powf(x[i], yexp); break;
            default: y[i] = 0.0f;
        }
    }
}

// sutil/Camera.cpp:33-45
void ptor_camera_uvw(const float* eye, const float* lookat, const float* up, float fovY, float aspect, float* U,
                     float* V, float* W)
{
    V3 e = v3(eye[0], eye[1], eye[2]), l = v3(lookat[0], lookat[1], lookat[2]), u = v3(up[0], up[1], up[2]);
    V3 w = l - e;
    float wlen = length(w);
    V3 uu = normalize(cross(w, u));
    V3 vv = normalize(cross(uu, w));
    float vlen = wlen * tanf(0.5f * fovY * kPi / 180.0f);
    vv = vv * vlen;
    float ulen = vlen * aspect;
    uu = uu * ulen;
    U[0] = uu.x; U[1] = uu.y; U[2] = uu.z;
    V[0] = vv.x; V[1] = vv.y; V[2] = vv.z;
    W[0] = w.x; W[1] = w.y; W[2] = w.z;
}

// ---- scene handle ----------------------------------------------------------------------------------
void* ptor_scene_create(const float* verts_xyzw, uint32_t nV, const uint32_t* indices, uint32_t nT,
                        const uint32_t* matIds, const OrMaterial* mats, uint32_t nM)
{
    OrScene* sc = new OrScene();
    sc->verts.assign(verts_xyzw, verts_xyzw + 4 * (size_t)nV);
    sc->indices.assign(indices, indices + 3 * (size_t)nT);
    sc->mat_ids.assign(matIds, matIds + nT);
    sc->mats.assign(mats, mats + nM);
    sc->bvh = lbvh_build(sc->verts.data(), nV, sc->indices.data(), nT, sc->mat_ids.data());
    sah_build(sc->bvh);
    ploc_build(sc->bvh);
    select_tree(sc->bvh, 0);
    return sc;
}
void ptor_scene_destroy(void* s) { delete (OrScene*)s; }
void ptor_scene_set_background(void* s, float r, float g, float b)
{
    OrScene* sc = (OrScene*)s;
    sc->background[0] = r; sc->background[1] = g; sc->background[2] = b; sc->background[3] = 0.0f;
}

void ptor_scene_set_jitter_order(void* s, int order) { ((OrScene*)s)->jitter_order = order; }
void ptor_scene_set_decision_probe(void* s, uint32_t* buf) { ((OrScene*)s)->decision_probe = buf; }

// which tree the traversals walk: 0 auto (the library's rule), 1 LBVH, 2 SAH, 3 PLOC; returns the tree in use
int ptor_scene_set_tree(void* s, int tree) { return select_tree(((OrScene*)s)->bvh, tree); }
uint32_t ptor_bvh_export_ploc(void* s, void* nodes64)
{
    const Lbvh& b = ((OrScene*)s)->bvh;
    if (nodes64) std::memcpy(nodes64, b.ploc_nodes.data(), b.ploc_nodes.size() * sizeof(BvhNode));
    return b.ploc_depth;
}
uint32_t ptor_bvh_export_sah(void* s, void* nodes64)
{
    const Lbvh& b = ((OrScene*)s)->bvh;
    if (nodes64) std::memcpy(nodes64, b.sah_nodes.data(), b.sah_nodes.size() * sizeof(BvhNode));
    return b.sah_depth;
}

// LBVH export: sizes via NULL pointers
uint32_t ptor_bvh_num_nodes(void* s) { return (uint32_t)((OrScene*)s)->bvh.nodes.size(); }
void ptor_bvh_export(void* s, uint32_t* morton, uint64_t* sorted_keys, uint32_t* sorted_prim, int32_t* parent,
                     void* nodes64, void* tris48, float* scene_bounds6)
{
    const Lbvh& b = ((OrScene*)s)->bvh;
    if (morton) std::memcpy(morton, b.morton.data(), b.morton.size() * 4);
    if (sorted_keys) std::memcpy(sorted_keys, b.sorted_keys.data(), b.sorted_keys.size() * 8);
    if (sorted_prim) std::memcpy(sorted_prim, b.sorted_prim.data(), b.sorted_prim.size() * 4);
    if (parent) std::memcpy(parent, b.parent.data(), b.parent.size() * 4);
    if (nodes64) std::memcpy(nodes64, b.nodes.data(), b.nodes.size() * sizeof(BvhNode));
    if (tris48) std::memcpy(tris48, b.tris.data(), b.tris.size() * sizeof(TriRec));
    if (scene_bounds6) {
        for (int a = 0; a < 3; ++a) {
            scene_bounds6[a] = b.scene_lo[a];
            scene_bounds6[3 + a] = b.scene_hi[a];
        }
    }
}

// Batch ray queries: closest (prim or -1, t) and any-hit, brute force or BVH.
void ptor_trace(void* s, int use_bvh, const float* org3, const float* dir3, const float* tmin, const float* tmax,
                uint64_t n, int32_t* prim_out, float* t_out, uint8_t* any_out, unsigned long long* counters2)
{
    const OrScene& sc = *(OrScene*)s;
    TraceCounters tc;
    for (uint64_t i = 0; i < n; ++i) {
        const float* o = org3 + 3 * i;
        const float* d = dir3 + 3 * i;
        if (prim_out || t_out) {
            Hit h = use_bvh ? bvh_closest(sc.bvh, o, d, tmin[i], tmax[i], &tc)
                            : brute_closest(sc.bvh.tris_input, o, d, tmin[i], tmax[i]);
            if (prim_out) prim_out[i] = h.hit ? (int32_t)h.prim : -1;
            if (t_out) t_out[i] = h.hit ? h.t : 0.0f;
        }
        if (any_out) {
            any_out[i] = (use_bvh ? bvh_any(sc.bvh, o, d, tmin[i], tmax[i], &tc)
                                  : brute_any(sc.bvh.tris_input, o, d, tmin[i], tmax[i]))
                             ? 1
                             : 0;
        }
    }
    if (counters2) {
        counters2[0] = tc.n_node;
        counters2[1] = tc.n_tri;
    }
}

// One launch == one optixLaunch(width,height,1) of the reference (PathTracerMain.cpp:197-206).
// accum is read (frame > 0) and written in place; fb may be NULL.  Region [x0,x1) x [y0,y1) lets a
// caller render a crop (bounded CPU-baseline samples); pass 0,w,0,h for the whole image.
int ptor_render(void* s, const OrParams* params, int math_mode, int use_bvh, int nthreads, float* accum, uint8_t* fb,
                OrCounters* counters, int32_t* probe_prim, float* probe_t, uint32_t* probe_seed_after_jitter,
                uint32_t* probe_final_path_seed, uint32_t* probe_final_pixel_seed, float* probe_primary_dir,
                uint32_t x0, uint32_t x1, uint32_t y0, uint32_t y1)
{
    const OrScene& sc = *(OrScene*)s;
    if (!params || params->samplesPerPixel == 0 || params->width == 0 || params->height == 0) return 1;
    if (x1 > params->width) x1 = params->width;
    if (y1 > params->height) y1 = params->height;
    PixelProbe pp{probe_prim, probe_t, probe_seed_after_jitter, probe_final_path_seed, probe_final_pixel_seed,
                  probe_primary_dir};
    const bool any_probe = probe_prim || probe_t || probe_seed_after_jitter || probe_final_path_seed ||
                           probe_final_pixel_seed || probe_primary_dir;
    const PixelProbe* ppp = any_probe ? &pp : nullptr;
    if (math_mode == 0) {
        if (use_bvh) render_rows<MathLibm, true>(sc, *params, accum, fb, nthreads, counters, ppp, y0, y1, x0, x1);
        else render_rows<MathLibm, false>(sc, *params, accum, fb, nthreads, counters, ppp, y0, y1, x0, x1);
    } else {
        if (use_bvh) render_rows<MathDet, true>(sc, *params, accum, fb, nthreads, counters, ppp, y0, y1, x0, x1);
        else render_rows<MathDet, false>(sc, *params, accum, fb, nthreads, counters, ppp, y0, y1, x0, x1);
    }
    return 0;
}

int ptor_hardware_threads() { return (int)std::thread::hardware_concurrency(); }

}  // extern "C"
