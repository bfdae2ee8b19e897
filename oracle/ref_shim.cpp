// oracle/ref_shim.cpp -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// Builds the reference's OWN device programs for the host: this translation unit #includes
// $PT_REFERENCE_DIR/PathTracer_Optix/pathTracerPrograms.cu unmodified (it is located at build
// time under /root/reference and is never copied into this repository) on top of
// oracle/shim/optix.h, and drives __raygen__rg() once per pixel like optixLaunch(w,h,1) does
// (PathTracerMain.cpp:197-206).  The result, oracle/_ref/libptref.so, is
//   (1) what the oracle port (pt_oracle.cpp) is pinned against, bit for bit, and
//   (2) the "reference" CPU baseline of bench.py (--impl reference).
// OptiX's traversal / ray-triangle arithmetic is not in the reference tree; it is supplied by the
// oracle's own routines in oracle_geom.h (brute force or the CPU LBVH).
#include <atomic>
#include <cstdint>
#include <cstring>
#include <thread>
#include <vector>

#include <optix.h>   // oracle/shim/optix.h

thread_local ShimContext g_shim;

// The reference's programs, verbatim from the reference checkout.
#include "PathTracer_Optix/pathTracerPrograms.cu"

#include "oracle_geom.h"

static_assert(sizeof(PathTraceParams) == 168, "reference PathTraceParams is 168 bytes");
static_assert(sizeof(Material) == 40, "reference Material is 40 bytes");
static_assert(sizeof(HitGroupData) == 56, "reference HitGroupData is 56 bytes");

namespace {

struct RefScene {
    std::vector<float4> verts;
    std::vector<uint3> indices;
    std::vector<uint32_t> mat_ids;
    std::vector<HitGroupData> records;   // one per material, like the SBT (PathTracerMain.cpp:587-608)
    MissData miss;
    ora::Lbvh bvh;
    bool use_bvh = false;
};

const RefScene* g_scene = nullptr;   // read-only while rendering
thread_local ora::TraceCounters g_tc;

}  // namespace

ShimHit shim_closest_hit(float3 o, float3 d, float tmin, float tmax)
{
    const float oo[3] = {o.x, o.y, o.z}, dd[3] = {d.x, d.y, d.z};
    ora::Hit h = g_scene->use_bvh ? ora::bvh_closest(g_scene->bvh, oo, dd, tmin, tmax, &g_tc)
                                  : ora::brute_closest(g_scene->bvh.tris_input, oo, dd, tmin, tmax);
    return ShimHit{h.hit, h.prim, h.t};
}
bool shim_any_hit(float3 o, float3 d, float tmin, float tmax)
{
    const float oo[3] = {o.x, o.y, o.z}, dd[3] = {d.x, d.y, d.z};
    return g_scene->use_bvh ? ora::bvh_any(g_scene->bvh, oo, dd, tmin, tmax, &g_tc)
                            : ora::brute_any(g_scene->bvh.tris_input, oo, dd, tmin, tmax);
}
// SBT offset of a triangle == its material id (PathTracerMain.cpp:324-327)
const void* shim_hit_record(unsigned int prim) { return &g_scene->records[g_scene->mat_ids[prim]]; }
const void* shim_miss_record() { return &g_scene->miss; }
void shim_run_closesthit() { __closesthit__diffuse__ch(); }
void shim_run_miss() { __miss__ms(); }

struct RefCounters {
    unsigned long long n_samples, n_radiance, n_shadow, n_node, n_tri, n_hit[3], n_miss;
};

extern "C" {

unsigned ptref_sizeof_params() { return (unsigned)sizeof(PathTraceParams); }
unsigned ptref_offsetof_params(int field)
{
    switch (field) {
        case 0: return offsetof(PathTraceParams, currentFrameIdx);
        case 1: return offsetof(PathTraceParams, accumulationBuffer);
        case 2: return offsetof(PathTraceParams, frameBuffer);
        case 3: return offsetof(PathTraceParams, width);
        case 4: return offsetof(PathTraceParams, height);
        case 5: return offsetof(PathTraceParams, samplesPerPixel);
        case 6: return offsetof(PathTraceParams, maxDepth);
        case 7: return offsetof(PathTraceParams, cameraEye);
        case 8: return offsetof(PathTraceParams, cameraU);
        case 9: return offsetof(PathTraceParams, cameraV);
        case 10: return offsetof(PathTraceParams, cameraW);
        case 11: return offsetof(PathTraceParams, areaLight);
        case 12: return offsetof(PathTraceParams, handle);
        case 13: return offsetof(PathTraceParams, useDirectLighting);
        case 14: return offsetof(PathTraceParams, useImportanceSampling);
        default: return 0xffffffffu;
    }
}
unsigned ptref_sizeof_material() { return (unsigned)sizeof(Material); }
unsigned ptref_sizeof_hitgroup() { return (unsigned)sizeof(HitGroupData); }

// Known-answer probes straight into the reference's functions.
uint32_t ptref_tea4(uint32_t a, uint32_t b) { return tea<4>(a, b); }
float ptref_rnd(uint32_t* s) { return rnd(*s); }
void ptref_cosine_sample_hemisphere(float u1, float u2, float* out)
{
    float3 p;
    cosine_sample_hemisphere(u1, u2, p);
    out[0] = p.x; out[1] = p.y; out[2] = p.z;
}
void ptref_uniform_sample_hemisphere(float u1, float u2, float* out)
{
    float3 p;
    uniform_sample_hemisphere(u1, u2, p);
    out[0] = p.x; out[1] = p.y; out[2] = p.z;
}
void ptref_sample_ggx(float u1, float u2, float roughness, const float* n, float* out)
{
    float3 p = sampleGGX(u1, u2, roughness, make_float3(n[0], n[1], n[2]));
    out[0] = p.x; out[1] = p.y; out[2] = p.z;
}
void ptref_fresnel_conductor(float c, const float* eta, const float* k, float* out)
{
    float3 r = fresnelSchlickConductor(c, make_float3(eta[0], eta[1], eta[2]), make_float3(k[0], k[1], k[2]));
    out[0] = r.x; out[1] = r.y; out[2] = r.z;
}
float ptref_fr_dielectric(float c, float etaI, float etaT) { return FrDielectric(c, etaI, etaT); }
int ptref_refract(const float* i, const float* n, float ior, float* out)
{
    float3 r;
    bool ok = refract(r, make_float3(i[0], i[1], i[2]), make_float3(n[0], n[1], n[2]), ior);
    out[0] = r.x; out[1] = r.y; out[2] = r.z;
    return ok ? 1 : 0;
}
void ptref_onb(const float* n, float* tangent, float* binormal)
{
    OrthonormalBasis o(make_float3(n[0], n[1], n[2]));
    tangent[0] = o.m_tangent.x; tangent[1] = o.m_tangent.y; tangent[2] = o.m_tangent.z;
    binormal[0] = o.m_binormal.x; binormal[1] = o.m_binormal.y; binormal[2] = o.m_binormal.z;
}
void ptref_make_color(const float* c, uint8_t* out)
{
    uchar4 q = make_color(make_float3(c[0], c[1], c[2]));
    out[0] = q.x; out[1] = q.y; out[2] = q.z; out[3] = q.w;
}

// One launch of the reference raygen over [x0,x1) x [y0,y1) (whole image: 0,w,0,h).
// `params168` is the caller's PathTraceParams image (byte-identical layout); its buffer pointers
// and handle are ignored and replaced by `accum` / `fb`.
int ptref_render(const void* params168, const float* verts_xyzw, uint32_t nV, const uint32_t* indices, uint32_t nT,
                 const uint32_t* matIds, const void* mats40, uint32_t nM, const float* bg4, int use_bvh, int nthreads,
                 float* accum, uint8_t* fb, RefCounters* counters, uint32_t x0, uint32_t x1, uint32_t y0, uint32_t y1)
{
    static RefScene scene;   // single-scene harness: not re-entrant
    scene = RefScene();
    scene.verts.resize(nV);
    for (uint32_t i = 0; i < nV; ++i)
        scene.verts[i] = make_float4(verts_xyzw[4 * i], verts_xyzw[4 * i + 1], verts_xyzw[4 * i + 2], verts_xyzw[4 * i + 3]);
    scene.indices.resize(nT);
    for (uint32_t i = 0; i < nT; ++i) scene.indices[i] = make_uint3(indices[3 * i], indices[3 * i + 1], indices[3 * i + 2]);
    scene.mat_ids.assign(matIds, matIds + nT);
    const Material* mats = (const Material*)mats40;
    scene.records.resize(nM);
    for (uint32_t m = 0; m < nM; ++m) {   // PathTracerMain.cpp:587-608
        HitGroupData& r = scene.records[m];
        r.bsdfType = mats[m].bsdfType;
        r.emissionColor = mats[m].emission;
        r.diffuseColor = mats[m].diffuse;
        r.IOR = mats[m].ior;
        r.roughness = mats[m].roughness;
        r.metallic = mats[m].metallic;
        r.vertices = scene.verts.data();
        r.indices = scene.indices.data();
    }
    scene.miss.backgroundColor = make_float4(bg4[0], bg4[1], bg4[2], bg4[3]);
    scene.bvh = ora::lbvh_build(verts_xyzw, nV, indices, nT, matIds);
    ora::sah_build(scene.bvh);   // same tree choice as the port and the library (work counters are compared)
    ora::ploc_build(scene.bvh);
    ora::select_tree(scene.bvh, 0);
    scene.use_bvh = use_bvh != 0;
    g_scene = &scene;

    std::memcpy(&params, params168, sizeof(PathTraceParams));
    if (params.samplesPerPixel == 0 || params.width == 0 || params.height == 0) return 1;
    std::vector<uint8_t> scratch_fb;
    if (!fb) {
        scratch_fb.resize(4 * (size_t)params.width * params.height);
        fb = scratch_fb.data();
    }
    params.accumulationBuffer = (float4*)accum;
    params.frameBuffer = (uchar4*)fb;
    params.handle = 1;
    if (x1 > params.width) x1 = params.width;
    if (y1 > params.height) y1 = params.height;

    if (nthreads < 1) nthreads = 1;
    std::atomic<uint32_t> next_row(y0);
    std::vector<RefCounters> ctrs(nthreads);
    auto work = [&](int tid) {
        g_shim = ShimContext();
        g_tc = ora::TraceCounters();
        for (;;) {
            uint32_t y = next_row.fetch_add(1);
            if (y >= y1) break;
            for (uint32_t x = x0; x < x1; ++x) {
                g_shim.launch_index = make_uint3(x, y, 0);
                __raygen__rg();
            }
        }
        RefCounters c;
        std::memset(&c, 0, sizeof(c));
        c.n_radiance = g_shim.n_radiance;
        c.n_shadow = g_shim.n_shadow;
        c.n_node = g_tc.n_node;
        c.n_tri = g_tc.n_tri;
        ctrs[tid] = c;
    };
    std::vector<std::thread> pool;
    for (int t = 1; t < nthreads; ++t) pool.emplace_back(work, t);
    work(0);
    for (auto& th : pool) th.join();
    if (counters) {
        std::memset(counters, 0, sizeof(*counters));
        for (const RefCounters& c : ctrs) {
            counters->n_radiance += c.n_radiance;
            counters->n_shadow += c.n_shadow;
            counters->n_node += c.n_node;
            counters->n_tri += c.n_tri;
        }
        counters->n_samples = (unsigned long long)(x1 - x0) * (y1 - y0) * params.samplesPerPixel;
    }
    g_scene = nullptr;
    return 0;
}

// pathTracerPrograms.cu:730 draws the two jitter values inside one argument list,
// make_float2(rnd(seed), rnd(seed)); C++ leaves the evaluation order unspecified.  nvcc's device code
// draws x first; g++ on x86-64 evaluates right to left, so in THIS host build y gets the
// first draw.  Returns 0 when x got the first draw, 1 when y did (same expression, same compiler,
// same flags as the #included raygen).
int ptref_jitter_order()
{
    unsigned int seed = 12345u, probe = 12345u;
    const float first = rnd(probe);
    const float2 j = make_float2(rnd(seed), rnd(seed));
    return j.x == first ? 0 : 1;
}

int ptref_hardware_threads() { return (int)std::thread::hardware_concurrency(); }

}  // extern "C"
