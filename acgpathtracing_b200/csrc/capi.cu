// capi.cu -- the extern "C" boundary declared in include/ptb200.h.
//
// Replaces what PathTracerMain.cpp does through the OptiX host API between "arrays from TinyObjWrapper" and
// "pixels in two buffers" (upload + optixAccelBuild :260-398, SBT :544-608, optixLaunch :184-210).  Nothing throws
// across this boundary; every entry point returns a PT_* status and records a message for pt_last_error().
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <chrono>
#include <mutex>
#include <string>
#include <unordered_set>
#include <vector>

#include "pt_scene.h"
#include "pt_shade.cuh"

namespace ptb {

static thread_local std::string g_error;
void set_error(const std::string& msg) { g_error = msg; }

static std::mutex g_scene_mutex;
static std::unordered_set<const Scene*> g_scenes;

int fail(int code, const std::string& msg)
{
    set_error(msg);
    return code;
}
int cuda_fail(cudaError_t e, const char* what)
{
    return fail(PT_ERR_CUDA, std::string(what) + ": " + cudaGetErrorName(e) + " (" + cudaGetErrorString(e) + ")");
}
#define PT_CUDA(expr)                                     \
    do {                                                  \
        cudaError_t e__ = (expr);                         \
        if (e__ != cudaSuccess) return cuda_fail(e__, #expr); \
    } while (0)

Scene* scene_from_handle(uint64_t handle)
{
    Scene* s = reinterpret_cast<Scene*>(static_cast<uintptr_t>(handle));
    std::lock_guard<std::mutex> lock(g_scene_mutex);
    if (!s || g_scenes.find(s) == g_scenes.end() || s->magic != kSceneMagic) return nullptr;
    return s;
}

struct DeviceGuard {
    int prev = -1;
    bool switched = false;
    explicit DeviceGuard(int dev)
    {
        if (cudaGetDevice(&prev) == cudaSuccess && prev != dev) switched = cudaSetDevice(dev) == cudaSuccess;
    }
    ~DeviceGuard()
    {
        if (switched) cudaSetDevice(prev);
    }
};

template <class T>
static cudaError_t dev_alloc(T** p, size_t count)
{
    *p = nullptr;
    return cudaMalloc((void**)p, (count ? count : 1) * sizeof(T));
}

static void scene_free(Scene* s)
{
    void* ptrs[] = {s->d_verts, s->d_indices, s->d_mat_ids, s->d_mats, s->d_tri_in, s->d_tri_bb, s->d_tri_shade,
                    s->d_bounds_ordered, s->d_bounds, s->d_morton, s->d_keys[0], s->d_keys[1], s->d_vals[0], s->d_vals[1],
                    s->d_block_hist, s->d_scan_scratch, s->d_left, s->d_right, s->d_node_parent, s->d_leaf_parent, s->d_visit, s->d_node_lo,
                    s->d_node_hi, s->d_nodes, s->d_nodes_sah, s->d_nodes_ploc, s->d_tris, s->d_qnodes, s->d_wnodes, s->d_work_counter, s->d_counters, s->d_stage_accum, s->d_stage_frame,
                    s->d_stack_over, s->d_team_log};
    for (void* p : ptrs)
        if (p) cudaFree(p);
    for (TileOrder& t : s->tile_orders) {
        if (t.d) cudaFree(t.d);
        if (t.h) cudaFreeHost(t.h);
        if (t.copied) cudaEventDestroy(t.copied);
    }
    s->magic = 0;
    delete s;
}

// Builds the PLOC tree of a scene that does not have one yet (synchronous; the time goes into ploc_ms and build_ms).
cudaError_t ensure_ploc(Scene& s)
{
    if (s.d_nodes_ploc || s.ploc_tried || s.n_tris < 2) return cudaSuccess;
    s.ploc_tried = true;
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    cudaError_t e = cudaEventCreate(&e0);
    if (e == cudaSuccess) e = cudaEventCreate(&e1);
    cudaStream_t stream = 0;
    if (e == cudaSuccess) e = cudaEventRecord(e0, stream);
    if (e == cudaSuccess) e = ploc_build(s, stream);
    if (e == cudaSuccess) e = cudaEventRecord(e1, stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(stream);
    if (e == cudaSuccess) {
        cudaEventElapsedTime(&s.ploc_ms, e0, e1);
        s.build_ms += s.ploc_ms;
    }
    if (e0) cudaEventDestroy(e0);
    if (e1) cudaEventDestroy(e1);
    return e;
}

static int validate_params(const pt_params* p, Scene** out)
{
    if (!p) return fail(PT_ERR_INVALID, "params is NULL");
    if (p->width == 0 || p->height == 0) return fail(PT_ERR_INVALID, "width and height must be > 0");
    if (p->samplesPerPixel == 0) return fail(PT_ERR_INVALID, "samplesPerPixel must be > 0");
    if ((uint64_t)p->width * p->height > 0x7fffffffull) return fail(PT_ERR_INVALID, "image too large");
    if (!p->accumulationBuffer) return fail(PT_ERR_INVALID, "accumulationBuffer is NULL");
    Scene* s = scene_from_handle(p->handle);
    if (!s) return fail(PT_ERR_INVALID, "params.handle is not a live pt_scene handle");
    *out = s;
    return PT_OK;
}

}  // namespace ptb

using namespace ptb;

extern "C" {

const char* pt_last_error(void) { return g_error.c_str(); }

int pt_scene_create(const float* verts_xyzw, uint32_t n_verts, const uint32_t* indices, uint32_t n_tris,
                    const uint32_t* mat_ids, const pt_material* materials, uint32_t n_materials, pt_scene** out)
{
    if (!out) return fail(PT_ERR_INVALID, "out is NULL");
    *out = nullptr;
    if (n_tris > 0 && (!verts_xyzw || !indices || !mat_ids || !materials))
        return fail(PT_ERR_INVALID, "NULL input array");
    if (n_tris > 0 && (n_verts == 0 || n_materials == 0)) return fail(PT_ERR_INVALID, "triangles without vertices or materials");
    if (n_tris > 0x3fffffffu) return fail(PT_ERR_INVALID, "too many triangles");
    for (uint64_t i = 0; i < 3ull * n_tris; ++i)
        if (indices[i] >= n_verts) return fail(PT_ERR_INVALID, "vertex index out of range at index-buffer position " + std::to_string(i));
    for (uint32_t k = 0; k < n_tris; ++k)
        if (mat_ids[k] >= n_materials)   // tinyobj's "no usemtl" id -1 arrives here as 0xffffffff
            return fail(PT_ERR_INVALID, "material id out of range at triangle " + std::to_string(k));
    for (uint32_t m = 0; m < n_materials; ++m)
        if (materials[m].bsdfType < 0 || materials[m].bsdfType > 2)
            return fail(PT_ERR_INVALID, "bad bsdfType in material " + std::to_string(m));

    int dev = 0;
    PT_CUDA(cudaGetDevice(&dev));
    cudaDeviceProp prop;
    PT_CUDA(cudaGetDeviceProperties(&prop, dev));
    if (prop.major < 10) return fail(PT_ERR_CUDA, std::string("device is not sm_100-class: ") + prop.name);

    Scene* s = new Scene();
    s->device = dev;
    s->n_verts = n_verts;
    s->n_tris = n_tris;
    s->n_mats = n_materials;
    s->n_nodes = n_tris == 0 ? 0 : (n_tris > 1 ? n_tris - 1 : 1);
    const size_t n = n_tris;
    const size_t n_blocks = (n + 2047) / 2048;
    cudaError_t e = cudaSuccess;
    auto A = [&](cudaError_t r) {
        if (e == cudaSuccess) e = r;
    };
    A(dev_alloc(&s->d_verts, n_verts));
    A(dev_alloc(&s->d_indices, 3 * n));
    A(dev_alloc(&s->d_mat_ids, n));
    A(dev_alloc(&s->d_mats, n_materials));
    A(dev_alloc(&s->d_tri_in, 3 * n));
    A(dev_alloc(&s->d_tri_bb, 2 * n));
    A(dev_alloc(&s->d_tri_shade, n));
    A(dev_alloc(&s->d_bounds_ordered, 6));
    A(dev_alloc(&s->d_bounds, 6));
    A(dev_alloc(&s->d_morton, n));
    for (int b = 0; b < 2; ++b) {
        A(dev_alloc(&s->d_keys[b], n));
        A(dev_alloc(&s->d_vals[b], n));
    }
    A(dev_alloc(&s->d_block_hist, 256 * n_blocks));
    A(dev_alloc(&s->d_scan_scratch, lbvh_scan_scratch_entries(n)));
    A(dev_alloc(&s->d_left, n));
    A(dev_alloc(&s->d_right, n));
    A(dev_alloc(&s->d_node_parent, n));
    A(dev_alloc(&s->d_leaf_parent, n));
    A(dev_alloc(&s->d_visit, n + 1));
    A(dev_alloc(&s->d_node_lo, n));
    A(dev_alloc(&s->d_node_hi, n));
    A(dev_alloc(&s->d_nodes, 4 * (size_t)(s->n_nodes ? s->n_nodes : 1)));
    A(dev_alloc(&s->d_tris, 3 * n));
    A(dev_alloc(&s->d_work_counter, 4));
    A(dev_alloc(&s->d_counters, 32));
    if (e != cudaSuccess) {
        scene_free(s);
        return cuda_fail(e, "cudaMalloc(scene)");
    }
    std::vector<ptd::DevMaterial> dm(n_materials);
    for (uint32_t m = 0; m < n_materials; ++m) {
        const pt_material& src = materials[m];
        ptd::DevMaterial& d = dm[m];
        memset(&d, 0, sizeof(d));
        d.diffuse = make_float3(src.diffuse.x, src.diffuse.y, src.diffuse.z);
        d.emission = make_float3(src.emission.x, src.emission.y, src.emission.z);
        d.ior = src.ior;
        d.bsdf = src.bsdfType;
        // pathTracerPrograms.cu:992  length(rt_data->emissionColor) > 0.0f
        const float l2 = src.emission.x * src.emission.x + src.emission.y * src.emission.y + src.emission.z * src.emission.z;
        d.emissive = sqrtf(l2) > 0.0f ? 1 : 0;
    }
    {
        bool seen[3] = {false, false, false};
        for (uint32_t m = 0; m < n_materials; ++m) seen[materials[m].bsdfType] = true;
        s->n_bsdf_types = (uint32_t)seen[0] + (uint32_t)seen[1] + (uint32_t)seen[2];
    }
    cudaStream_t stream = 0;
    A(cudaMemsetAsync(s->d_counters, 0, 32 * sizeof(unsigned long long), stream));
    if (n_verts) A(cudaMemcpyAsync(s->d_verts, verts_xyzw, sizeof(float4) * n_verts, cudaMemcpyHostToDevice, stream));
    if (n) {
        A(cudaMemcpyAsync(s->d_indices, indices, sizeof(uint32_t) * 3 * n, cudaMemcpyHostToDevice, stream));
        A(cudaMemcpyAsync(s->d_mat_ids, mat_ids, sizeof(uint32_t) * n, cudaMemcpyHostToDevice, stream));
    }
    if (n_materials) A(cudaMemcpyAsync(s->d_mats, dm.data(), sizeof(ptd::DevMaterial) * n_materials, cudaMemcpyHostToDevice, stream));
    cudaEvent_t e0 = nullptr, e1 = nullptr;
    A(cudaEventCreate(&e0));
    A(cudaEventCreate(&e1));
    if (e == cudaSuccess) {
        A(cudaEventRecord(e0, stream));
        A(lbvh_build(*s, stream));
        A(cudaEventRecord(e1, stream));
        A(cudaStreamSynchronize(stream));
        if (e == cudaSuccess) cudaEventElapsedTime(&s->build_ms, e0, e1);
        if (e == cudaSuccess && n) A(cudaMemcpy(&s->bvh_depth, s->d_work_counter + 1, sizeof(uint32_t), cudaMemcpyDeviceToHost));
        if (e == cudaSuccess && n) A(cudaMemcpy(s->h_bounds, s->d_bounds, 6 * sizeof(float), cudaMemcpyDeviceToHost));
    }
    if (e0) cudaEventDestroy(e0);
    if (e1) cudaEventDestroy(e1);
    // small scenes: sweep-SAH tree over the LBVH's leaves, built on the host (sah.cu); sah_ms is host wall-clock
    if (e == cudaSuccess && n >= 2 && n <= kSahMaxTris) {
        const auto t0 = std::chrono::steady_clock::now();
        std::vector<float4> bb(2 * n), nodes;
        std::vector<uint32_t> prim(n);
        A(cudaMemcpy(bb.data(), s->d_tri_bb, sizeof(float4) * 2 * n, cudaMemcpyDeviceToHost));
        A(cudaMemcpy(prim.data(), s->d_vals[s->sorted_buf], sizeof(uint32_t) * n, cudaMemcpyDeviceToHost));
        if (e == cudaSuccess) {
            sah_build_host(bb.data(), prim.data(), (uint32_t)n, nodes, s->sah_depth);
            A(dev_alloc(&s->d_nodes_sah, nodes.size()));
            A(cudaMemcpy(s->d_nodes_sah, nodes.data(), sizeof(float4) * nodes.size(), cudaMemcpyHostToDevice));
        }
        s->sah_ms = std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - t0).count();
    }
    // PLOC tree over the same leaves (GPU): built now only where it is the tree the kernels will walk (no usable SAH tree);
    // smaller scenes get it on demand (pt_scene_set_tree(scene, 3), pt_bvh_export_ploc)
    if (e == cudaSuccess && n >= 2 && !(s->d_nodes_sah && s->sah_depth <= kSahMaxDepth)) A(ensure_ploc(*s));
    if (e != cudaSuccess) {
        scene_free(s);
        return cuda_fail(e, "scene upload / LBVH build");
    }
    {
        std::lock_guard<std::mutex> lock(g_scene_mutex);
        g_scenes.insert(s);
    }
    *out = reinterpret_cast<pt_scene*>(s);
    return PT_OK;
}

void pt_scene_destroy(pt_scene* scene)
{
    Scene* s = reinterpret_cast<Scene*>(scene);
    if (!s) return;
    {
        std::lock_guard<std::mutex> lock(g_scene_mutex);
        if (g_scenes.erase(s) == 0) return;
    }
    DeviceGuard g(s->device);
    cudaDeviceSynchronize();
    scene_free(s);
}

int pt_scene_set_background(pt_scene* scene, float r, float g, float b)
{
    Scene* s = scene_from_handle((uint64_t)(uintptr_t)scene);
    if (!s) return fail(PT_ERR_INVALID, "not a live pt_scene");
    s->background[0] = r;
    s->background[1] = g;
    s->background[2] = b;
    return PT_OK;
}

uint64_t pt_scene_handle(const pt_scene* scene) { return (uint64_t)(uintptr_t)scene; }

int pt_scene_info(const pt_scene* scene, uint32_t* n_tris, uint32_t* n_nodes, float* build_ms)
{
    Scene* s = scene_from_handle((uint64_t)(uintptr_t)scene);
    if (!s) return fail(PT_ERR_INVALID, "not a live pt_scene");
    if (n_tris) *n_tris = s->n_tris;
    if (n_nodes) *n_nodes = s->n_nodes;
    if (build_ms) *build_ms = s->build_ms;
    return PT_OK;
}

int pt_launch_partition(const pt_params* params, uint32_t rank, uint32_t n_ranks, void* cuda_stream)
{
    Scene* s = nullptr;
    int rc = validate_params(params, &s);
    if (rc != PT_OK) return rc;
    if (n_ranks == 0 || rank >= n_ranks) return fail(PT_ERR_INVALID, "rank out of range");
    DeviceGuard g(s->device);
    PT_CUDA(render_launch(*s, *params, rank, n_ranks, (cudaStream_t)cuda_stream, nullptr));
    return PT_OK;
}

int pt_launch(const pt_params* params, void* cuda_stream) { return pt_launch_partition(params, 0, 1, cuda_stream); }

int pt_launch_raw(const pt_params* params, uint32_t rank, uint32_t n_ranks, void* cuda_stream)
{
    Scene* s = nullptr;
    int rc = validate_params(params, &s);
    if (rc != PT_OK) return rc;
    if (n_ranks == 0 || rank >= n_ranks) return fail(PT_ERR_INVALID, "rank out of range");
    DeviceGuard g(s->device);
    PT_CUDA(render_launch(*s, *params, rank, n_ranks, (cudaStream_t)cuda_stream, nullptr, true));
    return PT_OK;
}

int pt_fold_frames(void* accum, void* frame_out, const void* const* frame_means, uint32_t first_frame_idx, uint32_t n_frames,
                   uint32_t width, uint32_t height, void* cuda_stream)
{
    if (!accum || !frame_means || n_frames == 0 || width == 0 || height == 0) return fail(PT_ERR_INVALID, "pt_fold_frames: bad argument");
    if (n_frames > kMaxFoldFrames) return fail(PT_ERR_INVALID, "pt_fold_frames: at most 64 frames per call");
    for (uint32_t k = 0; k < n_frames; ++k)
        if (!frame_means[k]) return fail(PT_ERR_INVALID, "pt_fold_frames: NULL frame");
    PT_CUDA(fold_frames(accum, frame_out, frame_means, first_frame_idx, n_frames, width, height, (cudaStream_t)cuda_stream));
    return PT_OK;
}

uint64_t pt_partition_capacity(uint32_t width, uint32_t height, uint32_t n_ranks)
{
    if (n_ranks == 0) return 0;
    return (uint64_t)partition_max_owned_tiles(width, height, n_ranks) * 32u;
}

int pt_partition_pack(const void* accum_full, const void* frame_full, void* accum_compact, void* frame_compact,
                      uint32_t width, uint32_t height, uint32_t rank, uint32_t n_ranks, void* cuda_stream)
{
    if (!accum_full || !accum_compact || n_ranks == 0 || rank >= n_ranks || width == 0 || height == 0)
        return fail(PT_ERR_INVALID, "pt_partition_pack: bad argument");
    PT_CUDA(partition_pack(accum_full, frame_full, accum_compact, frame_compact, width, height, rank, n_ranks,
                           (cudaStream_t)cuda_stream));
    return PT_OK;
}

int pt_partition_unpack(const void* accum_compact_all, const void* frame_compact_all, void* accum_full,
                        void* frame_full, uint32_t width, uint32_t height, uint32_t n_ranks, void* cuda_stream)
{
    if (!accum_compact_all || !accum_full || n_ranks == 0 || width == 0 || height == 0)
        return fail(PT_ERR_INVALID, "pt_partition_unpack: bad argument");
    PT_CUDA(partition_unpack(accum_compact_all, frame_compact_all, accum_full, frame_full, width, height, n_ranks,
                             (cudaStream_t)cuda_stream));
    return PT_OK;
}

int pt_launch_host(const pt_params* hp, uint64_t* h2d_bytes, uint64_t* d2h_bytes)
{
    Scene* s = nullptr;
    int rc = validate_params(hp, &s);
    if (rc != PT_OK) return rc;
    DeviceGuard g(s->device);
    const size_t pixels = (size_t)hp->width * hp->height;
    if (s->stage_pixels < pixels) {
        if (s->d_stage_accum) cudaFree(s->d_stage_accum);
        if (s->d_stage_frame) cudaFree(s->d_stage_frame);
        s->d_stage_accum = nullptr;
        s->d_stage_frame = nullptr;
        s->stage_pixels = 0;
        PT_CUDA(cudaMalloc((void**)&s->d_stage_accum, pixels * sizeof(float4)));
        PT_CUDA(cudaMalloc((void**)&s->d_stage_frame, pixels * sizeof(uchar4)));
        s->stage_pixels = pixels;
    }
    cudaStream_t stream = 0;
    uint64_t up = 0, down = 0;
    if (hp->currentFrameIdx > 0) {   // the running average reads the previous accumulation (pathTracerPrograms.cu:803-810)
        PT_CUDA(cudaMemcpyAsync(s->d_stage_accum, hp->accumulationBuffer, pixels * sizeof(float4), cudaMemcpyHostToDevice, stream));
        up += pixels * sizeof(float4);
    }
    // test aid (PTB_POISON=1 in the environment): frame 0 must write every pixel, so the staging image starts as NaNs --
    // a pixel the kernel failed to write cannot hide behind the previous launch's value
    static const bool poison = getenv("PTB_POISON") != nullptr;
    if (poison && hp->currentFrameIdx == 0) PT_CUDA(cudaMemsetAsync(s->d_stage_accum, 0xff, pixels * sizeof(float4), stream));
    pt_params dp = *hp;
    dp.accumulationBuffer = s->d_stage_accum;
    dp.frameBuffer = hp->frameBuffer ? s->d_stage_frame : nullptr;
    up += sizeof(pt_params);   // the launch record itself (reference: cudaMemcpyAsync(d_params), PathTracerMain.cpp:191)
    PT_CUDA(render_launch(*s, dp, 0, 1, stream, nullptr));
    PT_CUDA(cudaMemcpyAsync(hp->accumulationBuffer, s->d_stage_accum, pixels * sizeof(float4), cudaMemcpyDeviceToHost, stream));
    down += pixels * sizeof(float4);
    if (hp->frameBuffer) {
        PT_CUDA(cudaMemcpyAsync(hp->frameBuffer, s->d_stage_frame, pixels * sizeof(uchar4), cudaMemcpyDeviceToHost, stream));
        down += pixels * sizeof(uchar4);
    }
    unsigned long long watchdog = 0;
    PT_CUDA(cudaMemcpyAsync(&watchdog, s->d_counters + 15, sizeof(watchdog), cudaMemcpyDeviceToHost, stream));
    PT_CUDA(cudaStreamSynchronize(stream));
    if (watchdog) return fail(PT_ERR_CUDA, "render kernel watchdog fired (scheduler made no progress)");
    if (h2d_bytes) *h2d_bytes = up;
    if (d2h_bytes) *d2h_bytes = down;
    return PT_OK;
}

int pt_scene_enable_counters(pt_scene* scene, int enable)
{
    Scene* s = scene_from_handle((uint64_t)(uintptr_t)scene);
    if (!s) return fail(PT_ERR_INVALID, "not a live pt_scene");
    if (enable < 0 || enable > 2) return fail(PT_ERR_INVALID, "pt_scene_enable_counters: 0, 1 or 2");
    s->count_enabled = enable != 0;
    s->count_executed = enable == 2;
    return PT_OK;
}

int pt_scene_get_counters(pt_scene* scene, pt_counters* out, int reset)
{
    Scene* s = scene_from_handle((uint64_t)(uintptr_t)scene);
    if (!s || !out) return fail(PT_ERR_INVALID, "pt_scene_get_counters: bad argument");
    DeviceGuard g(s->device);
    PT_CUDA(cudaDeviceSynchronize());
    unsigned long long h[16];
    PT_CUDA(cudaMemcpy(h, s->d_counters, sizeof(h), cudaMemcpyDeviceToHost));
    out->n_samples = h[0];
    out->n_radiance = h[1];
    out->n_shadow = h[2];
    out->n_node = h[3];
    out->n_tri = h[4];
    out->n_hit[0] = h[5];
    out->n_hit[1] = h[6];
    out->n_hit[2] = h[7];
    out->n_miss = h[8];
    if (h[15]) return fail(PT_ERR_CUDA, "render kernel watchdog fired (scheduler made no progress)");
    if (reset) PT_CUDA(cudaMemset(s->d_counters, 0, sizeof(h)));
    return PT_OK;
}

int pt_scene_get_sched_stats(pt_scene* scene, uint64_t* out16, int reset)
{
    Scene* s = scene_from_handle((uint64_t)(uintptr_t)scene);
    if (!s || !out16) return fail(PT_ERR_INVALID, "pt_scene_get_sched_stats: bad argument");
    DeviceGuard g(s->device);
    PT_CUDA(cudaDeviceSynchronize());
    PT_CUDA(cudaMemcpy(out16, s->d_counters + 16, 16 * sizeof(uint64_t), cudaMemcpyDeviceToHost));
    if (reset) PT_CUDA(cudaMemset(s->d_counters + 16, 0, 16 * sizeof(uint64_t)));
    return PT_OK;
}

int pt_scene_set_pipeline(pt_scene* scene, int pipeline)
{
    Scene* s = scene_from_handle((uint64_t)(uintptr_t)scene);
    if (!s) return fail(PT_ERR_INVALID, "not a live pt_scene");
    if (pipeline < 0 || pipeline > 2) return fail(PT_ERR_INVALID, "pipeline must be 0, 1 or 2");
    s->pipeline = pipeline;
    return PT_OK;
}

int pt_scene_set_tune(pt_scene* scene, const char* spec)
{
    Scene* s = scene_from_handle((uint64_t)(uintptr_t)scene);
    if (!s) return fail(PT_ERR_INVALID, "not a live pt_scene");
    WfTune t = WfTune();                       // every call starts from the defaults: "" resets
    if (!wf_tune_apply(t, spec)) return fail(PT_ERR_INVALID, std::string("pt_scene_set_tune: unknown key in \"") + (spec ? spec : "") + "\"");
    s->wf = t;
    return PT_OK;
}

int pt_scene_check_error(pt_scene* scene)
{
    Scene* s = scene_from_handle((uint64_t)(uintptr_t)scene);
    if (!s) return fail(PT_ERR_INVALID, "not a live pt_scene");
    DeviceGuard g(s->device);
    PT_CUDA(cudaDeviceSynchronize());
    unsigned long long watchdog = 0;
    PT_CUDA(cudaMemcpy(&watchdog, s->d_counters + 15, sizeof(watchdog), cudaMemcpyDeviceToHost));
    if (watchdog) return fail(PT_ERR_CUDA, "render kernel watchdog fired (scheduler made no progress)");
    return PT_OK;
}

int pt_scene_set_traversal(pt_scene* scene, int traversal)
{
    Scene* s = scene_from_handle((uint64_t)(uintptr_t)scene);
    if (!s) return fail(PT_ERR_INVALID, "not a live pt_scene");
    if (traversal < 0 || traversal > 4) return fail(PT_ERR_INVALID, "traversal must be 0 .. 4");
    s->traversal = traversal;
    return PT_OK;
}

int pt_scene_set_tree(pt_scene* scene, int tree)
{
    Scene* s = scene_from_handle((uint64_t)(uintptr_t)scene);
    if (!s) return fail(PT_ERR_INVALID, "not a live pt_scene");
    if (tree < 0 || tree > 3) return fail(PT_ERR_INVALID, "tree must be 0, 1, 2 or 3");
    if (tree == 3) {
        DeviceGuard g(s->device);
        PT_CUDA(ensure_ploc(*s));
    }
    if (tree == 2 && !s->d_nodes_sah && s->n_tris >= 2) return fail(PT_ERR_INVALID, "no SAH tree for this scene (more than 8192 triangles)");
    if (tree == 3 && !s->d_nodes_ploc && s->n_tris >= 2) return fail(PT_ERR_INVALID, "no PLOC tree for this scene");
    if ((tree == 2 && s->d_nodes_sah && s->sah_depth > kSahMaxDepth) || (tree == 3 && s->d_nodes_ploc && s->ploc_depth > kSahMaxDepth))
        return fail(PT_ERR_INVALID, "that tree is more than 64 levels deep for this scene (degenerate input); the LBVH stays in use");
    s->tree = tree;
    return PT_OK;
}

int pt_scene_tree_in_use(const pt_scene* scene)
{
    Scene* s = scene_from_handle((uint64_t)(uintptr_t)scene);
    return s ? scene_tree(*s) : -1;
}

int pt_bvh_export_sah(const pt_scene* scene, void* nodes64, uint32_t* depth)
{
    Scene* s = scene_from_handle((uint64_t)(uintptr_t)scene);
    if (!s) return fail(PT_ERR_INVALID, "not a live pt_scene");
    if (!s->d_nodes_sah) return fail(PT_ERR_INVALID, "no SAH tree for this scene");
    DeviceGuard g(s->device);
    if (depth) *depth = s->sah_depth;
    if (nodes64) PT_CUDA(cudaMemcpy(nodes64, s->d_nodes_sah, 64 * (size_t)s->n_nodes, cudaMemcpyDeviceToHost));
    return PT_OK;
}

int pt_bvh_export_ploc(const pt_scene* scene, void* nodes64, uint32_t* depth, float* build_ms)
{
    Scene* s = scene_from_handle((uint64_t)(uintptr_t)scene);
    if (!s) return fail(PT_ERR_INVALID, "not a live pt_scene");
    DeviceGuard g(s->device);
    PT_CUDA(ensure_ploc(*s));
    if (!s->d_nodes_ploc) return fail(PT_ERR_INVALID, "no PLOC tree for this scene");
    if (depth) *depth = s->ploc_depth;
    if (build_ms) *build_ms = s->ploc_ms;
    if (nodes64) PT_CUDA(cudaMemcpy(nodes64, s->d_nodes_ploc, 64 * (size_t)s->n_nodes, cudaMemcpyDeviceToHost));
    return PT_OK;
}

int pt_bvh_export_quant(pt_scene* scene, void* qnodes32, float* lo3, float* step3)
{
    Scene* s = scene_from_handle((uint64_t)(uintptr_t)scene);
    if (!s) return fail(PT_ERR_INVALID, "not a live pt_scene");
    if (s->n_tris < 2) return fail(PT_ERR_INVALID, "no quantised nodes for a scene of fewer than 2 triangles");
    DeviceGuard g(s->device);
    PT_CUDA(quantize_nodes(*s, 0));
    if (!s->d_qnodes || s->qnodes_tree != scene_tree(*s)) return fail(PT_ERR_CUDA, "quantisation could not contain a box");
    for (int a = 0; a < 3; ++a) {
        if (lo3) lo3[a] = s->q_lo[a];
        if (step3) step3[a] = s->q_step[a];
    }
    if (qnodes32) PT_CUDA(cudaMemcpy(qnodes32, s->d_qnodes, 32 * (size_t)s->n_nodes, cudaMemcpyDeviceToHost));
    return PT_OK;
}

int pt_bvh_export_wide(pt_scene* scene, void* wnodes64, uint32_t capacity, uint32_t* n_wide, uint32_t* depth, float* build_ms)
{
    Scene* s = scene_from_handle((uint64_t)(uintptr_t)scene);
    if (!s) return fail(PT_ERR_INVALID, "not a live pt_scene");
    if (s->n_tris < 2) return fail(PT_ERR_INVALID, "no wide nodes for a scene of fewer than 2 triangles");
    DeviceGuard g(s->device);
    PT_CUDA(build_wide_nodes(*s, 0));
    if (!s->d_wnodes || s->wnodes_tree != scene_tree(*s)) return fail(PT_ERR_CUDA, "the wide nodes could not be built");
    if (n_wide) *n_wide = s->n_wnodes;
    if (depth) *depth = s->wide_depth;
    if (build_ms) *build_ms = s->wide_ms;
    if (wnodes64) {
        if (capacity < s->n_wnodes) return fail(PT_ERR_INVALID, "pt_bvh_export_wide: capacity too small");
        PT_CUDA(cudaMemcpy(wnodes64, s->d_wnodes, 64 * (size_t)s->n_wnodes, cudaMemcpyDeviceToHost));
    }
    return PT_OK;
}

int pt_sah_build_host(const float* tri_boxes8, const uint32_t* sorted_prim, uint32_t n_tris, void* nodes64, uint32_t* depth)
{
    if (!tri_boxes8 || !sorted_prim || !nodes64 || n_tris < 2) return fail(PT_ERR_INVALID, "pt_sah_build_host: bad argument");
    std::vector<float4> nodes;
    uint32_t d = 0;
    sah_build_host(reinterpret_cast<const float4*>(tri_boxes8), sorted_prim, n_tris, nodes, d);
    memcpy(nodes64, nodes.data(), sizeof(float4) * nodes.size());
    if (depth) *depth = d;
    return PT_OK;
}

int pt_scene_last_traversal(const pt_scene* scene)
{
    Scene* s = scene_from_handle((uint64_t)(uintptr_t)scene);
    return s ? (int)s->last_trav : -1;
}

int pt_scene_last_launch_info(const pt_scene* scene, uint32_t* grid, uint32_t* block, uint32_t* smem_bytes, uint32_t* kernels_launched)
{
    Scene* s = scene_from_handle((uint64_t)(uintptr_t)scene);
    if (!s) return fail(PT_ERR_INVALID, "not a live pt_scene");
    if (grid) *grid = s->last_grid;
    if (block) *block = s->last_block;
    if (smem_bytes) *smem_bytes = s->last_smem;
    if (kernels_launched) *kernels_launched = s->last_kernels;
    return PT_OK;
}

// ---- parity probes ------------------------------------------------------------------------------------------------
int pt_bvh_export(const pt_scene* scene, uint32_t* morton, uint64_t* sorted_keys, uint32_t* sorted_prim, int32_t* parent,
                  void* nodes64, void* tris48, float* bounds6)
{
    Scene* s = scene_from_handle((uint64_t)(uintptr_t)scene);
    if (!s) return fail(PT_ERR_INVALID, "not a live pt_scene");
    DeviceGuard g(s->device);
    PT_CUDA(cudaDeviceSynchronize());
    const size_t n = s->n_tris;
    if (n == 0) return PT_OK;
    if (morton) PT_CUDA(cudaMemcpy(morton, s->d_morton, 4 * n, cudaMemcpyDeviceToHost));
    if (sorted_keys || sorted_prim) {
        std::vector<uint32_t> k(n), v(n);
        PT_CUDA(cudaMemcpy(k.data(), s->d_keys[s->sorted_buf], 4 * n, cudaMemcpyDeviceToHost));
        PT_CUDA(cudaMemcpy(v.data(), s->d_vals[s->sorted_buf], 4 * n, cudaMemcpyDeviceToHost));
        for (size_t i = 0; i < n; ++i) {
            if (sorted_keys) sorted_keys[i] = ((uint64_t)k[i] << 32) | v[i];
            if (sorted_prim) sorted_prim[i] = v[i];
        }
    }
    if (parent) {
        if (n > 1) PT_CUDA(cudaMemcpy(parent, s->d_node_parent, 4 * (n - 1), cudaMemcpyDeviceToHost));
        else parent[0] = -1;
    }
    if (nodes64) PT_CUDA(cudaMemcpy(nodes64, s->d_nodes, 64 * (size_t)s->n_nodes, cudaMemcpyDeviceToHost));
    if (tris48) PT_CUDA(cudaMemcpy(tris48, s->d_tris, 48 * n, cudaMemcpyDeviceToHost));
    if (bounds6) PT_CUDA(cudaMemcpy(bounds6, s->d_bounds, 24, cudaMemcpyDeviceToHost));
    return PT_OK;
}

int pt_trace_rays(const pt_scene* scene, const float* org3, const float* dir3, const float* tmin, const float* tmax,
                  uint64_t n, int32_t* prim_out, float* t_out, uint8_t* any_out, uint64_t* counters2)
{
    Scene* s = scene_from_handle((uint64_t)(uintptr_t)scene);
    if (!s) return fail(PT_ERR_INVALID, "not a live pt_scene");
    if (n == 0) return PT_OK;
    if (!org3 || !dir3 || !tmin || !tmax) return fail(PT_ERR_INVALID, "pt_trace_rays: NULL input");
    DeviceGuard g(s->device);
    float *d_org = nullptr, *d_dir = nullptr, *d_tmin = nullptr, *d_tmax = nullptr, *d_t = nullptr;
    int32_t* d_prim = nullptr;
    uint8_t* d_any = nullptr;
    unsigned long long* d_c = nullptr;
    cudaError_t e = cudaSuccess;
    auto A = [&](cudaError_t r) {
        if (e == cudaSuccess) e = r;
    };
    A(dev_alloc(&d_org, 3 * n));
    A(dev_alloc(&d_dir, 3 * n));
    A(dev_alloc(&d_tmin, n));
    A(dev_alloc(&d_tmax, n));
    if (prim_out || t_out) {
        A(dev_alloc(&d_prim, n));
        A(dev_alloc(&d_t, n));
    }
    if (any_out) A(dev_alloc(&d_any, n));
    A(dev_alloc(&d_c, 2));
    if (e == cudaSuccess) {
        A(cudaMemcpy(d_org, org3, 12 * n, cudaMemcpyHostToDevice));
        A(cudaMemcpy(d_dir, dir3, 12 * n, cudaMemcpyHostToDevice));
        A(cudaMemcpy(d_tmin, tmin, 4 * n, cudaMemcpyHostToDevice));
        A(cudaMemcpy(d_tmax, tmax, 4 * n, cudaMemcpyHostToDevice));
        A(cudaMemset(d_c, 0, 16));
        A(trace_rays_device(*s, d_org, d_dir, d_tmin, d_tmax, n, d_prim, d_t, d_any, d_c, 0));
        A(cudaDeviceSynchronize());
        if (prim_out && e == cudaSuccess) A(cudaMemcpy(prim_out, d_prim, 4 * n, cudaMemcpyDeviceToHost));
        if (t_out && e == cudaSuccess) A(cudaMemcpy(t_out, d_t, 4 * n, cudaMemcpyDeviceToHost));
        if (any_out && e == cudaSuccess) A(cudaMemcpy(any_out, d_any, n, cudaMemcpyDeviceToHost));
        if (counters2 && e == cudaSuccess) A(cudaMemcpy(counters2, d_c, 16, cudaMemcpyDeviceToHost));
    }
    void* ptrs[] = {d_org, d_dir, d_tmin, d_tmax, d_t, d_prim, d_any, d_c};
    for (void* p : ptrs)
        if (p) cudaFree(p);
    if (e != cudaSuccess) return cuda_fail(e, "pt_trace_rays");
    return PT_OK;
}

// One probed launch into scratch buffers; any output may be NULL.
static int debug_launch(const pt_params* params, uint32_t* seed_after_jitter, float* primary_dir3, int32_t* primary_prim,
                        float* primary_t, uint32_t* final_path_seed, uint32_t* final_pixel_seed)
{
    Scene* s = nullptr;
    pt_params p;
    if (!params) return fail(PT_ERR_INVALID, "params is NULL");
    p = *params;
    if (!p.accumulationBuffer) p.accumulationBuffer = (void*)1;   // replaced below; only validated for non-NULL
    int rc = validate_params(&p, &s);
    if (rc != PT_OK) return rc;
    DeviceGuard g(s->device);
    const size_t px = (size_t)p.width * p.height;
    float4* d_acc = nullptr;
    uint32_t *d_sj = nullptr, *d_fp = nullptr, *d_fx = nullptr;
    float *d_dir = nullptr, *d_t = nullptr;
    int32_t* d_prim = nullptr;
    cudaError_t e = cudaSuccess;
    auto A = [&](cudaError_t r) {
        if (e == cudaSuccess) e = r;
    };
    A(dev_alloc(&d_acc, px));
    if (seed_after_jitter) A(dev_alloc(&d_sj, px));
    if (primary_dir3) A(dev_alloc(&d_dir, 3 * px));
    if (primary_prim) A(dev_alloc(&d_prim, px));
    if (primary_t) A(dev_alloc(&d_t, px));
    if (final_path_seed) A(dev_alloc(&d_fp, px));
    if (final_pixel_seed) A(dev_alloc(&d_fx, px));
    if (e == cudaSuccess) {
        A(cudaMemset(d_acc, 0, px * sizeof(float4)));
        p.accumulationBuffer = d_acc;
        p.frameBuffer = nullptr;
        ptd::LaunchConst ov = {};
        ov.dbg_seed_after_jitter = d_sj;
        ov.dbg_primary_dir = d_dir;
        ov.dbg_primary_prim = d_prim;
        ov.dbg_primary_t = d_t;
        ov.dbg_final_path_seed = d_fp;
        ov.dbg_final_pixel_seed = d_fx;
        A(render_launch(*s, p, 0, 1, 0, &ov));
        A(cudaDeviceSynchronize());
        if (d_sj && e == cudaSuccess) A(cudaMemcpy(seed_after_jitter, d_sj, 4 * px, cudaMemcpyDeviceToHost));
        if (d_dir && e == cudaSuccess) A(cudaMemcpy(primary_dir3, d_dir, 12 * px, cudaMemcpyDeviceToHost));
        if (d_prim && e == cudaSuccess) A(cudaMemcpy(primary_prim, d_prim, 4 * px, cudaMemcpyDeviceToHost));
        if (d_t && e == cudaSuccess) A(cudaMemcpy(primary_t, d_t, 4 * px, cudaMemcpyDeviceToHost));
        if (d_fp && e == cudaSuccess) A(cudaMemcpy(final_path_seed, d_fp, 4 * px, cudaMemcpyDeviceToHost));
        if (d_fx && e == cudaSuccess) A(cudaMemcpy(final_pixel_seed, d_fx, 4 * px, cudaMemcpyDeviceToHost));
    }
    void* ptrs[] = {d_acc, d_sj, d_dir, d_prim, d_t, d_fp, d_fx};
    for (void* q : ptrs)
        if (q) cudaFree(q);
    if (e != cudaSuccess) return cuda_fail(e, "debug launch");
    return PT_OK;
}

int pt_debug_primary(const pt_params* params, uint32_t* seed_after_jitter, float* primary_dir3, int32_t* primary_prim,
                     float* primary_t)
{
    return debug_launch(params, seed_after_jitter, primary_dir3, primary_prim, primary_t, nullptr, nullptr);
}

int pt_debug_final_seeds(const pt_params* params, uint32_t* final_path_seed, uint32_t* final_pixel_seed)
{
    return debug_launch(params, nullptr, nullptr, nullptr, nullptr, final_path_seed, final_pixel_seed);
}

// ---- host helpers ---------------------------------------------------------------------------------------------------
// sutil::Camera::UVWFrame (sutil/Camera.cpp:33-45): W = lookat - eye (not normalised); U = normalize(W x up);
// V = normalize(U x W); vlen = |W| tan(fovY/2); V *= vlen; U *= vlen * aspect.
void pt_camera_uvw_frame(const float eye[3], const float lookat[3], const float up[3], float fov_y_deg, float aspect,
                         float U[3], float V[3], float W[3])
{
    auto cross3 = [](const float* a, const float* b, float* o) {
        o[0] = a[1] * b[2] - a[2] * b[1];
        o[1] = a[2] * b[0] - a[0] * b[2];
        o[2] = a[0] * b[1] - a[1] * b[0];
    };
    auto norm3 = [](float* v) {
        const float inv = 1.0f / sqrtf(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
        v[0] *= inv;
        v[1] *= inv;
        v[2] *= inv;
    };
    float w[3] = {lookat[0] - eye[0], lookat[1] - eye[1], lookat[2] - eye[2]};
    const float wlen = sqrtf(w[0] * w[0] + w[1] * w[1] + w[2] * w[2]);
    float u[3], v[3];
    cross3(w, up, u);
    norm3(u);
    cross3(u, w, v);
    norm3(v);
    const float vlen = wlen * tanf(0.5f * fov_y_deg * 3.14159265358979323846f / 180.0f);
    const float ulen = vlen * aspect;
    for (int a = 0; a < 3; ++a) {
        V[a] = v[a] * vlen;
        U[a] = u[a] * ulen;
        W[a] = w[a];
    }
}

int pt_device_info(int* sm_count, int* l2_bytes, int* clock_khz, char* name64)
{
    int dev = 0;
    PT_CUDA(cudaGetDevice(&dev));
    cudaDeviceProp prop;
    PT_CUDA(cudaGetDeviceProperties(&prop, dev));
    if (sm_count) *sm_count = prop.multiProcessorCount;
    if (l2_bytes) *l2_bytes = prop.l2CacheSize;
    if (clock_khz) {
        int khz = 0;
        cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, dev);
        *clock_khz = khz;
    }
    if (name64) {
        strncpy(name64, prop.name, 63);
        name64[63] = 0;
    }
    return PT_OK;
}

int pt_bench_fp32_peak(double* tflops)
{
    if (!tflops) return fail(PT_ERR_INVALID, "NULL output");
    PT_CUDA(bench_fp32_peak(tflops));
    return PT_OK;
}
int pt_bench_l2_bandwidth(double* gbps)
{
    if (!gbps) return fail(PT_ERR_INVALID, "NULL output");
    PT_CUDA(bench_l2_bandwidth(gbps));
    return PT_OK;
}
int pt_bench_hbm_bandwidth(double* gbps)
{
    if (!gbps) return fail(PT_ERR_INVALID, "NULL output");
    PT_CUDA(bench_hbm_bandwidth(gbps));
    return PT_OK;
}

}  // extern "C"
