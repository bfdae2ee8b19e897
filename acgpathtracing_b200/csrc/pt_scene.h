// pt_scene.h -- host-side scene object behind the opaque pt_scene handle (include/ptb200.h).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <string>
#include <vector>

#include "../../include/ptb200.h"

namespace ptd {
struct DevMaterial;
struct LaunchConst;
}

namespace ptb {

constexpr uint64_t kSceneMagic = 0x5054423230305343ull;   // "PTB200SC"

// Scheduling constants of render_wavefront_kernel; the defaults are the measured optimum on C2 (profiles/README.md).
// PTB_TUNE="key=value,..." overrides them per process for experiments (wavefront.cu wavefront_launch).
struct WfTune {
    int node_votes = 6;      // votes (of two slab-test steps each) per NODE batch
    int node_votes_g = 3;    // ... for scenes read from global memory: shorter batches regroup more often, which keeps more lanes
                             // (and more loads) in flight where a step is a cache miss (C4: 1396 -> 1519 with full_cols_g = 28)
    int node_keep8 = 2;      // a NODE batch continues while lanes_at_node * 8 >= lanes_at_start * node_keep8
    int full_cols = 30;      // a downstream stage is taken as soon as this many columns have a waiting path (32 before the
                             // scheduler affinity of round 2)
    int full_cols_g = 28;    // ... for scenes read from global memory
    int sticky = 4;          // repeats of the same stage without a full vote while it fills the whole warp
    int tail_cols = 24;      // tail mode when the non-empty columns of all queues add up to at most this
    int miss_loop = 16;      // samples of an all-miss pixel finished per END visit
    int paths32 = 0;         // paths per CTA / 32 (0 = as many as fit)
    int block = 0;           // threads per CTA (0 = derived from the paths)
    int smem_kb = 196;       // shared-memory budget for the paths when the scene is read from global memory: the largest
                             // carve-out step (..., 164, 196, 228 KB) that still leaves 60 KB of L1 (one step more costs C3 8 %)
    int shortlist = 1;       // per-pixel candidate lists for camera rays (tiny scenes)
    int idle_ns = 64;        // sleep of a warp that found every queue empty ...
    int idle_shift = 4;      // ... doubled for every further empty poll in a row, at most this many times (C3 +4 %: its idle polls
                             // were 21 % of the issued instructions; C2 and C4 unchanged)
    int stack_smem = 12;     // scenes in global memory: traversal-stack levels kept in shared memory, deeper ones spill to
                             // global memory (0 = all in shared memory).  C4: 672 -> 992 paths per CTA, +21 %
    int quant = -1;          // -1: follow Scene::traversal; 0 / 1: exact 64-byte / quantised 32-byte nodes for scenes in global memory
    int team = 1;            // sample teams (idle path slots take samples of a pixel another slot of their column is working on):
                             // 0 off, 1 auto (scenes read from global memory whose pixels differ in cost or that are small shares, wavefront_launch), 2 on, 3 on + state dump when a team is left open
    int team_min = 2;        // ... when that pixel has at least this many samples left to hand out
    int affinity = 2;        // schedulers (warp % 4 < affinity) whose warps look at the traversal queues first, the others at the
                             // shading queues (instruction-cache locality per scheduler; C2 +5.7 % together with full_cols = 30)
    int affinity_g = 3;      // ... for scenes read from global memory (C3 +7 %, C4 +2 %)
    int prefetch = 0;        // scenes in global memory: prefetch deferred children towards L2 (measured: C4 -6 %, C3 -1 %)
    int msort = 0;           // material-sorted SHADE: one queue per BSDF type (scenes with more than one type); measured, see DESIGN.md
};

struct TileOrder {           // one cached work order (render.cu tile_order_update)
    uint32_t* d = nullptr;
    uint32_t* h = nullptr;       // pinned staging copy
    cudaEvent_t copied = nullptr;
    size_t capacity = 0;
    uint32_t count = 0;
    float key[17] = {-1.f};
    uint64_t stamp = 0;
};

struct Scene {
    uint64_t magic = kSceneMagic;
    int device = 0;
    uint32_t n_verts = 0, n_tris = 0, n_mats = 0, n_nodes = 0;
    uint32_t n_bsdf_types = 0;     // distinct BSDF types in the material table
    uint32_t bvh_depth = 0;        // internal nodes on the longest root-to-leaf chain
    float background[3] = {0.f, 0.f, 0.f};
    float build_ms = 0.f;
    int pipeline = 0;              // 0 auto (wavefront when it can hold the scene), 1 persistent lane-per-path, 2 CTA-local wavefront
    WfTune wf;                     // scheduling constants of the wavefront kernel (measured defaults)
    bool count_enabled = false;
    bool count_executed = false;   // counted launches run the kernel configuration of uncounted ones (candidate lists, quantised nodes)
    int sched[5] = {1, 1, 1, 1, 20};   // render-kernel stage weights (NODE, LEAF, SHADE, END) and NODE fast-path lane count

    // inputs (device copies of the TinyObjWrapper arrays)
    float4* d_verts = nullptr;
    uint32_t* d_indices = nullptr;
    uint32_t* d_mat_ids = nullptr;
    ptd::DevMaterial* d_mats = nullptr;

    // LBVH build products
    float4* d_tri_in = nullptr;       // 3 per primitive (primitive order)
    float4* d_tri_bb = nullptr;       // 2 per primitive
    float4* d_tri_shade = nullptr;    // (N_0, material) per primitive
    uint32_t* d_bounds_ordered = nullptr;   // 6
    float* d_bounds = nullptr;        // 6
    uint32_t* d_morton = nullptr;     // primitive order
    uint32_t* d_keys[2] = {nullptr, nullptr};
    uint32_t* d_vals[2] = {nullptr, nullptr};
    int sorted_buf = 0;
    uint32_t* d_block_hist = nullptr;
    uint32_t* d_scan_scratch = nullptr;   // tile totals of the multi-block scan (lbvh.cu scan_exclusive)
    int* d_left = nullptr;
    int* d_right = nullptr;
    int* d_node_parent = nullptr;
    int* d_leaf_parent = nullptr;
    uint32_t* d_visit = nullptr;
    float4* d_node_lo = nullptr;
    float4* d_node_hi = nullptr;
    float4* d_nodes = nullptr;        // 4 per node
    float4* d_tris = nullptr;         // 3 per leaf (leaf order)
    // sweep-SAH tree over the same leaves (sah.cu), built on the host for scenes of 2 .. kSahMaxTris triangles
    float4* d_nodes_sah = nullptr;    // 4 per node, same record as d_nodes; null when not built
    uint32_t sah_depth = 0;
    // PLOC tree over the same leaves (lbvh.cu ploc_build), built on the GPU for every scene of >= 2 triangles
    float4* d_nodes_ploc = nullptr;   // 4 per node, same record; null when not built
    uint32_t ploc_depth = 0;
    float ploc_ms = 0.f;
    bool ploc_tried = false;          // the PLOC build ran (it is lazy for scenes that have a SAH tree)
    float sah_ms = 0.f;               // host time of the sweep-SAH build
    int tree = 0;                     // pt_scene_set_tree: 0 auto, 1 LBVH, 2 SAH, 3 PLOC
    // quantised 32-byte nodes of the tree in use (lbvh.cu quantize_nodes), built at the first launch that wants them
    float4* d_qnodes = nullptr;
    int qnodes_tree = 0;              // which tree (1 LBVH, 2 SAH, 3 PLOC) d_qnodes was made from; 0 = none
    float q_lo[3] = {0.f, 0.f, 0.f}, q_step[3] = {1.f, 1.f, 1.f};
    // 4-wide quantised 64-byte nodes of the tree in use (lbvh.cu build_wide_nodes), built at the first launch that wants them
    float4* d_wnodes = nullptr;
    int wnodes_tree = 0;              // which tree d_wnodes was made from; 0 = none
    uint32_t n_wnodes = 0, wide_depth = 0;
    float wide_ms = 0.f;              // host time of the collapse
    int traversal = 0;                // pt_scene_set_traversal: 0 auto, 1 exact nodes (staged when the scene fits), 2 exact nodes from
                                      // global memory whatever the size, 3 quantised nodes (global memory), 4 4-wide quantised nodes

    // launch-time scratch
    uint32_t* d_work_counter = nullptr;         // [0] pixel fetch counter  [1] LBVH depth (build)
    unsigned long long* d_counters = nullptr;   // pt_counters layout (10 x u64)
    float4* d_stage_accum = nullptr;            // pt_launch_host staging
    uchar4* d_stage_frame = nullptr;
    size_t stage_pixels = 0;

    // work order of the render kernels (render.cu tile_order_update): this rank's tiles, those a camera ray can hit
    // the scene through first.  Rebuilt when the key (image size, camera, partition) changes.
    float h_bounds[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};   // scene AABB (lo xyz, hi xyz)
    TileOrder tile_orders[4];
    uint64_t tile_order_clock = 0;
    bool tile_order_enabled = true;

    int* d_stack_over = nullptr;                // deeper stack levels of the wavefront kernel (scenes in global memory)
    size_t stack_over_capacity = 0;
    void* d_team_log = nullptr;                 // term logs + marks of the wavefront kernel's sample teams
    size_t team_log_bytes = 0;

    // last launch
    uint32_t last_grid = 0, last_block = 0, last_smem = 0, last_kernels = 0, last_trav = 0;   // last_trav: 0 exact 64-byte nodes, 1 quantised 32-byte nodes, 2 4-wide quantised nodes
};

constexpr uint32_t kSahMaxTris = 8192;    // the host build is O(n log^2 n); larger scenes keep the GPU-built LBVH
constexpr uint32_t kSahMaxDepth = 64;

// which tree the render / trace kernels walk: 1 LBVH, 2 SAH, 3 PLOC.  auto: the host-built SAH tree where there is one
// (2..kSahMaxTris triangles), else the GPU-built PLOC tree, else the LBVH; a tree deeper than kSahMaxDepth is passed over.
inline int scene_tree(const Scene& s)
{
    // (pt_scene_set_tree refuses a tree deeper than kSahMaxDepth: the traversal stacks of the fallback and trace
    // kernels hold 64 entries, which bounds the LBVH -- 62-bit keys -- but not a SAH or PLOC tree of degenerate input)
    if (s.tree == 2 && s.d_nodes_sah && s.sah_depth <= kSahMaxDepth) return 2;
    if (s.tree == 3 && s.d_nodes_ploc && s.ploc_depth <= kSahMaxDepth) return 3;
    if (s.tree != 0) return 1;
    if (s.d_nodes_sah && s.sah_depth <= kSahMaxDepth) return 2;
    if (s.d_nodes_ploc && s.ploc_depth <= kSahMaxDepth) return 3;
    return 1;
}
inline const float4* scene_nodes(const Scene& s)
{
    const int t = scene_tree(s);
    return t == 2 ? s.d_nodes_sah : (t == 3 ? s.d_nodes_ploc : s.d_nodes);
}
inline uint32_t scene_depth(const Scene& s)
{
    const int t = scene_tree(s);
    return t == 2 ? s.sah_depth : (t == 3 ? s.ploc_depth : s.bvh_depth);
}

// sah.cu
void sah_build_host(const float4* tri_bb, const uint32_t* sorted_prim, uint32_t n, std::vector<float4>& nodes, uint32_t& depth);

// lbvh.cu
size_t lbvh_scan_scratch_entries(size_t n_tris);
cudaError_t lbvh_build(Scene& s, cudaStream_t stream);
cudaError_t ploc_build(Scene& s, cudaStream_t stream);
cudaError_t quantize_nodes(Scene& s, cudaStream_t stream);
cudaError_t build_wide_nodes(Scene& s, cudaStream_t stream);

// render.cu
constexpr uint32_t kMaxFoldFrames = 64;
cudaError_t render_launch(Scene& s, const pt_params& params, uint32_t rank, uint32_t n_ranks, cudaStream_t stream,
                          const ptd::LaunchConst* probe_overrides, bool raw_mean = false, uint32_t span = 1);
cudaError_t fold_frames(void* accum, void* frame_out, const void* const* means, uint32_t first_frame, uint32_t n_frames,
                        uint32_t width, uint32_t height, cudaStream_t stream);
cudaError_t trace_rays_device(Scene& s, const float* d_org, const float* d_dir, const float* d_tmin, const float* d_tmax,
                              uint64_t n, int32_t* d_prim, float* d_t, uint8_t* d_any, unsigned long long* d_counters2,
                              cudaStream_t stream);
cudaError_t partition_pack(const void* accum_full, const void* frame_full, void* accum_compact, void* frame_compact,
                           uint32_t width, uint32_t height, uint32_t rank, uint32_t n_ranks, cudaStream_t stream);
cudaError_t partition_unpack(const void* accum_compact_all, const void* frame_compact_all, void* accum_full,
                             void* frame_full, uint32_t width, uint32_t height, uint32_t n_ranks, cudaStream_t stream);
uint32_t partition_owned_tiles(uint32_t width, uint32_t height, uint32_t rank, uint32_t n_ranks, uint32_t span);
uint32_t partition_max_owned_tiles(uint32_t width, uint32_t height, uint32_t n_ranks);

// wavefront.cu
bool wavefront_supported(const Scene& s);
bool wf_tune_apply(WfTune& t, const char* spec);   // "key=value,..."; false when a key is unknown
cudaError_t wavefront_launch(Scene& s, const ptd::LaunchConst& lc, bool probe, cudaStream_t stream);

// microbench.cu
cudaError_t bench_fp32_peak(double* tflops);
cudaError_t bench_l2_bandwidth(double* gbps);
cudaError_t bench_hbm_bandwidth(double* gbps);

// error plumbing and the registry of live scenes (capi.cu)
void set_error(const std::string& msg);
int fail(int code, const std::string& msg);             // records the message, returns the code
int cuda_fail(cudaError_t e, const char* what);         // PT_ERR_CUDA with the CUDA error's name and text
Scene* scene_from_handle(uint64_t handle);              // null unless `handle` is a live pt_scene
cudaError_t ensure_ploc(Scene& s);                      // builds the PLOC tree if the scene does not have it yet (current device)

}  // namespace ptb
