/* ptb200.h -- C ABI of the B200-native path tracer (libptb200.so).
 *
 * Drop-in boundary for the hot path of fallinbryan/ACGPathTracing: everything between "arrays from
 * TinyObjWrapper" and "pixels in two buffers".  The reference has no FFI layer; the calls below replace
 * what PathTracer_Optix/PathTracerMain.cpp does through the OptiX host API (file:line cited per entry).
 * Plain C types only: pointers, sizes, opaque handles.  Every function returns 0 on success or a
 * PT_ERR_* code; pt_last_error() gives the message (thread-local).  Nothing throws across the boundary
 * (the reference throws sutil::Exception, sutil/Exception.h:136, caught once in main :742).
 *
 * There is no CPU fallback: calls that need the GPU fail with PT_ERR_CUDA when no sm_100 device is usable.
 */
#ifndef PTB200_H
#define PTB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PT_OK 0
#define PT_ERR_INVALID 1   /* bad argument (null pointer, zero size, index out of range, bad BSDF type) */
#define PT_ERR_CUDA 2      /* CUDA runtime error / no device */
#define PT_ERR_IO 3        /* OBJ/MTL file could not be read or parsed */

/* ---- launch contract: byte-identical to the reference ------------------------------------------------ */
typedef struct { float x, y, z; } pt_float3;

/* pathTracer.h:77-83 (60 bytes) */
typedef struct {
    pt_float3 corner, v1, v2, normal, emission;
} pt_area_light;

/* pathTracer.h:85-108 (168 bytes, align 8).  `handle` carries pt_scene_handle() instead of an
 * OptixTraversableHandle.  accumulationBuffer (float4*) and frameBuffer (uchar4*) are DEVICE pointers owned
 * by the caller for pt_launch* (ref: cudaMalloc PathTracerMain.cpp:145 / mapped PBO :186) and HOST pointers
 * for pt_launch_host.  frameBuffer may be NULL (headless extension). */
typedef struct {
    uint32_t currentFrameIdx;
    void* accumulationBuffer;
    void* frameBuffer;
    uint32_t width, height, samplesPerPixel, maxDepth;
    pt_float3 cameraEye, cameraU, cameraV, cameraW;
    pt_area_light areaLight;
    uint64_t handle;
    uint8_t useDirectLighting;      /* bool */
    uint8_t useImportanceSampling;  /* bool */
} pt_params;

/* TinyObjWrapper.h:27-40 (40 bytes) */
enum { PT_BSDF_DIFFUSE = 0, PT_BSDF_METALLIC = 1, PT_BSDF_REFRACTION = 2 };
typedef struct {
    pt_float3 diffuse, emission;
    float roughness, metallic, ior;
    int32_t bsdfType;
} pt_material;

/* Work counters: per launch, cumulative until reset. */
typedef struct {
    uint64_t n_samples, n_radiance, n_shadow, n_node, n_tri, n_hit[3], n_miss;
} pt_counters;

typedef struct pt_scene pt_scene;

const char* pt_last_error(void);

/* ---- scene: replaces buildTheAccelarationStructure (PathTracerMain.cpp:260-398: upload + optixAccelBuild
 * + compaction) and createShaderBindingTable (:544-608: one hit record per material).  Copies the four
 * TinyObjWrapper arrays (TinyObjWrapper.h:87-93) to the current CUDA device and builds the LBVH there.
 * The caller keeps ownership of the host arrays. ------------------------------------------------------- */
int pt_scene_create(const float* verts_xyzw, uint32_t n_verts, const uint32_t* indices, uint32_t n_tris,
                    const uint32_t* mat_ids, const pt_material* materials, uint32_t n_materials, pt_scene** out);
void pt_scene_destroy(pt_scene* scene);                                   /* CleanAllTheThings :629-646 */
int pt_scene_set_background(pt_scene* scene, float r, float g, float b);  /* MissData.backgroundColor :568 */
uint64_t pt_scene_handle(const pt_scene* scene);                          /* value for pt_params.handle */
int pt_scene_info(const pt_scene* scene, uint32_t* n_tris, uint32_t* n_nodes, float* build_ms);

/* ---- launch: replaces cudaMemcpyAsync(d_params) + optixLaunch(pipeline, stream, d_params, ..., w, h, 1)
 * (PathTracerMain.cpp:191-206).  One call = one subframe of samplesPerPixel samples per pixel folded into
 * the running average with weight 1/(currentFrameIdx+1) (pathTracerPrograms.cu:803-811); the caller
 * increments currentFrameIdx (:714).  Asynchronous on `cuda_stream` (a cudaStream_t; NULL = default).
 * A pt_scene holds the launch scratch (work counter, work order, spill stacks, watchdog flag): at most ONE launch per
 * scene may be in flight, i.e. all launches of a scene go to the same stream (or are separated by a synchronisation),
 * as in the reference (one stream, :161).  Different scenes and devices may be driven concurrently. ---- */
int pt_launch(const pt_params* params, void* cuda_stream);

/* Pixel-partitioned launch for multi-GPU rendering: renders only the 8x4-pixel tiles owned by `rank` of
 * `n_ranks` (tile L is owned by rank L % n_ranks, rows rotated) into the full-size buffers.
 * The RNG stream of a pixel depends only on (pixel index, currentFrameIdx) (pathTracerPrograms.cu:721), so
 * the union over ranks is bit-identical to pt_launch. */
int pt_launch_partition(const pt_params* params, uint32_t rank, uint32_t n_ranks, void* cuda_stream);

/* Frame-parallel rendering (multi-GPU, several launches per image).  pt_launch_raw renders the tiles of `rank` like
 * pt_launch_partition but writes the launch's OWN mean (result / samplesPerPixel, alpha 1) into accumulationBuffer:
 * nothing is read, nothing is folded, frameBuffer is ignored.  The mean of frame k depends only on (pixel, k)
 * (pathTracerPrograms.cu:721), so different ranks can render different frames.  pt_fold_frames then applies the
 * reference's running average (:803-811: acc += (mean_k - acc) / (k + 1)) to n_frames (<= 64) device images in frame
 * order k = first_frame_idx .., reading `accum` first when first_frame_idx > 0, and writes accum + the sRGB frame
 * buffer (may be NULL): bit-identical to the same frames rendered one after the other with pt_launch. */
int pt_launch_raw(const pt_params* params, uint32_t rank, uint32_t n_ranks, void* cuda_stream);
int pt_fold_frames(void* accum, void* frame_out, const void* const* frame_means, uint32_t first_frame_idx,
                   uint32_t n_frames, uint32_t width, uint32_t height, void* cuda_stream);

/* Owned-pixel exchange helpers (one NCCL all-gather per image sits between them).
 * pack: full-size accumulation/frame buffers -> this rank's compact tile-major buffers
 *       (pt_partition_capacity() pixels: float4 and uchar4 each).
 * unpack: n_ranks compact buffers laid out back to back -> full-size buffers. */
uint64_t pt_partition_capacity(uint32_t width, uint32_t height, uint32_t n_ranks);   /* pixels per rank */
int pt_partition_pack(const void* accum_full, const void* frame_full, void* accum_compact, void* frame_compact,
                      uint32_t width, uint32_t height, uint32_t rank, uint32_t n_ranks, void* cuda_stream);
int pt_partition_unpack(const void* accum_compact_all, const void* frame_compact_all, void* accum_full,
                        void* frame_full, uint32_t width, uint32_t height, uint32_t n_ranks, void* cuda_stream);

/* ---- multi-GPU: one image on several B200s through this ABI (SURVEY.md 8e, Appendix D pt_launch_multi).  The reference
 * is single-GPU: LaunchCurrentFrame (PathTracerMain.cpp:184-210) calls optixLaunch once per subframe on one device; a
 * host that wants N GPUs calls pt_launch_multi where it called pt_launch.  Scene and acceleration structure are replicated
 * (the builds are deterministic), pixel-samples are partitioned, and the result is bit-identical to the same launches on
 * one GPU because a pixel's RNG stream depends only on (pixel index, currentFrameIdx) (pathTracerPrograms.cu:721).
 *
 * Work split for n_launches subframes ("frames") on n_ranks GPUs (pt_multi_plan is that function, host only):
 *   - n_launches / n_ranks rounds of WHOLE frames, frame k -> rank k % n_ranks, each rendered as its own mean
 *     (pt_launch_raw), so no GPU is ever left with a small share of a frame;
 *   - the n_launches % n_ranks = L leftover frames are cut into L * n_ranks sub-shares (frame, tile class) and rank r
 *     takes sub-shares [r L, (r + 1) L): consecutive tile classes of one frame (sometimes the end of one frame and the
 *     start of the next), e.g. 4 frames on 8 GPUs = half a frame each, 1 frame on 8 GPUs = the classic 1/8 tile share.
 * Exchange: the image is cut into n_ranks contiguous pixel slices (pt_multi_slice); every rendered frame (or share)
 * sends slice j to rank j, rank j applies the reference's running average (:803-811) to its slice in frame order, and
 * the folded slices are gathered: K/N + 1 images received per rank instead of K.
 * Transport: NCCL send/recv between processes (one process per GPU: pt_group_create_rank; ncclCommInitRank) or inside
 * one process (pt_group_create_local with transport 1; ncclCommInitAll); peer-to-peer copy engines inside one process
 * by default (transport 0: cudaMemcpyPeerAsync over NVLink, no SM taken from the render kernel). ---------------------- */
typedef struct pt_group pt_group;
typedef struct {
    uint32_t frame;        /* 0 .. n_launches-1, added to params.currentFrameIdx */
    uint32_t rank;         /* who renders it */
    uint32_t first_class;  /* tile classes first_class .. first_class + n_classes - 1 of n_ranks (tile L has class L % n_ranks) */
    uint32_t n_classes;    /* n_ranks = the whole frame */
} pt_multi_job;
/* jobs in exchange order; *n_jobs receives the count (<= n_launches + 2 * n_ranks); capacity 0 + jobs NULL = count only */
int pt_multi_plan(uint32_t n_launches, uint32_t n_ranks, pt_multi_job* jobs, uint32_t capacity, uint32_t* n_jobs);
/* pixel slice [first, first + count) of the row-major image that `rank` folds */
int pt_multi_slice(uint32_t width, uint32_t height, uint32_t rank, uint32_t n_ranks, uint64_t* first_pixel, uint64_t* n_pixels);

#define PT_GROUP_ID_BYTES 128
#define PT_TRANSPORT_PEER 0   /* cudaMemcpyPeerAsync between the devices of one process */
#define PT_TRANSPORT_NCCL 1
/* One process drives n_devices GPUs (the reference's single-threaded host).  The same device may be listed more than once
 * with PT_TRANSPORT_PEER (tests on one GPU).  Ranks are 0 .. n_devices-1 in list order; rank 0's device holds the caller's
 * buffers. */
int pt_group_create_local(const int* devices, uint32_t n_devices, int transport, pt_group** out);
/* One process per GPU (torchrun, MPI): rank 0 calls pt_group_unique_id and ships the PT_GROUP_ID_BYTES bytes to the
 * other ranks by its own means; then EVERY rank calls pt_group_create_rank with its current CUDA device (collective). */
int pt_group_unique_id(void* id_bytes);
int pt_group_create_rank(const void* id_bytes, uint32_t rank, uint32_t n_ranks, pt_group** out);
void pt_group_destroy(pt_group* group);                    /* also destroys the group's scene replicas */
int pt_group_info(const pt_group* group, uint32_t* n_ranks, uint32_t* n_local, uint32_t* first_rank, int* transport);
/* pt_scene_create on every local device of the group (replaces the group's previous scene).  *out = the replica on the
 * group's first local device: its handle goes into pt_params.handle, and its settings (background, pipeline, tree,
 * traversal, tune, counters on/off) are applied to every local replica at each pt_launch_multi; pt_group_scene(group, i) =
 * the replica of local member i (to read its counters).  In a process-per-GPU group every rank sets its own replica. */
int pt_group_scene_create(pt_group* group, const float* verts_xyzw, uint32_t n_verts, const uint32_t* indices, uint32_t n_tris,
                          const uint32_t* mat_ids, const pt_material* materials, uint32_t n_materials, pt_scene** out);
pt_scene* pt_group_scene(const pt_group* group, uint32_t local_index);
/* Renders subframes currentFrameIdx .. currentFrameIdx + n_launches - 1 and folds them into accumulationBuffer /
 * frameBuffer exactly as n_launches calls of pt_launch on one GPU would (replaces the do-while of main :700-730 for a
 * multi-GPU host).  params: DEVICE buffers on this rank's device (process-per-GPU: every rank passes its own and every
 * rank ends up with the complete image; local group: buffers on rank 0's device), handle = pt_group_scene_create's.
 * Collective over the group; stream-ordered on `cuda_stream` (a stream of the first local device) like pt_launch. */
int pt_launch_multi(pt_group* group, const pt_params* params, uint32_t n_launches, void* cuda_stream);
/* The same with HOST buffers (the end-to-end entry point): uploads this process's slices of the previous accumulation
 * when currentFrameIdx > 0, renders, downloads the image (process-per-GPU: the complete image on every rank; local
 * group: each device writes its own slice into the one host image); synchronous.  Reports the bytes this process moved. */
int pt_launch_multi_host(pt_group* group, const pt_params* params_host_buffers, uint32_t n_launches, uint64_t* h2d_bytes,
                         uint64_t* d2h_bytes);
/* Device time of the phases of the last pt_launch_multi* call on local member `local_index` (CUDA events; synchronises):
 * ms[0] render, ms[1] slice exchange + fold, ms[2] final gather, ms[3] whole call; n_kernels = kernels of this library
 * it launched there. */
int pt_group_last_timing(pt_group* group, uint32_t local_index, float ms[4], uint32_t* n_kernels);

/* Same launch with HOST buffers: uploads the previous accumulation when currentFrameIdx > 0, renders,
 * downloads accumulation (and frame buffer if non-NULL); synchronous.  This is the call the end-to-end
 * benchmark times.  Reports the bytes moved. */
int pt_launch_host(const pt_params* params_host_buffers, uint64_t* h2d_bytes, uint64_t* d2h_bytes);

/* ---- counters and tuning ----------------------------------------------------------------------------- */
/* Counted launches use an instrumented kernel.  enable = 1: the SPECIFIED work -- every ray walks the exact nodes of the tree
 * in use (the counts equal the CPU oracle's); enable = 2: the EXECUTED work of an uncounted launch -- camera rays of tiny
 * scenes resolved from per-pixel candidate lists (their triangle tests counted, no node visits), quantised nodes where
 * the launch would use them; 0: off.  The pixels are the same in all three. */
int pt_scene_enable_counters(pt_scene* scene, int enable);
int pt_scene_get_counters(pt_scene* scene, pt_counters* out, int reset);
/* Scheduler statistics of counted launches (wavefront pipeline): out16[0..3] batches run per stage (NODE, LEAF,
 * SHADE, END), [4..7] lanes that held a path in those batches, [8] NODE steps, [9] lanes active in NODE steps,
 * [10] votes that found no work; cumulative until reset. */
int pt_scene_get_sched_stats(pt_scene* scene, uint64_t* out16, int reset);
/* kernel variant: 0 = auto, 1 = persistent thread-per-path, 2 = CTA-local wavefront */
int pt_scene_set_pipeline(pt_scene* scene, int pipeline);
/* Which binary tree the render and trace kernels walk.  Three trees over the same leaf records, all made of the same
 * 64-byte two-child nodes, so the pixels, primary hits and any-hit answers do not depend on the choice (the work
 * counters do):
 *   1 = the Morton-order LBVH, built on the GPU (always built; pt_bvh_export);
 *   2 = a sweep-SAH tree, built on the host at pt_scene_create for scenes of 2..8192 triangles (18-27 % fewer node
 *       visits than the LBVH);
 *   3 = a PLOC tree (parallel locally-ordered clustering, radius 8), built on the GPU: at pt_scene_create where it is
 *       the tree the kernels will walk (no usable SAH tree), else the first time it is asked for (22-27 % fewer node
 *       visits than the LBVH on the 2 M-triangle scene);
 *   0 = auto: 2 where it exists, else 3, else 1 (a tree more than 64 levels deep is passed over).
 * Asking for a tree the scene does not have, or one deeper than 64 levels, is PT_ERR_INVALID.  Replaces the build
 * flags of optixAccelBuild (PathTracerMain.cpp:334-339). */
int pt_scene_set_tree(pt_scene* scene, int tree);
int pt_scene_tree_in_use(const pt_scene* scene);   /* 1 LBVH, 2 SAH, 3 PLOC */
/* How the render kernel reads the acceleration structure:
 *   0 = auto: 1, except 4 for scenes of more than 8192 triangles (the fastest measured on B200: +5 % at 29 K triangles, +12 % at 2 M);
 *   1 = the exact 64-byte nodes of the tree in use, every lane fetches its own (from shared memory when the scene fits);
 *   2 = the same, but read from global memory whatever the scene's size (tests, A/B runs);
 *   3 = quantised nodes: 32 bytes per node (twelve 16-bit plane positions on the grid of the scene's bounding box,
 *       rounded outwards and checked, + two child references), one 256-bit load per node; scene read from global memory;
 *   4 = 4-wide quantised nodes: 64 bytes per node (four child boxes on the same grid + four references), every node stands
 *       for up to two levels of the binary tree (the internal child with the largest surface area is opened, twice);
 *       collapsed on the host at the first launch that asks for it; scene read from global memory.
 * The pixels do not depend on it: every variant culls with boxes that contain the tree's own and every hit comes from
 * the same triangle test (OptiX's acceleration structure, PathTracerMain.cpp:311-398, is closed). */
int pt_scene_set_traversal(pt_scene* scene, int traversal);
int pt_scene_last_traversal(const pt_scene* scene);   /* last launch: 0 exact nodes, 1 quantised nodes, 2 4-wide quantised nodes */
/* Scheduling constants of the render kernel for ONE scene: "key=value,key=value" (keys: csrc/pt_scene.h WfTune).  They
 * change how paths are batched, never a pixel (tests/test_gpu_parity.py::test_schedule_perturbation).  The environment
 * variable PTB_TUNE, read once per process, overrides them for every scene.  PT_ERR_INVALID for an unknown key. */
int pt_scene_set_tune(pt_scene* scene, const char* spec);
/* pt_launch* are asynchronous: a kernel watchdog that fired (the scheduler made no progress -- never expected) is only
 * seen here or by pt_launch_host / pt_scene_get_counters.  Synchronises the scene's device; PT_ERR_CUDA if it fired. */
int pt_scene_check_error(pt_scene* scene);
int pt_scene_last_launch_info(const pt_scene* scene, uint32_t* grid, uint32_t* block, uint32_t* smem_bytes,
                              uint32_t* kernels_launched);

/* ---- parity probes (tests only; not on the hot path) --------------------------------------------------- */
/* LBVH export; any pointer may be NULL.  morton[n_tris] in primitive order; sorted_keys[n_tris] =
 * morton<<32|prim ascending; sorted_prim[n_tris]; parent[n_nodes]; nodes: n_nodes x 64 B; tris: n_tris x 48 B
 * (leaf order); bounds: lo xyz, hi xyz. */
int pt_bvh_export(const pt_scene* scene, uint32_t* morton, uint64_t* sorted_keys, uint32_t* sorted_prim,
                  int32_t* parent, void* nodes64, void* tris48, float* bounds6);
/* SAH tree export (tests): nodes: n_nodes x 64 B (same record as pt_bvh_export's); depth: internal nodes on the
 * longest root-to-leaf chain. */
int pt_bvh_export_sah(const pt_scene* scene, void* nodes64, uint32_t* depth);
int pt_bvh_export_ploc(const pt_scene* scene, void* nodes64, uint32_t* depth, float* build_ms);
/* Quantised nodes of the tree in use (built on demand): n_nodes x 32 B = six words of two 16-bit plane positions each
 * (lo0.x|lo0.y<<16, lo0.z|hi0.x<<16, hi0.y|hi0.z<<16, then the same for child 1) + the two child references; a plane is
 * fma(q, step[axis], lo[axis]). */
int pt_bvh_export_quant(pt_scene* scene, void* qnodes32, float* lo3, float* step3);
/* 4-wide quantised nodes of the tree in use (built on first use): n_wide records of 16 words -- words 3k..3k+2 the box of child
 * k (lo | hi << 16 per axis, planes of the binary quantised nodes), words 12..15 the child references (>= 0 wide node, < 0 ~leaf
 * slot); numbered breadth first.  wnodes64 may be NULL (sizes only); capacity in records. */
int pt_bvh_export_wide(pt_scene* scene, void* wnodes64, uint32_t capacity, uint32_t* n_wide, uint32_t* depth, float* build_ms);
/* The host builder on its own (no device needed): tri_boxes8 = per PRIMITIVE (lo.xyz, -, hi.xyz, -), sorted_prim =
 * leaf slot -> primitive (pt_bvh_export), n_tris >= 2; writes (n_tris - 1) x 64 B. */
int pt_sah_build_host(const float* tri_boxes8, const uint32_t* sorted_prim, uint32_t n_tris, void* nodes64, uint32_t* depth);
/* Batch ray queries through the device traversal kernels (host arrays in/out). */
int pt_trace_rays(const pt_scene* scene, const float* org3, const float* dir3, const float* tmin, const float* tmax,
                  uint64_t n, int32_t* prim_out, float* t_out, uint8_t* any_out, uint64_t* counters2);
/* Per-pixel state of the FIRST sample of a launch (host arrays, w*h each; any may be NULL): pixel seed after
 * the two jitter draws, primary ray direction (3 floats), primary hit primitive (-1 = miss) and distance. */
int pt_debug_primary(const pt_params* params, uint32_t* seed_after_jitter, float* primary_dir3,
                     int32_t* primary_prim, float* primary_t);
/* Final RNG states of a launch (host arrays, w*h each): path seed when the last sample ended, pixel seed after
 * all jitter draws.  Re-runs the launch into scratch buffers. */
int pt_debug_final_seeds(const pt_params* params, uint32_t* final_path_seed, uint32_t* final_pixel_seed);

/* ---- host helpers mirroring the reference's host code ------------------------------------------------ */
/* sutil::Camera::UVWFrame (sutil/Camera.cpp:33-45) */
void pt_camera_uvw_frame(const float eye[3], const float lookat[3], const float up[3], float fov_y_deg,
                         float aspect, float U[3], float V[3], float W[3]);
/* TinyObjWrapper-equivalent OBJ/MTL ingest (TinyObjWrapper.cpp:38-244).  Two-call protocol: counts first
 * (array pointers NULL), then fill. */
typedef struct pt_obj pt_obj;
int pt_obj_load(const char* path, pt_obj** out);
void pt_obj_free(pt_obj* obj);
int pt_obj_counts(const pt_obj* obj, uint32_t* n_vert_floats, uint32_t* n_indices, uint32_t* n_mat_ids,
                  uint32_t* n_materials);
int pt_obj_copy(const pt_obj* obj, float* verts_xyzw, uint32_t* indices, uint32_t* mat_ids, pt_material* materials);
int pt_scene_create_from_obj(const pt_obj* obj, pt_scene** out);

/* ---- image output and checkpoint (host only) -------------------------------------------------------------- */
/* sutil::saveImage for a UNSIGNED_BYTE4 buffer and a ".ppm" file name (sutil/sutil.cpp:543-570): binary P6 with
 * the vertical flip (row 0 of frameBuffer is the bottom of the picture), alpha dropped.  Host pointer. */
int pt_image_write_ppm(const char* path, const void* frame_rgba8_host, uint32_t width, uint32_t height);
/* The ".png" branch of sutil::saveImage (sutil/sutil.cpp:622-638: stbi_write_png, 4 components, flipped vertically): 8-bit
 * RGBA, same flip.  (Stored deflate blocks: the file is a valid PNG of the same pixels, not the same bytes as stb's.) */
int pt_image_write_png(const char* path, const void* frame_rgba8_host, uint32_t width, uint32_t height);
/* saveImage's own dispatch on the last three characters of the name ("ppm" / "png", either case; sutil.cpp:548-549, :622) */
int pt_image_write(const char* path, const void* frame_rgba8_host, uint32_t width, uint32_t height);
/* The running average + the next currentFrameIdx are the renderer's whole resumable state
 * (pathTracerPrograms.cu:803-811).  Host pointers; pt_accum_load checks the size. */
int pt_accum_save(const char* path, const void* accum_float4_host, uint32_t width, uint32_t height, uint32_t next_frame_idx);
int pt_accum_load(const char* path, void* accum_float4_host, uint32_t width, uint32_t height, uint32_t* next_frame_idx);

/* ---- device microbenchmarks for the roofline denominators (bench.py) ----------------------------------- */
int pt_device_info(int* sm_count, int* l2_bytes, int* clock_khz, char* name64);
int pt_bench_fp32_peak(double* tflops);          /* dependent-free FFMA loop on every SM */
int pt_bench_l2_bandwidth(double* gbps);         /* streaming read of an L2-resident buffer */
int pt_bench_hbm_bandwidth(double* gbps);        /* streaming copy of a buffer much larger than L2 */

#ifdef __cplusplus
}
#endif
#endif /* PTB200_H */
