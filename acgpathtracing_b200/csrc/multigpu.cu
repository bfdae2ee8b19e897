// multigpu.cu -- pt_group / pt_launch_multi: one image on several B200s behind the C ABI (include/ptb200.h).
//
// The reference renders on one GPU: main's do-while (PathTracerMain.cpp:700-730) calls LaunchCurrentFrame (:184-210) once
// per subframe.  A multi-GPU host calls pt_launch_multi instead; what happens inside is stated in ptb200.h and DESIGN.md 5:
//   plan      whole frames dealt to ranks, leftover frames cut into (frame, tile class range) sub-shares   [make_plan]
//   render    every job as the launch's OWN mean (render_launch raw mode) into a full-size scratch frame
//   exchange  slice j of every job's frame goes to rank j: NCCL send/recv, or copy engines between the devices of one
//             process (cudaMemcpyPeerAsync: no SM is taken from the persistent render kernel)
//   fold      rank j applies the reference's running average (pathTracerPrograms.cu:803-811) to slice j in frame order,
//             picking, for a frame rendered in shares, the piece whose tile classes own the pixel   [fold_slice_kernel]
//   gather    the folded slices are collected (every rank, or rank 0 of a local group)
// Everything is enqueued on per-member streams fenced against the caller's stream: pt_launch_multi is as asynchronous as
// pt_launch.  One thread drives all local devices (NCCL calls grouped), as the reference's single host thread would.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nccl.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <mutex>
#include <type_traits>
#include <string>
#include <unordered_set>
#include <vector>

#include "pt_scene.h"
#include "pt_shade.cuh"

namespace ptb {

using namespace ptd;

namespace {

constexpr uint64_t kGroupMagic = 0x5054423230304752ull;   // "PTB200GR"
constexpr uint32_t kMaxRanks = 16;
constexpr uint32_t kRawFramesPerPass = 8;                 // whole frames a rank renders between two folds (bounds the scratch)
constexpr uint32_t kMaxPassFrames = kMaxRanks * kRawFramesPerPass;
constexpr uint32_t kMaxPieces = kMaxPassFrames + 2 * kMaxRanks;

struct Job {
    uint32_t frame, rank, c0, nc;
};

// See ptb200.h.  Jobs come out in exchange order: round i of the whole frames (rank 0 .. N-1), then the leftover
// sub-share jobs rank by rank.
std::vector<Job> make_plan(uint32_t K, uint32_t N)
{
    std::vector<Job> jobs;
    const uint32_t rounds = K / N, L = K % N;
    for (uint32_t i = 0; i < rounds; ++i)
        for (uint32_t r = 0; r < N; ++r) jobs.push_back(Job{i * N + r, r, 0, N});
    for (uint32_t r = 0; r < N && L; ++r) {
        uint32_t s0 = r * L;
        const uint32_t s1 = s0 + L;                        // sub-share s = (frame s / N, class s % N)
        while (s0 < s1) {
            const uint32_t f = s0 / N, c = s0 % N, n = std::min(N - c, s1 - s0);
            jobs.push_back(Job{rounds * N + f, r, c, n});
            s0 += n;
        }
    }
    return jobs;
}

inline uint64_t slice_begin(uint64_t px, uint32_t j, uint32_t N) { return j >= N ? px : ((px * j / N) & ~(uint64_t)63); }

// ---- fold of one pixel slice --------------------------------------------------------------------------------------
struct FoldPiece {
    const float4* base;     // pixel p of the slice is base[p - slice_begin]
    uint32_t c0, c1;        // tile classes [c0, c1) this piece holds
};
struct FoldTable {
    FoldPiece piece[kMaxPieces];
    uint16_t first[kMaxPassFrames + 1];   // pieces of frame k: first[k] .. first[k + 1] - 1
    uint32_t n_frames, first_frame, n_classes, width, tiles_x;
};

__global__ void __launch_bounds__(256) fold_slice_kernel(float4* __restrict__ accum, uchar4* __restrict__ frame_out,
                                                         const __grid_constant__ FoldTable t, uint64_t begin, uint32_t count)
{
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    const uint64_t p = begin + i;
    const uint32_t y = (uint32_t)(p / t.width), x = (uint32_t)(p - (uint64_t)y * t.width);
    const uint32_t cls = ((y >> 2) * t.tiles_x + (x >> 3)) % t.n_classes;
    float3 acc;
    uint32_t k = 0;
    auto mean_of = [&](uint32_t frame) {
        uint32_t q = t.first[frame];
        const uint32_t q1 = t.first[frame + 1];
        for (; q + 1 < q1; ++q)
            if (cls >= t.piece[q].c0 && cls < t.piece[q].c1) break;
        return t.piece[q].base[i];
    };
    if (t.first_frame == 0u) {                       // frame 0 overwrites (:803: only frames > 0 read the buffer)
        const float4 m = mean_of(0);
        acc = f3(m.x, m.y, m.z);
        k = 1;
    } else {
        const float4 prev = accum[p];
        acc = f3(prev.x, prev.y, prev.z);
    }
    for (; k < t.n_frames; ++k) {
        const float w = 1.0f / (float)((int)(t.first_frame + k) + 1);
        const float4 m = mean_of(k);
        acc = acc + w * (f3(m.x, m.y, m.z) - acc);   // lerp(prev, accum, a), vec_math.h:500 -- same order as write_pixel
    }
    accum[p] = make_float4(acc.x, acc.y, acc.z, 1.0f);
    if (frame_out) frame_out[p] = make_color(acc);
}

// ---- group ------------------------------------------------------------------------------------------------------------
struct Member {
    int device = 0;
    uint32_t rank = 0;
    Scene* scene = nullptr;
    cudaStream_t stream = nullptr, xfer = nullptr;
    ncclComm_t comm = nullptr;
    std::vector<float4*> raw;        // scratch frames (one per job of a pass)
    size_t raw_px = 0;
    float4* rx = nullptr;            // received slices: piece q at rx + q * slice_cap
    size_t rx_px = 0;
    float4* acc = nullptr;           // internal image (host mode; members other than the first of a local group)
    uchar4* fb = nullptr;
    size_t img_px = 0;
    std::vector<cudaEvent_t> ev_job;
    cudaEvent_t ev_pass = nullptr, ev_xfer = nullptr, ev_end = nullptr, ev_t[4] = {nullptr, nullptr, nullptr, nullptr};
    uint32_t n_kernels = 0;
    bool timed = false;
};

struct Group {
    uint64_t magic = kGroupMagic;
    uint32_t n_ranks = 0, first_rank = 0;
    int transport = PT_TRANSPORT_PEER;
    bool local = true;               // every rank is a member of this process
    std::vector<Member> m;
    cudaEvent_t ev_user = nullptr;
};

std::mutex g_group_mutex;
std::unordered_set<const Group*> g_groups;

Group* group_from(const pt_group* g)
{
    Group* G = reinterpret_cast<Group*>(const_cast<pt_group*>(g));
    std::lock_guard<std::mutex> lock(g_group_mutex);
    if (!G || g_groups.find(G) == g_groups.end() || G->magic != kGroupMagic) return nullptr;
    return G;
}

// NCCL is bound at run time, the first time a group needs it, not at link time: a process may already hold an NCCL of its
// own (PyTorch bundles one under the same soname, newer than the system's) and two copies of libnccl.so.2 cannot coexist --
// whichever is loaded first wins for everyone.  Order: $PTB_NCCL_LIB, the libnccl.so.2 already in the process, the
// loader's default search.  Only API that has been stable since NCCL 2.7 (send/recv) is used.
struct NcclApi {
    decltype(&ncclGetErrorString) GetErrorString = nullptr;
    decltype(&ncclGetVersion) GetVersion = nullptr;
    decltype(&ncclGetUniqueId) GetUniqueId = nullptr;
    decltype(&ncclCommInitRank) CommInitRank = nullptr;
    decltype(&ncclCommInitAll) CommInitAll = nullptr;
    decltype(&ncclCommDestroy) CommDestroy = nullptr;
    decltype(&ncclGroupStart) GroupStart = nullptr;
    decltype(&ncclGroupEnd) GroupEnd = nullptr;
    decltype(&ncclSend) Send = nullptr;
    decltype(&ncclRecv) Recv = nullptr;
    void* handle = nullptr;
    std::string error;
};

NcclApi& nccl_api()
{
    static NcclApi api;
    static std::once_flag once;
    std::call_once(once, [] {
        const char* env = getenv("PTB_NCCL_LIB");
        if (env && *env) api.handle = dlopen(env, RTLD_NOW | RTLD_LOCAL);
        if (!api.handle) api.handle = dlopen("libnccl.so.2", RTLD_NOW | RTLD_LOCAL | RTLD_NOLOAD);
        if (!api.handle) api.handle = dlopen("libnccl.so.2", RTLD_NOW | RTLD_LOCAL);
        if (!api.handle) {
            api.error = std::string("cannot load libnccl.so.2: ") + dlerror();
            return;
        }
        bool ok = true;
        auto bind = [&](auto& fn, const char* name) {
            fn = reinterpret_cast<std::remove_reference_t<decltype(fn)>>(dlsym(api.handle, name));
            if (!fn) ok = false, api.error = std::string("libnccl.so.2 lacks ") + name;
        };
        bind(api.GetErrorString, "ncclGetErrorString");
        bind(api.GetVersion, "ncclGetVersion");
        bind(api.GetUniqueId, "ncclGetUniqueId");
        bind(api.CommInitRank, "ncclCommInitRank");
        bind(api.CommInitAll, "ncclCommInitAll");
        bind(api.CommDestroy, "ncclCommDestroy");
        bind(api.GroupStart, "ncclGroupStart");
        bind(api.GroupEnd, "ncclGroupEnd");
        bind(api.Send, "ncclSend");
        bind(api.Recv, "ncclRecv");
        if (!ok) {
            dlclose(api.handle);
            api.handle = nullptr;
        }
    });
    return api;
}

int nccl_fail(ncclResult_t r, const char* what)
{
    return fail(PT_ERR_CUDA, std::string(what) + ": NCCL " + nccl_api().GetErrorString(r));
}
#define PT_NEED_NCCL()                                                   \
    do {                                                                 \
        if (!nccl_api().handle) return fail(PT_ERR_CUDA, nccl_api().error); \
    } while (0)

#define PT_CUDA(expr)                                         \
    do {                                                      \
        cudaError_t e__ = (expr);                             \
        if (e__ != cudaSuccess) return cuda_fail(e__, #expr); \
    } while (0)
#define PT_NCCL(expr)                                         \
    do {                                                      \
        ncclResult_t r__ = (expr);                            \
        if (r__ != ncclSuccess) return nccl_fail(r__, #expr); \
    } while (0)

struct DeviceRestore {
    int prev = -1;
    DeviceRestore() { cudaGetDevice(&prev); }
    ~DeviceRestore()
    {
        if (prev >= 0) cudaSetDevice(prev);
    }
};

int member_init(Member& m)
{
    PT_CUDA(cudaSetDevice(m.device));
    cudaDeviceProp prop;
    PT_CUDA(cudaGetDeviceProperties(&prop, m.device));
    if (prop.major < 10) return fail(PT_ERR_CUDA, std::string("device is not sm_100-class: ") + prop.name);
    PT_CUDA(cudaStreamCreateWithFlags(&m.stream, cudaStreamNonBlocking));
    PT_CUDA(cudaStreamCreateWithFlags(&m.xfer, cudaStreamNonBlocking));
    PT_CUDA(cudaEventCreateWithFlags(&m.ev_pass, cudaEventDisableTiming));
    PT_CUDA(cudaEventCreateWithFlags(&m.ev_xfer, cudaEventDisableTiming));
    PT_CUDA(cudaEventCreateWithFlags(&m.ev_end, cudaEventDisableTiming));
    for (cudaEvent_t& e : m.ev_t) PT_CUDA(cudaEventCreate(&e));
    return PT_OK;
}

void member_free(Member& m)
{
    cudaSetDevice(m.device);
    if (m.stream) cudaStreamSynchronize(m.stream);
    if (m.xfer) cudaStreamSynchronize(m.xfer);
    if (m.comm) nccl_api().CommDestroy(m.comm);
    if (m.scene) pt_scene_destroy(reinterpret_cast<pt_scene*>(m.scene));
    for (float4* p : m.raw)
        if (p) cudaFree(p);
    if (m.rx) cudaFree(m.rx);
    if (m.acc) cudaFree(m.acc);
    if (m.fb) cudaFree(m.fb);
    for (cudaEvent_t e : m.ev_job) cudaEventDestroy(e);
    for (cudaEvent_t e : {m.ev_pass, m.ev_xfer, m.ev_end, m.ev_t[0], m.ev_t[1], m.ev_t[2], m.ev_t[3]})
        if (e) cudaEventDestroy(e);
    if (m.stream) cudaStreamDestroy(m.stream);
    if (m.xfer) cudaStreamDestroy(m.xfer);
}

void group_free(Group* G)
{
    DeviceRestore restore;
    for (Member& m : G->m) member_free(m);
    if (G->ev_user) {
        cudaSetDevice(G->m.empty() ? 0 : G->m[0].device);
        cudaEventDestroy(G->ev_user);
    }
    G->magic = 0;
    delete G;
}

// grow-only scratch; cudaMalloc / cudaFree synchronise, so this only happens when a launch is larger than any before
int member_scratch(Member& m, size_t px, size_t n_raw, size_t rx_px, bool want_image)
{
    if (m.raw_px < px) {
        for (float4*& p : m.raw) {
            if (p) cudaFree(p);
            p = nullptr;
        }
        m.raw.clear();
        m.raw_px = px;
    }
    while (m.raw.size() < n_raw) {
        float4* p = nullptr;
        PT_CUDA(cudaMalloc((void**)&p, m.raw_px * sizeof(float4)));
        m.raw.push_back(p);
    }
    if (m.rx_px < rx_px) {
        if (m.rx) cudaFree(m.rx);
        m.rx = nullptr;
        m.rx_px = 0;
        PT_CUDA(cudaMalloc((void**)&m.rx, rx_px * sizeof(float4)));
        m.rx_px = rx_px;
    }
    if (want_image && m.img_px < px) {
        if (m.acc) cudaFree(m.acc);
        if (m.fb) cudaFree(m.fb);
        m.acc = nullptr;
        m.fb = nullptr;
        m.img_px = 0;
        PT_CUDA(cudaMalloc((void**)&m.acc, px * sizeof(float4)));
        PT_CUDA(cudaMalloc((void**)&m.fb, px * sizeof(uchar4)));
        m.img_px = px;
    }
    while (m.ev_job.size() < n_raw) {
        cudaEvent_t e = nullptr;
        PT_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        m.ev_job.push_back(e);
    }
    return PT_OK;
}

struct Target {            // where a member's slice of the image lives on its device
    float4* acc;
    uchar4* fb;
};

int launch_multi(Group& G, const pt_params& up, uint32_t K, cudaStream_t user_stream, bool host, uint64_t* h2d_out, uint64_t* d2h_out)
{
    if (K == 0) return fail(PT_ERR_INVALID, "pt_launch_multi: n_launches must be > 0");
    if (up.width == 0 || up.height == 0 || up.samplesPerPixel == 0) return fail(PT_ERR_INVALID, "pt_launch_multi: width, height and samplesPerPixel must be > 0");
    if ((uint64_t)up.width * up.height > 0x7fffffffull) return fail(PT_ERR_INVALID, "image too large");
    if (!up.accumulationBuffer) return fail(PT_ERR_INVALID, "accumulationBuffer is NULL");
    if ((uint64_t)up.currentFrameIdx + K > 0x7fffffffull) return fail(PT_ERR_INVALID, "frame index overflow");
    if (G.m.empty() || !G.m[0].scene) return fail(PT_ERR_INVALID, "pt_launch_multi: the group has no scene (pt_group_scene_create)");
    if (scene_from_handle(up.handle) != G.m[0].scene) return fail(PT_ERR_INVALID, "params.handle is not the handle of the group's scene");
    DeviceRestore restore;
    const uint32_t N = G.n_ranks;
    const uint64_t px = (uint64_t)up.width * up.height;
    const uint32_t tiles_x = (up.width + 7) / 8;
    const bool want_fb = up.frameBuffer != nullptr;
    uint64_t h2d = 0, d2h = 0;
    uint64_t slice_cap = 0;
    for (uint32_t j = 0; j < N; ++j) slice_cap = std::max(slice_cap, slice_begin(px, j + 1, N) - slice_begin(px, j, N));
    const bool nccl = G.transport == PT_TRANSPORT_NCCL;

    // ---- targets, scratch, begin fence ------------------------------------------------------------------------------
    const uint32_t pass_cap = N * kRawFramesPerPass;
    const uint32_t first_pass = std::min(K, pass_cap);
    const size_t n_raw = N == 1 ? 0 : (first_pass / N + 2);
    const size_t n_pieces_max = N == 1 ? 0 : (first_pass + 2 * N);
    std::vector<Target> tgt(G.m.size());
    for (size_t i = 0; i < G.m.size(); ++i) {
        Member& m = G.m[i];
        PT_CUDA(cudaSetDevice(m.device));
        if (i > 0) {   // the settings of the first replica (the one the caller holds) apply to every replica
            const Scene& a = *G.m[0].scene;
            Scene& b = *m.scene;
            memcpy(b.background, a.background, sizeof(b.background));
            b.pipeline = a.pipeline;
            b.wf = a.wf;
            b.count_enabled = a.count_enabled;
            b.count_executed = a.count_executed;
            b.traversal = a.traversal;
            b.tree = a.tree;
            b.tile_order_enabled = a.tile_order_enabled;
            if (b.tree == 3) PT_CUDA(ensure_ploc(b));
        }
        const bool internal = host || (G.local && i > 0);
        int rc = member_scratch(m, px, n_raw, n_pieces_max * slice_cap, internal);
        if (rc != PT_OK) return rc;
        tgt[i] = internal ? Target{m.acc, want_fb ? m.fb : nullptr} : Target{(float4*)up.accumulationBuffer, (uchar4*)up.frameBuffer};
        m.n_kernels = 0;
        m.timed = true;
    }
    if (!host) {
        PT_CUDA(cudaSetDevice(G.m[0].device));
        PT_CUDA(cudaEventRecord(G.ev_user, user_stream));
    }
    for (size_t i = 0; i < G.m.size(); ++i) {
        Member& m = G.m[i];
        PT_CUDA(cudaSetDevice(m.device));
        if (!host) PT_CUDA(cudaStreamWaitEvent(m.stream, G.ev_user, 0));
        PT_CUDA(cudaEventRecord(m.ev_t[0], m.stream));
        // the previous running average of this member's slice
        const uint64_t b = slice_begin(px, m.rank, N), e = slice_begin(px, m.rank + 1, N);
        if (up.currentFrameIdx > 0 && e > b) {
            if (host) {
                const uint64_t lo = N == 1 ? 0 : b, hi = N == 1 ? px : e;
                PT_CUDA(cudaMemcpyAsync(m.acc + lo, (const float4*)up.accumulationBuffer + lo, (hi - lo) * sizeof(float4), cudaMemcpyHostToDevice, m.stream));
                h2d += (hi - lo) * sizeof(float4);
            } else if (G.local && i > 0) {
                PT_CUDA(cudaMemcpyPeerAsync(m.acc + b, m.device, (const float4*)up.accumulationBuffer + b, G.m[0].device, (e - b) * sizeof(float4), m.stream));
            }
        }
    }
    h2d += sizeof(pt_params);   // the launch record itself (reference: cudaMemcpyAsync(d_params), PathTracerMain.cpp:191)

    if (N == 1) {
        // one GPU: the launches in sequence, folded in place by the render kernel itself
        Member& m = G.m[0];
        PT_CUDA(cudaSetDevice(m.device));
        for (uint32_t k = 0; k < K; ++k) {
            pt_params p = up;
            p.currentFrameIdx = up.currentFrameIdx + k;
            p.accumulationBuffer = tgt[0].acc;
            p.frameBuffer = tgt[0].fb;
            PT_CUDA(render_launch(*m.scene, p, 0, 1, m.stream, nullptr));
            m.n_kernels++;
        }
        for (int t = 1; t < 4; ++t) PT_CUDA(cudaEventRecord(m.ev_t[t], m.stream));
    }

    // ---- passes: render, exchange, fold --------------------------------------------------------------------------------
    uint32_t done = 0;
    while (N > 1 && done < K) {
        const uint32_t Kp = std::min(K - done, pass_cap);
        const bool last_pass = done + Kp == K;
        const std::vector<Job> jobs = make_plan(Kp, N);
        // the i-th job of every rank
        std::vector<std::vector<uint32_t>> of_rank(N);
        for (uint32_t q = 0; q < jobs.size(); ++q) of_rank[jobs[q].rank].push_back(q);
        size_t max_jobs = 0;
        for (const auto& v : of_rank) max_jobs = std::max(max_jobs, v.size());
        if (jobs.size() > kMaxPieces || max_jobs > n_raw) return fail(PT_ERR_INVALID, "pt_launch_multi: internal plan overflow");

        for (Member& m : G.m) {
            PT_CUDA(cudaSetDevice(m.device));
            PT_CUDA(cudaEventRecord(m.ev_pass, m.stream));
        }
        // incoming slices overwrite rx: not before the receiver's previous fold is done
        for (Member& m : G.m) {
            PT_CUDA(cudaSetDevice(m.device));
            if (nccl) {
                PT_CUDA(cudaStreamWaitEvent(m.xfer, m.ev_pass, 0));
            } else {
                for (Member& o : G.m) PT_CUDA(cudaStreamWaitEvent(m.xfer, o.ev_pass, 0));
            }
        }
        // exchange round i: slice d of every job rendered in round i goes to rank d
        auto exchange_round = [&](size_t i) -> int {
            for (Member& m : G.m) {
                PT_CUDA(cudaSetDevice(m.device));
                const bool mine = i < of_rank[m.rank].size();
                for (uint32_t r = 0; r < N; ++r) {
                    if (r == m.rank) continue;
                    if (mine) {   // send slice r of my job to rank r
                        const uint64_t b = slice_begin(px, r, N), e = slice_begin(px, r + 1, N);
                        const uint32_t q = of_rank[m.rank][i];
                        if (e > b) {
                            if (nccl) {
                                PT_NCCL(nccl_api().Send(m.raw[i] + b, (e - b) * sizeof(float4), ncclChar, (int)r, m.comm, m.xfer));
                            } else {
                                Member& dst = G.m[r - G.first_rank];
                                PT_CUDA(cudaMemcpyPeerAsync(dst.rx + (uint64_t)q * slice_cap, dst.device, m.raw[i] + b, m.device,
                                                            (e - b) * sizeof(float4), m.xfer));
                            }
                        }
                    }
                    if (nccl && i < of_rank[r].size()) {   // receive my slice of rank r's job
                        const uint64_t b = slice_begin(px, m.rank, N), e = slice_begin(px, m.rank + 1, N);
                        const uint32_t q = of_rank[r][i];
                        if (e > b) PT_NCCL(nccl_api().Recv(m.rx + (uint64_t)q * slice_cap, (e - b) * sizeof(float4), ncclChar, (int)r, m.comm, m.xfer));
                    }
                }
            }
            return PT_OK;
        };
        for (size_t i = 0; i < max_jobs; ++i) {
            // render job i of every local member (round-major: all devices get work before the host looks at the next round)
            for (Member& m : G.m) {
                if (i >= of_rank[m.rank].size()) continue;
                const Job& J = jobs[of_rank[m.rank][i]];
                PT_CUDA(cudaSetDevice(m.device));
                pt_params p = up;
                p.currentFrameIdx = up.currentFrameIdx + done + J.frame;
                p.accumulationBuffer = m.raw[i];
                p.frameBuffer = nullptr;
                p.handle = pt_scene_handle(reinterpret_cast<pt_scene*>(m.scene));
                const bool whole = J.nc == N;
                PT_CUDA(render_launch(*m.scene, p, whole ? 0 : J.c0, whole ? 1 : N, m.stream, nullptr, true, whole ? 1 : J.nc));
                m.n_kernels++;
                PT_CUDA(cudaEventRecord(m.ev_job[i], m.stream));
                PT_CUDA(cudaStreamWaitEvent(m.xfer, m.ev_job[i], 0));
            }
            // copy engines move a finished frame's slices while the next frame renders; they take nothing from the render kernel
            if (!nccl) {
                const int rc = exchange_round(i);
                if (rc != PT_OK) return rc;
            }
        }
        if (nccl) {
            // NCCL's send / recv kernels need SMs, and the persistent render kernel owns all of them: issued between two
            // frames they slow the next frame down (measured: +4.5 % on the render phase at 2 GPUs) for a transfer that takes
            // well under a millisecond.  All rounds of the pass go out together once the pass has been rendered.
            PT_NCCL(nccl_api().GroupStart());
            for (size_t i = 0; i < max_jobs; ++i) {
                const int rc = exchange_round(i);
                if (rc != PT_OK) {
                    nccl_api().GroupEnd();
                    return rc;
                }
            }
            PT_NCCL(nccl_api().GroupEnd());
        }
        for (Member& m : G.m) {
            PT_CUDA(cudaSetDevice(m.device));
            PT_CUDA(cudaEventRecord(m.ev_xfer, m.xfer));
            if (last_pass) PT_CUDA(cudaEventRecord(m.ev_t[1], m.stream));
        }
        // fold this member's slice
        for (size_t mi = 0; mi < G.m.size(); ++mi) {
            Member& m = G.m[mi];
            PT_CUDA(cudaSetDevice(m.device));
            if (nccl) {
                PT_CUDA(cudaStreamWaitEvent(m.stream, m.ev_xfer, 0));
            } else {
                for (Member& o : G.m) PT_CUDA(cudaStreamWaitEvent(m.stream, o.ev_xfer, 0));
            }
            const uint64_t b = slice_begin(px, m.rank, N), e = slice_begin(px, m.rank + 1, N);
            if (e > b) {
                FoldTable t;
                memset(&t, 0, sizeof(t));
                uint32_t n = 0;
                for (uint32_t k = 0; k < Kp; ++k) {
                    t.first[k] = (uint16_t)n;
                    for (uint32_t q = 0; q < jobs.size(); ++q) {
                        if (jobs[q].frame != k) continue;
                        const Job& J = jobs[q];
                        const float4* base;
                        if (J.rank == m.rank) {
                            const auto& mine = of_rank[m.rank];
                            const size_t idx = std::find(mine.begin(), mine.end(), q) - mine.begin();
                            base = m.raw[idx] + b;                            // my own frame: no copy
                        } else {
                            base = m.rx + (uint64_t)q * slice_cap;
                        }
                        t.piece[n++] = FoldPiece{base, J.c0, J.c0 + J.nc};
                    }
                }
                t.first[Kp] = (uint16_t)n;
                t.n_frames = Kp;
                t.first_frame = up.currentFrameIdx + done;
                t.n_classes = N;
                t.width = up.width;
                t.tiles_x = tiles_x;
                const uint32_t count = (uint32_t)(e - b);
                fold_slice_kernel<<<(count + 255u) / 256u, 256, 0, m.stream>>>(tgt[mi].acc, last_pass ? tgt[mi].fb : nullptr, t, b, count);
                PT_CUDA(cudaGetLastError());
                m.n_kernels++;
            }
            if (last_pass) PT_CUDA(cudaEventRecord(m.ev_t[2], m.stream));
        }
        done += Kp;
    }

    // ---- gather the folded slices -------------------------------------------------------------------------------------
    if (N > 1) {
        if (!G.local) {
            // one process per GPU: every rank gets every slice (in place in its own buffers)
            Member& m = G.m[0];
            PT_CUDA(cudaSetDevice(m.device));
            PT_NCCL(nccl_api().GroupStart());
            const uint64_t mb = slice_begin(px, m.rank, N), me = slice_begin(px, m.rank + 1, N);
            for (uint32_t r = 0; r < N; ++r) {
                if (r == m.rank) continue;
                const uint64_t b = slice_begin(px, r, N), e = slice_begin(px, r + 1, N);
                if (me > mb) {
                    PT_NCCL(nccl_api().Send(tgt[0].acc + mb, (me - mb) * sizeof(float4), ncclChar, (int)r, m.comm, m.stream));
                    if (tgt[0].fb) PT_NCCL(nccl_api().Send(tgt[0].fb + mb, (me - mb) * sizeof(uchar4), ncclChar, (int)r, m.comm, m.stream));
                }
                if (e > b) {
                    PT_NCCL(nccl_api().Recv(tgt[0].acc + b, (e - b) * sizeof(float4), ncclChar, (int)r, m.comm, m.stream));
                    if (tgt[0].fb) PT_NCCL(nccl_api().Recv(tgt[0].fb + b, (e - b) * sizeof(uchar4), ncclChar, (int)r, m.comm, m.stream));
                }
            }
            PT_NCCL(nccl_api().GroupEnd());
        } else if (!host) {
            // local group: the slices go to the caller's buffers on the first device
            if (nccl) PT_NCCL(nccl_api().GroupStart());
            for (size_t i = 1; i < G.m.size(); ++i) {
                Member& m = G.m[i];
                const uint64_t b = slice_begin(px, m.rank, N), e = slice_begin(px, m.rank + 1, N);
                if (e == b) continue;
                PT_CUDA(cudaSetDevice(m.device));
                if (nccl) {
                    PT_NCCL(nccl_api().Send(tgt[i].acc + b, (e - b) * sizeof(float4), ncclChar, 0, m.comm, m.stream));
                    if (tgt[i].fb) PT_NCCL(nccl_api().Send(tgt[i].fb + b, (e - b) * sizeof(uchar4), ncclChar, 0, m.comm, m.stream));
                    PT_CUDA(cudaSetDevice(G.m[0].device));
                    PT_NCCL(nccl_api().Recv(tgt[0].acc + b, (e - b) * sizeof(float4), ncclChar, (int)m.rank, G.m[0].comm, G.m[0].stream));
                    if (tgt[i].fb) PT_NCCL(nccl_api().Recv(tgt[0].fb + b, (e - b) * sizeof(uchar4), ncclChar, (int)m.rank, G.m[0].comm, G.m[0].stream));
                } else {
                    PT_CUDA(cudaMemcpyPeerAsync(tgt[0].acc + b, G.m[0].device, tgt[i].acc + b, m.device, (e - b) * sizeof(float4), m.stream));
                    if (tgt[i].fb)
                        PT_CUDA(cudaMemcpyPeerAsync(tgt[0].fb + b, G.m[0].device, tgt[i].fb + b, m.device, (e - b) * sizeof(uchar4), m.stream));
                }
            }
            if (nccl) PT_NCCL(nccl_api().GroupEnd());
        }
        for (Member& m : G.m) {
            PT_CUDA(cudaSetDevice(m.device));
            PT_CUDA(cudaEventRecord(m.ev_t[3], m.stream));
        }
    }

    // ---- results out ----------------------------------------------------------------------------------------------------
    if (host) {
        for (size_t i = 0; i < G.m.size(); ++i) {
            Member& m = G.m[i];
            PT_CUDA(cudaSetDevice(m.device));
            // local group: every device writes its own slice of the one host image; process per GPU: the complete image
            uint64_t b = 0, e = px;
            if (G.local && N > 1) b = slice_begin(px, m.rank, N), e = slice_begin(px, m.rank + 1, N);
            if (e > b) {
                PT_CUDA(cudaMemcpyAsync((float4*)up.accumulationBuffer + b, m.acc + b, (e - b) * sizeof(float4), cudaMemcpyDeviceToHost, m.stream));
                d2h += (e - b) * sizeof(float4);
                if (want_fb) {
                    PT_CUDA(cudaMemcpyAsync((uchar4*)up.frameBuffer + b, m.fb + b, (e - b) * sizeof(uchar4), cudaMemcpyDeviceToHost, m.stream));
                    d2h += (e - b) * sizeof(uchar4);
                }
            }
        }
        bool watchdog = false;
        for (Member& m : G.m) {
            PT_CUDA(cudaSetDevice(m.device));
            PT_CUDA(cudaStreamSynchronize(m.stream));
            unsigned long long flag = 0;
            PT_CUDA(cudaMemcpy(&flag, m.scene->d_counters + 15, sizeof(flag), cudaMemcpyDeviceToHost));
            watchdog = watchdog || flag != 0;
        }
        if (watchdog) return fail(PT_ERR_CUDA, "render kernel watchdog fired (scheduler made no progress)");
    } else {
        for (Member& m : G.m) {
            PT_CUDA(cudaSetDevice(m.device));
            PT_CUDA(cudaEventRecord(m.ev_end, m.stream));
        }
        PT_CUDA(cudaSetDevice(G.m[0].device));
        for (Member& m : G.m) PT_CUDA(cudaStreamWaitEvent(user_stream, m.ev_end, 0));
    }
    if (h2d_out) *h2d_out = h2d;
    if (d2h_out) *d2h_out = d2h;
    return PT_OK;
}

}  // namespace

}  // namespace ptb

using namespace ptb;

extern "C" {

int pt_multi_plan(uint32_t n_launches, uint32_t n_ranks, pt_multi_job* jobs, uint32_t capacity, uint32_t* n_jobs)
{
    if (n_ranks == 0 || n_ranks > kMaxRanks || !n_jobs) return fail(PT_ERR_INVALID, "pt_multi_plan: bad argument");
    const std::vector<Job> plan = make_plan(n_launches, n_ranks);
    *n_jobs = (uint32_t)plan.size();
    if (!jobs) return PT_OK;
    if (capacity < plan.size()) return fail(PT_ERR_INVALID, "pt_multi_plan: capacity too small");
    for (size_t i = 0; i < plan.size(); ++i) jobs[i] = pt_multi_job{plan[i].frame, plan[i].rank, plan[i].c0, plan[i].nc};
    return PT_OK;
}

int pt_multi_slice(uint32_t width, uint32_t height, uint32_t rank, uint32_t n_ranks, uint64_t* first_pixel, uint64_t* n_pixels)
{
    if (n_ranks == 0 || rank >= n_ranks || width == 0 || height == 0) return fail(PT_ERR_INVALID, "pt_multi_slice: bad argument");
    const uint64_t px = (uint64_t)width * height, b = slice_begin(px, rank, n_ranks), e = slice_begin(px, rank + 1, n_ranks);
    if (first_pixel) *first_pixel = b;
    if (n_pixels) *n_pixels = e - b;
    return PT_OK;
}

int pt_group_create_local(const int* devices, uint32_t n_devices, int transport, pt_group** out)
{
    if (!out) return fail(PT_ERR_INVALID, "out is NULL");
    *out = nullptr;
    if (!devices || n_devices == 0 || n_devices > kMaxRanks) return fail(PT_ERR_INVALID, "pt_group_create_local: 1 .. 16 devices");
    if (transport != PT_TRANSPORT_PEER && transport != PT_TRANSPORT_NCCL) return fail(PT_ERR_INVALID, "pt_group_create_local: bad transport");
    if (transport == PT_TRANSPORT_NCCL && n_devices > 1) PT_NEED_NCCL();
    int count = 0;
    PT_CUDA(cudaGetDeviceCount(&count));
    for (uint32_t i = 0; i < n_devices; ++i) {
        if (devices[i] < 0 || devices[i] >= count) return fail(PT_ERR_INVALID, "pt_group_create_local: no such device");
        for (uint32_t j = 0; j < i && transport == PT_TRANSPORT_NCCL; ++j)
            if (devices[j] == devices[i]) return fail(PT_ERR_INVALID, "pt_group_create_local: NCCL needs distinct devices");
    }
    DeviceRestore restore;
    Group* G = new Group();
    G->n_ranks = n_devices;
    G->first_rank = 0;
    G->transport = transport;
    G->local = true;
    G->m.resize(n_devices);
    int rc = PT_OK;
    for (uint32_t i = 0; i < n_devices && rc == PT_OK; ++i) {
        G->m[i].device = devices[i];
        G->m[i].rank = i;
        rc = member_init(G->m[i]);
    }
    if (rc == PT_OK) {
        cudaSetDevice(devices[0]);
        if (cudaEventCreateWithFlags(&G->ev_user, cudaEventDisableTiming) != cudaSuccess) rc = fail(PT_ERR_CUDA, "cudaEventCreate");
    }
    if (rc == PT_OK && transport == PT_TRANSPORT_NCCL && n_devices > 1) {
        std::vector<ncclComm_t> comms(n_devices);
        ncclResult_t r = nccl_api().CommInitAll(comms.data(), (int)n_devices, devices);
        if (r != ncclSuccess) rc = nccl_fail(r, "ncclCommInitAll");
        else
            for (uint32_t i = 0; i < n_devices; ++i) G->m[i].comm = comms[i];
    }
    if (rc == PT_OK && transport == PT_TRANSPORT_PEER) {
        // direct NVLink copies where the devices can reach each other; cudaMemcpyPeerAsync stages through the host otherwise
        for (uint32_t i = 0; i < n_devices; ++i)
            for (uint32_t j = 0; j < n_devices; ++j) {
                if (devices[i] == devices[j]) continue;
                int can = 0;
                if (cudaDeviceCanAccessPeer(&can, devices[i], devices[j]) == cudaSuccess && can) {
                    cudaSetDevice(devices[i]);
                    cudaError_t e = cudaDeviceEnablePeerAccess(devices[j], 0);
                    if (e != cudaSuccess) cudaGetLastError();   // already enabled (by us or the application): fine
                }
            }
    }
    if (rc != PT_OK) {
        group_free(G);
        return rc;
    }
    {
        std::lock_guard<std::mutex> lock(g_group_mutex);
        g_groups.insert(G);
    }
    *out = reinterpret_cast<pt_group*>(G);
    return PT_OK;
}

int pt_group_unique_id(void* id_bytes)
{
    if (!id_bytes) return fail(PT_ERR_INVALID, "id_bytes is NULL");
    static_assert(sizeof(ncclUniqueId) == PT_GROUP_ID_BYTES, "ncclUniqueId is 128 bytes");
    PT_NEED_NCCL();
    ncclUniqueId id;
    PT_NCCL(nccl_api().GetUniqueId(&id));
    memcpy(id_bytes, &id, sizeof(id));
    return PT_OK;
}

int pt_group_create_rank(const void* id_bytes, uint32_t rank, uint32_t n_ranks, pt_group** out)
{
    if (!out) return fail(PT_ERR_INVALID, "out is NULL");
    *out = nullptr;
    if (!id_bytes || n_ranks == 0 || n_ranks > kMaxRanks || rank >= n_ranks) return fail(PT_ERR_INVALID, "pt_group_create_rank: bad argument");
    if (n_ranks > 1) PT_NEED_NCCL();
    int dev = 0;
    PT_CUDA(cudaGetDevice(&dev));
    Group* G = new Group();
    G->n_ranks = n_ranks;
    G->first_rank = rank;
    G->transport = PT_TRANSPORT_NCCL;
    G->local = n_ranks == 1;
    G->m.resize(1);
    G->m[0].device = dev;
    G->m[0].rank = rank;
    int rc = member_init(G->m[0]);
    if (rc == PT_OK && cudaEventCreateWithFlags(&G->ev_user, cudaEventDisableTiming) != cudaSuccess) rc = fail(PT_ERR_CUDA, "cudaEventCreate");
    if (rc == PT_OK && n_ranks > 1) {
        ncclUniqueId id;
        memcpy(&id, id_bytes, sizeof(id));
        ncclResult_t r = nccl_api().CommInitRank(&G->m[0].comm, (int)n_ranks, id, (int)rank);
        if (r != ncclSuccess) rc = nccl_fail(r, "ncclCommInitRank");
    }
    if (rc != PT_OK) {
        group_free(G);
        return rc;
    }
    {
        std::lock_guard<std::mutex> lock(g_group_mutex);
        g_groups.insert(G);
    }
    *out = reinterpret_cast<pt_group*>(G);
    return PT_OK;
}

void pt_group_destroy(pt_group* group)
{
    Group* G = reinterpret_cast<Group*>(group);
    if (!G) return;
    {
        std::lock_guard<std::mutex> lock(g_group_mutex);
        if (g_groups.erase(G) == 0) return;
    }
    group_free(G);
}

int pt_group_info(const pt_group* group, uint32_t* n_ranks, uint32_t* n_local, uint32_t* first_rank, int* transport)
{
    Group* G = group_from(group);
    if (!G) return fail(PT_ERR_INVALID, "not a live pt_group");
    if (n_ranks) *n_ranks = G->n_ranks;
    if (n_local) *n_local = (uint32_t)G->m.size();
    if (first_rank) *first_rank = G->first_rank;
    if (transport) *transport = G->transport;
    return PT_OK;
}

int pt_group_scene_create(pt_group* group, const float* verts_xyzw, uint32_t n_verts, const uint32_t* indices, uint32_t n_tris,
                          const uint32_t* mat_ids, const pt_material* materials, uint32_t n_materials, pt_scene** out)
{
    Group* G = group_from(group);
    if (!G) return fail(PT_ERR_INVALID, "not a live pt_group");
    if (!out) return fail(PT_ERR_INVALID, "out is NULL");
    *out = nullptr;
    DeviceRestore restore;
    std::vector<pt_scene*> made;
    int rc = PT_OK;
    for (Member& m : G->m) {
        pt_scene* s = nullptr;
        if (cudaSetDevice(m.device) != cudaSuccess) rc = fail(PT_ERR_CUDA, "cudaSetDevice");
        if (rc == PT_OK) rc = pt_scene_create(verts_xyzw, n_verts, indices, n_tris, mat_ids, materials, n_materials, &s);
        if (rc != PT_OK) break;
        made.push_back(s);
    }
    if (rc != PT_OK) {
        for (pt_scene* s : made) pt_scene_destroy(s);
        return rc;
    }
    for (size_t i = 0; i < G->m.size(); ++i) {
        Member& m = G->m[i];
        if (m.scene) {
            cudaSetDevice(m.device);
            cudaStreamSynchronize(m.stream);
            pt_scene_destroy(reinterpret_cast<pt_scene*>(m.scene));
        }
        m.scene = reinterpret_cast<Scene*>(made[i]);
    }
    *out = made[0];
    return PT_OK;
}

pt_scene* pt_group_scene(const pt_group* group, uint32_t local_index)
{
    Group* G = group_from(group);
    if (!G || local_index >= G->m.size()) return nullptr;
    return reinterpret_cast<pt_scene*>(G->m[local_index].scene);
}

int pt_launch_multi(pt_group* group, const pt_params* params, uint32_t n_launches, void* cuda_stream)
{
    Group* G = group_from(group);
    if (!G) return fail(PT_ERR_INVALID, "not a live pt_group");
    if (!params) return fail(PT_ERR_INVALID, "params is NULL");
    return launch_multi(*G, *params, n_launches, (cudaStream_t)cuda_stream, false, nullptr, nullptr);
}

int pt_launch_multi_host(pt_group* group, const pt_params* params, uint32_t n_launches, uint64_t* h2d_bytes, uint64_t* d2h_bytes)
{
    Group* G = group_from(group);
    if (!G) return fail(PT_ERR_INVALID, "not a live pt_group");
    if (!params) return fail(PT_ERR_INVALID, "params is NULL");
    return launch_multi(*G, *params, n_launches, nullptr, true, h2d_bytes, d2h_bytes);
}

int pt_group_last_timing(pt_group* group, uint32_t local_index, float ms[4], uint32_t* n_kernels)
{
    Group* G = group_from(group);
    if (!G || local_index >= G->m.size()) return fail(PT_ERR_INVALID, "pt_group_last_timing: bad argument");
    Member& m = G->m[local_index];
    if (!m.timed) return fail(PT_ERR_INVALID, "pt_group_last_timing: no launch yet");
    DeviceRestore restore;
    PT_CUDA(cudaSetDevice(m.device));
    PT_CUDA(cudaStreamSynchronize(m.stream));
    if (ms) {
        PT_CUDA(cudaEventElapsedTime(&ms[0], m.ev_t[0], m.ev_t[1]));
        PT_CUDA(cudaEventElapsedTime(&ms[1], m.ev_t[1], m.ev_t[2]));
        PT_CUDA(cudaEventElapsedTime(&ms[2], m.ev_t[2], m.ev_t[3]));
        PT_CUDA(cudaEventElapsedTime(&ms[3], m.ev_t[0], m.ev_t[3]));
    }
    if (n_kernels) *n_kernels = m.n_kernels;
    return PT_OK;
}

}  // extern "C"
