// wavefront.cu -- pipeline 2, the CTA-local wavefront render kernel (the default for every scene it can hold).
//
// Same arithmetic per path as render_persistent_kernel (render.cu) and as the reference programs
// (pathTracerPrograms.cu:707-1031) -- only the scheduling differs, so results are bit-identical.
//
// Why: a path alternates between four kinds of work of very different cost -- NODE (two slab tests), LEAF (one
// triangle test), SHADE (closest-hit shading, BSDF sample, NEE set-up) and END (miss / shadow result, accumulate,
// Russian roulette, next ray / next sample / next pixel).  With one path per lane held in registers the 32 lanes
// of a warp are spread over the four kinds and each kind runs with a handful of lanes (ncu: 25 % useful SIMT
// efficiency, profiles/r1_v2_*).  Here paths live in SHARED MEMORY and lanes are just workers:
//   * one persistent CTA per SM (up to 768 threads) owns NP = 32 * P paths (P <= 32; 1024 for the Cornell box);
//     path id = 32 * j + c belongs to "column" c and is only ever processed by lane c of some warp;
//   * the queues are four 32-bit words per column (one per stage): bit j of mask[stage][c] says that path
//     32 * j + c waits for that stage.  Push = atomicOr, pop = atomicAnd on the lane's own word: no ring buffers,
//     no slots to spin on, nothing to overflow, and every queue word sits in its own bank;
//   * a warp repeatedly votes (ballot + popc) for the most downstream stage that can fill all 32 lanes, else the
//     stage with the most non-empty columns; every lane pops one path of its column, the warp runs that stage's code
//     ONCE for (measured) 27-28 lanes, and every lane pushes its path to the stage it needs next.  NODE batches run
//     several slab-test steps and stop early when too many lanes have left the NODE state.  When only a few paths
//     are left in the CTA (end of the launch) traversals run to completion instead of hopping between queues;
//   * pixels are taken from a global counter, expensive tiles first (render.cu tile_order_update);
//   * bank conflicts are designed out: all per-path state is float4 arrays indexed by path id (ids of a batch are
//     distinct mod 32, hence distinct mod 8 inside every quarter-warp => conflict-free LDS.128/STS.128), traversal
//     stacks are level-major word arrays (bank = column), and a scene small enough is staged as 8 interleaved
//     copies so that a divergent node/triangle fetch costs exactly its data wavefronts (pt_device.cuh).
// Per-pixel order is untouched: one path per pixel at a time, samples in sequence, segments in path order.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <string>

#include "pt_scene.h"
#include "pt_shade.cuh"

namespace ptb {

using namespace ptd;

namespace {

constexpr uint32_t kFull = 0xffffffffu;
constexpr int kWfBlock = 768;          // most threads per CTA (24 warps, 80 registers); one CTA per SM
constexpr int kCtlF4 = 169;            // float4s in front of the state arrays: mask[7][32] + 4 control words + idle[32] + team[13][32]
constexpr int kStateF4 = 9;            // float4 state arrays per path
constexpr size_t kMaxSmem = 227 * 1024;

// stages 0..4 have a queue row each; rows 5 and 6 are the SHADE queues of metallic and refractive hits when shading is sorted by
// material (MSORT: row 2 then holds the diffuse hits only); TAIL serves NODE + LEAF
enum : int { ST_NODE = 0, ST_LEAF = 1, ST_SHADE = 2, ST_END = 3, ST_PIXEL = 4, ST_SHADE1 = 5, ST_SHADE2 = 6, ST_TAIL = 7 };
constexpr int kQueues = 7;
constexpr uint32_t kCandNone = 0xffffffffu;       // per-pixel candidate list: no triangle can be hit by a camera ray
constexpr uint32_t kCandOverflow = 0xfffffffeu;   // more than 4 candidates (or lists disabled): traverse the BVH
enum : int { MODE_RAD = 0, MODE_SHADOW = 1, MODE_SHADED = 2, MODE_FRESH = 3, MODE_NEWPIX = 4, MODE_JOIN = 5 };
constexpr int kTeamBit = 1 << 9;       // P1.x: this slot runs a sample of its column's team (depth | done << 8 | team << 9 | material << 16)

struct WfConst {
    uint32_t np;             // paths per CTA
    uint32_t stack_depth;    // words per path in shared memory
    uint32_t stack_total;    // words per path the tree can need; levels >= stack_depth live in `stack_over`
    int* stack_over;         // [stack_total - stack_depth][grid * np], or null
    int prefetch;            // scenes in global memory: prefetch deferred children towards L2
    uint32_t copies_shift;   // staged scene copies = 1 << copies_shift
    int node_steps;          // K: slab-test steps per NODE batch
    int node_keep8;          // a NODE batch continues while lanes_at_node * 8 >= lanes_at_start * node_keep8
    int full_cols;           // columns that count as "can fill a warp"
    int sticky;              // repeats of the same stage without a full vote while it fills the whole warp
    int tail_cols;           // tail mode when the non-empty columns of all four queues add up to at most this
    int miss_loop;           // samples of an all-miss pixel finished per END visit
    int block_override;      // host only: threads per CTA when set by PTB_TUNE
    int idle_ns;             // sleep of a warp that found every queue empty ...
    int idle_shift;          // ... doubled for every further empty poll, at most this many times
    int affinity;            // schedulers (warp % 4 below this) whose warps prefer the traversal stages (0 = none)
    // sample teams (see "Sample teams" below)
    int team;                // 0 = off
    int team_min;            // a pixel is handed to a team when it has at least this many samples left
    uint32_t team_wmask;     // log window (samples in flight beyond the commit frontier) - 1; power of two - 1
    uint32_t team_cap;       // log entries per sample = maxDepth + 3 (emission, one term per segment, terminator)
    float4* team_log;        // [grid][32 columns][window][team_cap]
};

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

__device__ __forceinline__ uint32_t vol_ld(const uint32_t* p) { return *reinterpret_cast<const volatile uint32_t*>(p); }

// Work item j of this rank -> pixel (same map as render.cu): owned tile j / 32, pixel j % 32 of the 8x4 tile.
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

// Pop one waiting path of this lane's column: returns j (path id = 32 * j + lane) or -1.
__device__ __forceinline__ int wf_pop(uint32_t* word, uint32_t snapshot, int rot)
{
    uint32_t m = snapshot;
    while (m) {
        const uint32_t mr = __funnelshift_r(m, m, rot);   // rotate right: fairness between the column's paths
        const int j = (__ffs(mr) - 1 + rot) & 31;
        const uint32_t bit = 1u << j;
        const uint32_t old = atomicAnd(word, ~bit);
        if (old & bit) {
            __threadfence_block();                          // acquire: the path's state was written before the bit was set
            return j;
        }
        m = old & ~bit;
    }
    return -1;
}

// wf_pop loops a different number of times per lane; without an explicit reconvergence point the compiler lets the
// lanes that won early run the whole stage body before the others (measured: 2.4 executions of the LEAF body per
// batch).  Every stage therefore pops through this wrapper.
__device__ __forceinline__ int wf_pop_converged(uint32_t* word, uint32_t snapshot, int rot)
{
    const int j = wf_pop(word, snapshot, rot);
    __syncwarp();
    return j;
}

__device__ __forceinline__ void wf_push(uint32_t* mask, int stage, int lane, int j)
{
    __threadfence_block();                                  // release: state stores before the bit becomes visible
    atomicOr(mask + stage * 32 + lane, 1u << j);
}

// SCENE = -1: scene records in global memory; SCENE = s >= 0: 1 << s interleaved copies staged in shared memory
template <int SCENE> struct WfFetch { typedef FetchSharedBanked<SCENE> type; };
template <> struct WfFetch<-1> { typedef FetchGlobal type; };

// traversal stack of a path: whole in shared memory when the scene is staged there (shallow trees), else the first
// wc.stack_depth levels in shared memory and the rest in global memory
template <int SCENE> struct WfStack {
    typedef StackStrided type;
    static __device__ __forceinline__ type make(int* stack_mem, int pid, uint32_t np, const WfConst&) { return type{stack_mem + pid, (int)np}; }
};
template <> struct WfStack<-1> {
    typedef StackSpill type;
    static __device__ __forceinline__ type make(int* stack_mem, int pid, uint32_t np, const WfConst& wc)
    {
        return type{stack_mem + pid, (int)np, (int)wc.stack_depth, wc.stack_over + (size_t)blockIdx.x * np + pid, (int)(gridDim.x * np)};
    }
};

__device__ __forceinline__ float4 pack(float3 v, float w) { return make_float4(v.x, v.y, v.z, w); }
__device__ __forceinline__ float3 xyz(float4 v) { return make_float3(v.x, v.y, v.z); }

// Candidate list of a pixel: the (at most 4) leaves whose triangle can be reached by ANY camera ray of the pixel.
// All sample rays of pixel (ix, iy) start at the eye and point into the pyramid spanned by the un-normalised
// directions d(dx, dy) = dx U + dy V + W (:732-737) at the pixel's four corners (jitter in [0, 1)).  A triangle whose
// three vertices lie outside one of the pyramid's four side planes, by more than a margin of 2e-4 of their distance
// (a quarter of a 1080p pixel; float error of the ray/triangle test is ~1e-6), cannot be hit.  Only a superset is
// needed: the list decides which triangles are TESTED, never what a test returns.
template <class FETCH>
__device__ __forceinline__ uint32_t pixel_candidates(const LaunchConst& lc, const FETCH& f, uint32_t ix, uint32_t iy)
{
    const float w = (float)(int)lc.width, h = (float)(int)lc.height;
    const float dx0 = 2.0f * (((float)ix + 0.0f) / w) - 1.0f, dx1 = 2.0f * (((float)ix + 1.0f) / w) - 1.0f;
    const float dy0 = 2.0f * (((float)iy + 0.0f) / h) - 1.0f, dy1 = 2.0f * (((float)iy + 1.0f) / h) - 1.0f;
    const float3 c00 = dx0 * lc.U + dy0 * lc.V + lc.W, c10 = dx1 * lc.U + dy0 * lc.V + lc.W;
    const float3 c11 = dx1 * lc.U + dy1 * lc.V + lc.W, c01 = dx0 * lc.U + dy1 * lc.V + lc.W;
    const float3 cc = (0.5f * (dx0 + dx1)) * lc.U + (0.5f * (dy0 + dy1)) * lc.V + lc.W;
    float3 n[4] = {cross(c00, c10), cross(c10, c11), cross(c11, c01), cross(c01, c00)};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        n[i] = normalize(n[i]);
        if (dot(n[i], cc) < 0.0f) n[i] = -n[i];             // inside the pyramid = positive side
    }
    uint32_t cand = kCandNone;
    int count = 0;
    for (uint32_t leaf = 0; leaf < lc.bvh.n_tris; ++leaf) {
        const float4 t0 = f.tri((int)leaf, 0), t1 = f.tri((int)leaf, 1), t2 = f.tri((int)leaf, 2);
        const float3 v0 = xyz(t0) - lc.eye, v1 = v0 + xyz(t1), v2 = v0 + xyz(t2);
        const float reach = fmaxf(fmaxf(fabsf(v0.x) + fabsf(v0.y) + fabsf(v0.z), fabsf(v1.x) + fabsf(v1.y) + fabsf(v1.z)),
                                  fabsf(v2.x) + fabsf(v2.y) + fabsf(v2.z));
        const float m = -2e-4f * reach;
        bool outside = false;
#pragma unroll
        for (int i = 0; i < 4; ++i) outside = outside || (dot(n[i], v0) < m && dot(n[i], v1) < m && dot(n[i], v2) < m);
        if (!outside) {
            if (count == 4 || leaf >= 0xfeu) return kCandOverflow;
            cand = (cand & ~(0xffu << (8 * count))) | (leaf << (8 * count));
            ++count;
        }
    }
    return cand;
}

// ---- Sample teams -----------------------------------------------------------------------------------------------------
// The reference runs a pixel's samples one after the other (pathTracerPrograms.cu:724-780) and so does a path slot here.
// That chain is the latency floor of a launch: a slot needs ~25 us per sample, 3.5 ms for 128 of them, whatever else the
// GPU is doing -- the tail of every launch, most of a launch that gives a GPU only a small share of the frame, and a
// pixel on the rim of the glass sphere (28-bounce paths) holds its slot for a large part of a C3 launch.
// The samples are independent, though: the pixel seed of sample k is the seed of sample 0 advanced by 2k draws
// (only the two jitter draws move it) and the LCG can jump there in O(log k) integer steps (lcg_skip).  What is
// ordered is only the float sum `result += emission; result += radiance * attenuation` (:760-761).  So, when a column
// has slots with nothing to do (no pixels left to take), a pixel can be handed to the column's TEAM:
//   * the team's state lives in thirteen shared-memory words of the column: how many samples it has and how many were
//     handed out (one word, one CAS), the commit frontier, the pixel's seed / coordinates / candidate list, the running
//     sum, and one "ended" bit per sample of the log window;
//   * every slot of the column that has no pixel takes the next sample index k, runs that path exactly like any other,
//     and instead of adding its (non-zero) terms to a result it appends them to the log of sample k in global memory
//     (adding +-0 never changes this sum: it starts at +0 and can never become -0);
//   * when its sample ends the slot closes the log with a terminator, sets the sample's bit, and whoever gets the team's
//     lock adds the logged terms of all ended samples at the commit frontier, IN SAMPLE ORDER, to the running sum -- the
//     same additions in the same order as the sequential loop, hence the same bits.  Nobody waits: the slot takes the
//     next sample at once; a long sample only delays the frontier, bounded by the log window (then slots park until it
//     moves);
//   * the slot that commits the last sample carries the pixel to the PIXEL stage, which writes it as usual.
// One team per column at a time; lanes only touch their own column's words, like the queues.  All of it is out of line
// (noinline functions that rebuild their view from the kernel parameters): a slot that works alone pays one flag test
// per stage visit and one shared-memory load per sample.
enum : int { T_TAKE = 0, T_COMMIT, T_LOCK, T_SEED, T_PIX, T_CAND, T_RX, T_RY, T_RZ, T_DONE, kTeamWords = T_DONE + 4 };
constexpr uint32_t kTeamWindowMax = 128;   // 4 words of "ended" bits
static_assert(kCtlF4 * 4 >= kQueues * 32 + 4 + 32 + kTeamWords * 32, "control block too small");
constexpr uint32_t kTeamSetup = 0x0000ffffu;        // T_TAKE while a team is being set up: no samples, nothing to take or commit
constexpr float kLogEnd = 1.0f;                     // .w of the entry that closes a sample's log (terms carry 0)

struct Team {              // this lane's column
    uint32_t* w;           // team word i at w[i * 32]:  T_TAKE = samples << 16 | handed out
    uint32_t* idle;        // bit j: slot 32 j + lane is parked (its TS already says MODE_JOIN)
    float4* log;           // [window][cap]
    uint32_t wmask, cap;
};

__device__ __forceinline__ Team team_view(const WfConst& wc, uint32_t* ctl, int lane)
{
    Team t;
    t.idle = ctl + 4 + lane;
    t.w = ctl + 4 + 32 + lane;
    t.wmask = wc.team_wmask;
    t.cap = wc.team_cap;
    t.log = wc.team_log + ((size_t)blockIdx.x * 32 + lane) * (wc.team_wmask + 1u) * wc.team_cap;
    return t;
}

// n draws ahead: s -> A_n s + C_n, by repeated squaring of the affine map s -> 1664525 s + 1013904223 (cuda/random.h:49-56)
__device__ __forceinline__ uint32_t lcg_skip(uint32_t s, uint32_t n)
{
    uint32_t A = 1u, C = 0u, a = 1664525u, c = 1013904223u;
    while (n) {
        if (n & 1u) {
            A = a * A;
            C = a * C + c;
        }
        c = (a + 1u) * c;
        a = a * a;
        n >>= 1;
    }
    return A * s + C;
}

// next sample of the column's team, or -1: no team, every sample handed out, or the log window is full
__device__ __forceinline__ int team_take(const Team& t)
{
    for (;;) {
        const uint32_t w = vol_ld(t.w + T_TAKE * 32), n = w >> 16, i = w & 0xffffu;
        if (i >= n || i > vol_ld(t.w + T_COMMIT * 32) + t.wmask) return -1;
        if (atomicCAS(t.w + T_TAKE * 32, w, w + 1u) == w) return (int)i;
    }
}

// parked slots of this column back to work (END stage, MODE_JOIN)
__device__ __forceinline__ void team_wake(const Team& t, uint32_t* mask, uint32_t* ctl, int lane)
{
    __threadfence_block();                            // "samples available, then look for parked slots" against team_park's
    if (vol_ld(t.idle) == 0u) return;                 // "parked, then look for samples": sequentially consistent on both sides
    const uint32_t m = atomicExch(t.idle, 0u);
    if (m) {
        atomicAdd(ctl, (uint32_t)__popc(m));          // before they become visible: the exit test must never see 0 too early
        __threadfence_block();
        atomicOr(mask + ST_END * 32 + lane, m);
    }
}

// Park slot j (its TS must already say MODE_JOIN).  False: samples turned up meanwhile and the slot was not woken by
// anyone else -- the caller still owns it and tries again.
__device__ __forceinline__ bool team_park(const Team& t, uint32_t* ctl, int j)
{
    __threadfence_block();
    atomicOr(t.idle, 1u << j);
    atomicSub(ctl, 1u);
    __threadfence_block();
    const uint32_t w = vol_ld(t.w + T_TAKE * 32), n = w >> 16, i = w & 0xffffu;
    if (i < n && i <= vol_ld(t.w + T_COMMIT * 32) + t.wmask) {
        if (atomicAnd(t.idle, ~(1u << j)) & (1u << j)) {
            atomicAdd(ctl, 1u);
            return false;
        }
    }
    return true;
}

// entry `at` of sample k's log
__device__ __noinline__ void team_log(const WfConst& wc, uint32_t* ctl, int lane, uint32_t k, uint32_t at, float3 v)
{
    const Team t = team_view(wc, ctl, lane);
    t.log[(size_t)(k & t.wmask) * t.cap + at] = make_float4(v.x, v.y, v.z, 0.0f);
}

// Commits whatever has ended at the frontier of the column's team.  True: this call committed the last sample of a pixel
// -- `result` is its sum, `pix` its coordinates, the team is closed and the caller carries the pixel to the PIXEL stage.
// (The lock holder works for whatever team is open: a slot whose own sample was committed by someone else may find the
// next team here and complete that one.)
__device__ __forceinline__ bool team_commit(const Team& t, float3& result, uint32_t& pix, uint32_t* mask, uint32_t* ctl, int lane)
{
    for (;;) {
        // "set my bit, then look at the lock" here against "unlock, then look at the frontier" in the holder: a store followed by a
        // load of another location on both sides.  Without a sequentially consistent fence between them both can read the old
        // value -- the holder leaves, this slot gives up, and if it was the team's last sample nobody ever commits it.
        __threadfence_block();
        // The log of the sample at the frontier is read BEFORE the lock is taken (it cannot change any more once its bit
        // is set): the global-memory latency stays out of the critical section, which is what bounds a team's commit rate.
        const uint32_t cmg = vol_ld(t.w + T_COMMIT * 32), slotg = cmg & t.wmask;
        const bool ahead = (vol_ld(t.w + (T_DONE + (slotg >> 5)) * 32) & (1u << (slotg & 31u))) != 0u;
        float4 g0 = make_float4(0.f, 0.f, 0.f, 0.f), g1 = g0, g2 = g0;
        if (ahead) {
            __threadfence_block();                                         // acquire: the log of a sample whose bit is set
            const float4* src = t.log + (size_t)slotg * t.cap;
            g0 = src[0], g1 = src[1], g2 = src[2];
        }
        if (atomicCAS(t.w + T_LOCK * 32, 0u, 1u) != 0u) return false;      // the holder re-checks the frontier after unlocking
        __threadfence_block();
        const uint32_t n = vol_ld(t.w + T_TAKE * 32) >> 16;
        if (n == 0u) {                                                     // no team open: mine has been completed
            atomicExch(t.w + T_LOCK * 32, 0u);
            return false;
        }
        const uint32_t cm0 = vol_ld(t.w + T_COMMIT * 32);
        uint32_t cm = cm0;
        float3 r = f3(__uint_as_float(vol_ld(t.w + T_RX * 32)), __uint_as_float(vol_ld(t.w + T_RY * 32)), __uint_as_float(vol_ld(t.w + T_RZ * 32)));
        while (cm < n) {
            const uint32_t slot = cm & t.wmask, bit = 1u << (slot & 31u);
            uint32_t* dw = t.w + (T_DONE + (slot >> 5)) * 32;
            if ((vol_ld(dw) & bit) == 0u) break;
            __threadfence_block();                                         // acquire: the log of a sample whose bit is set
            atomicAnd(dw, ~bit);                                           // the window slot is free again once T_COMMIT has passed it
            const float4* src = t.log + (size_t)slot * t.cap;
            const bool pre = ahead && cm == cmg;                           // the team (and its frontier) is the one looked at above
            const float4 e0 = pre ? g0 : src[0], e1 = pre ? g1 : src[1], e2 = pre ? g2 : src[2];   // most samples have at most two terms
            for (uint32_t i = 0;; ++i) {
                const float4 v = i == 0 ? e0 : (i == 1 ? e1 : (i == 2 ? e2 : src[i]));
                if (v.w != 0.0f) break;                                    // the terminator
                r = r + f3(v.x, v.y, v.z);                                 // :760-761, in sample order, term by term
            }
            ++cm;
        }
        volatile uint32_t* w = t.w;
        if (cm == n) {
            result = r;
            pix = w[T_PIX * 32];
            w[T_TAKE * 32] = 0u;                                           // the team is closed (all n samples were taken long ago)
            __threadfence_block();
            atomicExch(t.w + T_LOCK * 32, 0u);
            return true;
        }
        w[T_RX * 32] = __float_as_uint(r.x);
        w[T_RY * 32] = __float_as_uint(r.y);
        w[T_RZ * 32] = __float_as_uint(r.z);
        __threadfence_block();
        w[T_COMMIT * 32] = cm;
        __threadfence_block();
        atomicExch(t.w + T_LOCK * 32, 0u);
        __threadfence_block();                                             // (see above)
        if (cm != cm0) team_wake(t, mask, ctl, lane);                      // the window moved
        const uint32_t slot = cm & t.wmask;
        if ((vol_ld(t.w + (T_DONE + (slot >> 5)) * 32) & (1u << (slot & 31u))) == 0u) return false;
    }
}

// What a slot of a team does between two samples.  ended >= 0: sample `ended` of the column's team has just finished with
// n_terms logged terms -- close its log, publish it, commit what can be committed.  Then take the next sample.
// Everything it hands back goes through the slot's shared-memory record (no reference parameters: a variable whose
// address is passed to an out-of-line function lives in local memory for the whole kernel):
//   k >= 0  the slot runs sample k now: P1[pid] = (team bit, k, pixel seed BEFORE the sample's two jitter draws, pixel),
//           CAND[pid] = 0 terms logged so far; the pixel's candidate list is the team word T_CAND;
//   -1      nothing to take -- the slot is parked (TS / P1 say so) and must not be touched any more;
//   -2      this slot committed a pixel's last sample: P1 / RS hold the pixel and the slot is queued for the PIXEL stage.
__device__ __noinline__ int team_next(const WfConst& wc, float4* smem4, int lane, int j, int ended, uint32_t n_terms)
{
    uint32_t* mask = reinterpret_cast<uint32_t*>(smem4);
    uint32_t* ctl = mask + kQueues * 32;
    const uint32_t NP = wc.np;
    const int pid = j * 32 + lane;
    float4* TS = smem4 + kCtlF4 + 3 * NP;
    float4* P1 = smem4 + kCtlF4 + 5 * NP;
    float4* RS = smem4 + kCtlF4 + 6 * NP;
    uint32_t* CAND = reinterpret_cast<uint32_t*>(smem4 + kCtlF4 + kStateF4 * NP);
    const Team t = team_view(wc, ctl, lane);
    if (ended >= 0) {
        const uint32_t slot = (uint32_t)ended & t.wmask;
        t.log[(size_t)slot * t.cap + n_terms] = make_float4(0.0f, 0.0f, 0.0f, kLogEnd);
        __threadfence_block();                                             // release: the log before the bit (every reader is in this CTA)
        atomicOr(t.w + (T_DONE + (slot >> 5)) * 32, 1u << (slot & 31u));
        float3 result;
        uint32_t pix;
        if (team_commit(t, result, pix, mask, ctl, lane)) {
            P1[pid] = make_float4(__int_as_float(0), __uint_as_float(0u), __uint_as_float(0u), __uint_as_float(pix));
            RS[pid] = pack(result, 0.0f);
            wf_push(mask, ST_PIXEL, lane, j);
            return -2;
        }
    }
    for (;;) {
        const int k = team_take(t);
        if (k >= 0) {
            const uint32_t seed = lcg_skip(vol_ld(t.w + T_SEED * 32), 2u * (uint32_t)k);
            P1[pid] = make_float4(__int_as_float(kTeamBit), __uint_as_float((uint32_t)k), __uint_as_float(seed), __uint_as_float(vol_ld(t.w + T_PIX * 32)));
            CAND[pid] = 0u;
            return k;
        }
        TS[pid] = make_float4(0.0f, __uint_as_float(kNoPrim), __int_as_float(kTravDone), __int_as_float(MODE_JOIN << 8));
        P1[pid] = make_float4(__int_as_float(0), __uint_as_float(0u), __uint_as_float(0u), __uint_as_float(0u));
        if (team_park(t, ctl, j)) return -1;
    }
}

// A slot that works alone on a pixel (`left` samples to go, pixel seed before the next sample's jitter draws) hands it to
// the column's team.  False: a team is open already.  The claim is a CAS on T_TAKE itself (closed -> "being set up": no
// samples, so nobody takes or commits anything) and never touches T_LOCK: whoever holds T_LOCK is a committer that
// re-checks the frontier when it unlocks -- team_commit relies on that when it finds the lock taken.
__device__ __noinline__ bool team_create(const WfConst& wc, uint32_t* mask, uint32_t* ctl, int lane, uint32_t left, uint32_t pixel_seed,
                                         uint32_t pix, uint32_t cand, float3 result)
{
    const Team t = team_view(wc, ctl, lane);
    if (vol_ld(t.w + T_TAKE * 32) != 0u || atomicCAS(t.w + T_TAKE * 32, 0u, kTeamSetup) != 0u) return false;
    volatile uint32_t* w = t.w;
    w[T_COMMIT * 32] = 0u;
    w[T_SEED * 32] = pixel_seed;
    w[T_PIX * 32] = pix;
    w[T_CAND * 32] = cand;
    w[T_RX * 32] = __float_as_uint(result.x);
    w[T_RY * 32] = __float_as_uint(result.y);
    w[T_RZ * 32] = __float_as_uint(result.z);
    __threadfence_block();                                // the words above before the team opens
    w[T_TAKE * 32] = left << 16;
    team_wake(t, mask, ctl, lane);
    return true;
}

// MSORT: which SHADE queue a hit goes to -- the BSDF type of the primitive's material (the switch at :905-983)
template <bool MSORT, class FETCH>
__device__ __forceinline__ int shade_row_of(const FETCH& f, uint32_t prim)
{
    if (!MSORT) return ST_SHADE;
    const int b = __float_as_int(f.mat(__float_as_int(f.tri_shade(prim).w), 1).w);
    return b == 0 ? ST_SHADE : (b == 1 ? ST_SHADE1 : ST_SHADE2);
}

// ---- END stage, one slot: finish the segment (__miss__ms :833-847, shadow result, :760-778), Russian roulette, then the
// next ray / the next sample / the finished pixel.  Two instantiations:
//   ANY = false  every slot of the batch works alone on its pixel (or the lane holds none: j < 0).  Inlined into the kernel:
//                the hot path, with nothing of the teams in it.  Used until the work counter runs dry -- no slot can be
//                without a pixel before that, so no team can exist;
//   ANY = true   from then on ("team phase", a flag of the CTA): slots that serve their column's team, slots that have no
//                pixel and want to (MODE_JOIN) and slots still alone on a pixel, told apart per lane.  Out of line
//                (end_any_slot) and used by WHOLE warps, so that exactly one copy of the END code is live in either phase:
//                the first version sent only the team lanes of a batch out of line -- both copies were hot at once, the
//                hot code no longer fitted the instruction caches, and a launch in its team phase stalled on instruction
//                fetch 5.2 times per issued instruction (profiles/r2_icache_notes.txt).
// Returns the work counted (COUNT): misses | samples << 16 | radiance rays << 32 | triangle tests << 48.
template <bool COUNT, bool PROBE, bool SHORT, bool MSORT, bool TEAM, bool ANY, class FETCH>
__device__ __forceinline__ unsigned long long end_slot(const LaunchConst& lc, const WfConst& wc, float4* smem4, const FETCH& f, int lane, int j,
                                                       bool have_tris)
{
    const uint32_t NP = wc.np;
    uint32_t* mask = reinterpret_cast<uint32_t*>(smem4);
    uint32_t* ctl = mask + kQueues * 32;
    float4* R0 = smem4 + kCtlF4;
    float4* R1 = R0 + NP;
    float4* R2 = R1 + NP;
    float4* TS = R2 + NP;
    float4* P0 = TS + NP;
    float4* P1 = P0 + NP;
    float4* RS = P1 + NP;
    float4* S0 = RS + NP;
    float4* S1 = S0 + NP;
    uint32_t* CAND = reinterpret_cast<uint32_t*>(S1 + NP);
    uint32_t c_miss = 0u, c_samples = 0u, c_radiance = 0u, c_tri = 0u;

    const int pid = (j >= 0 ? j : 0) * 32 + lane;
    Path p;
    p.o = p.d = f3(0.0f);
    p.att = f3(1.0f);
    p.seed = 0u;
    p.depth = 0;
    // left: samples this slot still owes its pixel -- or, for a slot that serves its column's team, the sample index
    uint32_t left = 0u, pixel_seed = 0u, ix = 0u, iy = 0u;
    float3 result = f3(0.0f);
    bool alive = j >= 0, need_sample = false, team = false;
    if (j >= 0) {
        const float4 ts = TS[pid];
        const int mode = __float_as_int(ts.w) >> 8;
        const float4 p1 = P1[pid];
        left = __float_as_uint(p1.y);
        pixel_seed = __float_as_uint(p1.z);
        const uint32_t pix = __float_as_uint(p1.w);
        ix = pix & 0xffffu;
        iy = pix >> 16;
        int ended = -2;                              // team slots: -1 = take a sample, k >= 0 = sample k has just ended
        uint32_t n_terms = 0u;
        team = ANY && (__float_as_int(p1.x) & kTeamBit) != 0;            // this slot runs sample `left` of its column's team
        if (mode == MODE_NEWPIX) {
            need_sample = true;                      // first sample of a new pixel; result starts at 0
        } else if (ANY && mode == MODE_JOIN) {
            team = true;                             // a slot without a pixel: take a sample of the column's team below
            ended = -1;
        } else {
            const float4 p0 = P0[pid], rs = RS[pid];
            p.att = xyz(p0);
            p.seed = __float_as_uint(p0.w);
            const int dw = __float_as_int(p1.x);
            p.depth = dw & 0xff;
            bool done = ((dw >> 8) & 1) != 0;
            result = xyz(rs);
            float3 rad;
            if (mode == MODE_RAD) {                  // miss
                if (PROBE && left == lc.spp && p.depth == 0) {
                    const uint32_t image_index = iy * lc.width + ix;
                    if (lc.dbg_primary_prim) lc.dbg_primary_prim[image_index] = -1;
                    if (lc.dbg_primary_t) lc.dbg_primary_t[image_index] = 0.0f;
                }
                rad = lc.background;
                done = true;
                result = result + f3(0.0f);                                            // emissionColor = 0, :760
                if (COUNT) c_miss++;
            } else {
                const float4 s0 = S0[pid], s1 = S1[pid];
                p.o = xyz(s0);
                p.d = xyz(s1);
                rad = f3(0.0f);
                if (done) rad = xyz(f.mat((int)((uint32_t)dw >> 16), 1));               // emissive hit, :992-996
                if (mode == MODE_SHADOW && __float_as_uint(ts.y) == kNoPrim)
                    rad = rad + f3(s0.w, s1.w, rs.w);                                  // unoccluded, :1019-1024
            }
            const float3 term = rad * p.att;
            if (ANY && team) {                                                         // :761, into the sample's log
                n_terms = CAND[pid];
                if (term.x != 0.0f || term.y != 0.0f || term.z != 0.0f) team_log(wc, ctl, lane, left, n_terms++, term);
            } else {
                result = result + term;                                                // :761
            }
            const float pr = dot(p.att, f3(0.30f, 0.59f, 0.11f));                      // :763
            const bool rr = rnd(p.seed) > pr;                                          // :765
            if (done || rr || (uint32_t)p.depth >= lc.max_depth) {                     // :767
                if (ANY && team) {
                    ended = (int)left;               // the team's sample `left` has ended (published below)
                } else if (--left == 0u) {           // pixel finished: the PIXEL stage writes it
                    P0[pid] = pack(p.att, __uint_as_float(p.seed));
                    P1[pid] = make_float4(p1.x, __uint_as_float(0u), p1.z, p1.w);
                    RS[pid] = pack(result, 0.0f);
                    wf_push(mask, ST_PIXEL, lane, j);
                    alive = false;
                } else {
                    need_sample = true;
                    // slots of this column are parked and the pixel has samples to spare: hand it to the column's team and go
                    // on as one of its members
                    if (ANY && wc.team != 0 && vol_ld(ctl + 4 + lane) != 0u && left >= (uint32_t)wc.team_min && left < 0xffffu &&
                        (!SHORT || CAND[pid] != kCandNone) &&
                        team_create(wc, mask, ctl, lane, left, pixel_seed, pix, SHORT ? CAND[pid] : kCandOverflow, result)) {
                        team = true;
                        ended = -1;
                    }
                }
            } else {
                if (ANY && team) CAND[pid] = n_terms;
                p.att = f3(safe_divide(p.att.x, pr), safe_divide(p.att.y, pr), safe_divide(p.att.z, pr));   // :773
                ++p.depth;
            }
        }
        if (ANY && team && ended >= -1) {
            // between two samples of the column's team: publish the one that ended, commit, take the next one
            if (team_next(wc, smem4, lane, j, ended, n_terms) >= 0) {
                const float4 q1 = P1[pid];           // (team bit, sample, pixel seed before its jitter draws, pixel)
                left = __float_as_uint(q1.y);
                pixel_seed = __float_as_uint(q1.z);
                ix = __float_as_uint(q1.w) & 0xffffu;
                iy = __float_as_uint(q1.w) >> 16;
                result = f3(0.0f);
                need_sample = true;
            } else {
                need_sample = alive = false;         // parked (a waker pushes the slot back to END), or gone to PIXEL with a finished pixel
            }
        }
    }
    // "continue the path" and "next sample" took different branches above; reconverge before the part they share
    // (set up the next ray), or it runs once per branch at half occupancy
    __syncwarp();
    if (alive) {
        uint32_t cand = kCandOverflow;
        if (need_sample) {
            if (SHORT) cand = (ANY && team) ? vol_ld(ctl + 4 + 32 + T_CAND * 32 + lane) : CAND[pid];
            // The next sample's camera ray, in place.  SHORT: when no triangle can be reached from this pixel (empty
            // candidate list) every camera ray is a miss; up to wc.miss_loop samples are then finished right here, each
            // exactly as the miss branch above does (:833-847, :760-772): background times attenuation 1, the
            // Russian-roulette draw, next sample.  Background tiles are handed out together (tile order), so the lanes
            // of a batch take this loop together.  (One copy of the ray generation for both uses: instruction cache.)
            for (int rep = 0;; ++rep) {
                generate_camera_ray(lc, ix, iy, pixel_seed, p);
                if (COUNT) c_samples++;
                if (PROBE && left == lc.spp) {
                    const uint32_t image_index = iy * lc.width + ix;
                    if (lc.dbg_seed_after_jitter) lc.dbg_seed_after_jitter[image_index] = pixel_seed;
                    if (lc.dbg_primary_dir) {
                        lc.dbg_primary_dir[3 * image_index + 0] = p.d.x;
                        lc.dbg_primary_dir[3 * image_index + 1] = p.d.y;
                        lc.dbg_primary_dir[3 * image_index + 2] = p.d.z;
                    }
                }
                if (!(SHORT && !(ANY && team) && cand == kCandNone && rep < wc.miss_loop && left > 1u)) break;
                if (PROBE && left == lc.spp) {
                    const uint32_t image_index = iy * lc.width + ix;
                    if (lc.dbg_primary_prim) lc.dbg_primary_prim[image_index] = -1;
                    if (lc.dbg_primary_t) lc.dbg_primary_t[image_index] = 0.0f;
                }
                result = result + f3(0.0f);                                            // :760
                result = result + lc.background * p.att;                               // :761
                (void)rnd(p.seed);                                                     // :765 (done anyway)
                --left;
                if (COUNT) c_miss++, c_radiance++;
            }
        }
        if (COUNT) c_radiance++;
        P0[pid] = pack(p.att, __uint_as_float(p.seed));
        P1[pid] = make_float4(__int_as_float(p.depth | ((ANY && team) ? kTeamBit : 0)), __uint_as_float(left), __uint_as_float(pixel_seed),
                              __uint_as_float(ix | (iy << 16)));
        RS[pid] = pack(result, 0.0f);
        if (SHORT && cand != kCandOverflow) {
            // camera ray of a pixel with a candidate list: closest hit among <= 4 triangles, no traversal.  The
            // list holds every triangle any camera ray of the pixel can reach (pixel_candidates), and the
            // closest hit under (t, primIdx) does not depend on the order of the tests => same (t, prim).
            Ray r;
            r.o = p.o;
            r.d = p.d;
            r.inv = r.ood = f3(0.0f);
            float best_t = 1e16f;                                                      // tmax :755
            uint32_t best_prim = kNoPrim;
            for (uint32_t c = cand; (c & 0xffu) != 0xffu; c = (c >> 8) | 0xff000000u) {
                const int leaf = (int)(c & 0xffu);
                const float4 t0 = f.tri(leaf, 0), t1 = f.tri(leaf, 1), t2 = f.tri(leaf, 2);
                float th;
                if (COUNT) c_tri++;
                if (ray_tri(r, xyz(t0), xyz(t1), xyz(t2), 0.01f, 1e16f, th)) {         // tmin :754
                    const uint32_t prim = __float_as_uint(t0.w);
                    if (th < best_t || (th == best_t && prim < best_prim)) {
                        best_t = th;
                        best_prim = prim;
                    }
                }
            }
            R0[pid] = pack(r.o, 0.0f);
            R1[pid] = pack(r.d, 0.0f);
            TS[pid] = make_float4(best_t, __uint_as_float(best_prim), __int_as_float(kTravDone), __int_as_float(MODE_RAD << 8));
            wf_push(mask, best_prim != kNoPrim ? shade_row_of<MSORT>(f, best_prim) : ST_END, lane, j);
        } else {
            const Ray r = make_ray(p.o, p.d);                                          // traceRadiance :577-635
            R0[pid] = pack(r.o, 0.0f);
            R1[pid] = pack(r.d, 0.0f);
            R2[pid] = pack(r.inv, 0.0f);
            TS[pid] = make_float4(1e16f, __uint_as_float(kNoPrim), __int_as_float(have_tris ? 0 : kTravDone),
                                  __int_as_float(MODE_RAD << 8));                      // tmax :755
            wf_push(mask, have_tris ? ST_NODE : ST_END, lane, j);
        }
    }
    return COUNT ? ((unsigned long long)c_miss | ((unsigned long long)c_samples << 16) | ((unsigned long long)c_radiance << 32) |
                    ((unsigned long long)c_tri << 48))
                 : 0ull;
}

template <bool COUNT, bool PROBE, bool SHORT, bool MSORT, class FETCH>
__device__ __noinline__ unsigned long long end_any_slot(const LaunchConst& lc, const WfConst& wc, float4* smem4, const FETCH f, int lane, int j,
                                                        bool have_tris)
{
    return end_slot<COUNT, PROBE, SHORT, MSORT, true, true>(lc, wc, smem4, f, lane, j, have_tris);
}

// TRAV (scenes in global memory only; TRAV_DIRECT elsewhere): how NODE batches read the tree.
//   TRAV_DIRECT  every lane reads its own 64-byte node (two LDG.E.256)
//   TRAV_QUANT   quantised nodes: twelve 16-bit planes + two child references in 32 bytes, ONE LDG.E.256 per lane and node
//                (lbvh.cu quantize_nodes_kernel): half the L1 requests and half the node bytes of TRAV_DIRECT
//   TRAV_WIDE    4-wide quantised nodes: four child boxes + four references in 64 bytes (lbvh.cu build_wide_nodes), one step
//                resolves up to two binary levels: half the dependent fetches and half the pops / pushes / votes per ray
// (Round 1 also measured uncompressed two-level nodes and a warp-cooperative transposed fetch; both lost on B200 --
// profiles/r1_c3_c4_traversal_variants.txt -- and were removed.)
enum : int { TRAV_DIRECT = 0, TRAV_QUANT = 1, TRAV_WIDE = 2 };

template <int TRAV, class FETCH, class STK>
__device__ __forceinline__ void wf_node_step(Trav& t, const STK& stack, const FETCH& f, WideSel sel)
{
    if constexpr (TRAV == TRAV_WIDE) trav_wnode_step(t, stack, f, 0.01f, sel);
    else if constexpr (TRAV == TRAV_QUANT) trav_qnode_step(t, stack, f, 0.01f);
    else trav_node_step(t, stack, f, 0.01f);                                     // tmin :754
}

template <bool COUNT, bool PROBE, int SCENE, bool SHORT, int TRAV, bool MSORT, bool TEAM>
__global__ void __launch_bounds__(kWfBlock, 1)
render_wavefront_kernel(const __grid_constant__ LaunchConst lc, const __grid_constant__ WfConst wc)
{
    extern __shared__ float4 smem4[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t NP = wc.np;
    uint32_t* mask = reinterpret_cast<uint32_t*>(smem4);            // [7][32]
    uint32_t* ctl = mask + kQueues * 32;                            // [0] slots that are not parked  [1] error flag
    static_assert(!(TEAM && PROBE), "the probes record per-pixel sequential state");
    float4* R0 = smem4 + kCtlF4;        // (ray origin, -)
    float4* R1 = R0 + NP;               // (ray direction, -)
    float4* R2 = R1 + NP;               // (1 / direction, -)
    float4* TS = R2 + NP;               // traversal: (tfar, prim, cur, sp | mode << 8)
    float4* P0 = TS + NP;               // (attenuation, path seed)
    float4* P1 = P0 + NP;               // (depth | done << 8 | material << 16, samples left, pixel seed, ix | iy << 16)
    float4* RS = P1 + NP;               // (result, add.z)
    float4* S0 = RS + NP;               // (next ray origin, add.x)     add = NEE term, applied when the shadow ray is free
    float4* S1 = S0 + NP;               // (next ray direction, add.y)
    uint32_t* CAND = reinterpret_cast<uint32_t*>(S1 + NP);          // per pixel: up to 4 leaves a camera ray can hit
    int* stack_mem = reinterpret_cast<int*>(CAND + NP);             // [stack_depth][NP]
    float4* scene = reinterpret_cast<float4*>(stack_mem + (size_t)wc.stack_depth * NP);   // staged scene

    typename WfFetch<SCENE>::type f;
    f.nodes = lc.bvh.nodes;
    f.tris = lc.bvh.tris;
    f.shade = lc.tri_shade;
    f.mats = reinterpret_cast<const float4*>(lc.mats);
    if constexpr (SCENE < 0) {
        f.qnodes = lc.qnodes;
        f.q_lo = lc.q_lo;
        f.q_step = lc.q_step;
        f.prefetch = wc.prefetch != 0;
    }
    if (SCENE >= 0) {
        const uint32_t shift = SCENE >= 0 ? (uint32_t)SCENE : 0u, copies = 1u << shift;
        const uint32_t n_nodes4 = 4 * lc.n_nodes, n_tris4 = 3 * lc.bvh.n_tris, n_sh = lc.bvh.n_tris, n_m4 = 3 * lc.n_mats;
        float4* s_nodes = scene;
        float4* s_tris = s_nodes + ((size_t)n_nodes4 << shift);
        float4* s_shade = s_tris + ((size_t)n_tris4 << shift);
        float4* s_mats = s_shade + ((size_t)n_sh << shift);
        for (uint32_t i = threadIdx.x; i < (n_nodes4 << shift); i += blockDim.x) s_nodes[i] = __ldg(f.nodes + (i >> shift));
        for (uint32_t i = threadIdx.x; i < (n_tris4 << shift); i += blockDim.x) s_tris[i] = __ldg(f.tris + (i >> shift));
        for (uint32_t i = threadIdx.x; i < (n_sh << shift); i += blockDim.x) s_shade[i] = __ldg(f.shade + (i >> shift));
        for (uint32_t i = threadIdx.x; i < (n_m4 << shift); i += blockDim.x) s_mats[i] = __ldg(f.mats + (i >> shift));
        const uint32_t k = (uint32_t)lane & (copies - 1u);
        f.nodes = s_nodes + k;
        f.tris = s_tris + k;
        f.shade = s_shade + k;
        f.mats = s_mats + k;
    }

    // every path starts "fresh": waiting in PIXEL for its first pixel
    for (uint32_t i = threadIdx.x; i < kCtlF4 * 4; i += blockDim.x) mask[i] = 0u;
    for (uint32_t pid = threadIdx.x; pid < NP; pid += blockDim.x)
        TS[pid] = make_float4(0.0f, __uint_as_float(kNoPrim), __int_as_float(kTravDone), __int_as_float(MODE_FRESH << 8));
    __syncthreads();
    if (warp == 0) {
        const uint32_t P = NP >> 5;
        mask[ST_PIXEL * 32 + lane] = P >= 32u ? 0xffffffffu : ((1u << P) - 1u);
        if (lane == 0) ctl[0] = NP;
    }
    __syncthreads();

    const float l_area = length(cross(lc.l_v1, lc.l_v2));   // :1021
    const bool have_tris = lc.bvh.n_tris != 0;
    const uint32_t n_work = lc.n_owned_tiles * 32u;
    ThreadCounters ctr = {};
    uint32_t st_batches[4] = {0, 0, 0, 0}, st_lanes[4] = {0, 0, 0, 0}, st_steps = 0, st_step_lanes = 0, st_idle = 0;
    uint32_t it = 0, idle_polls = 0;
    int stage = -1, sticky = 0, row = 0;     // row: the queue the stage pops from (its own; SHADE has three when MSORT)

    for (;;) {
        // ---- vote ---------------------------------------------------------------------------------------------------
        // cols = columns with a waiting path = the lanes a batch of that stage would fill.
        // Sticky: the stage just run is repeated (up to wc.sticky times) while it can fill the whole warp -- one LDS and
        // one VOTE.ALL instead of the full vote, and the other stages' queues fill up meanwhile.
        uint32_t msnap = 0u, msnap2 = 0u;
        bool decided = false;
        if (stage >= 0 && stage != ST_TAIL && sticky < wc.sticky) {
            msnap = vol_ld(mask + (MSORT ? row : stage) * 32 + lane);
            if (__all_sync(kFull, msnap != 0u)) {
                ++sticky;
                decided = true;
            }
        }
        if (!decided) {
            const uint32_t m0 = vol_ld(mask + 0 * 32 + lane), m1 = vol_ld(mask + 1 * 32 + lane);
            const uint32_t m2 = vol_ld(mask + 2 * 32 + lane), m3 = vol_ld(mask + 3 * 32 + lane);
            const uint32_t m4 = vol_ld(mask + 4 * 32 + lane);
            const uint32_t b0 = __ballot_sync(kFull, m0 != 0u), b1 = __ballot_sync(kFull, m1 != 0u);
            uint32_t b2 = __ballot_sync(kFull, m2 != 0u);
            const uint32_t b3 = __ballot_sync(kFull, m3 != 0u);
            const uint32_t b4 = __ballot_sync(kFull, m4 != 0u);
            uint32_t mshade = m2;
            int shade_row = ST_SHADE;
            if (MSORT) {                                     // the fullest of the three SHADE queues stands for the stage
                const uint32_t m5 = vol_ld(mask + ST_SHADE1 * 32 + lane), m6 = vol_ld(mask + ST_SHADE2 * 32 + lane);
                const uint32_t b5 = __ballot_sync(kFull, m5 != 0u), b6 = __ballot_sync(kFull, m6 != 0u);
                if (__popc(b5) > __popc(b2)) b2 = b5, mshade = m5, shade_row = ST_SHADE1;
                if (__popc(b6) > __popc(b2)) b2 = b6, mshade = m6, shade_row = ST_SHADE2;
            }
            if ((b0 | b1 | b2 | b3 | b4) == 0u) {
                if (vol_ld(ctl) == 0u) break;                // every slot is parked: no pixel left to take, no team to serve
                if (++idle_polls > (1u << 26)) {             // watchdog: never spin forever
                    ctl[1] = 1u;
                    break;
                }
                if (COUNT) st_idle++;
                stage = -1;
                // back off: a warp that keeps finding nothing sleeps longer (up to 16x), so that its polls stop taking issue
                // slots from the warps that hold the paths
                __nanosleep((unsigned)wc.idle_ns << min(idle_polls - 1u, (uint32_t)wc.idle_shift));
                continue;
            }
            idle_polls = 0;
            sticky = 0;
            // the most downstream stage that can fill a warp (PIXEL > END > SHADE > LEAF > NODE): drains paths towards
            // completion and leaves the NODE backlog as the reservoir; else the stage with the most columns
            const uint32_t fullmask = wc.full_cols >= 32 ? 0xffffffffu : 0u;
            // Affinity: the warps of the first `affinity` schedulers (warp % 4) look at the traversal queues first, the others
            // at the shading queues first.  A scheduler's warps then share most of their instruction working set (each
            // scheduler has a 6 KB L0 instruction cache in front of the SM's 32 KB; the kernel's hot code is 43 KB) and the
            // traversal queues are drained before they are full.  Two of four schedulers and a "full" threshold of 30 columns:
            // C2 5457 -> 5766 Mpixel-samples/s; three of four for scenes in global memory: C3 2127 -> 2276; with all four
            // (a plain change of priorities) both lose.
            const bool trav_first = wc.affinity >= 0 ? (warp & 3) < wc.affinity : ((warp >> 2) & 3) < -wc.affinity;   // < 0: experiment, classes across schedulers
            if (trav_first && (fullmask ? (b1 == fullmask) : (__popc(b1) >= wc.full_cols))) {
                stage = ST_LEAF;
            } else if (trav_first && (fullmask ? (b0 == fullmask) : (__popc(b0) >= wc.full_cols))) {
                stage = ST_NODE;
            } else if (__popc(b4) >= 16) {                   // long and rare batches: half a warp is worth running
                stage = ST_PIXEL;
            } else if (fullmask ? (b3 == fullmask) : (__popc(b3) >= wc.full_cols)) {
                stage = ST_END;
            } else if (fullmask ? (b2 == fullmask) : (__popc(b2) >= wc.full_cols)) {
                stage = ST_SHADE;
            } else if (fullmask ? (b1 == fullmask) : (__popc(b1) >= wc.full_cols)) {
                stage = ST_LEAF;
            } else if (fullmask ? (b0 == fullmask) : (__popc(b0) >= wc.full_cols)) {
                stage = ST_NODE;
            } else {
                const int c0 = __popc(b0), c1 = __popc(b1), c2 = __popc(b2), c3 = __popc(b3), c4 = __popc(b4);
                int best = c0;
                stage = ST_NODE;
                if (c1 > best) stage = ST_LEAF, best = c1;
                if (c2 > best) stage = ST_SHADE, best = c2;
                if (c3 > best) stage = ST_END, best = c3;
                if (c4 > best) stage = ST_PIXEL, best = c4;
                // Tail of the launch (few paths left in this CTA): regrouping cannot fill warps any more and every queue
                // hop only adds latency to the last pixels, so traversals run to completion in one go.
                if (stage <= ST_LEAF && c0 + c1 + c2 + c3 + c4 <= wc.tail_cols) {
                    stage = ST_TAIL;
                    msnap = m0;
                    msnap2 = m1;
                }
            }
            if (stage != ST_TAIL)
                msnap = stage == ST_NODE ? m0 : (stage == ST_LEAF ? m1 : (stage == ST_SHADE ? mshade : (stage == ST_END ? m3 : m4)));
            row = stage == ST_SHADE ? shade_row : stage;
        }
        const int rot = (warp * 7 + (int)(it++)) & 31;

        if (stage == ST_TAIL) {
            // ---- TAIL: every lane takes a waiting traversal (at a node or at a leaf) and finishes it ----------------
            int j = wf_pop(mask + ST_NODE * 32 + lane, msnap, rot);
            if (j < 0) j = wf_pop(mask + ST_LEAF * 32 + lane, msnap2, rot);
            __syncwarp();
            const bool active = j >= 0;
            const int pid = (active ? j : 0) * 32 + lane;
            Trav t;
            int mode = 0;
            t.cur = kTravDone;
            t.sp = 0;
            t.tfar = 0.0f;
            t.prim = kNoPrim;
            t.r.o = t.r.d = t.r.inv = t.r.ood = f3(0.0f);
            if (active) {
                const float4 r0 = R0[pid], r1 = R1[pid], r2 = R2[pid], ts = TS[pid];
                t.r.o = xyz(r0);
                t.r.d = xyz(r1);
                t.r.inv = xyz(r2);
                t.r.ood = f3(r0.x * r2.x, r0.y * r2.y, r0.z * r2.z);          // make_ray: o * inv
                t.tfar = ts.x;
                t.prim = __float_as_uint(ts.y);
                t.cur = __float_as_int(ts.z);
                const int spm = __float_as_int(ts.w);
                t.sp = spm & 0xff;
                mode = spm >> 8;
            }
            const typename WfStack<SCENE>::type stack = WfStack<SCENE>::make(stack_mem, pid, NP, wc);
            const WideSel sel = wide_selectors(t.r.inv);
            for (;;) {
                const bool at_node = t.cur >= 0;
                const bool at_leaf = t.cur < 0 && t.cur != kTravDone;
                const uint32_t bn = __ballot_sync(kFull, at_node), bl = __ballot_sync(kFull, at_leaf);
                if ((bn | bl) == 0u) break;
                if (__popc(bn) >= __popc(bl)) {
                    if (at_node) {
                        if (COUNT) ctr.n_node++;
                        wf_node_step<TRAV>(t, stack, f, sel);
                    }
                } else if (at_leaf) {
                    if (COUNT) ctr.n_tri++;
                    trav_leaf_step(t, stack, f, mode == MODE_SHADOW, 0.01f, 1e16f);
                }
            }
            if (active) {
                TS[pid] = make_float4(t.tfar, __uint_as_float(t.prim), __int_as_float(kTravDone), __int_as_float(t.sp | (mode << 8)));
                wf_push(mask, (mode == MODE_RAD && t.prim != kNoPrim) ? shade_row_of<MSORT>(f, t.prim) : ST_END, lane, j);
            }
            stage = -1;
        } else if (stage == ST_NODE) {
            // ---- NODE: up to K slab-test steps for every lane that holds a path at an internal node ----------------
            const int j = wf_pop_converged(mask + ST_NODE * 32 + lane, msnap, rot);
            const bool active = j >= 0;
            const int pid = (active ? j : 0) * 32 + lane;
            Trav t;
            int mode = 0;
            t.cur = kTravDone;
            t.sp = 0;
            t.tfar = 0.0f;
            t.prim = kNoPrim;
            t.r.inv = t.r.ood = f3(0.0f);
            if (active) {
                const float4 r0 = R0[pid], r2 = R2[pid], ts = TS[pid];
                t.r.inv = xyz(r2);
                t.r.ood = f3(r0.x * r2.x, r0.y * r2.y, r0.z * r2.z);          // make_ray: o * inv
                t.tfar = ts.x;
                t.prim = __float_as_uint(ts.y);
                t.cur = __float_as_int(ts.z);
                const int spm = __float_as_int(ts.w);
                t.sp = spm & 0xff;
                mode = spm >> 8;
            }
            const typename WfStack<SCENE>::type stack = WfStack<SCENE>::make(stack_mem, pid, NP, wc);
            const int n0 = __popc(__ballot_sync(kFull, active));
            const WideSel sel = wide_selectors(t.r.inv);
            if (COUNT) st_batches[ST_NODE]++, st_lanes[ST_NODE] += n0;
            for (int step = 0; step < wc.node_steps; ++step) {
                const bool at = t.cur >= 0;
                const int n = __popc(__ballot_sync(kFull, at));
                if (n == 0 || n * 8 < n0 * wc.node_keep8) break;
                if (COUNT) st_steps++, st_step_lanes += n;
                if (at) {
                    if (COUNT) ctr.n_node++;
                    wf_node_step<TRAV>(t, stack, f, sel);
                }
                if (TRAV != TRAV_WIDE && t.cur >= 0) {                           // second step under the same vote
                    if (COUNT) ctr.n_node++;
                    wf_node_step<TRAV>(t, stack, f, sel);
                }
            }
            if (active) {
                TS[pid] = make_float4(t.tfar, __uint_as_float(t.prim), __int_as_float(t.cur), __int_as_float(t.sp | (mode << 8)));
                const int next = t.cur >= 0 ? ST_NODE
                                            : (t.cur != kTravDone ? ST_LEAF
                                                                  : ((mode == MODE_RAD && t.prim != kNoPrim) ? shade_row_of<MSORT>(f, t.prim) : ST_END));
                wf_push(mask, next, lane, j);
            }
        } else if (stage == ST_LEAF) {
            // ---- LEAF: one triangle test; a shadow ray ends on its first hit (traceOcclusion :651-684) ------------
            const int j = wf_pop_converged(mask + ST_LEAF * 32 + lane, msnap, rot);
            if (COUNT) st_batches[ST_LEAF]++, st_lanes[ST_LEAF] += __popc(__ballot_sync(kFull, j >= 0));
            if (j >= 0) {
                const int pid = j * 32 + lane;
                const float4 r0 = R0[pid], r1 = R1[pid], ts = TS[pid];
                Trav t;
                t.r.o = xyz(r0);
                t.r.d = xyz(r1);
                t.tfar = ts.x;
                t.prim = __float_as_uint(ts.y);
                t.cur = __float_as_int(ts.z);
                const int spm = __float_as_int(ts.w);
                t.sp = spm & 0xff;
                const int mode = spm >> 8;
                const typename WfStack<SCENE>::type stack = WfStack<SCENE>::make(stack_mem, pid, NP, wc);
                if (COUNT) ctr.n_tri++;
                trav_leaf_step(t, stack, f, mode == MODE_SHADOW, 0.01f, 1e16f);
                TS[pid] = make_float4(t.tfar, __uint_as_float(t.prim), __int_as_float(t.cur), __int_as_float(t.sp | (mode << 8)));
                const int next = t.cur >= 0 ? ST_NODE
                                            : (t.cur != kTravDone ? ST_LEAF
                                                                  : ((mode == MODE_RAD && t.prim != kNoPrim) ? shade_row_of<MSORT>(f, t.prim) : ST_END));
                wf_push(mask, next, lane, j);
            }
        } else if (stage == ST_SHADE) {
            // ---- SHADE: __closesthit__diffuse__ch :866-1031 up to the occlusion trace -------------------------------
            const int j = wf_pop_converged(mask + (MSORT ? row : (int)ST_SHADE) * 32 + lane, msnap, rot);
            if (COUNT) st_batches[ST_SHADE]++, st_lanes[ST_SHADE] += __popc(__ballot_sync(kFull, j >= 0));
            if (j >= 0) {
                const int pid = j * 32 + lane;
                const float4 r0 = R0[pid], r1 = R1[pid], ts = TS[pid], p0 = P0[pid], p1 = P1[pid], rs = RS[pid];
                Path p;
                p.o = xyz(r0);
                p.d = xyz(r1);
                p.att = xyz(p0);
                p.seed = __float_as_uint(p0.w);
                p.depth = __float_as_int(p1.x) & 0xff;
                const int team_bit = TEAM ? (__float_as_int(p1.x) & kTeamBit) : 0;
                const float t_hit = ts.x;
                const uint32_t prim = __float_as_uint(ts.y);
                if (PROBE && __float_as_uint(p1.y) == lc.spp && p.depth == 0) {
                    const uint32_t pix = __float_as_uint(p1.w);
                    const uint32_t image_index = (pix >> 16) * lc.width + (pix & 0xffffu);
                    if (lc.dbg_primary_prim) lc.dbg_primary_prim[image_index] = (int32_t)prim;
                    if (lc.dbg_primary_t) lc.dbg_primary_t[image_index] = t_hit;
                }
                ShadowQuery sq;
                int hit_type, mat_id;
                float3 emission, rad;
                bool done;
                shade_hit(lc, f, l_area, prim, t_hit, p, emission, rad, done, sq, hit_type, mat_id);
                float3 result = xyz(rs);
                if (TEAM && __builtin_expect(team_bit != 0, 0)) {                               // :760, into the sample's log
                    if (emission.x != 0.0f || emission.y != 0.0f || emission.z != 0.0f) {
                        const uint32_t n_terms = CAND[pid];
                        team_log(wc, ctl, lane, __float_as_uint(p1.y), n_terms, emission);
                        CAND[pid] = n_terms + 1u;
                    }
                } else {
                    result = result + emission;                                                 // :760
                }
                if (COUNT) {
                    ctr.n_hit0 += hit_type == 0;
                    ctr.n_hit1 += hit_type == 1;
                    ctr.n_hit2 += hit_type == 2;
                }
                // rad is the material's emission when it is emissive (done), else 0: END re-reads it from the table
                P0[pid] = pack(p.att, __uint_as_float(p.seed));
                P1[pid] = make_float4(__int_as_float(p.depth | ((done ? 1 : 0) << 8) | team_bit | (mat_id << 16)), p1.y, p1.z, p1.w);
                const float3 add = sq.want ? sq.add : f3(0.0f);
                S0[pid] = pack(p.o, add.x);
                S1[pid] = pack(p.d, add.y);
                RS[pid] = pack(result, add.z);
                if (sq.want) {
                    if (COUNT) ctr.n_shadow++;
                    const Ray r = make_ray(sq.org, sq.dir);
                    R0[pid] = pack(r.o, 0.0f);
                    R1[pid] = pack(r.d, 0.0f);
                    R2[pid] = pack(r.inv, 0.0f);
                    TS[pid] = make_float4(sq.tmax, __uint_as_float(kNoPrim), __int_as_float(0), __int_as_float(MODE_SHADOW << 8));
                    wf_push(mask, ST_NODE, lane, j);
                } else {
                    TS[pid] = make_float4(0.0f, __uint_as_float(kNoPrim), __int_as_float(kTravDone), __int_as_float(MODE_SHADED << 8));
                    wf_push(mask, ST_END, lane, j);
                }
            }
        } else if (stage == ST_PIXEL) {
            // ---- PIXEL: write a finished pixel (:784-814), take the next one from the work counter, build its list of
            // triangles a camera ray can reach.  Rare per sample (once per samplesPerPixel) and long, so it has its own
            // queue instead of running for one or two lanes inside END.
            const int j = wf_pop_converged(mask + ST_PIXEL * 32 + lane, msnap, rot);
            if (j >= 0) {
                const int pid = j * 32 + lane;
                const int mode = __float_as_int(TS[pid].w) >> 8;
                if (mode != MODE_FRESH) {
                    const float4 p0 = P0[pid], p1 = P1[pid], rs = RS[pid];
                    const uint32_t pix = __float_as_uint(p1.w);
                    const uint32_t image_index = (pix >> 16) * lc.width + (pix & 0xffffu);
                    write_pixel(lc, image_index, xyz(rs));
                    if (PROBE) {
                        if (lc.dbg_final_path_seed) lc.dbg_final_path_seed[image_index] = __float_as_uint(p0.w);
                        if (lc.dbg_final_pixel_seed) lc.dbg_final_pixel_seed[image_index] = __float_as_uint(p1.z);
                    }
                }
                uint32_t ix = 0u, iy = 0u;
                if (take_pixel(lc, n_work, atomicAdd(lc.work_counter, 1u), ix, iy)) {
                    const uint32_t pixel_seed = tea4(iy * lc.width + ix, lc.frame);            // :721
                    CAND[pid] = SHORT ? pixel_candidates(lc, f, ix, iy) : kCandOverflow;
                    P1[pid] = make_float4(__int_as_float(0), __uint_as_float(lc.spp), __uint_as_float(pixel_seed),
                                          __uint_as_float(ix | (iy << 16)));
                    TS[pid] = make_float4(0.0f, __uint_as_float(kNoPrim), __int_as_float(kTravDone), __int_as_float(MODE_NEWPIX << 8));
                    wf_push(mask, ST_END, lane, j);
                } else if (TEAM) {
                    // no pixel left: serve the column's team, or park (END, MODE_JOIN); the CTA is in its team phase from now on
                    *reinterpret_cast<volatile uint32_t*>(ctl + 2) = 1u;
                    P1[pid] = make_float4(__int_as_float(0), __uint_as_float(0u), __uint_as_float(0u), __uint_as_float(0u));
                    TS[pid] = make_float4(0.0f, __uint_as_float(kNoPrim), __int_as_float(kTravDone), __int_as_float(MODE_JOIN << 8));
                    wf_push(mask, ST_END, lane, j);
                } else {
                    atomicSub(ctl, 1u);                       // this slot is done for the launch
                }
            }
        } else {
            // ---- END: finish the segment, Russian roulette, next ray / sample / pixel (end_slot above) ------------------
            const int j = wf_pop_converged(mask + ST_END * 32 + lane, msnap, rot);
            if (COUNT) st_batches[ST_END]++, st_lanes[ST_END] += __popc(__ballot_sync(kFull, j >= 0));
            // once the work counter has run dry (ctl[2]) slots without a pixel and teams can exist: the whole warp takes the
            // out-of-line END that knows about them (see end_slot)
            unsigned long long counted;
            if (TEAM && __any_sync(kFull, vol_ld(ctl + 2) != 0u)) counted = end_any_slot<COUNT, PROBE, SHORT, MSORT>(lc, wc, smem4, f, lane, j, have_tris);
            else counted = end_slot<COUNT, PROBE, SHORT, MSORT, TEAM, false>(lc, wc, smem4, f, lane, j, have_tris);
            if (COUNT) {
                ctr.n_miss += (uint32_t)(counted & 0xffffu);
                ctr.n_samples += (uint32_t)((counted >> 16) & 0xffffu);
                ctr.n_radiance += (uint32_t)((counted >> 32) & 0xffffu);
                ctr.n_tri += (uint32_t)(counted >> 48);
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
        if (lane == 0) {
            for (int i = 0; i < 4; ++i) {
                atomicAdd(lc.counters + 16 + i, (unsigned long long)st_batches[i]);
                atomicAdd(lc.counters + 20 + i, (unsigned long long)st_lanes[i]);
            }
            atomicAdd(lc.counters + 24, (unsigned long long)st_steps);
            atomicAdd(lc.counters + 25, (unsigned long long)st_step_lanes);
            atomicAdd(lc.counters + 26, (unsigned long long)st_idle);
        }
    }
    // a team that is still open when every slot has parked would be a lost pixel: never expected, reported like the watchdog
    // (tune "team=2" also leaves its words in d_counters[27..30], pt_scene_get_sched_stats 11..14)
    if (TEAM && wc.team != 0 && warp == 0) {
        const Team t = team_view(wc, ctl, lane);
        if (vol_ld(t.w + T_TAKE * 32) != 0u) {
            ctl[1] = 1u;
            if (wc.team > 1) {
                atomicAdd(lc.error_flag + 12, 1ull);
                lc.error_flag[13] = vol_ld(t.w + T_TAKE * 32);
                lc.error_flag[14] = vol_ld(t.w + T_COMMIT * 32);
                lc.error_flag[15] = vol_ld(t.w + T_DONE * 32) | ((unsigned long long)vol_ld(t.w + T_LOCK * 32) << 32) | ((unsigned long long)vol_ld(t.idle) << 33);
            }
        }
    }
    __syncwarp();
    if (threadIdx.x == 0 && vol_ld(ctl + 1)) atomicAdd(lc.error_flag, 1ull);   // watchdog fired
}

template <bool COUNT, bool PROBE, int SCENE, bool SHORT, int TRAV, bool MSORT = false, bool TEAM = false>
cudaError_t launch_wavefront(Scene& s, const LaunchConst& lc, const WfConst& wc, size_t smem_bytes, cudaStream_t stream)
{
    auto kernel = render_wavefront_kernel<COUNT, PROBE, SCENE, SHORT, TRAV, MSORT, TEAM>;
    cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes);
    if (e != cudaSuccess) return e;
    int dev = 0, sms = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const uint64_t n_work = (uint64_t)lc.n_owned_tiles * 32u;
    uint32_t grid = (uint32_t)sms;                                       // persistent: one CTA per SM
    // fewer CTAs when the launch is small: one pixel per path slot -- or, with sample teams, down to 64 pixels per CTA as
    // long as every slot still gets about four samples (slots without a pixel serve their column's team)
    uint64_t grid_needed = (n_work + wc.np - 1) / wc.np;
    if (wc.team) grid_needed = std::max<uint64_t>(grid_needed, std::min<uint64_t>(n_work / 64u, n_work * lc.spp / (4ull * wc.np)));
    if (grid > grid_needed) grid = (uint32_t)grid_needed;
    if (grid == 0) grid = 1;
    e = cudaMemsetAsync(s.d_work_counter, 0, sizeof(uint32_t), stream);
    if (e != cudaSuccess) return e;
    // threads: at most 3/4 of the paths when the scene is on chip, so that every column keeps waiting paths to form
    // batches from; a scene read from global memory is latency-bound (ncu: long-scoreboard stalls, 33 % issue), there
    // every extra warp in flight pays more than a fuller batch (C4: 957 -> 1090 Mpixel-samples/s)
    uint32_t block = SCENE >= 0 ? ((wc.np * 3u / 4u) & ~31u) : (wc.np & ~31u);
    if (block > (uint32_t)kWfBlock) block = kWfBlock;
    if (block < 128u) block = 128u;
    if (wc.block_override >= 32 && wc.block_override <= kWfBlock && wc.block_override % 32 == 0) block = (uint32_t)wc.block_override;
    kernel<<<grid, block, smem_bytes, stream>>>(lc, wc);
    s.last_grid = grid;
    s.last_block = block;
    s.last_smem = (uint32_t)smem_bytes;
    s.last_kernels = 1;
    return cudaGetLastError();
}

}  // namespace

// How NODE batches read a scene that stays in global memory (pt_scene_set_traversal; PTB_TUNE quant= overrides).
// Measured on B200: the 4-wide quantised nodes win from a few thousand triangles on (+5 % at 29 K, +8 % at
// 180 K and 500 K, +12 % at 2 M triangles: C4) and tie with per-lane 256-bit loads of the exact 64-byte nodes on the smallest
// scenes that are not staged (C3, 2592 triangles), which keep the exact nodes.  The boundary is the one of the trees: above
// 8192 triangles (PLOC tree) the wide nodes.  All variants are bit-identical in their pixels.
static int choose_trav(const Scene& s, const WfTune& tune)
{
    // quantised nodes: not for launches that count the SPECIFIED work (their counters are those of the exact boxes)
    const bool wide = tune.quant >= 0 ? tune.quant == 2 : (s.traversal == 4 || (s.traversal == 0 && s.n_tris > 8192u));
    const bool quant = tune.quant >= 0 ? tune.quant != 0 : s.traversal == 3;
    return ((!s.count_enabled || s.count_executed) && s.n_tris >= 2 && (quant || wide)) ? (wide ? TRAV_WIDE : TRAV_QUANT) : TRAV_DIRECT;
}

// PTB_TUNE="node_votes=6,tail_cols=40,..." -- experiment overrides of WfTune from the environment, parsed once per
// process; pt_scene_set_tune() applies the same "key=value,..." syntax to one scene.
struct TuneKey {
    const char* name;
    int WfTune::*field;
};
static const TuneKey kTuneKeys[] = {{"node_votes_g", &WfTune::node_votes_g}, {"full_cols_g", &WfTune::full_cols_g},
                                    {"node_votes", &WfTune::node_votes}, {"node_keep8", &WfTune::node_keep8},
                                    {"full_cols", &WfTune::full_cols},   {"sticky", &WfTune::sticky},
                                    {"tail_cols", &WfTune::tail_cols},   {"miss_loop", &WfTune::miss_loop},
                                    {"paths32", &WfTune::paths32},       {"block", &WfTune::block},
                                    {"smem_kb", &WfTune::smem_kb},       {"shortlist", &WfTune::shortlist},
                                    {"quant", &WfTune::quant},           {"idle_ns", &WfTune::idle_ns},
                                    {"idle_shift", &WfTune::idle_shift}, {"stack_smem", &WfTune::stack_smem},
                                    {"team", &WfTune::team},             {"team_min", &WfTune::team_min},
                                    {"msort", &WfTune::msort},           {"affinity", &WfTune::affinity},
                                    {"affinity_g", &WfTune::affinity_g}, {"prefetch", &WfTune::prefetch}};

bool wf_tune_apply(WfTune& t, const char* spec)
{
    if (!spec) return true;
    bool ok = true;
    const std::string str(spec);
    size_t pos = 0;
    while (pos < str.size()) {
        size_t end = str.find(',', pos);
        if (end == std::string::npos) end = str.size();
        const std::string item = str.substr(pos, end - pos);
        const size_t eq = item.find('=');
        bool found = item.empty();
        if (eq != std::string::npos)
            for (const TuneKey& k : kTuneKeys)
                if (item.compare(0, eq, k.name) == 0 && eq == strlen(k.name)) t.*(k.field) = atoi(item.c_str() + eq + 1), found = true;
        ok = ok && found;
        pos = end + 1;
    }
    // "node_votes=" / "full_cols=" without their "_g" twins apply to every scene (sweeps)
    if (str.find("node_votes=") != std::string::npos && str.find("node_votes_g=") == std::string::npos) t.node_votes_g = t.node_votes;
    if (str.find("full_cols=") != std::string::npos && str.find("full_cols_g=") == std::string::npos) t.full_cols_g = t.full_cols;
    if (str.find("affinity=") != std::string::npos && str.find("affinity_g=") == std::string::npos) t.affinity_g = t.affinity;
    return ok;
}

// scene settings + environment overrides, clamped to what the kernel can run with
static WfTune tuned(const WfTune& base)
{
    static const std::string env = [] {
        const char* e = getenv("PTB_TUNE");
        return std::string(e ? e : "");
    }();
    WfTune t = base;
    if (!env.empty()) wf_tune_apply(t, env.c_str());
    auto clamp = [](int& v, int lo, int hi) { v = v < lo ? lo : (v > hi ? hi : v); };
    clamp(t.node_votes, 1, 64);
    clamp(t.node_votes_g, 1, 64);
    clamp(t.node_keep8, 0, 8);
    clamp(t.full_cols, 1, 32);
    clamp(t.full_cols_g, 1, 32);
    clamp(t.sticky, 0, 64);
    clamp(t.tail_cols, 0, 160);
    clamp(t.miss_loop, 0, 4096);
    clamp(t.idle_ns, 0, 1 << 16);
    clamp(t.idle_shift, 0, 10);
    return t;
}

// Can the wavefront pipeline hold this scene?  Fills the launch layout when it can.
static bool wavefront_plan(const Scene& s, const WfTune& tune, WfConst& wc, size_t& smem_bytes, bool& smem_scene, int& trav)
{
    if (s.n_mats > 0xffffu) return false;                 // material id is packed into 16 bits of the path record
    const uint32_t depth = scene_depth(s) + 1u;           // deferred children never exceed the tree depth
    if (depth > 255u) return false;
    trav = TRAV_DIRECT;
    const bool force_global = s.traversal >= 2;
    const size_t scene_f4 = 4 * (size_t)s.n_nodes + 4 * (size_t)s.n_tris + 3 * (size_t)s.n_mats;
    const size_t fixed = (size_t)kCtlF4 * 16;
    const size_t per_path = (size_t)kStateF4 * 16 + 4 + 4 * (size_t)depth;   // state, candidate word, stack
    // scene copies: 8 if the paths still get >= 512 slots, else 1; larger scenes stay in global memory
    for (int shift = 3; shift >= 0; shift -= 3) {
        const size_t scene_bytes = (scene_f4 * 16) << shift;
        if (s.n_tris == 0 || force_global || scene_bytes + fixed + 512 * per_path > kMaxSmem) continue;
        uint32_t P = (uint32_t)((kMaxSmem - fixed - scene_bytes) / (32 * per_path));
        if (P > 32) P = 32;
        if (tune.paths32 >= 1 && (uint32_t)tune.paths32 < P) P = (uint32_t)tune.paths32;
        wc.np = 32 * P;
        wc.stack_depth = depth;
        wc.stack_total = depth;
        wc.copies_shift = (uint32_t)shift;
        smem_bytes = fixed + wc.np * per_path + scene_bytes;
        smem_scene = true;
        return true;
    }
    // global-memory scene: node/triangle records come through L1/L2; the paths take most of the shared memory
    // (batch occupancy needs many more paths than threads) and the rest stays L1
    trav = choose_trav(s, tune);
    // shared-memory part of the stack: tune.stack_smem levels (deeper ones spill to global memory, StackSpill)
    const uint32_t smem_levels = (tune.stack_smem >= 2 && (uint32_t)tune.stack_smem < depth) ? (uint32_t)tune.stack_smem : depth;
    const size_t per_path_g = (size_t)kStateF4 * 16 + 4 + 4 * (size_t)smem_levels;
    const size_t budget = (size_t)(tune.smem_kb >= 32 && tune.smem_kb <= 227 ? tune.smem_kb : 200) * 1024;
    uint32_t P = (uint32_t)((budget - fixed) / (32 * per_path_g));
    if (P > 32) P = 32;
    if (tune.paths32 >= 1 && (uint32_t)tune.paths32 < P) P = (uint32_t)tune.paths32;
    if (P < 4 && tune.paths32 < 1) return false;
    wc.np = 32 * P;
    wc.stack_depth = smem_levels;
    wc.stack_total = depth;
    wc.copies_shift = 0;
    smem_bytes = fixed + wc.np * per_path_g;
    smem_scene = false;
    return true;
}

bool wavefront_supported(const Scene& s)
{
    WfConst wc;
    size_t b;
    bool ss;
    int trav;
    return wavefront_plan(s, tuned(s.wf), wc, b, ss, trav);
}

cudaError_t wavefront_launch(Scene& s, const LaunchConst& lc, bool probe, cudaStream_t stream)
{
    const WfTune tune = tuned(s.wf);
    WfConst wc = {};
    size_t smem_bytes = 0;
    bool smem_scene = false;
    int trav = TRAV_DIRECT;
    if (!wavefront_plan(s, tune, wc, smem_bytes, smem_scene, trav)) return cudaErrorInvalidConfiguration;
    wc.node_steps = smem_scene ? tune.node_votes : tune.node_votes_g;
    wc.node_keep8 = tune.node_keep8;
    wc.full_cols = smem_scene ? tune.full_cols : tune.full_cols_g;
    wc.sticky = tune.sticky;
    wc.tail_cols = tune.tail_cols;
    wc.miss_loop = tune.miss_loop;
    wc.block_override = tune.block;
    wc.idle_ns = tune.idle_ns;
    wc.idle_shift = tune.idle_shift;
    wc.affinity = smem_scene ? tune.affinity : tune.affinity_g;
    wc.prefetch = tune.prefetch;
    wc.stack_over = nullptr;
    if (!smem_scene && trav == TRAV_WIDE) {
        // a wide step defers up to three children: the stack is sized from the wide tree's own depth
        cudaError_t we = build_wide_nodes(s, stream);
        if (we == cudaErrorMemoryAllocation) {           // no room for the build's scratch: the binary quantised nodes will do
            cudaGetLastError();
            we = cudaSuccess;
        }
        if (we != cudaSuccess) return we;
        if (s.d_wnodes && s.wnodes_tree == scene_tree(s) && 3u * s.wide_depth + 1u <= 255u) wc.stack_total = 3u * s.wide_depth + 1u;
        else trav = TRAV_QUANT;
    }
    if (!smem_scene && wc.stack_total > wc.stack_depth) {
        int dev = 0, sms = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        const size_t need = (size_t)(wc.stack_total - wc.stack_depth) * wc.np * (size_t)sms;
        if (s.stack_over_capacity < need) {
            if (s.d_stack_over) cudaFree(s.d_stack_over);
            s.d_stack_over = nullptr;
            s.stack_over_capacity = 0;
            if (cudaMalloc((void**)&s.d_stack_over, need * sizeof(int)) != cudaSuccess) return cudaErrorMemoryAllocation;
            s.stack_over_capacity = need;
        }
        wc.stack_over = s.d_stack_over;
    }
    const bool count = s.count_enabled;
    const int scene = smem_scene ? (int)wc.copies_shift : -1;
    // Sample teams (tune "team": 0 off, 1 auto, 2 on, 3 on + debug dump).  Auto = scenes read from global memory: there a
    // launch is latency-bound, pixels differ widely in cost and the teams are worth +28 % (C3); a scene staged in shared
    // memory runs at its instruction-fetch limit and the kernel WITHOUT the team code is the faster one on full frames
    // (C2: the hot code exceeds the 32 KB L1.5 instruction cache either way; measurements in DESIGN.md 4.2b).
    // ... and within those, launches whose pixels differ widely in cost (more than one BSDF type, or deep paths) or that have
    // fewer than four pixels per path slot (small shares of a frame: the tail is most of the launch).  A full frame of uniform
    // pixels has no tail to speak of and runs 2 % faster without the team code (C4: 1736 -> 1770).
    int dev_t = 0, sms_t = 0;
    cudaGetDevice(&dev_t);
    cudaDeviceGetAttribute(&sms_t, cudaDevAttrMultiProcessorCount, dev_t);
    const uint64_t pixels = (uint64_t)lc.n_owned_tiles * 32u, slots = (uint64_t)wc.np * (uint64_t)sms_t;
    const bool team_auto = scene < 0 && (s.n_bsdf_types > 1 || lc.max_depth > 16u || pixels < 4u * slots);
    const bool team_on = !probe && (tune.team >= 2 || (tune.team == 1 && team_auto));
    wc.team = 0;
    wc.team_min = tune.team_min < 1 ? 1 : tune.team_min;
    wc.team_cap = lc.max_depth + 3u;                     // emission + one term per segment (maxDepth + 1 segments) + terminator
    wc.team_wmask = (wc.team_cap <= 32u ? 128u : (wc.team_cap <= 64u ? 64u : (wc.team_cap <= 128u ? 32u : 16u))) - 1u;
    if (team_on) {
        // the term log, sized for every SM
        int dev = 0, sms = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        const size_t per_column = (size_t)(wc.team_wmask + 1u) * wc.team_cap * sizeof(float4);
        const size_t need = (size_t)sms * 32u * per_column;
        if (s.team_log_bytes < need) {
            if (s.d_team_log) cudaFree(s.d_team_log);
            s.d_team_log = nullptr;
            s.team_log_bytes = 0;
            if (cudaMalloc((void**)&s.d_team_log, need) == cudaSuccess) s.team_log_bytes = need;
            else cudaGetLastError();                     // no memory for the log: this launch runs without teams
        }
        if (s.team_log_bytes >= need) {
            wc.team = tune.team == 3 ? 2 : 1;
            wc.team_log = reinterpret_cast<float4*>(s.d_team_log);
        }
    }
    const bool team = wc.team != 0;
    // Per-pixel candidate lists for camera rays: only where building one (a loop over all triangles per pixel) is
    // cheap, i.e. tiny scenes staged with 8 copies.  The counted kernel always traverses, so that its node / triangle
    // counters stay the ones of the specified traversal (and equal to the oracle's).
    const bool use_short = scene == 3 && s.n_tris <= 64 && tune.shortlist != 0;
    s.last_trav = 0;
    if (scene < 0 && trav == TRAV_WIDE) {
        LaunchConst lq = lc;
        lq.qnodes = s.d_wnodes;
        lq.q_lo = make_float3(s.q_lo[0], s.q_lo[1], s.q_lo[2]);
        lq.q_step = make_float3(s.q_step[0], s.q_step[1], s.q_step[2]);
        s.last_trav = 2;
        if (probe) return launch_wavefront<false, true, -1, false, TRAV_WIDE>(s, lq, wc, smem_bytes, stream);
        if (count)
            return team ? launch_wavefront<true, false, -1, false, TRAV_WIDE, false, true>(s, lq, wc, smem_bytes, stream)
                        : launch_wavefront<true, false, -1, false, TRAV_WIDE>(s, lq, wc, smem_bytes, stream);
        return team ? launch_wavefront<false, false, -1, false, TRAV_WIDE, false, true>(s, lq, wc, smem_bytes, stream)
                    : launch_wavefront<false, false, -1, false, TRAV_WIDE>(s, lq, wc, smem_bytes, stream);
    }
    if (scene < 0 && trav == TRAV_QUANT) {
        cudaError_t qe = quantize_nodes(s, stream);
        if (qe != cudaSuccess) return qe;
        if (s.d_qnodes && s.qnodes_tree == scene_tree(s)) {
            LaunchConst lq = lc;
            lq.qnodes = s.d_qnodes;
            lq.q_lo = make_float3(s.q_lo[0], s.q_lo[1], s.q_lo[2]);
            lq.q_step = make_float3(s.q_step[0], s.q_step[1], s.q_step[2]);
            s.last_trav = 1;
            if (probe) return launch_wavefront<false, true, -1, false, TRAV_QUANT>(s, lq, wc, smem_bytes, stream);
            if (count)
                return team ? launch_wavefront<true, false, -1, false, TRAV_QUANT, false, true>(s, lq, wc, smem_bytes, stream)
                            : launch_wavefront<true, false, -1, false, TRAV_QUANT>(s, lq, wc, smem_bytes, stream);
            return team ? launch_wavefront<false, false, -1, false, TRAV_QUANT, false, true>(s, lq, wc, smem_bytes, stream)
                        : launch_wavefront<false, false, -1, false, TRAV_QUANT>(s, lq, wc, smem_bytes, stream);
        }
        trav = TRAV_DIRECT;                              // a box could not be contained (never expected): exact nodes
    }
    // material-sorted SHADE batches (pt_scene_set_tune "msort=1"; A/B in profiles/): scenes with more than one BSDF type
    if (tune.msort != 0 && s.n_bsdf_types > 1 && !probe && !count && !use_short) {
        if (scene == 0)
            return team ? launch_wavefront<false, false, 0, false, TRAV_DIRECT, true, true>(s, lc, wc, smem_bytes, stream)
                        : launch_wavefront<false, false, 0, false, TRAV_DIRECT, true, false>(s, lc, wc, smem_bytes, stream);
        if (scene < 0)
            return team ? launch_wavefront<false, false, -1, false, TRAV_DIRECT, true, true>(s, lc, wc, smem_bytes, stream)
                        : launch_wavefront<false, false, -1, false, TRAV_DIRECT, true, false>(s, lc, wc, smem_bytes, stream);
    }
#define PTB_WF_DISPATCH(COUNT, PROBE, SHORT, TEAM)                                                                    \
    (scene == 3   ? launch_wavefront<COUNT, PROBE, 3, SHORT, TRAV_DIRECT, false, TEAM>(s, lc, wc, smem_bytes, stream)   \
     : scene == 0 ? launch_wavefront<COUNT, PROBE, 0, false, TRAV_DIRECT, false, TEAM>(s, lc, wc, smem_bytes, stream)   \
                  : launch_wavefront<COUNT, PROBE, -1, false, TRAV_DIRECT, false, TEAM>(s, lc, wc, smem_bytes, stream))
    if (probe) return use_short ? PTB_WF_DISPATCH(false, true, true, false) : PTB_WF_DISPATCH(false, true, false, false);
    if (count) {
        if (use_short && s.count_executed) return team ? PTB_WF_DISPATCH(true, false, true, true) : PTB_WF_DISPATCH(true, false, true, false);
        return team ? PTB_WF_DISPATCH(true, false, false, true) : PTB_WF_DISPATCH(true, false, false, false);
    }
    if (use_short) return team ? PTB_WF_DISPATCH(false, false, true, true) : PTB_WF_DISPATCH(false, false, true, false);
    return team ? PTB_WF_DISPATCH(false, false, false, true) : PTB_WF_DISPATCH(false, false, false, false);
#undef PTB_WF_DISPATCH
}

}  // namespace ptb
