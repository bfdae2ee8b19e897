// oracle/oracle_geom.h -- TEST INFRASTRUCTURE (CPU oracle), NOT PRODUCT CODE.
//
// What the reference delegates to the closed OptiX driver -- acceleration-structure build
// (PathTracerMain.cpp:311-398, optixAccelBuild) and traversal + ray/triangle test
// (pathTracerPrograms.cu:600,660, optixTraverse) -- has no source in the reference tree, so
// parity at that boundary is UNPINNED by the reference.  This file is the scalar CPU statement
// of the arithmetic the B200 kernels are specified to perform (DESIGN.md "Intersection spec"
// and "LBVH spec"), written independently of the CUDA code:
//   * Moeller-Trumbore ray/triangle test on (v0, e1, e2) records, explicit fma placement
//   * slab test against child boxes with a 2-ulp padded far bound
//   * closest hit = minimum over the total order (t, primIdx); any hit = boolean
//   * LBVH: 30-bit Morton code of the triangle-AABB centre, key = morton<<32 | primIdx,
//     sorted ascending (std::sort here; radix sort on the GPU), Karras-2012 radix tree,
//     bottom-up AABB union, 64-byte two-child nodes
// plus a brute-force loop over all triangles used to check that BVH culling is conservative.
#pragma once
#include <algorithm>
#include <cfloat>
#include <cstdint>
#include <vector>

#include "oracle_vec.h"

namespace ora {

struct TriRec {   // 48 bytes, same layout as the device triangle record
    float v0[3]; uint32_t prim;
    float e1[3]; uint32_t mat;
    float e2[3]; uint32_t pad;
};

struct BvhNode {  // 64 bytes, same layout as the device node
    float lo0[3]; int32_t c0;
    float hi0[3]; int32_t c1;
    float lo1[3]; int32_t pad0;
    float hi1[3]; int32_t pad1;
};

struct Hit {
    bool hit;
    uint32_t prim;
    float t;
};

struct TraceCounters {
    unsigned long long n_node = 0, n_tri = 0;
};

// ---- ray / triangle -----------------------------------------------------------------------
static inline float fms(float a, float b, float c, float d) { return fmaf(a, b, -(c * d)); }   // a*b - c*d
static inline float dot_fma(const float* a, const float* b) { return fmaf(a[0], b[0], fmaf(a[1], b[1], a[2] * b[2])); }

static inline bool ray_tri(const float o[3], const float d[3], const TriRec& tr, float tmin, float tmax, float& t_out)
{
    float p[3] = {fms(d[1], tr.e2[2], d[2], tr.e2[1]), fms(d[2], tr.e2[0], d[0], tr.e2[2]),
                  fms(d[0], tr.e2[1], d[1], tr.e2[0])};
    float det = dot_fma(tr.e1, p);
    if (det == 0.0f) return false;
    float inv = 1.0f / det;
    float tv[3] = {o[0] - tr.v0[0], o[1] - tr.v0[1], o[2] - tr.v0[2]};
    float u = dot_fma(tv, p) * inv;
    if (!(u >= 0.0f && u <= 1.0f)) return false;
    float q[3] = {fms(tv[1], tr.e1[2], tv[2], tr.e1[1]), fms(tv[2], tr.e1[0], tv[0], tr.e1[2]),
                  fms(tv[0], tr.e1[1], tv[1], tr.e1[0])};
    float v = dot_fma(d, q) * inv;
    if (!(v >= 0.0f && u + v <= 1.0f)) return false;
    float t = dot_fma(tr.e2, q) * inv;
    if (!(t > tmin && t < tmax)) return false;
    t_out = t;
    return true;
}

// ---- ray / box -----------------------------------------------------------------------------
struct RayPre {
    float o[3], d[3], inv[3], ood[3];
};

static inline RayPre ray_prepare(const float o[3], const float d[3])
{
    RayPre r;
    for (int a = 0; a < 3; ++a) {
        r.o[a] = o[a];
        r.d[a] = d[a];
        float dd = fabsf(d[a]) > 1e-20f ? d[a] : copysignf(1e-20f, d[a]);
        r.inv[a] = 1.0f / dd;
        r.ood[a] = o[a] * r.inv[a];
    }
    return r;
}

static inline bool ray_box(const RayPre& r, const float lo[3], const float hi[3], float tmin, float tfar, float& tnear)
{
    float t0x = fmaf(lo[0], r.inv[0], -r.ood[0]), t1x = fmaf(hi[0], r.inv[0], -r.ood[0]);
    float t0y = fmaf(lo[1], r.inv[1], -r.ood[1]), t1y = fmaf(hi[1], r.inv[1], -r.ood[1]);
    float t0z = fmaf(lo[2], r.inv[2], -r.ood[2]), t1z = fmaf(hi[2], r.inv[2], -r.ood[2]);
    float tn = fmaxf(fmaxf(fminf(t0x, t1x), fminf(t0y, t1y)), fmaxf(fminf(t0z, t1z), tmin));
    float tf = fminf(fminf(fmaxf(t0x, t1x), fmaxf(t0y, t1y)), fminf(fmaxf(t0z, t1z), tfar));
    tnear = tn;
    return tn <= tf * 1.0000004f;
}

// ---- brute force ---------------------------------------------------------------------------
static inline Hit brute_closest(const std::vector<TriRec>& tris, const float o[3], const float d[3], float tmin, float tmax)
{
    Hit h{false, 0xffffffffu, tmax};
    for (const TriRec& tr : tris) {
        float t;
        if (ray_tri(o, d, tr, tmin, tmax, t)) {
            if (t < h.t || (t == h.t && tr.prim < h.prim)) {
                h.hit = true;
                h.t = t;
                h.prim = tr.prim;
            }
        }
    }
    return h;
}

static inline bool brute_any(const std::vector<TriRec>& tris, const float o[3], const float d[3], float tmin, float tmax)
{
    for (const TriRec& tr : tris) {
        float t;
        if (ray_tri(o, d, tr, tmin, tmax, t)) return true;
    }
    return false;
}

// ---- LBVH ----------------------------------------------------------------------------------
static inline uint32_t expand_bits10(uint32_t v)
{
    v &= 0x3ffu;
    v = (v | (v << 16)) & 0x030000ffu;
    v = (v | (v << 8)) & 0x0300f00fu;
    v = (v | (v << 4)) & 0x030c30c3u;
    v = (v | (v << 2)) & 0x09249249u;
    return v;
}

struct Lbvh {
    uint32_t n_tri = 0;
    float scene_lo[3], scene_hi[3];
    std::vector<uint32_t> morton;       // per primitive (input order)
    std::vector<uint64_t> sorted_keys;  // morton<<32 | prim, ascending
    std::vector<uint32_t> sorted_prim;  // leaf i -> primitive
    std::vector<int32_t> parent;        // per internal node (root: -1)
    std::vector<BvhNode> nodes;         // max(n_tri - 1, 1) two-child nodes (empty when n_tri == 0)
    std::vector<TriRec> tris;           // leaf order
    std::vector<TriRec> tris_input;     // input (primitive) order, for brute force
    std::vector<float> leaf_lo, leaf_hi; // 3 per leaf slot: the triangle's box (leaf order)
    // sweep-SAH tree over the same leaves (sah_build below); `use_sah` selects which tree the traversals walk
    std::vector<BvhNode> sah_nodes;
    uint32_t sah_depth = 0;
    bool use_sah = false;
    // PLOC tree over the same leaves (ploc_build below)
    std::vector<BvhNode> ploc_nodes;
    uint32_t ploc_depth = 0;
    bool use_ploc = false;
    const std::vector<BvhNode>& walk() const { return use_sah ? sah_nodes : (use_ploc ? ploc_nodes : nodes); }
};

static inline int delta64(const std::vector<uint64_t>& k, int i, int j)
{
    if (j < 0 || j >= (int)k.size()) return -1;
    return __builtin_clzll(k[i] ^ k[j]);  // keys are unique -> xor never 0
}

static inline Lbvh lbvh_build(const float* verts_xyzw, uint32_t /*n_verts*/, const uint32_t* indices, uint32_t n_tri,
                              const uint32_t* mat_ids)
{
    Lbvh b;
    b.n_tri = n_tri;
    for (int a = 0; a < 3; ++a) {
        b.scene_lo[a] = FLT_MAX;
        b.scene_hi[a] = -FLT_MAX;
    }
    std::vector<float> bbmin(3 * (size_t)n_tri), bbmax(3 * (size_t)n_tri);
    b.tris_input.resize(n_tri);
    for (uint32_t k = 0; k < n_tri; ++k) {
        const float* v0 = verts_xyzw + 4 * (size_t)indices[3 * k + 0];
        const float* v1 = verts_xyzw + 4 * (size_t)indices[3 * k + 1];
        const float* v2 = verts_xyzw + 4 * (size_t)indices[3 * k + 2];
        TriRec& tr = b.tris_input[k];
        for (int a = 0; a < 3; ++a) {
            tr.v0[a] = v0[a];
            tr.e1[a] = v1[a] - v0[a];
            tr.e2[a] = v2[a] - v0[a];
            float mn = fminf(fminf(v0[a], v1[a]), v2[a]);
            float mx = fmaxf(fmaxf(v0[a], v1[a]), v2[a]);
            bbmin[3 * k + a] = mn;
            bbmax[3 * k + a] = mx;
            b.scene_lo[a] = fminf(b.scene_lo[a], mn);
            b.scene_hi[a] = fmaxf(b.scene_hi[a], mx);
        }
        tr.prim = k;
        tr.mat = mat_ids ? mat_ids[k] : 0u;
        tr.pad = 0;
    }
    if (n_tri == 0) return b;

    float inv_e[3];
    for (int a = 0; a < 3; ++a) {
        float e = b.scene_hi[a] - b.scene_lo[a];
        inv_e[a] = e > 0.0f ? 1.0f / e : 0.0f;
    }
    b.morton.resize(n_tri);
    b.sorted_keys.resize(n_tri);
    for (uint32_t k = 0; k < n_tri; ++k) {
        uint32_t q[3];
        for (int a = 0; a < 3; ++a) {
            float c = (bbmin[3 * k + a] + bbmax[3 * k + a]) * 0.5f;
            float n = (c - b.scene_lo[a]) * inv_e[a];
            q[a] = (uint32_t)fminf(fmaxf(n * 1024.0f, 0.0f), 1023.0f);
        }
        uint32_t m = (expand_bits10(q[0]) << 2) | (expand_bits10(q[1]) << 1) | expand_bits10(q[2]);
        b.morton[k] = m;
        b.sorted_keys[k] = ((uint64_t)m << 32) | k;
    }
    std::sort(b.sorted_keys.begin(), b.sorted_keys.end());
    b.sorted_prim.resize(n_tri);
    b.tris.resize(n_tri);
    for (uint32_t i = 0; i < n_tri; ++i) {
        b.sorted_prim[i] = (uint32_t)(b.sorted_keys[i] & 0xffffffffu);
        b.tris[i] = b.tris_input[b.sorted_prim[i]];
    }
    b.leaf_lo.resize(3 * (size_t)n_tri);
    b.leaf_hi.resize(3 * (size_t)n_tri);
    for (uint32_t i = 0; i < n_tri; ++i)
        for (int a = 0; a < 3; ++a) {
            b.leaf_lo[3 * (size_t)i + a] = bbmin[3 * (size_t)b.sorted_prim[i] + a];
            b.leaf_hi[3 * (size_t)i + a] = bbmax[3 * (size_t)b.sorted_prim[i] + a];
        }

    const int N = (int)n_tri;
    const int n_internal = N > 1 ? N - 1 : 1;
    b.nodes.resize(n_internal);
    b.parent.assign(n_internal, -1);
    if (N == 1) {
        // degenerate tree: one node whose two children are the same leaf
        BvhNode& nd = b.nodes[0];
        for (int a = 0; a < 3; ++a) {
            nd.lo0[a] = nd.lo1[a] = bbmin[a];
            nd.hi0[a] = nd.hi1[a] = bbmax[a];
        }
        nd.c0 = nd.c1 = ~0;
        nd.pad0 = nd.pad1 = 0;
        return b;
    }

    // Karras 2012, "Maximizing Parallelism in the Construction of BVHs, Octrees, and k-d Trees"
    std::vector<int32_t> left(N - 1), right(N - 1);
    std::vector<int32_t> leaf_parent(N);
    const std::vector<uint64_t>& K = b.sorted_keys;
    for (int i = 0; i < N - 1; ++i) {
        int d = (delta64(K, i, i + 1) - delta64(K, i, i - 1)) >= 0 ? 1 : -1;
        int dmin = delta64(K, i, i - d);
        int lmax = 2;
        while (delta64(K, i, i + lmax * d) > dmin) lmax *= 2;
        int l = 0;
        for (int t = lmax / 2; t >= 1; t /= 2)
            if (delta64(K, i, i + (l + t) * d) > dmin) l += t;
        int j = i + l * d;
        int dnode = delta64(K, i, j);
        int s = 0;
        int t = l;
        do {
            t = (t + 1) >> 1;
            if (delta64(K, i, i + (s + t) * d) > dnode) s += t;
        } while (t > 1);
        int gamma = i + s * d + std::min(d, 0);
        int lo = std::min(i, j), hi = std::max(i, j);
        left[i] = (lo == gamma) ? ~gamma : gamma;
        right[i] = (hi == gamma + 1) ? ~(gamma + 1) : gamma + 1;
        if (left[i] >= 0) b.parent[left[i]] = i; else leaf_parent[~left[i]] = i;
        if (right[i] >= 0) b.parent[right[i]] = i; else leaf_parent[~right[i]] = i;
    }

    // bottom-up boxes (children before parents: process by recursion from the root)
    std::vector<float> nlo(3 * (size_t)(N - 1)), nhi(3 * (size_t)(N - 1));
    std::vector<uint8_t> done(N - 1, 0);
    auto child_box = [&](int32_t ref, float* lo, float* hi) {
        if (ref < 0) {
            uint32_t p = b.sorted_prim[~ref];
            for (int a = 0; a < 3; ++a) {
                lo[a] = bbmin[3 * p + a];
                hi[a] = bbmax[3 * p + a];
            }
        } else {
            for (int a = 0; a < 3; ++a) {
                lo[a] = nlo[3 * ref + a];
                hi[a] = nhi[3 * ref + a];
            }
        }
    };
    // iterative post-order
    std::vector<int32_t> stack;
    stack.push_back(0);
    while (!stack.empty()) {
        int32_t n = stack.back();
        bool ready = true;
        if (left[n] >= 0 && !done[left[n]]) { stack.push_back(left[n]); ready = false; }
        if (right[n] >= 0 && !done[right[n]]) { stack.push_back(right[n]); ready = false; }
        if (!ready) continue;
        stack.pop_back();
        BvhNode& nd = b.nodes[n];
        child_box(left[n], nd.lo0, nd.hi0);
        child_box(right[n], nd.lo1, nd.hi1);
        nd.c0 = left[n];
        nd.c1 = right[n];
        nd.pad0 = nd.pad1 = 0;
        for (int a = 0; a < 3; ++a) {
            nlo[3 * n + a] = fminf(nd.lo0[a], nd.lo1[a]);
            nhi[3 * n + a] = fmaxf(nd.hi0[a], nd.hi1[a]);
        }
        done[n] = 1;
    }
    return b;
}

// ---- sweep-SAH tree over the LBVH's leaf slots ----------------------------------------------------
// Our specification (csrc/sah.cu states it; OptiX's builder is closed): a subtree over m >= 2 slots tries, for the
// axes x, y, z in turn, the order by (lo + hi in float, slot) and every split "first k | rest", cost = A(first k) * k +
// A(rest) * (m - k) with A = ex*ey + ey*ez + ez*ex in double; the first strict minimum wins; nodes in preorder.
// Written recursively here (the library walks an explicit job stack) -- same tree, bit for bit, or the tests fail.
struct SahBuilder {
    const Lbvh& b;
    std::vector<BvhNode>& out;
    uint32_t depth = 0;
    SahBuilder(const Lbvh& bb, std::vector<BvhNode>& o) : b(bb), out(o) {}

    struct Aabb {
        float lo[3], hi[3];
    };
    Aabb slot_box(uint32_t s) const
    {
        Aabb x;
        for (int a = 0; a < 3; ++a) {
            x.lo[a] = b.leaf_lo[3 * (size_t)s + a];
            x.hi[a] = b.leaf_hi[3 * (size_t)s + a];
        }
        return x;
    }
    static Aabb join(Aabb x, const Aabb& y)
    {
        for (int a = 0; a < 3; ++a) {
            x.lo[a] = fminf(x.lo[a], y.lo[a]);
            x.hi[a] = fmaxf(x.hi[a], y.hi[a]);
        }
        return x;
    }
    static double area(const Aabb& x)
    {
        double e[3];
        for (int a = 0; a < 3; ++a) e[a] = (double)x.hi[a] - (double)x.lo[a];
        return e[0] * e[1] + e[1] * e[2] + e[2] * e[0];
    }

    // returns the child reference of the subtree over `slots`
    int32_t build(std::vector<uint32_t> slots, uint32_t level)
    {
        if (slots.size() == 1) return ~(int32_t)slots[0];
        if (level > depth) depth = level;
        const size_t m = slots.size();
        const int32_t me = (int32_t)out.size();
        out.emplace_back();
        bool found = false;
        double best = 0.0;
        size_t best_k = 0;
        std::vector<uint32_t> best_sorted;
        Aabb best_left{}, best_right{};
        for (int a = 0; a < 3; ++a) {
            std::vector<uint32_t> sorted = slots;
            std::stable_sort(sorted.begin(), sorted.end(), [&](uint32_t x, uint32_t y) {
                const float kx = b.leaf_lo[3 * (size_t)x + a] + b.leaf_hi[3 * (size_t)x + a];
                const float ky = b.leaf_lo[3 * (size_t)y + a] + b.leaf_hi[3 * (size_t)y + a];
                if (kx != ky) return kx < ky;
                return x < y;
            });
            std::vector<Aabb> left(m), right(m);      // left[i] = boxes 0..i, right[i] = boxes i..m-1
            left[0] = slot_box(sorted[0]);
            for (size_t i = 1; i < m; ++i) left[i] = join(left[i - 1], slot_box(sorted[i]));
            right[m - 1] = slot_box(sorted[m - 1]);
            for (size_t i = m - 1; i-- > 0;) right[i] = join(right[i + 1], slot_box(sorted[i]));
            bool took = false;
            for (size_t k = 1; k < m; ++k) {
                const double c = area(left[k - 1]) * (double)k + area(right[k]) * (double)(m - k);
                if (!found || c < best) {
                    found = true;
                    took = true;
                    best = c;
                    best_k = k;
                    best_left = left[k - 1];
                    best_right = right[k];
                }
            }
            if (took) best_sorted = sorted;
        }
        const int32_t l = build(std::vector<uint32_t>(best_sorted.begin(), best_sorted.begin() + best_k), level + 1);
        const int32_t r = build(std::vector<uint32_t>(best_sorted.begin() + best_k, best_sorted.end()), level + 1);
        BvhNode& nd = out[me];
        for (int a = 0; a < 3; ++a) {
            nd.lo0[a] = best_left.lo[a];
            nd.hi0[a] = best_left.hi[a];
            nd.lo1[a] = best_right.lo[a];
            nd.hi1[a] = best_right.hi[a];
        }
        nd.c0 = l;
        nd.c1 = r;
        nd.pad0 = nd.pad1 = 0;
        return me;
    }
};

// ---- PLOC tree over the LBVH's leaf slots (csrc/lbvh.cu ploc_* kernels state the rounds; serial here) ----------------
// Clusters start as the leaves in Morton order.  Round: cluster i chooses, among j in [i - 8, i + 8], j != i, the one
// whose union box has the smallest ex*ey + ey*ez + ez*ex (float; first strict minimum for ascending j; default i - 1, or
// i + 1 for i = 0); mutual choices merge into a node at the lower index, survivors keep their order.  Node ids count
// down from n - 2 in creation order (round by round, ascending position), so the last node -- the root -- is node 0.
static inline void ploc_build(Lbvh& b)
{
    b.ploc_nodes.clear();
    b.ploc_depth = 0;
    b.use_ploc = false;
    const int n = (int)b.n_tri;
    if (n < 2) return;
    struct Cluster {
        float lo[3], hi[3];
        int32_t ref;
    };
    std::vector<Cluster> cur(n), next;
    for (int i = 0; i < n; ++i) {
        for (int a = 0; a < 3; ++a) {
            cur[i].lo[a] = b.leaf_lo[3 * (size_t)i + a];
            cur[i].hi[a] = b.leaf_hi[3 * (size_t)i + a];
        }
        cur[i].ref = ~i;
    }
    b.ploc_nodes.resize(n - 1);
    std::vector<uint32_t> node_level(n - 1, 0);          // height of the subtree under a node, in internal nodes
    int next_id = n - 2;
    std::vector<int> nn;
    while (cur.size() > 1) {
        const int c = (int)cur.size();
        nn.assign(c, 0);
        for (int i = 0; i < c; ++i) {
            int best_j = i > 0 ? i - 1 : i + 1;
            float best = INFINITY;
            const int first = std::max(0, i - 8), last = std::min(c - 1, i + 8);
            for (int j = first; j <= last; ++j) {
                if (j == i) continue;
                float e[3];
                for (int a = 0; a < 3; ++a) e[a] = fmaxf(cur[i].hi[a], cur[j].hi[a]) - fminf(cur[i].lo[a], cur[j].lo[a]);
                const float area = e[0] * e[1] + e[1] * e[2] + e[2] * e[0];
                if (area < best) {
                    best = area;
                    best_j = j;
                }
            }
            nn[i] = best_j;
        }
        next.clear();
        int made = 0;
        for (int i = 0; i < c; ++i) {
            const int j = nn[i];
            const bool mutual = nn[j] == i;
            if (mutual && i > j) continue;
            if (!mutual) {
                next.push_back(cur[i]);
                continue;
            }
            const int id = next_id - made;
            ++made;
            BvhNode& nd = b.ploc_nodes[id];
            Cluster m;
            for (int a = 0; a < 3; ++a) {
                nd.lo0[a] = cur[i].lo[a];
                nd.hi0[a] = cur[i].hi[a];
                nd.lo1[a] = cur[j].lo[a];
                nd.hi1[a] = cur[j].hi[a];
                m.lo[a] = fminf(cur[i].lo[a], cur[j].lo[a]);
                m.hi[a] = fmaxf(cur[i].hi[a], cur[j].hi[a]);
            }
            nd.c0 = cur[i].ref;
            nd.c1 = cur[j].ref;
            nd.pad0 = nd.pad1 = 0;
            const uint32_t l0 = cur[i].ref >= 0 ? node_level[cur[i].ref] : 0u, l1 = cur[j].ref >= 0 ? node_level[cur[j].ref] : 0u;
            node_level[id] = 1u + std::max(l0, l1);
            m.ref = id;
            next.push_back(m);
        }
        next_id -= made;
        cur.swap(next);
    }
    b.ploc_depth = node_level[0];
}

constexpr uint32_t kSahMaxTris = 8192, kSahMaxDepth = 64;   // the library's limits (csrc/pt_scene.h)

// Builds b.sah_nodes for scenes of 2 .. kSahMaxTris triangles and applies the library's "auto" rule to b.use_sah.
static inline void sah_build(Lbvh& b)
{
    b.sah_nodes.clear();
    b.sah_depth = 0;
    b.use_sah = false;
    if (b.n_tri < 2 || b.n_tri > kSahMaxTris) return;
    b.sah_nodes.reserve(b.n_tri - 1);
    SahBuilder sb(b, b.sah_nodes);
    std::vector<uint32_t> all(b.n_tri);
    for (uint32_t i = 0; i < b.n_tri; ++i) all[i] = i;
    sb.build(all, 1);
    b.sah_depth = sb.depth;
    b.use_sah = b.sah_depth <= kSahMaxDepth;
}

// The library's rule (csrc/pt_scene.h scene_tree): tree = 0 auto (SAH where there is one, else PLOC, else LBVH; a tree
// deeper than 64 is passed over, also when asked for by name), 1 LBVH, 2 SAH, 3 PLOC.  Returns the tree in use (1 / 2 / 3).
static inline int select_tree(Lbvh& b, int tree)
{
    const bool have_sah = !b.sah_nodes.empty(), have_ploc = !b.ploc_nodes.empty();
    int t = 1;
    if (tree == 2 && have_sah && b.sah_depth <= kSahMaxDepth) t = 2;
    else if (tree == 3 && have_ploc && b.ploc_depth <= kSahMaxDepth) t = 3;
    else if (tree == 0) {
        if (have_sah && b.sah_depth <= kSahMaxDepth) t = 2;
        else if (have_ploc && b.ploc_depth <= kSahMaxDepth) t = 3;
    }
    b.use_sah = t == 2;
    b.use_ploc = t == 3;
    return t;
}

// ---- traversal -----------------------------------------------------------------------------
// Logical sequence (mirrored by the device kernels): start at node 0; at an internal node test
// both child boxes against [tmin, best_t]; visit the nearer hit child first (child 1 only when
// strictly nearer), push the other; a leaf child is intersected when visited.
static inline Hit bvh_closest(const Lbvh& b, const float o[3], const float d[3], float tmin, float tmax,
                              TraceCounters* ctr = nullptr)
{
    Hit h{false, 0xffffffffu, tmax};
    if (b.n_tri == 0) return h;
    RayPre r = ray_prepare(o, d);
    int32_t stack[128];
    int sp = 0;
    int32_t cur = 0;
    for (;;) {
        if (cur >= 0) {
            const BvhNode& nd = b.walk()[cur];
            if (ctr) ctr->n_node++;
            float tn0, tn1;
            bool h0 = ray_box(r, nd.lo0, nd.hi0, tmin, h.t, tn0);
            bool h1 = ray_box(r, nd.lo1, nd.hi1, tmin, h.t, tn1);
            if (h0 && h1) {
                bool swap = tn1 < tn0;
                int32_t near = swap ? nd.c1 : nd.c0;
                int32_t far = swap ? nd.c0 : nd.c1;
                stack[sp++] = far;
                cur = near;
                continue;
            } else if (h0) {
                cur = nd.c0;
                continue;
            } else if (h1) {
                cur = nd.c1;
                continue;
            }
        } else {
            const TriRec& tr = b.tris[~cur];
            if (ctr) ctr->n_tri++;
            float t;
            if (ray_tri(o, d, tr, tmin, tmax, t)) {
                if (t < h.t || (t == h.t && tr.prim < h.prim)) {
                    h.hit = true;
                    h.t = t;
                    h.prim = tr.prim;
                }
            }
        }
        if (sp == 0) break;
        cur = stack[--sp];
    }
    return h;
}

static inline bool bvh_any(const Lbvh& b, const float o[3], const float d[3], float tmin, float tmax,
                           TraceCounters* ctr = nullptr)
{
    if (b.n_tri == 0) return false;
    RayPre r = ray_prepare(o, d);
    int32_t stack[128];
    int sp = 0;
    int32_t cur = 0;
    for (;;) {
        if (cur >= 0) {
            const BvhNode& nd = b.walk()[cur];
            if (ctr) ctr->n_node++;
            float tn0, tn1;
            bool h0 = ray_box(r, nd.lo0, nd.hi0, tmin, tmax, tn0);
            bool h1 = ray_box(r, nd.lo1, nd.hi1, tmin, tmax, tn1);
            if (h0 && h1) {
                bool swap = tn1 < tn0;
                stack[sp++] = swap ? nd.c0 : nd.c1;
                cur = swap ? nd.c1 : nd.c0;
                continue;
            } else if (h0) {
                cur = nd.c0;
                continue;
            } else if (h1) {
                cur = nd.c1;
                continue;
            }
        } else {
            const TriRec& tr = b.tris[~cur];
            if (ctr) ctr->n_tri++;
            float t;
            if (ray_tri(o, d, tr, tmin, tmax, t)) return true;
        }
        if (sp == 0) break;
        cur = stack[--sp];
    }
    return false;
}

}  // namespace ora
