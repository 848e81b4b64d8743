// jg_lbvh.cuh -- linear BVH over the triangles of the merged obstacle mesh, built on the GPU at scene load
// (Morton codes -> radix sort -> Karras 2012 hierarchy -> bottom-up box refit), and the tests of its traversal (the
// traversal kernel itself, k_mesh_traverse, is in jg_kernels.cuh: a group of lanes per query walks the tree together).
//
// The reference exports every scenario's obstacles as ONE triangle mesh (create_cpp_export.py:40-69,
// scenarios/<id>/export/obstacles.obj) for the sister planners; "mesh mode" answers the same query against that mesh
// (distance link hull <-> triangle soup) instead of against the convex pieces.
#pragma once
#include <cub/device/device_radix_sort.cuh>

#include "jg_math.cuh"

namespace jg {

struct BvhNode {
    float lo[3], hi[3];
    int left, right;    // >= 0: internal node index; < 0: leaf, triangle slot = ~child (slot in Morton order)
};

struct Lbvh {
    int n_tris = 0;
    BvhNode* nodes = nullptr;     // [max(n-1, 1)] internal nodes, root = 0
    float4* tri_v = nullptr;      // [3n] triangle vertices in Morton order (w unused)
    int* tri_id = nullptr;        // [n] original triangle index of slot i
};

__device__ __forceinline__ unsigned expand_bits(unsigned v) {
    v = (v * 0x00010001u) & 0xFF0000FFu;
    v = (v * 0x00000101u) & 0x0F00F00Fu;
    v = (v * 0x00000011u) & 0xC30C30C3u;
    v = (v * 0x00000005u) & 0x49249249u;
    return v;
}

__global__ void k_lbvh_morton(const float* __restrict__ verts, const int* __restrict__ tris, int n, float3 lo, float3 inv,
                              unsigned long long* __restrict__ keys, int* __restrict__ vals) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float c[3] = {0.f, 0.f, 0.f};
    for (int k = 0; k < 3; ++k)
        for (int a = 0; a < 3; ++a) c[a] += verts[3 * tris[3 * i + k] + a] * (1.f / 3.f);
    const unsigned x = (unsigned)fminf(fmaxf((c[0] - lo.x) * inv.x * 1024.f, 0.f), 1023.f);
    const unsigned y = (unsigned)fminf(fmaxf((c[1] - lo.y) * inv.y * 1024.f, 0.f), 1023.f);
    const unsigned z = (unsigned)fminf(fmaxf((c[2] - lo.z) * inv.z * 1024.f, 0.f), 1023.f);
    const unsigned code = (expand_bits(x) << 2) | (expand_bits(y) << 1) | expand_bits(z);
    keys[i] = ((unsigned long long)code << 32) | (unsigned)i;   // index appended: keys are unique
    vals[i] = i;
}

__device__ __forceinline__ int lbvh_delta(const unsigned long long* keys, int n, int i, int j) {
    if (j < 0 || j >= n) return -1;
    return __clzll(keys[i] ^ keys[j]);
}

// one thread per internal node (Karras 2012, fig. 4)
__global__ void k_lbvh_hierarchy(const unsigned long long* __restrict__ keys, int n, BvhNode* __restrict__ nodes,
                                 int* __restrict__ parent /* [2n-1]: internal 0..n-2, leaves n-1.. */) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n - 1) return;
    const int d = lbvh_delta(keys, n, i, i + 1) - lbvh_delta(keys, n, i, i - 1) >= 0 ? 1 : -1;
    const int dmin = lbvh_delta(keys, n, i, i - d);
    int lmax = 2;
    while (lbvh_delta(keys, n, i, i + lmax * d) > dmin) lmax <<= 1;
    int l = 0;
    for (int t = lmax >> 1; t >= 1; t >>= 1)
        if (lbvh_delta(keys, n, i, i + (l + t) * d) > dmin) l += t;
    const int j = i + l * d;
    const int dnode = lbvh_delta(keys, n, i, j);
    int s = 0;
    for (int t = (l + 1) >> 1;; t = (t + 1) >> 1) {
        if (lbvh_delta(keys, n, i, i + (s + t) * d) > dnode) s += t;
        if (t <= 1) break;
    }
    const int gamma = i + s * d + min(d, 0);
    const int lo = min(i, j), hi = max(i, j);
    const int left = lo == gamma ? ~gamma : gamma;            // leaf slots are encoded as ~slot
    const int right = hi == gamma + 1 ? ~(gamma + 1) : gamma + 1;
    nodes[i].left = left;
    nodes[i].right = right;
    parent[left < 0 ? (n - 1) + ~left : left] = i;
    parent[right < 0 ? (n - 1) + ~right : right] = i;
    if (i == 0) parent[0] = -1;
}

// gather triangles into Morton order; then one thread per leaf climbs to the root, the second arrival at a node
// merges the child boxes
__global__ void k_lbvh_gather(const float* __restrict__ verts, const int* __restrict__ tris, const int* __restrict__ order,
                              int n, float4* __restrict__ tri_v, int* __restrict__ tri_id) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int t = order[i];
    tri_id[i] = t;
    for (int k = 0; k < 3; ++k) {
        const int v = tris[3 * t + k];
        tri_v[3 * i + k] = make_float4(verts[3 * v], verts[3 * v + 1], verts[3 * v + 2], 0.f);
    }
}

__global__ void k_lbvh_refit(const float4* __restrict__ tri_v, int n, BvhNode* __restrict__ nodes,
                             const int* __restrict__ parent, int* __restrict__ visits) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    int node = parent[(n - 1) + i];
    while (node >= 0) {
        if (atomicAdd(visits + node, 1) == 0) return;   // first arrival: the sibling subtree is not ready yet
        __threadfence();
        BvhNode& N = nodes[node];
        float lo[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, hi[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
        for (int c = 0; c < 2; ++c) {
            const int ch = c == 0 ? N.left : N.right;
            if (ch < 0) {
                for (int k = 0; k < 3; ++k) {
                    const float4 p = tri_v[3 * (~ch) + k];
                    lo[0] = fminf(lo[0], p.x); lo[1] = fminf(lo[1], p.y); lo[2] = fminf(lo[2], p.z);
                    hi[0] = fmaxf(hi[0], p.x); hi[1] = fmaxf(hi[1], p.y); hi[2] = fmaxf(hi[2], p.z);
                }
            } else {
                const volatile BvhNode& C = nodes[ch];
                for (int a = 0; a < 3; ++a) { lo[a] = fminf(lo[a], C.lo[a]); hi[a] = fmaxf(hi[a], C.hi[a]); }
            }
        }
        for (int a = 0; a < 3; ++a) { N.lo[a] = lo[a]; N.hi[a] = hi[a]; }
        __threadfence();
        node = parent[node];
    }
}

// ------------------------------------------------------------------------------------ group-cooperative traversal
// The query of one (configuration, link shape): the shape's own (oriented) box, inflated by everything a reported
// distance may span.  Nodes are culled with the 6 face axes of the two boxes (3 world axes, 3 axes of the shape's
// box): conservative (the 9 edge-edge axes are left out) and much tighter than world box vs world box for a rotated link.
struct ObbQuery {
    V3 c;            // centre (robot frame)
    V3 a0, a1, a2;   // unit axes of the shape's box (columns of the shape's rotation)
    V3 h;            // half extents along a0, a1, a2 INCLUDING the inflation
    V3 e;            // half extents of its axis-aligned bounding box INCLUDING the inflation
};

__device__ __forceinline__ bool obb_vs_aabb(const ObbQuery& Q, const float* lo, const float* hi) {
    const V3 nh = mk(0.5f * (hi[0] - lo[0]), 0.5f * (hi[1] - lo[1]), 0.5f * (hi[2] - lo[2]));
    const V3 d = mk(0.5f * (hi[0] + lo[0]) - Q.c.x, 0.5f * (hi[1] + lo[1]) - Q.c.y, 0.5f * (hi[2] + lo[2]) - Q.c.z);
    if (fabsf(d.x) > Q.e.x + nh.x || fabsf(d.y) > Q.e.y + nh.y || fabsf(d.z) > Q.e.z + nh.z) return false;
    if (fabsf(dot(Q.a0, d)) > Q.h.x + fmaf(fabsf(Q.a0.x), nh.x, fmaf(fabsf(Q.a0.y), nh.y, fabsf(Q.a0.z) * nh.z))) return false;
    if (fabsf(dot(Q.a1, d)) > Q.h.y + fmaf(fabsf(Q.a1.x), nh.x, fmaf(fabsf(Q.a1.y), nh.y, fabsf(Q.a1.z) * nh.z))) return false;
    if (fabsf(dot(Q.a2, d)) > Q.h.z + fmaf(fabsf(Q.a2.x), nh.x, fmaf(fabsf(Q.a2.y), nh.y, fabsf(Q.a2.z) * nh.z))) return false;
    return true;
}

// Exact box-vs-triangle overlap (separating axes: 3 box faces, the triangle's plane, 9 edge cross products); the box is
// Q's (centre, axes, half extents h).  Closed sets: touching counts as overlapping.
__device__ __forceinline__ bool box_tri_overlap(const ObbQuery& Q, V3 p0, V3 p1, V3 p2) {
    const V3 d0 = p0 - Q.c, d1 = p1 - Q.c, d2 = p2 - Q.c;
    const V3 v0 = mk(dot(Q.a0, d0), dot(Q.a1, d0), dot(Q.a2, d0));
    const V3 v1 = mk(dot(Q.a0, d1), dot(Q.a1, d1), dot(Q.a2, d1));
    const V3 v2 = mk(dot(Q.a0, d2), dot(Q.a1, d2), dot(Q.a2, d2));
    const V3 h = Q.h;
    if (fminf(v0.x, fminf(v1.x, v2.x)) > h.x || fmaxf(v0.x, fmaxf(v1.x, v2.x)) < -h.x) return false;
    if (fminf(v0.y, fminf(v1.y, v2.y)) > h.y || fmaxf(v0.y, fmaxf(v1.y, v2.y)) < -h.y) return false;
    if (fminf(v0.z, fminf(v1.z, v2.z)) > h.z || fmaxf(v0.z, fmaxf(v1.z, v2.z)) < -h.z) return false;
    const V3 e0 = v1 - v0, e1 = v2 - v1, e2 = v0 - v2;
    const V3 n = cross(e0, e1);
    if (fabsf(dot(n, v0)) > fmaf(h.x, fabsf(n.x), fmaf(h.y, fabsf(n.y), h.z * fabsf(n.z)))) return false;
    // axis = (box axis u_j) x (edge e): projections of the three vertices (two of them coincide) against the box radius
#define JG_AXIS(ex, ey, ez, va, vb)                                                                              \
    {                                                                                                            \
        const float fx = fabsf(ex), fy = fabsf(ey), fz = fabsf(ez);                                               \
        { const float pa = ey * va.z - ez * va.y, pb = ey * vb.z - ez * vb.y;   /* u = x */                       \
          if (fminf(pa, pb) > h.y * fz + h.z * fy || fmaxf(pa, pb) < -(h.y * fz + h.z * fy)) return false; }     \
        { const float pa = ez * va.x - ex * va.z, pb = ez * vb.x - ex * vb.z;   /* u = y */                       \
          if (fminf(pa, pb) > h.x * fz + h.z * fx || fmaxf(pa, pb) < -(h.x * fz + h.z * fx)) return false; }     \
        { const float pa = ex * va.y - ey * va.x, pb = ex * vb.y - ey * vb.x;   /* u = z */                       \
          if (fminf(pa, pb) > h.x * fy + h.y * fx || fmaxf(pa, pb) < -(h.x * fy + h.y * fx)) return false; }     \
    }
    JG_AXIS(e0.x, e0.y, e0.z, v0, v2)   // along e0 the vertices v0 and v1 project alike
    JG_AXIS(e1.x, e1.y, e1.z, v1, v0)
    JG_AXIS(e2.x, e2.y, e2.z, v2, v1)
#undef JG_AXIS
    return true;
}

}  // namespace jg
