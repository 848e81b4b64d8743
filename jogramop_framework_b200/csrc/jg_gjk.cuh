// jg_gjk.cuh -- per-thread GJK distance between two convex vertex sets held in shared memory.
//
// Layout contract: every lane of a warp works on the SAME (A, B) vertex ranges (pairs are bucketed by
// (link shape, obstacle piece) before this runs), so every shared-memory read in the support scans is a
// warp-wide broadcast: one LDS.128 feeds 32 lanes x 3 FMA.  Only the direction, the rigid transform of A
// and the simplex differ per lane and live in registers.
//
// Semantics follow Bullet's btGjkPairDetector as restated in oracle/oracle.c (gjk()): distance between
// the *core* shapes; early exit once a separating axis proves the distance exceeds dmax; overlap (origin
// inside the simplex / closer than 1e-7 m) is reported as JG_GJK_PENETRATING.
#pragma once
#include <float.h>

#include "jg_math.cuh"

namespace jg {

enum { JG_GJK_SEPARATED = 0, JG_GJK_FAR = 1, JG_GJK_PENETRATING = 2 };

#ifndef JG_GJK_MAX_ITER
#define JG_GJK_MAX_ITER 24
#endif
// stop when the duality gap |v| - v.w/|v| is below this many metres (fp32 noise floor is ~2e-7)
#ifndef JG_GJK_TOL
#define JG_GJK_TOL 1e-6f
#endif

// argmax_i dot(verts[i], d) over a shared-memory vertex range (warp-uniform range, per-lane direction).
// Vertex sets are padded to a multiple of 4 with copies of their last vertex (host side).  Four dot products per
// step are reduced with a max tree and only the group index is tracked: 5.5 issue slots per vertex instead of 7 and a
// loop-carried dependency once per 4 vertices.  The winner inside the winning group is located afterwards by
// repeating the same arithmetic (bitwise identical values); ties resolve to the lowest index, like a sequential scan.
#ifndef JG_SCAN_UNROLL
#define JG_SCAN_UNROLL 2
#endif
#define JG_STR_(x) #x
#define JG_UNROLL(n) _Pragma(JG_STR_(unroll n))
__device__ __forceinline__ float dot4(const float4 p, V3 d) { return fmaf(p.x, d.x, fmaf(p.y, d.y, p.z * d.z)); }

__device__ __forceinline__ int support_scan(const float4* __restrict__ verts, int n, V3 d) {
    const int ng = (n + 3) >> 2;
    JG_WORK(JG_WORK_DOTS, (unsigned)(4 * ng + 3));
    float best = -FLT_MAX;
    int bg = 0;
    JG_UNROLL(JG_SCAN_UNROLL)
    for (int g = 0; g < ng; ++g) {
        const float s0 = dot4(verts[4 * g], d), s1 = dot4(verts[4 * g + 1], d);
        const float s2 = dot4(verts[4 * g + 2], d), s3 = dot4(verts[4 * g + 3], d);
        const float m = fmaxf(fmaxf(s0, s1), fmaxf(s2, s3));
        if (m > best) {
            best = m;
            bg = g;
        }
    }
    const float s0 = dot4(verts[4 * bg], d), s1 = dot4(verts[4 * bg + 1], d), s2 = dot4(verts[4 * bg + 2], d);
    int k = 3;
    if (s2 == best) k = 2;
    if (s1 == best) k = 1;
    if (s0 == best) k = 0;
    return 4 * bg + k;
}

// closest point of triangle (a,b,c) to the origin; mask bit0/1/2 = a/b/c belongs to the supporting feature
__device__ __forceinline__ V3 tri_closest(V3 a, V3 b, V3 c, int& mask) {
    const V3 ab = b - a, ac = c - a;
    const float d1 = -dot(ab, a), d2 = -dot(ac, a);
    if (d1 <= 0.f && d2 <= 0.f) { mask = 1; return a; }
    const float d3 = -dot(ab, b), d4 = -dot(ac, b);
    if (d3 >= 0.f && d4 <= d3) { mask = 2; return b; }
    const float vc = d1 * d4 - d3 * d2;
    if (vc <= 0.f && d1 >= 0.f && d3 <= 0.f) { mask = 3; return a + ab * (d1 / (d1 - d3)); }
    const float d5 = -dot(ab, c), d6 = -dot(ac, c);
    if (d6 >= 0.f && d5 <= d6) { mask = 4; return c; }
    const float vb = d5 * d2 - d1 * d6;
    if (vb <= 0.f && d2 >= 0.f && d6 <= 0.f) { mask = 5; return a + ac * (d2 / (d2 - d6)); }
    const float va = d3 * d6 - d5 * d4;
    if (va <= 0.f && (d4 - d3) >= 0.f && (d5 - d6) >= 0.f) {
        mask = 6;
        return b + (c - b) * ((d4 - d3) / ((d4 - d3) + (d5 - d6)));
    }
    const V3 n = cross(ab, ac);
    const float nn = norm2(n);
    if (!(nn > 0.f)) { mask = 3; return a + ab * (d1 / (d1 - d3)); }
    mask = 7;
    return n * (dot(n, a) / nn);  // project the origin onto the plane: stable when the origin is close to it
}

struct Simplex {
    V3 a, b, c, d;         // a = newest vertex
    uint32_t ia, ib, ic, id;  // (index in A) << 16 | index in B
    int n;
};

// keep the vertices of (p0,p1,p2) selected by mask, in order
__device__ __forceinline__ void simplex_set3(Simplex& S, V3 p0, V3 p1, V3 p2, uint32_t i0, uint32_t i1, uint32_t i2,
                                             int mask) {
    int n = 0;
    if (mask & 1) { S.a = p0; S.ia = i0; n = 1; }
    if (mask & 2) {
        if (n == 0) { S.a = p1; S.ia = i1; } else { S.b = p1; S.ib = i1; }
        ++n;
    }
    if (mask & 4) {
        if (n == 0) { S.a = p2; S.ia = i2; } else if (n == 1) { S.b = p2; S.ib = i2; } else { S.c = p2; S.ic = i2; }
        ++n;
    }
    S.n = n;
}

enum { JG_GJK_RUNNING = -1 };

struct GjkState {
    Simplex S;
    V3 v;        // current closest point of the simplex to the origin (separating direction)
    float vv;    // |v|^2 (FLT_MAX before the first vertex)
    float hmin;  // smallest support value h(d) = max_x d.x over the unit directions tried: if the sets overlap, an upper
                 // bound of their penetration depth (free by-product of the iterations)
    int it;
};

__device__ __forceinline__ void gjk_init(GjkState& G, V3 v0) {
    G.S.n = 0;
    G.S.a = G.S.b = G.S.c = G.S.d = mk(0.f, 0.f, 0.f);
    G.S.ia = G.S.ib = G.S.ic = G.S.id = 0xFFFFFFFFu;
    G.v = v0;
    if (!(norm2(G.v) > 1e-12f)) G.v = mk(1.f, 0.f, 0.f);
    G.vv = FLT_MAX;
    G.hmin = FLT_MAX;
    G.it = 0;
}

// One GJK iteration: one support evaluation of A - B plus the simplex update.  Returns JG_GJK_RUNNING or the
// final status; for JG_GJK_SEPARATED the distance is sqrtf(G.vv).
// support oracle of the static set B: a shared-memory vertex range (SetSupport) or a triangle held in registers
struct SetSupport {
    const float4* B; int nB;
    __device__ __forceinline__ V3 operator()(V3 d, int& idx) const {
        idx = support_scan(B, nB, d);
        const float4 p = B[idx];
        return mk(p.x, p.y, p.z);
    }
};
struct TriSupport {
    V3 p0, p1, p2;
    __device__ __forceinline__ V3 operator()(V3 d, int& idx) const {
        JG_WORK(JG_WORK_DOTS, 3u);
        const float s0 = dot(p0, d), s1 = dot(p1, d), s2 = dot(p2, d);
        idx = (s1 > s0) ? (s2 > s1 ? 2 : 1) : (s2 > s0 ? 2 : 0);
        return idx == 0 ? p0 : (idx == 1 ? p1 : p2);
    }
};

template <class SupB>
__device__ __forceinline__ int gjk_step(GjkState& G, const float4* __restrict__ A, int nA, const SupB& supB, const Tf& T,
                                        float dmax) {
    Simplex& S = G.S;
    const V3 v = G.v;
    JG_WORK(JG_WORK_GJK_ITERS, 1u);
    // support point of A - B in direction -v
    const V3 dl = tf_rot_t(T, -v);
    const int ia = support_scan(A, nA, dl);
    int ib;
    const V3 pb = supB(v, ib);
    const float4 pa = A[ia];
    const V3 w = tf_apply(T, mk(pa.x, pa.y, pa.z)) - pb;
    const uint32_t id = ((uint32_t)ia << 16) | (uint32_t)ib;
    const float vw = dot(v, w);
    const float v2 = norm2(v);
    G.it += 1;
    G.hmin = fminf(G.hmin, -vw * rsqrtf(v2));
    if (vw > 0.f && vw * vw > dmax * dmax * v2) return JG_GJK_FAR;
    if (S.n > 0) {
        const bool dupl = (id == S.ia) || (S.n > 1 && id == S.ib) || (S.n > 2 && id == S.ic);
        if (dupl || G.vv - vw <= JG_GJK_TOL * sqrtf(G.vv)) return JG_GJK_SEPARATED;
    }
    // push w as the newest vertex
    S.d = S.c; S.id = S.ic;
    S.c = S.b; S.ic = S.ib;
    S.b = S.a; S.ib = S.ia;
    S.a = w;   S.ia = id;
    S.n += 1;
    V3 nv;
    if (S.n == 1) {
        nv = w;
    } else if (S.n == 2) {
        const V3 ab = S.b - S.a;
        const float t = -dot(S.a, ab), den = norm2(ab);
        if (t <= 0.f || !(den > 0.f)) { nv = S.a; S.n = 1; }
        else if (t >= den) { nv = S.b; S.a = S.b; S.ia = S.ib; S.n = 1; }
        else nv = S.a + ab * (t / den);
    } else if (S.n == 3) {
        int mask;
        nv = tri_closest(S.a, S.b, S.c, mask);
        if (mask != 7) simplex_set3(S, S.a, S.b, S.c, S.ia, S.ib, S.ic, mask);
    } else {
        // tetrahedron (a,b,c,d): the closest feature always contains the newest vertex a, so only the
        // three faces through a are candidates; the origin is inside iff it is inside all three.
        const V3 A0 = S.a, B0 = S.b, C0 = S.c, D0 = S.d;
        const uint32_t ja = S.ia, jb = S.ib, jc = S.ic, jd = S.id;
        float best = FLT_MAX;
        bool any_out = false;
        nv = mk(0.f, 0.f, 0.f);
#pragma unroll
        for (int f = 0; f < 3; ++f) {
            const V3 x = f == 0 ? B0 : (f == 1 ? C0 : D0);
            const V3 y = f == 0 ? C0 : (f == 1 ? D0 : B0);
            const V3 z = f == 0 ? D0 : (f == 1 ? B0 : C0);
            const uint32_t jx = f == 0 ? jb : (f == 1 ? jc : jd);
            const uint32_t jy = f == 0 ? jc : (f == 1 ? jd : jb);
            const V3 nrm = cross(x - A0, y - A0);
            const float sp = -dot(A0, nrm);
            const float sd = dot(z - A0, nrm);
            // a flat tetrahedron (the support point repeats a simplex point through another vertex pair, or lies in the
            // plane of the others) has sd = rounding noise of either sign: never evidence that the origin is inside
            const bool flat = sd * sd <= 1e-12f * norm2(nrm) * norm2(z - A0);
            const bool outside = flat || !(sp * sd > 0.f);
            if (outside) {
                any_out = true;
                int mask;
                const V3 cand = tri_closest(A0, x, y, mask);
                const float c2 = norm2(cand);
                if (c2 < best) {
                    best = c2;
                    nv = cand;
                    simplex_set3(S, A0, x, y, ja, jx, jy, mask);
                }
            }
        }
        if (!any_out) { G.vv = 0.f; return JG_GJK_PENETRATING; }
    }
    const float nvv = norm2(nv);
    if (!(nvv > 1e-14f)) { G.vv = 0.f; return JG_GJK_PENETRATING; }   // touching: treat as overlap
    if (nvv >= G.vv) return JG_GJK_SEPARATED;                         // numerical stall
    G.v = nv;
    G.vv = nvv;
    return G.it >= JG_GJK_MAX_ITER ? (int)JG_GJK_SEPARATED : (int)JG_GJK_RUNNING;
}

// Distance between conv(A) placed by T and conv(B) (B frame).  v0: initial separating direction guess.
// Returns JG_GJK_*; dist is valid for JG_GJK_SEPARATED; S receives the end simplex.
__device__ __forceinline__ int gjk_distance(const float4* __restrict__ A, int nA, const float4* __restrict__ B, int nB,
                                            const Tf& T, V3 v0, float dmax, float& dist, int& iters, Simplex& S) {
    GjkState G;
    gjk_init(G, v0);
    int status;
    do {
        status = gjk_step(G, A, nA, SetSupport{B, nB}, T, dmax);
    } while (status == JG_GJK_RUNNING);
    iters = G.it;
    dist = sqrtf(G.vv);
    S = G.S;
    return status;
}

// lowest signed distance of the vertices of A (placed by T) to the plane n.p + d = 0
__device__ __forceinline__ float plane_distance(const float4* __restrict__ A, int nA, const Tf& T, V3 n, float d) {
    const V3 nl = tf_rot_t(T, n);
    float best = FLT_MAX;
#pragma unroll 4
    for (int i = 0; i < nA; ++i) {
        const float4 p = A[i];
        best = fminf(best, fmaf(p.x, nl.x, fmaf(p.y, nl.y, p.z * nl.z)));
    }
    return best + fmaf(n.x, T.tx, fmaf(n.y, T.ty, n.z * T.tz)) + d;
}

}  // namespace jg
