// jg_epa.cuh -- per-thread expanding-polytope penetration depth for two overlapping convex vertex sets.
//
// Runs only on pairs whose GJK cores overlap (Bullet: btGjkEpaSolver2::Penetration on the margin-inflated
// shapes; restated in oracle/oracle.c epa()).  Like the GJK, every lane of a warp works on the same two vertex
// sets, so the support scans are shared-memory broadcasts; the polytope (<= EPA_MAX_V vertices, <= EPA_MAX_F
// faces) lives in per-thread local memory.
#pragma once
#include "jg_gjk.cuh"

namespace jg {

#ifndef JG_EPA_MAX_V
#define JG_EPA_MAX_V 40
#endif
#ifndef JG_EPA_MAX_F
#define JG_EPA_MAX_F 80
#endif
#ifndef JG_EPA_MAX_ITER
#define JG_EPA_MAX_ITER 36
#endif
// expansion stops when the support point gains less than TOL over the closest face; a face counts as visible from
// the new vertex only beyond VIS_EPS (< TOL).  fp32 face normals are good to ~1e-6 rad, i.e. ~1e-6 m of plane-evaluation
// noise at the ~1 m extent of the Minkowski difference: faces coplanar with the new vertex must stay, otherwise the
// horizon cone folds over itself (the flood fill below keeps such folds local; folded faces are never selected).  At
// convergence the true depth lies between the face offset (lower bound) and the support value h(n) = n.w (upper bound):
// the midpoint is reported.
#ifndef JG_EPA_TOL
#define JG_EPA_TOL 1e-5f
#endif
#ifndef JG_EPA_VIS_EPS
#define JG_EPA_VIS_EPS 2e-6f
#endif

// support point of A - B in direction d (world / B frame)
__device__ __forceinline__ V3 support_md(const float4* __restrict__ A, int nA, const float4* __restrict__ B, int nB,
                                         const Tf& T, V3 d, uint32_t& id) {
    const int ia = support_scan(A, nA, tf_rot_t(T, d));
    const int ib = support_scan(B, nB, -d);
    const float4 pa = A[ia], pb = B[ib];
    id = ((uint32_t)ia << 16) | (uint32_t)ib;
    return tf_apply(T, mk(pa.x, pa.y, pa.z)) - mk(pb.x, pb.y, pb.z);
}

struct EpaPoly {
    V3 P[JG_EPA_MAX_V];
    uint32_t pid[JG_EPA_MAX_V];
    float fnx[JG_EPA_MAX_F], fny[JG_EPA_MAX_F], fnz[JG_EPA_MAX_F], fd[JG_EPA_MAX_F];
    uint32_t fv[JG_EPA_MAX_F];   // a | b<<8 | c<<16 | alive<<24 | mark<<25 (iteration that visited the face)
    uint32_t fa[JG_EPA_MAX_F];   // adjacency, 10 bits per edge k (c[k]->c[k+1]): neighbour face | neighbour edge << 8
    int nv, nf;
};

__device__ __forceinline__ int epa_vert(uint32_t fv, int k) { return (int)((fv >> (8 * k)) & 255u); }
__device__ __forceinline__ uint32_t epa_adj(uint32_t fa, int k) { return (fa >> (10 * k)) & 1023u; }
__device__ __forceinline__ void epa_bind(EpaPoly& E, int f0, int e0, int f1, int e1) {
    E.fa[f0] = (E.fa[f0] & ~(1023u << (10 * e0))) | (((uint32_t)f1 | ((uint32_t)e1 << 8)) << (10 * e0));
    E.fa[f1] = (E.fa[f1] & ~(1023u << (10 * e1))) | (((uint32_t)f0 | ((uint32_t)e0 << 8)) << (10 * e1));
}

__device__ __forceinline__ void epa_make_face(EpaPoly& E, int f, int a, int b, int c) {
    const V3 n = cross(E.P[b] - E.P[a], E.P[c] - E.P[a]);
    const float l2 = norm2(n);
    E.fv[f] = (uint32_t)a | ((uint32_t)b << 8) | ((uint32_t)c << 16) | (1u << 24);
    E.fa[f] = 0u;
    float inv = 0.f, d = FLT_MAX;
    if (l2 > 1e-30f) {
        inv = rsqrtf(l2);
        d = (n.x * E.P[a].x + n.y * E.P[a].y + n.z * E.P[a].z) * inv;
        // a face that looks at the origin from the wrong side is a numerical fold-over of near-coplanar faces:
        // keep it in the mesh (topology) but never select it as the closest face
        if (d < -1e-5f) d = FLT_MAX;
    }
    E.fnx[f] = n.x * inv; E.fny[f] = n.y * inv; E.fnz[f] = n.z * inv;
    E.fd[f] = d;
}

// Grow a lower-dimensional GJK end simplex (origin on it) to a tetrahedron; false if the difference is flat.
__device__ __forceinline__ bool epa_blow_up(const float4* __restrict__ A, int nA, const float4* __restrict__ B, int nB,
                                            const Tf& T, Simplex& S) {
    if (S.n == 1) {
#pragma unroll 1
        for (int k = 0; k < 6 && S.n == 1; ++k) {
            const V3 ax = mk(k == 0 ? 1.f : (k == 1 ? -1.f : 0.f), k == 2 ? 1.f : (k == 3 ? -1.f : 0.f),
                             k == 4 ? 1.f : (k == 5 ? -1.f : 0.f));
            uint32_t id;
            const V3 w = support_md(A, nA, B, nB, T, ax, id);
            if (norm2(w - S.a) > 1e-12f) { S.b = w; S.ib = id; S.n = 2; }
        }
        if (S.n == 1) return false;
    }
    if (S.n == 2) {
        const V3 ab = S.b - S.a;
        const V3 ax = (fabsf(ab.x) <= fabsf(ab.y) && fabsf(ab.x) <= fabsf(ab.z)) ? mk(1.f, 0.f, 0.f)
                      : (fabsf(ab.y) <= fabsf(ab.z) ? mk(0.f, 1.f, 0.f) : mk(0.f, 0.f, 1.f));
        V3 d = cross(ab, ax);
        const V3 u = ab * rsqrtf(norm2(ab));
#pragma unroll 1
        for (int k = 0; k < 6 && S.n == 2; ++k) {
            uint32_t id;
            const V3 w = support_md(A, nA, B, nB, T, d, id);
            if (norm2(cross(ab, w - S.a)) > 1e-14f * norm2(ab)) { S.c = w; S.ic = id; S.n = 3; break; }
            d = d * 0.5f + cross(u, d) * 0.8660254f;   // rotate 60 degrees about the segment
        }
        if (S.n == 2) return false;
    }
    if (S.n == 3) {
        const V3 n = cross(S.b - S.a, S.c - S.a);
        const float tol = 1e-6f * sqrtf(norm2(n));
        uint32_t id;
        V3 w = support_md(A, nA, B, nB, T, n, id);
        if (fabsf(dot(w - S.a, n)) <= tol) {
            w = support_md(A, nA, B, nB, T, -n, id);
            if (fabsf(dot(w - S.a, n)) <= tol) return false;
        }
        S.d = w; S.id = id; S.n = 4;
    }
    return true;
}

// Penetration depth of overlapping sets (origin inside conv(A - B)) by an expanding polytope with edge adjacency and
// a flood fill of the faces visible from each new support point (depth-first from the closest face, children in
// edge order, as in Bullet's btGjkEpa2 expand()): the removed region is edge-connected by construction, so its
// horizon is one closed loop even when near-coplanar faces are classified inconsistently by fp32 arithmetic.
//
// Every closest-face offset is a lower bound of the depth (the polytope lies inside the difference) and every
// support value h(n) an upper bound; the state keeps the best of both so a late numerical hiccup cannot lose them.
struct EpaState {
    EpaPoly E;
    float lb, ub;
    int it;
};

// S = GJK end simplex (origin inside or on it).  Returns false when the difference is flat (depth 0, nothing to do).
__device__ __forceinline__ bool epa_begin(EpaState& X, const float4* __restrict__ A, int nA,
                                          const float4* __restrict__ B, int nB, const Tf& T, Simplex S) {
    X.lb = 0.f;
    X.ub = FLT_MAX;
    X.it = 0;
    if (S.n < 4 && !epa_blow_up(A, nA, B, nB, T, S)) { X.ub = 0.f; return false; }
    EpaPoly& E = X.E;
    E.P[0] = S.a; E.P[1] = S.b; E.P[2] = S.c; E.P[3] = S.d;
    E.pid[0] = S.ia; E.pid[1] = S.ib; E.pid[2] = S.ic; E.pid[3] = S.id;
    // orient so that face (0,1,2) looks away from vertex 3
    if (dot(cross(E.P[1] - E.P[0], E.P[2] - E.P[0]), E.P[3] - E.P[0]) > 0.f) {
        const V3 tp = E.P[0]; E.P[0] = E.P[1]; E.P[1] = tp;
        const uint32_t ti = E.pid[0]; E.pid[0] = E.pid[1]; E.pid[1] = ti;
    }
    E.nv = 4;
    E.nf = 4;
    epa_make_face(E, 0, 0, 1, 2);
    epa_make_face(E, 1, 1, 0, 3);
    epa_make_face(E, 2, 2, 1, 3);
    epa_make_face(E, 3, 0, 2, 3);
    epa_bind(E, 0, 0, 1, 0); epa_bind(E, 0, 1, 2, 0); epa_bind(E, 0, 2, 3, 0);
    epa_bind(E, 1, 1, 3, 2); epa_bind(E, 1, 2, 2, 1); epa_bind(E, 2, 2, 3, 1);
    return true;
}

// One expansion (one support evaluation).  Returns true when finished.
__device__ __forceinline__ bool epa_step(EpaState& X, const float4* __restrict__ A, int nA, const float4* __restrict__ B,
                                         int nB, const Tf& T) {
    EpaPoly& E = X.E;
    int best = -1;
    float bd = FLT_MAX;
    for (int f = 0; f < E.nf; ++f)
        if (((E.fv[f] >> 24) & 1u) && E.fd[f] < bd) { bd = E.fd[f]; best = f; }
    if (best < 0) return true;
    X.lb = fmaxf(X.lb, bd);
    const V3 n = mk(E.fnx[best], E.fny[best], E.fnz[best]);
    uint32_t id;
    const V3 w = support_md(A, nA, B, nB, T, n, id);
    X.ub = fminf(X.ub, dot(n, w));
    X.it += 1;
    if (X.ub - X.lb <= JG_EPA_TOL || X.it >= JG_EPA_MAX_ITER) return true;
    bool dupl = false;
    for (int i = 0; i < E.nv; ++i) dupl |= (E.pid[i] == id);
    if (dupl || E.nv >= JG_EPA_MAX_V) return true;
    const int vi = E.nv;
    // ---- flood fill of the visible faces, horizon edges in loop order
    const uint32_t mark = (uint32_t)X.it;
    uint16_t stack[JG_EPA_MAX_F];      // face | edge << 8
    uint16_t horizon[JG_EPA_MAX_F];
    uint8_t removed[JG_EPA_MAX_F];
    int sp = 0, nh = 0, nr = 0;
    E.fv[best] = (E.fv[best] & 0x01FFFFFFu) | (mark << 25);
    removed[nr++] = (uint8_t)best;
    {
        const uint32_t fa = E.fa[best];
        stack[sp++] = (uint16_t)epa_adj(fa, 2);
        stack[sp++] = (uint16_t)epa_adj(fa, 1);
        stack[sp++] = (uint16_t)epa_adj(fa, 0);
    }
    while (sp > 0) {
        const uint32_t fe = stack[--sp];
        const int f = (int)(fe & 255u), e = (int)(fe >> 8);
        const uint32_t fv = E.fv[f];
        if ((fv >> 25) == mark) return true;   // reached twice: inconsistent classification, keep the bracket
        const bool vis = !(E.fd[f] < FLT_MAX) ||
                         fmaf(E.fnx[f], w.x, fmaf(E.fny[f], w.y, E.fnz[f] * w.z)) - E.fd[f] > JG_EPA_VIS_EPS;
        if (!vis) {
            horizon[nh++] = (uint16_t)fe;
        } else {
            E.fv[f] = (fv & 0x01FFFFFFu) | (mark << 25);
            removed[nr++] = (uint8_t)f;
            const uint32_t fa = E.fa[f];
            const int e1 = e == 2 ? 0 : e + 1, e2 = e == 0 ? 2 : e - 1;
            if (sp + 2 > JG_EPA_MAX_F) return true;
            stack[sp++] = (uint16_t)epa_adj(fa, e2);
            stack[sp++] = (uint16_t)epa_adj(fa, e1);
        }
    }
    if (nh < 3 || nh > nr + (JG_EPA_MAX_F - E.nf)) return true;
    E.P[vi] = w;
    E.pid[vi] = id;
    E.nv = vi + 1;
    int first = -1, prev = -1;
    for (int j = 0; j < nh; ++j) {
        const int f = (int)(horizon[j] & 255u), e = (int)(horizon[j] >> 8);
        const int e1 = e == 2 ? 0 : e + 1;
        const uint32_t fv = E.fv[f];
        const int slot = nr > 0 ? (int)removed[--nr] : E.nf++;
        epa_make_face(E, slot, epa_vert(fv, e1), epa_vert(fv, e), vi);
        epa_bind(E, slot, 0, f, e);
        if (prev >= 0) epa_bind(E, prev, 1, slot, 2); else first = slot;
        prev = slot;
    }
    epa_bind(E, prev, 1, first, 2);
    while (nr > 0) E.fv[removed[--nr]] &= ~(1u << 24);   // the remaining removed faces die
    return false;
}

// converged: midpoint of the bracket; otherwise (iteration / size caps, inconsistent mesh) the bracket is still valid:
// its midpoint when tight, else the guaranteed lower bound
__device__ __forceinline__ float epa_result(const EpaState& X) {
    if (X.ub < FLT_MAX && X.ub - X.lb <= 1e-4f) return fmaxf(0.5f * (X.lb + X.ub), 0.f);
    return X.lb;
}

__device__ __noinline__ float epa_depth(const float4* __restrict__ A, int nA, const float4* __restrict__ B, int nB,
                                        const Tf& T, Simplex S, int& iters) {
    EpaState X;
    iters = 0;
    if (!epa_begin(X, A, nA, B, nB, T, S)) return 0.f;
    while (!epa_step(X, A, nA, B, nB, T)) {}
    iters = X.it;
    return epa_result(X);
}

}  // namespace jg
