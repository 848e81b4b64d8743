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
#define JG_EPA_MAX_V 48
#endif
#ifndef JG_EPA_MAX_F
#define JG_EPA_MAX_F 96
#endif
#ifndef JG_EPA_MAX_ITER
#define JG_EPA_MAX_ITER 44
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

// Support oracle used by the polytope code: one thread scans both vertex sets (thread-per-pair kernels, emulation).
// The overflow kernel plugs in a warp-cooperative one with the same call signature.
struct ThreadSupport {
    const float4* A; int nA;
    const float4* B; int nB;
    Tf T;
    __device__ __forceinline__ V3 operator()(V3 d, uint32_t& id) const { return support_md(A, nA, B, nB, T, d, id); }
    __device__ __forceinline__ unsigned work_unit() const { return 1u; }   // (instrumented build) one expansion per thread
};

// Polytope storage.  All state lives in 32-bit words addressed as w[k * STRIDE]: STRIDE = 1 for a per-thread array
// (host emulation, overflow kernel: local memory) and STRIDE = threads-per-block for the fast kernel, where the block's
// polytopes are interleaved in SHARED memory (lanes that touch the same logical word hit consecutive banks).
// Face normals are not stored: they are recomputed from the three vertices when needed (closest face once per
// expansion, a handful of flood-fill visits), which halves the footprint.
//   V vertices: P (3 words) + id (1);  F faces: offset d, fv (a | b<<8 | c<<16 | alive<<24 | mark<<25),
//   fa (10 bits per edge k = c[k]->c[k+1]: neighbour face | neighbour edge << 8);  then flood-fill scratch.
template <int V_, int F_, int S_, int H_, int R_, int STRIDE>
struct EpaPolyT {
    static constexpr int V = V_, F = F_, S = S_, H = H_, R = R_;
    static constexpr int oP = 0, oId = 3 * V, oFd = 4 * V, oFv = 4 * V + F, oFa = 4 * V + 2 * F, oSt = 4 * V + 3 * F,
                         oHz = oSt + S, oRm = oHz + H, WORDS = oRm + R;
    uint32_t* w;
    int nv, nf;
    __device__ __forceinline__ uint32_t& at(int k) const { return w[(size_t)k * STRIDE]; }
    __device__ __forceinline__ V3 P(int i) const {
        return mk(__uint_as_float(at(oP + 3 * i)), __uint_as_float(at(oP + 3 * i + 1)), __uint_as_float(at(oP + 3 * i + 2)));
    }
    __device__ __forceinline__ void setP(int i, V3 p, uint32_t id) const {
        at(oP + 3 * i) = __float_as_uint(p.x); at(oP + 3 * i + 1) = __float_as_uint(p.y); at(oP + 3 * i + 2) = __float_as_uint(p.z);
        at(oId + i) = id;
    }
    __device__ __forceinline__ uint32_t pid(int i) const { return at(oId + i); }
    __device__ __forceinline__ float fd(int f) const { return __uint_as_float(at(oFd + f)); }
    __device__ __forceinline__ void set_fd(int f, float d) const { at(oFd + f) = __float_as_uint(d); }
    __device__ __forceinline__ uint32_t& fv(int f) const { return at(oFv + f); }
    __device__ __forceinline__ uint32_t& fa(int f) const { return at(oFa + f); }
    __device__ __forceinline__ uint32_t& st(int i) const { return at(oSt + i); }
    __device__ __forceinline__ uint32_t& hz(int i) const { return at(oHz + i); }
    __device__ __forceinline__ uint32_t& rm(int i) const { return at(oRm + i); }
};

__device__ __forceinline__ int epa_vert(uint32_t fv, int k) { return (int)((fv >> (8 * k)) & 255u); }
__device__ __forceinline__ uint32_t epa_adj(uint32_t fa, int k) { return (fa >> (10 * k)) & 1023u; }

template <class E>
__device__ __forceinline__ void epa_bind(const E& X, int f0, int e0, int f1, int e1) {
    X.fa(f0) = (X.fa(f0) & ~(1023u << (10 * e0))) | (((uint32_t)f1 | ((uint32_t)e1 << 8)) << (10 * e0));
    X.fa(f1) = (X.fa(f1) & ~(1023u << (10 * e1))) | (((uint32_t)f0 | ((uint32_t)e0 << 8)) << (10 * e1));
}

// unit outward normal of face f (zero vector for a degenerate sliver)
template <class E>
__device__ __forceinline__ V3 epa_normal(const E& X, uint32_t fv) {
    const V3 a = X.P(epa_vert(fv, 0));
    const V3 n = cross(X.P(epa_vert(fv, 1)) - a, X.P(epa_vert(fv, 2)) - a);
    const float l2 = norm2(n);
    return l2 > 1e-30f ? n * rsqrtf(l2) : mk(0.f, 0.f, 0.f);
}

template <class E>
__device__ __forceinline__ void epa_make_face(const E& X, int f, int a, int b, int c) {
    const V3 pa = X.P(a);
    const V3 n = cross(X.P(b) - pa, X.P(c) - pa);
    const float l2 = norm2(n);
    X.fv(f) = (uint32_t)a | ((uint32_t)b << 8) | ((uint32_t)c << 16) | (1u << 24);
    X.fa(f) = 0u;
    float d = FLT_MAX;
    if (l2 > 1e-30f) {
        d = dot(n, pa) * rsqrtf(l2);
        // a face that looks at the origin from the wrong side is a numerical fold-over of near-coplanar faces:
        // keep it in the mesh (topology) but never select it as the closest face
        if (d < -1e-5f) d = FLT_MAX;
    }
    X.set_fd(f, d);
}

// Grow a lower-dimensional GJK end simplex (origin on it) to a tetrahedron; false if the difference is flat.
template <class Sup>
__device__ __forceinline__ bool epa_blow_up(const Sup& sup, Simplex& S) {
    if (S.n == 1) {
#pragma unroll 1
        for (int k = 0; k < 6 && S.n == 1; ++k) {
            const V3 ax = mk(k == 0 ? 1.f : (k == 1 ? -1.f : 0.f), k == 2 ? 1.f : (k == 3 ? -1.f : 0.f),
                             k == 4 ? 1.f : (k == 5 ? -1.f : 0.f));
            uint32_t id;
            const V3 w = sup(ax, id);
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
            const V3 w = sup(d, id);
            if (norm2(cross(ab, w - S.a)) > 1e-14f * norm2(ab)) { S.c = w; S.ic = id; S.n = 3; break; }
            d = d * 0.5f + cross(u, d) * 0.8660254f;   // rotate 60 degrees about the segment
        }
        if (S.n == 2) return false;
    }
    if (S.n == 3) {
        const V3 n = cross(S.b - S.a, S.c - S.a);
        const float tol = 1e-6f * sqrtf(norm2(n));
        uint32_t id;
        V3 w = sup(n, id);
        if (fabsf(dot(w - S.a, n)) <= tol) {
            w = sup(-n, id);
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
enum { JG_EPA_RUNNING = 0, JG_EPA_DONE = 1, JG_EPA_OVERFLOW = 2 };

template <class E>
struct EpaStateT {
    E poly;
    float lb, ub;
    int it;
};

// S = GJK end simplex (origin inside or on it).  Returns false when the difference is flat (depth 0, nothing to do).
template <class E, class Sup>
__device__ __forceinline__ bool epa_begin(EpaStateT<E>& X, const Sup& sup, Simplex S) {
    X.lb = 0.f;
    X.ub = FLT_MAX;
    X.it = 0;
    if (S.n < 4 && !epa_blow_up(sup, S)) { X.ub = 0.f; return false; }
    E& P = X.poly;
    // orient so that face (0,1,2) looks away from vertex 3
    if (dot(cross(S.b - S.a, S.c - S.a), S.d - S.a) > 0.f) {
        const V3 tp = S.a; S.a = S.b; S.b = tp;
        const uint32_t ti = S.ia; S.ia = S.ib; S.ib = ti;
    }
    P.setP(0, S.a, S.ia); P.setP(1, S.b, S.ib); P.setP(2, S.c, S.ic); P.setP(3, S.d, S.id);
    P.nv = 4;
    P.nf = 4;
    epa_make_face(P, 0, 0, 1, 2);
    epa_make_face(P, 1, 1, 0, 3);
    epa_make_face(P, 2, 2, 1, 3);
    epa_make_face(P, 3, 0, 2, 3);
    epa_bind(P, 0, 0, 1, 0); epa_bind(P, 0, 1, 2, 0); epa_bind(P, 0, 2, 3, 0);
    epa_bind(P, 1, 1, 3, 2); epa_bind(P, 1, 2, 2, 1); epa_bind(P, 2, 2, 3, 1);
    return true;
}

// One expansion (one support evaluation).  JG_EPA_OVERFLOW: the polytope or the flood-fill scratch outgrew this
// storage class (nothing was modified in that step): the pair must be redone with a larger one.
template <class E, class Sup>
__device__ __forceinline__ int epa_step(EpaStateT<E>& X, const Sup& sup) {
    E& P = X.poly;
    int best = -1;
    float bd = FLT_MAX;
    for (int f = 0; f < P.nf; ++f) {
        const float d = P.fd(f);     // dead faces carry FLT_MAX
        if (d < bd) { bd = d; best = f; }
    }
    if (best < 0) return JG_EPA_DONE;
    X.lb = fmaxf(X.lb, bd);
    const V3 n = epa_normal(P, P.fv(best));
    uint32_t id;
    const V3 w = sup(n, id);
    X.ub = fminf(X.ub, dot(n, w));
    X.it += 1;
    JG_WORK(JG_WORK_EPA_ITERS, sup.work_unit());
    if (X.ub - X.lb <= JG_EPA_TOL || X.it >= JG_EPA_MAX_ITER) return JG_EPA_DONE;
    bool dupl = false;
    for (int i = 0; i < P.nv; ++i) dupl |= (P.pid(i) == id);
    if (dupl) return JG_EPA_DONE;
    if (P.nv >= E::V) return JG_EPA_OVERFLOW;
    const int vi = P.nv;
    // ---- flood fill of the visible faces, horizon edges in loop order
    const uint32_t mark = (uint32_t)X.it;
    int sp = 0, nh = 0, nr = 0;
    bool ok = true, over = false;
    P.fv(best) = (P.fv(best) & 0x01FFFFFFu) | (mark << 25);
    P.rm(nr++) = (uint32_t)best;
    {
        const uint32_t fa = P.fa(best);
        P.st(sp++) = epa_adj(fa, 2);
        P.st(sp++) = epa_adj(fa, 1);
        P.st(sp++) = epa_adj(fa, 0);
    }
    while (sp > 0) {
        const uint32_t fe = P.st(--sp);
        const int f = (int)(fe & 255u), e = (int)(fe >> 8);
        const uint32_t fv = P.fv(f);
        if ((fv >> 25) == mark) { ok = false; break; }   // reached twice: inconsistent classification
        bool vis = !(P.fd(f) < FLT_MAX);
        if (!vis) {   // w above the face plane by more than VIS_EPS (no normalisation: compare squares)
            const V3 pa = P.P(epa_vert(fv, 0));
            const V3 nu = cross(P.P(epa_vert(fv, 1)) - pa, P.P(epa_vert(fv, 2)) - pa);
            const float t = dot(nu, w - pa);
            vis = t > 0.f && t * t > (JG_EPA_VIS_EPS * JG_EPA_VIS_EPS) * norm2(nu);
        }
        if (!vis) {
            if (nh >= E::H) { over = true; break; }
            P.hz(nh++) = fe;
        } else {
            if (nr >= E::R || sp + 2 > E::S) { over = true; break; }
            P.fv(f) = (fv & 0x01FFFFFFu) | (mark << 25);
            P.rm(nr++) = (uint32_t)f;
            const uint32_t fa = P.fa(f);
            const int e1 = e == 2 ? 0 : e + 1, e2 = e == 0 ? 2 : e - 1;
            P.st(sp++) = epa_adj(fa, e2);
            P.st(sp++) = epa_adj(fa, e1);
        }
    }
    if (!over && ok && nh >= 3 && nh > nr + (E::F - P.nf)) over = true;
    if (over) {   // undo the marks of this step so that a restart / larger storage sees a clean polytope
        for (int i = 0; i < nr; ++i) P.fv((int)P.rm(i)) &= 0x01FFFFFFu;
        return JG_EPA_OVERFLOW;
    }
    if (!ok || nh < 3) return JG_EPA_DONE;   // keep the bracket
    P.setP(vi, w, id);
    P.nv = vi + 1;
    // slots of the new faces: recycle the removed ones, then grow.  hz(j) |= slot << 16
    {
        int nfree = nr;
        for (int j = 0; j < nh; ++j) {
            const int slot = nfree > 0 ? (int)P.rm(--nfree) : P.nf++;
            P.hz(j) |= (uint32_t)slot << 16;
        }
        while (nfree > 0) {   // the remaining removed faces die
            const int f = (int)P.rm(--nfree);
            P.fv(f) &= ~(1u << 24);
            P.set_fd(f, FLT_MAX);
        }
    }
    // new face j = (a_j, b_j, w): edge 0 <-> outside face, edge 1 <-> face j+1 (its edge 2), edge 2 <-> face j-1
    // (its edge 1).  Consecutive horizon faces share a vertex (b_j == a_{j+1}), so each face loads one vertex.
    {
        const uint32_t h0 = P.hz(0);
        uint32_t prev_slot = P.hz(nh - 1) >> 16;
        uint32_t hcur = h0;
        int f = (int)(hcur & 255u), e = (int)((hcur >> 8) & 3u);
        uint32_t fvo = P.fv(f);
        int ia = epa_vert(fvo, e == 2 ? 0 : e + 1);
        V3 pa = P.P(ia);
        for (int j = 0; j < nh; ++j) {
            const uint32_t hnext = j + 1 < nh ? P.hz(j + 1) : h0;
            const uint32_t slot = hcur >> 16, next_slot = hnext >> 16;
            const int ib = epa_vert(fvo, e);
            const V3 pb = P.P(ib);
            const V3 n = cross(pb - pa, w - pa);
            const float l2 = norm2(n);
            float d = FLT_MAX;
            if (l2 > 1e-30f) {
                d = dot(n, pa) * rsqrtf(l2);
                if (d < -1e-5f) d = FLT_MAX;   // folded face (see epa_make_face)
            }
            P.set_fd((int)slot, d);
            P.fv((int)slot) = (uint32_t)ia | ((uint32_t)ib << 8) | ((uint32_t)vi << 16) | (1u << 24);
            P.fa((int)slot) = ((uint32_t)f | ((uint32_t)e << 8)) | ((next_slot | (2u << 8)) << 10) | ((prev_slot | (1u << 8)) << 20);
            P.fa(f) = (P.fa(f) & ~(1023u << (10 * e))) | (slot << (10 * e));   // outside face now borders (slot, edge 0)
            prev_slot = slot;
            hcur = hnext;
            f = (int)(hcur & 255u); e = (int)((hcur >> 8) & 3u);
            fvo = P.fv(f);
            ia = ib;
            pa = pb;
        }
    }
    return JG_EPA_RUNNING;
}

// converged: midpoint of the bracket; otherwise (iteration caps, inconsistent mesh) the bracket is still valid:
// its midpoint when tight, else the guaranteed lower bound
template <class E>
__device__ __forceinline__ float epa_result(const EpaStateT<E>& X) {
    if (X.ub < FLT_MAX && X.ub - X.lb <= 1e-4f) return fmaxf(0.5f * (X.lb + X.ub), 0.f);
    return X.lb;
}

// ---------------------------------------------------------------------------------------------- compact polytope
// The fast kernel is bound by how many pairs fit in shared memory at once (its time scales 1/threads), so its
// polytope is packed to 472 bytes: 18 vertices (3 words each), 32 faces (2 words each: offset d and one word with
// three 5-bit vertex ids, three 5-bit neighbour faces and an alive bit).  The neighbour's edge index is not stored: it
// is found from the shared vertex when the neighbour is visited.  Visited marks and the flood-fill scratch live in
// registers.  14 expansions fit; larger polytopes overflow.
template <int STRIDE>
struct EpaCompact {
    static constexpr int V = 18, F = 32, S = 12, H = 24, R = 12;
    static constexpr int oP = 0, oFd = 3 * V, oFx = oFd + F, WORDS = oFx + F;
    uint32_t* w;
    int nv, nf;
    __device__ __forceinline__ uint32_t& at(int k) const { return w[(size_t)k * STRIDE]; }
    __device__ __forceinline__ V3 P(int i) const {
        return mk(__uint_as_float(at(oP + 3 * i)), __uint_as_float(at(oP + 3 * i + 1)), __uint_as_float(at(oP + 3 * i + 2)));
    }
    __device__ __forceinline__ void setP(int i, V3 p) const {
        at(oP + 3 * i) = __float_as_uint(p.x); at(oP + 3 * i + 1) = __float_as_uint(p.y); at(oP + 3 * i + 2) = __float_as_uint(p.z);
    }
    __device__ __forceinline__ float fd(int f) const { return __uint_as_float(at(oFd + f)); }
    __device__ __forceinline__ void set_fd(int f, float d) const { at(oFd + f) = __float_as_uint(d); }
    __device__ __forceinline__ uint32_t& fx(int f) const { return at(oFx + f); }
};

// Flood-fill scratch of one expansion, held in REGISTERS as packed shift registers (the shared-memory version spent a
// third of the kernel's instructions on 16-bit pack / unpack at 8 of 32 lanes): a 12-deep LIFO of 10-bit entries in
// two 64-bit words, the removed faces (12 x 5 bits) in one, the horizon (24 x 7 bits) in three.
struct EpaScratch {
    unsigned long long st_lo, st_hi, rm, h0, h1, h2;
    __device__ __forceinline__ void push(uint32_t e) { st_hi = (st_hi << 10) | (st_lo >> 54); st_lo = (st_lo << 10) | e; }
    __device__ __forceinline__ uint32_t pop() {
        const uint32_t e = (uint32_t)st_lo & 1023u;
        st_lo = (st_lo >> 10) | (st_hi << 54);
        st_hi >>= 10;
        return e;
    }
    __device__ __forceinline__ void rm_set(int i, uint32_t f) { rm |= (unsigned long long)f << (5 * i); }
    __device__ __forceinline__ int rm_get(int i) const { return (int)((rm >> (5 * i)) & 31ull); }
    __device__ __forceinline__ void hz_set(int i, uint32_t e) {
        const unsigned long long v = (unsigned long long)e << (7 * (i & 7));
        if (i < 8) h0 |= v; else if (i < 16) h1 |= v; else h2 |= v;
    }
    __device__ __forceinline__ uint32_t hz_get(int i) const {
        const unsigned long long wv = i < 8 ? h0 : (i < 16 ? h1 : h2);
        return (uint32_t)(wv >> (7 * (i & 7))) & 127u;
    }
};
// fx word: vertex k at bits 5k (k = 0..2), neighbour across edge k (c[k]->c[k+1]) at bits 15 + 5k, alive = bit 30
__device__ __forceinline__ int cx_vert(uint32_t fx, int k) { return (int)((fx >> (5 * k)) & 31u); }
__device__ __forceinline__ int cx_nbr(uint32_t fx, int k) { return (int)((fx >> (15 + 5 * k)) & 31u); }

template <int STRIDE>
struct EpaStateC {
    EpaCompact<STRIDE> poly;
    float lb, ub;
    int it;
};

template <int STRIDE, class Sup>
__device__ __forceinline__ bool epa_begin(EpaStateC<STRIDE>& X, const Sup& sup, Simplex S) {
    X.lb = 0.f;
    X.ub = FLT_MAX;
    X.it = 0;
    if (S.n < 4 && !epa_blow_up(sup, S)) { X.ub = 0.f; return false; }
    EpaCompact<STRIDE>& P = X.poly;
    if (dot(cross(S.b - S.a, S.c - S.a), S.d - S.a) > 0.f) { const V3 tp = S.a; S.a = S.b; S.b = tp; }
    P.setP(0, S.a); P.setP(1, S.b); P.setP(2, S.c); P.setP(3, S.d);
    P.nv = 4;
    P.nf = 4;
    // faces (0,1,2) (1,0,3) (2,1,3) (0,2,3); neighbours across edges 0,1,2 (see the generic epa_begin binds)
    const int fv[4][3] = {{0, 1, 2}, {1, 0, 3}, {2, 1, 3}, {0, 2, 3}};
    const int fn[4][3] = {{1, 2, 3}, {0, 3, 2}, {0, 1, 3}, {0, 2, 1}};
#pragma unroll
    for (int f = 0; f < 4; ++f) {
        const V3 pa = P.P(fv[f][0]);
        const V3 n = cross(P.P(fv[f][1]) - pa, P.P(fv[f][2]) - pa);
        const float l2 = norm2(n);
        float d = FLT_MAX;
        if (l2 > 1e-30f) {
            d = dot(n, pa) * rsqrtf(l2);
            if (d < -1e-5f) d = FLT_MAX;
        }
        P.set_fd(f, d);
        P.fx(f) = (uint32_t)fv[f][0] | ((uint32_t)fv[f][1] << 5) | ((uint32_t)fv[f][2] << 10) | ((uint32_t)fn[f][0] << 15) |
                  ((uint32_t)fn[f][1] << 20) | ((uint32_t)fn[f][2] << 25) | (1u << 30);
    }
    return true;
}

template <int STRIDE, class Sup>
__device__ __forceinline__ int epa_step(EpaStateC<STRIDE>& X, const Sup& sup) {
    typedef EpaCompact<STRIDE> E;
    E& P = X.poly;
    int best = -1;
    float bd = FLT_MAX;
    for (int f = 0; f < P.nf; ++f) {
        const float d = P.fd(f);     // dead faces carry FLT_MAX
        if (d < bd) { bd = d; best = f; }
    }
    if (best < 0) return JG_EPA_DONE;
    X.lb = fmaxf(X.lb, bd);
    const uint32_t fxb = P.fx(best);
    V3 n;
    {
        const V3 pa = P.P(cx_vert(fxb, 0));
        n = cross(P.P(cx_vert(fxb, 1)) - pa, P.P(cx_vert(fxb, 2)) - pa);
        n = n * rsqrtf(norm2(n));
    }
    uint32_t id;
    const V3 w = sup(n, id);
    X.ub = fminf(X.ub, dot(n, w));
    X.it += 1;
    JG_WORK(JG_WORK_EPA_ITERS, sup.work_unit());
    // (a support point that is already a polytope vertex cannot gain over the closest face: the bracket test ends it)
    if (X.ub - X.lb <= JG_EPA_TOL || X.it >= JG_EPA_MAX_ITER) return JG_EPA_DONE;
    if (P.nv >= E::V) return JG_EPA_OVERFLOW;
    const int vi = P.nv;
    // ---- flood fill of the visible faces, horizon edges in loop order.  stack entry: face | start vertex << 5
    uint32_t visited = 1u << best;
    EpaScratch sc;
    sc.st_lo = sc.st_hi = sc.rm = sc.h0 = sc.h1 = sc.h2 = 0ull;
    int sp = 0, nh = 0, nr = 0;
    bool ok = true, over = false;
    sc.rm_set(nr++, (uint32_t)best);
    sc.push((uint32_t)cx_nbr(fxb, 2) | ((uint32_t)cx_vert(fxb, 0) << 5));
    sc.push((uint32_t)cx_nbr(fxb, 1) | ((uint32_t)cx_vert(fxb, 2) << 5));
    sc.push((uint32_t)cx_nbr(fxb, 0) | ((uint32_t)cx_vert(fxb, 1) << 5));
    sp = 3;
    while (sp > 0) {
        const uint32_t se = sc.pop();
        --sp;
        const int f = (int)(se & 31u), v0 = (int)(se >> 5);
        if ((visited >> f) & 1u) { ok = false; break; }   // reached twice: inconsistent classification
        const uint32_t fx = P.fx(f);
        const int e = cx_vert(fx, 0) == v0 ? 0 : (cx_vert(fx, 1) == v0 ? 1 : 2);   // the edge we came through
        bool vis = !(P.fd(f) < FLT_MAX);
        if (!vis) {
            const V3 pa = P.P(cx_vert(fx, 0));
            const V3 nu = cross(P.P(cx_vert(fx, 1)) - pa, P.P(cx_vert(fx, 2)) - pa);
            const float t = dot(nu, w - pa);
            vis = t > 0.f && t * t > (JG_EPA_VIS_EPS * JG_EPA_VIS_EPS) * norm2(nu);
        }
        if (!vis) {
            if (nh >= E::H) { over = true; break; }
            sc.hz_set(nh++, (uint32_t)f | ((uint32_t)e << 5));
        } else {
            if (nr >= E::R || sp + 2 > E::S) { over = true; break; }
            visited |= 1u << f;
            sc.rm_set(nr++, (uint32_t)f);
            const int e1 = e == 2 ? 0 : e + 1, e2 = e == 0 ? 2 : e - 1;
            // neighbour across edge k starts that edge (in its own orientation) at our vertex k+1
            sc.push((uint32_t)cx_nbr(fx, e2) | ((uint32_t)cx_vert(fx, e2 == 2 ? 0 : e2 + 1) << 5));
            sc.push((uint32_t)cx_nbr(fx, e1) | ((uint32_t)cx_vert(fx, e1 == 2 ? 0 : e1 + 1) << 5));
            sp += 2;
        }
    }
    if (!over && ok && nh >= 3 && nh > nr + (E::F - P.nf)) over = true;
    if (over) return JG_EPA_OVERFLOW;       // nothing was modified
    if (!ok || nh < 3) return JG_EPA_DONE;   // keep the bracket
    P.setP(vi, w);
    P.nv = vi + 1;
    // slots of the new faces: recycle removed ones (last removed first), then grow; the rest of the removed die
    const int nr0 = nr, nf0 = P.nf;
    auto slot_of = [&](int j) -> int { return j < nr0 ? sc.rm_get(nr0 - 1 - j) : nf0 + (j - nr0); };
    if (nh > nr0) P.nf = nf0 + (nh - nr0);
    for (int i = 0; i < nr0 - nh; ++i) {
        const int f = sc.rm_get(i);
        P.fx(f) = 0u;
        P.set_fd(f, FLT_MAX);
    }
    {
        uint32_t hcur = sc.hz_get(0);
        const uint32_t h0 = hcur;
        int prev_slot = slot_of(nh - 1), slot = slot_of(0);
        int f = (int)(hcur & 31u), e = (int)(hcur >> 5);
        uint32_t fxo = P.fx(f);
        int ia = cx_vert(fxo, e == 2 ? 0 : e + 1);
        V3 pa = P.P(ia);
        for (int j = 0; j < nh; ++j) {
            const uint32_t hnext = j + 1 < nh ? sc.hz_get(j + 1) : h0;
            const int next_slot = j + 1 < nh ? slot_of(j + 1) : slot_of(0);
            const int ib = cx_vert(fxo, e);
            const V3 pb = P.P(ib);
            const V3 nn = cross(pb - pa, w - pa);
            const float l2 = norm2(nn);
            float d = FLT_MAX;
            if (l2 > 1e-30f) {
                d = dot(nn, pa) * rsqrtf(l2);
                if (d < -1e-5f) d = FLT_MAX;   // folded face: stays in the mesh, never selected
            }
            P.set_fd(slot, d);
            P.fx(slot) = (uint32_t)ia | ((uint32_t)ib << 5) | ((uint32_t)vi << 10) | ((uint32_t)f << 15) |
                         ((uint32_t)next_slot << 20) | ((uint32_t)prev_slot << 25) | (1u << 30);
            // outside face: its edge e now borders the new face (re-read: two horizon edges may share it)
            P.fx(f) = (P.fx(f) & ~(31u << (15 + 5 * e))) | ((uint32_t)slot << (15 + 5 * e));
            prev_slot = slot;
            slot = next_slot;
            hcur = hnext;
            f = (int)(hcur & 31u); e = (int)(hcur >> 5);
            fxo = P.fx(f);
            ia = ib;
            pa = pb;
        }
    }
    return JG_EPA_RUNNING;
}

template <int STRIDE>
__device__ __forceinline__ float epa_result(const EpaStateC<STRIDE>& X) {
    if (X.ub < FLT_MAX && X.ub - X.lb <= 1e-4f) return fmaxf(0.5f * (X.lb + X.ub), 0.f);
    return X.lb;
}

// storage classes: small = the fast kernel's shared-memory polytope; big = per-thread array (overflow kernel, tests)
typedef EpaPolyT<JG_EPA_MAX_V, JG_EPA_MAX_F, JG_EPA_MAX_F, JG_EPA_MAX_F, JG_EPA_MAX_F, 1> EpaPolyBig;

// whole query in one call on a per-thread array (overflow kernel, host emulation)
__device__ __noinline__ float epa_depth(const float4* __restrict__ A, int nA, const float4* __restrict__ B, int nB,
                                        const Tf& T, Simplex S, int& iters) {
    uint32_t words[EpaPolyBig::WORDS];
    EpaStateT<EpaPolyBig> X;
    X.poly.w = words;
    iters = 0;
    const ThreadSupport sup = {A, nA, B, nB, T};
    if (!epa_begin(X, sup, S)) return 0.f;
    while (epa_step(X, sup) == JG_EPA_RUNNING) {}
    iters = X.it;
    return epa_result(X);
}

}  // namespace jg
