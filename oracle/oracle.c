/*
 * oracle.c -- CPU restatement of the jogramop collision query (see oracle.h for scope, citations and
 * the parity-pin statement).  TEST INFRASTRUCTURE: never linked into or called from the product path.
 *
 * Plain C99, double precision, pthreads over configurations for the CPU-baseline timing
 * (libgomp is not available in this image).
 */
#include "oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>
#include <unistd.h>

/* ------------------------------------------------------------------------------------------ vec */
typedef struct { double x, y, z; } v3;
typedef struct { double m[12]; } tf; /* row-major 3x4 */

static inline v3 V(double x, double y, double z) { v3 r = {x, y, z}; return r; }
static inline v3 add(v3 a, v3 b) { return V(a.x + b.x, a.y + b.y, a.z + b.z); }
static inline v3 sub(v3 a, v3 b) { return V(a.x - b.x, a.y - b.y, a.z - b.z); }
static inline v3 mul(v3 a, double s) { return V(a.x * s, a.y * s, a.z * s); }
static inline v3 neg(v3 a) { return V(-a.x, -a.y, -a.z); }
static inline double dot(v3 a, v3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
static inline v3 cross(v3 a, v3 b) { return V(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
static inline double norm2(v3 a) { return dot(a, a); }

static const tf TF_ID = {{1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0}};

static inline v3 tf_apply(const tf* T, v3 p) {
    return V(T->m[0] * p.x + T->m[1] * p.y + T->m[2] * p.z + T->m[3],
             T->m[4] * p.x + T->m[5] * p.y + T->m[6] * p.z + T->m[7],
             T->m[8] * p.x + T->m[9] * p.y + T->m[10] * p.z + T->m[11]);
}
static inline v3 tf_rot_t(const tf* T, v3 d) { /* R^T d */
    return V(T->m[0] * d.x + T->m[4] * d.y + T->m[8] * d.z,
             T->m[1] * d.x + T->m[5] * d.y + T->m[9] * d.z,
             T->m[2] * d.x + T->m[6] * d.y + T->m[10] * d.z);
}
static inline tf tf_mul(const tf* A, const tf* B) {
    tf C;
    for (int r = 0; r < 3; ++r) {
        for (int c = 0; c < 4; ++c) {
            double s = A->m[4 * r] * B->m[c] + A->m[4 * r + 1] * B->m[4 + c] + A->m[4 * r + 2] * B->m[8 + c];
            if (c == 3) s += A->m[4 * r + 3];
            C.m[4 * r + c] = s;
        }
    }
    return C;
}
static inline tf tf_inv(const tf* A) {
    tf C;
    for (int r = 0; r < 3; ++r)
        for (int c = 0; c < 3; ++c) C.m[4 * r + c] = A->m[4 * c + r];
    for (int r = 0; r < 3; ++r)
        C.m[4 * r + 3] = -(C.m[4 * r] * A->m[3] + C.m[4 * r + 1] * A->m[7] + C.m[4 * r + 2] * A->m[11]);
    return C;
}

/* ------------------------------------------------------------------------------------- structures */
typedef struct {
    const double* v; /* [n,3] */
    int n;
    const tf* T;
} geom;

struct orc_robot {
    int L, S, K, dof;
    int *joint_type, *joint_parent, *joint_qidx;
    double *joint_origin, *joint_axis, *q_limits;
    int *shape_link, *shape_off, *plane_off, *self_pairs;
    double *shape_margin, *core, *plane_v, *plane_margin;
    double *aabb_c, *aabb_h; /* local AABB of core verts per shape [S,3] */
    int* link_shape;          /* first shape of link or -1 */
};

struct orc_scene {
    int P;
    int* off;
    double *v, *margin;
    double *aabb_lo, *aabb_hi; /* [P,3] */
    int has_plane;
    double plane[4];
};

static void* dup_mem(const void* p, size_t bytes) {
    void* q = malloc(bytes ? bytes : 1);
    if (p && bytes) memcpy(q, p, bytes);
    return q;
}

orc_robot* orc_robot_create(int L, const int* joint_type, const int* joint_parent, const double* joint_origin,
                            const double* joint_axis, const int* joint_qidx, const double* q_limits, int dof,
                            int S, const int* shape_link, const double* shape_margin, const int* shape_off,
                            const double* core, const int* plane_off, const double* plane_v,
                            const double* plane_margin, int K, const int* self_pairs) {
    orc_robot* r = (orc_robot*)calloc(1, sizeof(orc_robot));
    r->L = L; r->S = S; r->K = K; r->dof = dof;
    r->joint_type = dup_mem(joint_type, sizeof(int) * L);
    r->joint_parent = dup_mem(joint_parent, sizeof(int) * L);
    r->joint_qidx = dup_mem(joint_qidx, sizeof(int) * L);
    r->joint_origin = dup_mem(joint_origin, sizeof(double) * 12 * L);
    r->joint_axis = dup_mem(joint_axis, sizeof(double) * 3 * L);
    r->q_limits = dup_mem(q_limits, sizeof(double) * 2 * dof);
    r->shape_link = dup_mem(shape_link, sizeof(int) * S);
    r->shape_margin = dup_mem(shape_margin, sizeof(double) * S);
    r->shape_off = dup_mem(shape_off, sizeof(int) * (S + 1));
    r->core = dup_mem(core, sizeof(double) * 3 * shape_off[S]);
    r->plane_off = dup_mem(plane_off, sizeof(int) * (S + 1));
    r->plane_v = dup_mem(plane_v, sizeof(double) * 3 * plane_off[S]);
    r->plane_margin = dup_mem(plane_margin, sizeof(double) * S);
    r->self_pairs = dup_mem(self_pairs, sizeof(int) * 2 * K);
    r->aabb_c = malloc(sizeof(double) * 3 * S);
    r->aabb_h = malloc(sizeof(double) * 3 * S);
    r->link_shape = malloc(sizeof(int) * (L + 1));
    for (int i = 0; i <= L; ++i) r->link_shape[i] = -1;
    for (int s = 0; s < S; ++s) {
        double lo[3] = {1e300, 1e300, 1e300}, hi[3] = {-1e300, -1e300, -1e300};
        for (int i = shape_off[s]; i < shape_off[s + 1]; ++i)
            for (int k = 0; k < 3; ++k) {
                if (core[3 * i + k] < lo[k]) lo[k] = core[3 * i + k];
                if (core[3 * i + k] > hi[k]) hi[k] = core[3 * i + k];
            }
        for (int k = 0; k < 3; ++k) {
            r->aabb_c[3 * s + k] = 0.5 * (lo[k] + hi[k]);
            r->aabb_h[3 * s + k] = 0.5 * (hi[k] - lo[k]);
        }
        if (r->link_shape[shape_link[s] + 1] < 0) r->link_shape[shape_link[s] + 1] = s;
    }
    return r;
}

void orc_robot_free(orc_robot* r) {
    if (!r) return;
    free(r->joint_type); free(r->joint_parent); free(r->joint_qidx); free(r->joint_origin); free(r->joint_axis);
    free(r->q_limits); free(r->shape_link); free(r->shape_margin); free(r->shape_off); free(r->core);
    free(r->plane_off); free(r->plane_v); free(r->plane_margin); free(r->self_pairs); free(r->aabb_c);
    free(r->aabb_h); free(r->link_shape); free(r);
}

orc_scene* orc_scene_create(int P, const int* off, const double* v, const double* margin, const double* plane) {
    orc_scene* s = (orc_scene*)calloc(1, sizeof(orc_scene));
    s->P = P;
    s->off = dup_mem(off, sizeof(int) * (P + 1));
    s->v = dup_mem(v, sizeof(double) * 3 * (P ? off[P] : 0));
    s->margin = dup_mem(margin, sizeof(double) * P);
    s->aabb_lo = malloc(sizeof(double) * 3 * (P + 1));
    s->aabb_hi = malloc(sizeof(double) * 3 * (P + 1));
    for (int p = 0; p < P; ++p)
        for (int k = 0; k < 3; ++k) {
            double lo = 1e300, hi = -1e300;
            for (int i = off[p]; i < off[p + 1]; ++i) {
                if (v[3 * i + k] < lo) lo = v[3 * i + k];
                if (v[3 * i + k] > hi) hi = v[3 * i + k];
            }
            s->aabb_lo[3 * p + k] = lo;
            s->aabb_hi[3 * p + k] = hi;
        }
    s->has_plane = plane != NULL;
    if (plane) memcpy(s->plane, plane, sizeof(double) * 4);
    return s;
}

void orc_scene_free(orc_scene* s) {
    if (!s) return;
    free(s->off); free(s->v); free(s->margin); free(s->aabb_lo); free(s->aabb_hi); free(s);
}


/* ---------------------------------------------------------------------------------- parallel for */
typedef void (*range_fn)(int64_t begin, int64_t end, void* ctx);
typedef struct { int64_t n, grain; volatile int64_t next; range_fn fn; void* ctx; } pf_job;

static void* pf_worker(void* arg) {
    pf_job* j = (pf_job*)arg;
    for (;;) {
        int64_t b = __sync_fetch_and_add(&j->next, j->grain);
        if (b >= j->n) break;
        int64_t e = b + j->grain < j->n ? b + j->grain : j->n;
        j->fn(b, e, j->ctx);
    }
    return NULL;
}

static void parallel_for(int64_t n, int64_t grain, int n_threads, range_fn fn, void* ctx) {
    if (n_threads <= 0) {
        long c = sysconf(_SC_NPROCESSORS_ONLN);
        n_threads = c > 0 ? (int)c : 1;
    }
    if (n_threads > 256) n_threads = 256;
    if (n_threads == 1 || n <= grain) { fn(0, n, ctx); return; }
    pf_job job = {n, grain, 0, fn, ctx};
    pthread_t th[256];
    int started = 0;
    for (int t = 0; t < n_threads - 1; ++t)
        if (pthread_create(&th[started], NULL, pf_worker, &job) == 0) ++started;
    pf_worker(&job);
    for (int t = 0; t < started; ++t) pthread_join(th[t], NULL);
}

/* --------------------------------------------------------------------------------------------- FK */
/* Link i = child of joint i: T_i = T_parent * origin_i * motion_i(q)  (URDF convention, Appendix A). */
static void fk_links(const orc_robot* r, const double* q, double finger_open, tf* out) {
    for (int i = 0; i < r->L; ++i) {
        tf O;
        memcpy(O.m, r->joint_origin + 12 * i, sizeof(double) * 12);
        double val = 0.0;
        int qi = r->joint_qidx[i];
        if (qi >= 0) val = q[qi];
        else if (qi == -2) val = finger_open;
        tf M = TF_ID;
        const double* a = r->joint_axis + 3 * i;
        if (r->joint_type[i] == 1) { /* revolute: Rodrigues about the (unit) axis */
            double n = sqrt(a[0] * a[0] + a[1] * a[1] + a[2] * a[2]);
            double x = a[0] / n, y = a[1] / n, z = a[2] / n;
            double c = cos(val), s = sin(val), t = 1.0 - c;
            M.m[0] = t * x * x + c;     M.m[1] = t * x * y - s * z; M.m[2] = t * x * z + s * y;
            M.m[4] = t * x * y + s * z; M.m[5] = t * y * y + c;     M.m[6] = t * y * z - s * x;
            M.m[8] = t * x * z - s * y; M.m[9] = t * y * z + s * x; M.m[10] = t * z * z + c;
        } else if (r->joint_type[i] == 2) { /* prismatic */
            M.m[3] = a[0] * val; M.m[7] = a[1] * val; M.m[11] = a[2] * val;
        }
        tf J = tf_mul(&O, &M);
        int p = r->joint_parent[i];
        out[i] = (p >= 0) ? tf_mul(&out[p], &J) : J;
    }
}

int orc_fk(const orc_robot* r, const double* q, int64_t n, int dof, double finger_open, double* out) {
    if (dof != r->dof) return -1;
    for (int64_t i = 0; i < n; ++i) fk_links(r, q + i * dof, finger_open, (tf*)(out + i * r->L * 12));
    return 0;
}

/* ------------------------------------------------------------------------------------------- GJK */
static inline v3 support_local(const geom* g, v3 d, int* idx) {
    v3 dl = g->T ? tf_rot_t(g->T, d) : d;
    int best = 0;
    double bv = -1e300;
    for (int i = 0; i < g->n; ++i) {
        double s = g->v[3 * i] * dl.x + g->v[3 * i + 1] * dl.y + g->v[3 * i + 2] * dl.z;
        if (s > bv) { bv = s; best = i; }
    }
    if (idx) *idx = best;
    v3 p = V(g->v[3 * best], g->v[3 * best + 1], g->v[3 * best + 2]);
    return g->T ? tf_apply(g->T, p) : p;
}

/* support point of the Minkowski difference A - B in direction d */
static inline v3 support_md(const geom* A, const geom* B, v3 d, int* ia, int* ib) {
    v3 a = support_local(A, d, ia);
    v3 b = support_local(B, neg(d), ib);
    return sub(a, b);
}

typedef struct {
    v3 p[4];
    int ia[4], ib[4];
    int n;
} simplex;

/* closest point of segment ab to the origin; keeps the supporting sub-simplex */
static v3 closest_segment(simplex* s) {
    v3 a = s->p[0], b = s->p[1];
    v3 ab = sub(b, a);
    double t = -dot(a, ab);
    double den = norm2(ab);
    if (t <= 0.0 || den <= 0.0) { s->n = 1; return a; }
    if (t >= den) { s->p[0] = b; s->ia[0] = s->ia[1]; s->ib[0] = s->ib[1]; s->n = 1; return b; }
    return add(a, mul(ab, t / den));
}

/* closest point of triangle (a,b,c) to the origin (Voronoi regions); mask = vertices kept */
static v3 closest_triangle_pts(v3 a, v3 b, v3 c, int* mask) {
    v3 ab = sub(b, a), ac = sub(c, a);
    double d1 = -dot(ab, a), d2 = -dot(ac, a);
    if (d1 <= 0.0 && d2 <= 0.0) { *mask = 1; return a; }
    double d3 = -dot(ab, b), d4 = -dot(ac, b);
    if (d3 >= 0.0 && d4 <= d3) { *mask = 2; return b; }
    double vc = d1 * d4 - d3 * d2;
    if (vc <= 0.0 && d1 >= 0.0 && d3 <= 0.0) { *mask = 3; return add(a, mul(ab, d1 / (d1 - d3))); }
    double d5 = -dot(ab, c), d6 = -dot(ac, c);
    if (d6 >= 0.0 && d5 <= d6) { *mask = 4; return c; }
    double vb = d5 * d2 - d1 * d6;
    if (vb <= 0.0 && d2 >= 0.0 && d6 <= 0.0) { *mask = 5; return add(a, mul(ac, d2 / (d2 - d6))); }
    double va = d3 * d6 - d5 * d4;
    if (va <= 0.0 && (d4 - d3) >= 0.0 && (d5 - d6) >= 0.0) {
        *mask = 6;
        return add(b, mul(sub(c, b), (d4 - d3) / ((d4 - d3) + (d5 - d6))));
    }
    *mask = 7;
    v3 n = cross(ab, ac); /* interior: project the origin on the plane (stable when close to it) */
    double nn = norm2(n);
    if (nn <= 0.0) { *mask = 3; return add(a, mul(ab, d1 / (d1 - d3))); }
    return mul(n, dot(n, a) / nn);
}

static void simplex_keep(simplex* s, const int* ids, int mask) {
    simplex t = *s;
    int k = 0;
    for (int i = 0; i < 3; ++i)
        if (mask & (1 << i)) {
            s->p[k] = t.p[ids[i]]; s->ia[k] = t.ia[ids[i]]; s->ib[k] = t.ib[ids[i]];
            ++k;
        }
    s->n = k;
}

static v3 closest_triangle(simplex* s) {
    int mask;
    v3 v = closest_triangle_pts(s->p[0], s->p[1], s->p[2], &mask);
    static const int ids[3] = {0, 1, 2};
    simplex_keep(s, ids, mask);
    return v;
}

/* tetrahedron: returns 1 if the origin is inside (simplex unchanged), else reduces to the closest face feature */
static int closest_tetra(simplex* s, v3* vout) {
    static const int F[4][4] = {{0, 1, 2, 3}, {0, 2, 3, 1}, {0, 3, 1, 2}, {1, 3, 2, 0}};
    double best = 1e300;
    int bestf = -1, bestmask = 0;
    v3 bestv = V(0, 0, 0);
    int outside_any = 0;
    for (int f = 0; f < 4; ++f) {
        v3 a = s->p[F[f][0]], b = s->p[F[f][1]], c = s->p[F[f][2]], d = s->p[F[f][3]];
        v3 n = cross(sub(b, a), sub(c, a));
        double sp = -dot(a, n);          /* origin side */
        double sd = dot(sub(d, a), n);   /* opposite vertex side */
        /* A flat tetrahedron (the new support point repeats a simplex point through another vertex pair, or lies in
         * the plane of the other three) has sd = rounding noise of either sign: the "same side" test is then
         * meaningless and four lucky signs would report an overlap that is not there.  Such a face is treated like
         * sd == 0: a candidate for the closest feature, never evidence that the origin is inside. */
        int flat = fabs(sd) <= 1e-12 * sqrt(norm2(n) * norm2(sub(d, a)));
        int outside = flat ? 1 : (sp * sd < 0.0);
        if (!outside) continue;
        outside_any = 1;
        int mask;
        v3 v = closest_triangle_pts(a, b, c, &mask);
        double d2 = norm2(v);
        if (d2 < best) { best = d2; bestf = f; bestmask = mask; bestv = v; }
    }
    if (!outside_any) return 1;
    simplex_keep(s, F[bestf], bestmask);
    *vout = bestv;
    return 0;
}

#define GJK_MAX_ITER 128
enum { GJK_SEPARATED = 0, GJK_FAR = 1, GJK_PENETRATING = 2 };

/* Distance between the convex hulls A and B.  Returns the distance (0 when overlapping); *status tells which.
 * If a separating direction proves the distance exceeds dmax the routine stops early with GJK_FAR (return > dmax). */
static double gjk(const geom* A, const geom* B, double dmax, int* status, simplex* sx, int* iters) {
    simplex s;
    s.n = 0;
    v3 ca = A->T ? V(A->T->m[3], A->T->m[7], A->T->m[11]) : V(0, 0, 0);
    v3 cb = B->T ? V(B->T->m[3], B->T->m[7], B->T->m[11]) : V(0, 0, 0);
    v3 v = sub(ca, cb);
    if (norm2(v) < 1e-24) v = V(1, 0, 0);
    double vv = 1e300;
    int it = 0;
    *status = GJK_SEPARATED;
    for (; it < GJK_MAX_ITER; ++it) {
        int ia, ib;
        v3 w = support_md(A, B, neg(v), &ia, &ib);
        double vw = dot(v, w);
        double v2 = norm2(v);
        if (vw > 0.0 && vw * vw > dmax * dmax * v2) { *status = GJK_FAR; vv = vw * vw / v2; break; }
        if (s.n > 0) {
            int dupl = 0;
            for (int i = 0; i < s.n; ++i)
                if (s.ia[i] == ia && s.ib[i] == ib) dupl = 1;
            if (dupl || vv - vw <= 1e-14 * vv) break; /* no more progress possible: converged */
        }
        s.p[s.n] = w; s.ia[s.n] = ia; s.ib[s.n] = ib; s.n++;
        v3 nv;
        if (s.n == 1) nv = w;
        else if (s.n == 2) nv = closest_segment(&s);
        else if (s.n == 3) nv = closest_triangle(&s);
        else {
            if (closest_tetra(&s, &nv)) { *status = GJK_PENETRATING; vv = 0.0; break; }
        }
        double nvv = norm2(nv);
        if (nvv <= 1e-26) { *status = GJK_PENETRATING; vv = 0.0; v = nv; break; } /* touching: treat as overlap */
        if (nvv >= vv) break; /* numerical stall */
        v = nv;
        vv = nvv;
    }
    if (sx) *sx = s;
    if (iters) *iters = it + 1;
    return sqrt(vv);
}

/* ------------------------------------------------------------------------------------------- EPA */
#define EPA_MAX_V 512
#define EPA_MAX_F 2048
typedef struct { int a, b, c; v3 n; double d; int alive; } epa_face;

static int epa_make_face(epa_face* f, const v3* P, int a, int b, int c) {
    f->a = a; f->b = b; f->c = c; f->alive = 1;
    v3 n = cross(sub(P[b], P[a]), sub(P[c], P[a]));
    double l = sqrt(norm2(n));
    if (l < 1e-300) { f->n = V(0, 0, 0); f->d = 1e300; return 0; }
    f->n = mul(n, 1.0 / l);
    f->d = dot(f->n, P[a]);
    return 1;
}

/* Build a tetrahedron that contains the origin from a (possibly lower dimensional) GJK end simplex. */
static int epa_blow_up(const geom* A, const geom* B, simplex* s) {
    static const v3 axes[6] = {{1, 0, 0}, {-1, 0, 0}, {0, 1, 0}, {0, -1, 0}, {0, 0, 1}, {0, 0, -1}};
    if (s->n == 1) {
        for (int k = 0; k < 6 && s->n == 1; ++k) {
            int ia, ib;
            v3 w = support_md(A, B, axes[k], &ia, &ib);
            if (norm2(sub(w, s->p[0])) > 1e-20) { s->p[1] = w; s->ia[1] = ia; s->ib[1] = ib; s->n = 2; }
        }
        if (s->n == 1) return 0;
    }
    if (s->n == 2) {
        v3 ab = sub(s->p[1], s->p[0]);
        /* try directions perpendicular to the segment */
        v3 ax = fabs(ab.x) <= fabs(ab.y) && fabs(ab.x) <= fabs(ab.z) ? V(1, 0, 0)
                : (fabs(ab.y) <= fabs(ab.z) ? V(0, 1, 0) : V(0, 0, 1));
        v3 d = cross(ab, ax);
        for (int k = 0; k < 6 && s->n == 2; ++k) {
            int ia, ib;
            v3 w = support_md(A, B, d, &ia, &ib);
            v3 cr = cross(ab, sub(w, s->p[0]));
            if (norm2(cr) > 1e-24) { s->p[2] = w; s->ia[2] = ia; s->ib[2] = ib; s->n = 3; break; }
            /* rotate d by 60 degrees about ab */
            v3 u = mul(ab, 1.0 / sqrt(norm2(ab)));
            v3 dr = add(mul(d, 0.5), mul(cross(u, d), 0.8660254037844386));
            d = dr;
        }
        if (s->n == 2) return 0;
    }
    if (s->n == 3) {
        v3 n = cross(sub(s->p[1], s->p[0]), sub(s->p[2], s->p[0]));
        int ia, ib;
        v3 w = support_md(A, B, n, &ia, &ib);
        if (fabs(dot(sub(w, s->p[0]), n)) <= 1e-14 * sqrt(norm2(n))) {
            w = support_md(A, B, neg(n), &ia, &ib);
            if (fabs(dot(sub(w, s->p[0]), n)) <= 1e-14 * sqrt(norm2(n))) return 0;
        }
        s->p[3] = w; s->ia[3] = ia; s->ib[3] = ib; s->n = 4;
    }
    return 1;
}

/* Penetration depth (minimum translation distance) of overlapping polytopes; -1 if it fails to start. */
static double epa(const geom* A, const geom* B, const simplex* s0, int* iters) {
    static const int F[4][3] = {{0, 1, 2}, {0, 2, 3}, {0, 3, 1}, {1, 3, 2}};
    simplex s = *s0;
    if (s.n < 4 && !epa_blow_up(A, B, &s)) { if (iters) *iters = 0; return 0.0; }
    v3* P = (v3*)malloc(sizeof(v3) * EPA_MAX_V);
    int *pa = (int*)malloc(sizeof(int) * EPA_MAX_V), *pb = (int*)malloc(sizeof(int) * EPA_MAX_V);
    epa_face* Fc = (epa_face*)malloc(sizeof(epa_face) * EPA_MAX_F);
    int (*edges)[2] = malloc(sizeof(int[2]) * EPA_MAX_F * 3);
    int nv = 4, nf = 0;
    for (int i = 0; i < 4; ++i) { P[i] = s.p[i]; pa[i] = s.ia[i]; pb[i] = s.ib[i]; }
    /* orient faces outward (away from the 4th vertex) */
    for (int f = 0; f < 4; ++f) {
        int a = F[f][0], b = F[f][1], c = F[f][2], d = 6 - a - b - c;
        v3 n = cross(sub(P[b], P[a]), sub(P[c], P[a]));
        if (dot(n, sub(P[d], P[a])) > 0.0) { int t = b; b = c; c = t; }
        epa_make_face(&Fc[nf++], P, a, b, c);
    }
    double depth = 0.0;
    int it = 0;
    for (; it < 1000; ++it) {
        int best = -1;
        double bd = 1e300;
        for (int f = 0; f < nf; ++f)
            if (Fc[f].alive && Fc[f].d < bd) { bd = Fc[f].d; best = f; }
        if (best < 0) break;
        depth = bd < 0.0 ? 0.0 : bd;
        int ia, ib;
        v3 w = support_md(A, B, Fc[best].n, &ia, &ib);
        double gain = dot(Fc[best].n, w) - Fc[best].d;
        if (gain <= 1e-12) break;
        int dupl = 0;
        for (int i = 0; i < nv; ++i)
            if (pa[i] == ia && pb[i] == ib) dupl = 1;
        if (dupl || nv >= EPA_MAX_V) break;
        P[nv] = w; pa[nv] = ia; pb[nv] = ib;
        int ne = 0;
        for (int f = 0; f < nf; ++f) {
            if (!Fc[f].alive) continue;
            if (Fc[f].d < 1e299 && dot(Fc[f].n, w) - Fc[f].d <= 1e-14) continue; /* not visible (degenerate faces are dropped) */
            Fc[f].alive = 0;
            int e[3][2] = {{Fc[f].a, Fc[f].b}, {Fc[f].b, Fc[f].c}, {Fc[f].c, Fc[f].a}};
            for (int k = 0; k < 3; ++k) {
                int found = -1;
                for (int j = 0; j < ne; ++j)
                    if (edges[j][0] == e[k][1] && edges[j][1] == e[k][0]) { found = j; break; }
                if (found >= 0) { edges[found][0] = edges[ne - 1][0]; edges[found][1] = edges[ne - 1][1]; --ne; }
                else { edges[ne][0] = e[k][0]; edges[ne][1] = e[k][1]; ++ne; }
            }
        }
        if (nf + ne > EPA_MAX_F) break;
        for (int j = 0; j < ne; ++j) epa_make_face(&Fc[nf++], P, edges[j][0], edges[j][1], nv);
        ++nv;
    }
    if (iters) *iters = it + 1;
    free(P); free(pa); free(pb); free(Fc); free(edges);
    return depth;
}

double orc_gjk_distance(const double* va, int na, const double* Ta, const double* vb, int nb, const double* Tb,
                        double dmax, int* iterations) {
    geom A = {va, na, (const tf*)Ta}, B = {vb, nb, (const tf*)Tb};
    int st;
    double d = gjk(&A, &B, dmax, &st, NULL, iterations);
    return st == GJK_PENETRATING ? 0.0 : d;
}

double orc_epa_depth(const double* va, int na, const double* Ta, const double* vb, int nb, const double* Tb,
                     int* iterations) {
    geom A = {va, na, (const tf*)Ta}, B = {vb, nb, (const tf*)Tb};
    int st;
    simplex s;
    gjk(&A, &B, 1e300, &st, &s, NULL);
    if (st != GJK_PENETRATING) return -1.0;
    return epa(&A, &B, &s, iterations);
}

/* ----------------------------------------------------------------------------- contact distances */
/* contactDistance of a robot shape vs a convex piece (Appendix C.2/C.3); +inf if beyond the query distance */
static double shape_vs_hull(const orc_robot* r, int s, const tf* Ts, const double* pv, int pn, const tf* Tp,
                            double pmargin, const double* p_epa_v, int p_epa_n, double p_epa_margin,
                            double qd, uint32_t flags) {
    geom A = {r->core + 3 * r->shape_off[s], r->shape_off[s + 1] - r->shape_off[s], Ts};
    geom B = {pv, pn, Tp};
    double msum = r->shape_margin[s] + pmargin;
    int st;
    simplex sx;
    double d = gjk(&A, &B, qd + msum, &st, &sx, NULL);
    if (st == GJK_FAR) return INFINITY;
    if (st == GJK_SEPARATED) {
        double c = d - msum;
        return c <= qd ? c : INFINITY;
    }
    /* cores overlap: EPA on the margin-inflated shapes (box: exact box, hull: hull + sphere) */
    if (flags & ORC_NO_EPA) return -msum;
    geom A2 = {r->plane_v + 3 * r->plane_off[s], r->plane_off[s + 1] - r->plane_off[s], Ts};
    geom B2 = {p_epa_v, p_epa_n, Tp};
    double em = r->plane_margin[s] + p_epa_margin;
    if (r->plane_margin[s] != r->shape_margin[s] || p_epa_margin != pmargin) { /* box: EPA geometry differs from the core, restart GJK on it */
        gjk(&A2, &B2, 1e300, &st, &sx, NULL);
        if (st != GJK_PENETRATING) return -msum;
    }
    double depth = epa(&A2, &B2, &sx, NULL);
    return -(depth + em);
}

static inline void shape_world_aabb(const orc_robot* r, int s, const tf* T, double* lo, double* hi) {
    const double* c = r->aabb_c + 3 * s;
    const double* h = r->aabb_h + 3 * s;
    for (int k = 0; k < 3; ++k) {
        double cc = T->m[4 * k] * c[0] + T->m[4 * k + 1] * c[1] + T->m[4 * k + 2] * c[2] + T->m[4 * k + 3];
        double hh = fabs(T->m[4 * k]) * h[0] + fabs(T->m[4 * k + 1]) * h[1] + fabs(T->m[4 * k + 2]) * h[2];
        lo[k] = cc - hh; hi[k] = cc + hh;
    }
}

static double shape_vs_plane(const orc_robot* r, const orc_scene* sc, int s, const tf* T, double qd) {
    double best = 1e300;
    v3 n = V(sc->plane[0], sc->plane[1], sc->plane[2]);
    for (int i = r->plane_off[s]; i < r->plane_off[s + 1]; ++i) {
        v3 p = tf_apply(T, V(r->plane_v[3 * i], r->plane_v[3 * i + 1], r->plane_v[3 * i + 2]));
        double d = dot(n, p) + sc->plane[3];
        if (d < best) best = d;
    }
    best -= r->plane_margin[s];
    return best <= qd ? best : INFINITY;
}

/* min contact distance of the robot shapes placed at Ts[s] (s in list) against the scene */
static double scene_min_distance(const orc_robot* r, const orc_scene* sc, const tf* shapeT, const int* shapes,
                                 int ns, double qd, uint32_t flags, double* per_pair /* [S,P+1] or NULL */) {
    double dmin = INFINITY;
    for (int k = 0; k < ns; ++k) {
        int s = shapes ? shapes[k] : k;
        const tf* T = &shapeT[s];
        double lo[3], hi[3];
        shape_world_aabb(r, s, T, lo, hi);
        if (flags & ORC_SCENE) {
            for (int p = 0; p < sc->P; ++p) {
                double infl = qd + r->shape_margin[s] + sc->margin[p];
                int sep = 0;
                for (int a = 0; a < 3; ++a)
                    if (lo[a] > sc->aabb_hi[3 * p + a] + infl || hi[a] < sc->aabb_lo[3 * p + a] - infl) sep = 1;
                double d = INFINITY;
                if (!sep) {
                    const double* pv = sc->v + 3 * sc->off[p];
                    int pn = sc->off[p + 1] - sc->off[p];
                    d = shape_vs_hull(r, s, T, pv, pn, NULL, sc->margin[p], pv, pn, sc->margin[p], qd, flags);
                }
                if (per_pair) per_pair[s * (sc->P + 1) + p] = d;
                if (d < dmin) dmin = d;
            }
        }
        if ((flags & ORC_PLANE) && sc->has_plane) {
            double d = shape_vs_plane(r, sc, s, T, qd);
            if (per_pair) per_pair[s * (sc->P + 1) + sc->P] = d;
            if (d < dmin) dmin = d;
        }
    }
    return dmin;
}

static void shape_transforms(const orc_robot* r, const tf* links, tf* shapeT) {
    for (int s = 0; s < r->S; ++s) shapeT[s] = r->shape_link[s] >= 0 ? links[r->shape_link[s]] : TF_ID;
}

static double self_min_distance(const orc_robot* r, const tf* links, double qd, uint32_t flags, double* per_pair) {
    double dmin = INFINITY;
    for (int k = 0; k < r->K; ++k) {
        int la = r->self_pairs[2 * k], lb = r->self_pairs[2 * k + 1];
        double d = INFINITY;
        /* getClosestPoints(body, body, 0.01, linkA, linkB) reports points of EVERY collision shape of both links */
        for (int sa = 0; sa < r->S; ++sa) {
            if (r->shape_link[sa] != la) continue;
            for (int sb = 0; sb < r->S; ++sb) {
                if (r->shape_link[sb] != lb) continue;
                const tf* Ta = la >= 0 ? &links[la] : &TF_ID;
                const tf* Tb = lb >= 0 ? &links[lb] : &TF_ID;
                double dd = shape_vs_hull(r, sa, Ta, r->core + 3 * r->shape_off[sb], r->shape_off[sb + 1] - r->shape_off[sb], Tb,
                                          r->shape_margin[sb], r->plane_v + 3 * r->plane_off[sb],
                                          r->plane_off[sb + 1] - r->plane_off[sb], r->plane_margin[sb], qd, flags);
                if (dd < d) d = dd;
            }
        }
        if (per_pair) per_pair[k] = d;
        if (d < dmin) dmin = d;
    }
    return dmin;
}

#define MAX_LINKS 64

static void check_one(const orc_robot* r, const orc_scene* sc, const double* q, uint32_t flags, double finger_open,
                      double threshold, double qd, uint8_t* bit, double* dist, uint8_t* reason) {
    tf links[MAX_LINKS], shapeT[MAX_LINKS];
    fk_links(r, q, finger_open, links);
    shape_transforms(r, links, shapeT);
    int lim = 0;
    for (int j = 0; j < r->dof; ++j)
        if (!isfinite(q[j])) {   /* a NaN / infinite joint value has no pose: rejected whatever the flags say */
            if (bit) *bit = 1;
            if (dist) *dist = qd;
            if (reason) *reason = 1;
            return;
        }
    if (flags & ORC_LIMITS)
        for (int j = 0; j < r->dof; ++j)
            if (!(r->q_limits[2 * j] <= q[j] && q[j] <= r->q_limits[2 * j + 1])) lim = 1;
    double dself = (flags & ORC_SELF) ? self_min_distance(r, links, qd, flags, NULL) : INFINITY;
    double dscene = (flags & (ORC_SCENE | ORC_PLANE)) && sc ? scene_min_distance(r, sc, shapeT, NULL, r->S, qd, flags, NULL)
                                                           : INFINITY;
    double d = dself < dscene ? dself : dscene;
    int cself = dself < threshold, cscene = dscene < threshold;
    if (bit) *bit = (uint8_t)(lim || cself || cscene);
    if (dist) *dist = d < qd ? d : qd;
    if (reason) *reason = lim ? 1 : (cself ? 2 : (cscene ? 3 : 0));
}

typedef struct {
    const orc_robot* r; const orc_scene* sc; const double *q, *qb, *poses; int dof, substeps; uint32_t flags;
    double finger_open, threshold, qd; uint8_t* bits; double* dist; uint8_t* reason;
    const tf* rel; const int* shapes; int ns;
} job_ctx;

static void check_range(int64_t b, int64_t e, void* p) {
    job_ctx* c = (job_ctx*)p;
    for (int64_t i = b; i < e; ++i)
        check_one(c->r, c->sc, c->q + i * c->dof, c->flags, c->finger_open, c->threshold, c->qd,
                  c->bits ? c->bits + i : NULL, c->dist ? c->dist + i : NULL, c->reason ? c->reason + i : NULL);
}

int orc_check(const orc_robot* r, const orc_scene* sc, const double* q, int64_t n, int dof, uint32_t flags,
              double finger_open, double threshold, double qd, uint8_t* bits, double* dist, uint8_t* reason,
              int n_threads) {
    if (dof != r->dof || r->L > MAX_LINKS || r->S > MAX_LINKS) return -1;
    job_ctx c = {r, sc, q, NULL, NULL, dof, 0, flags, finger_open, threshold, qd, bits, dist, reason, NULL, NULL, 0};
    parallel_for(n, 64, n_threads, check_range, &c);
    return 0;
}

int orc_pair_distances(const orc_robot* r, const orc_scene* sc, const double* q, int dof, double finger_open,
                       double qd, uint32_t flags, double* out) {
    if (dof != r->dof || r->L > MAX_LINKS || r->S > MAX_LINKS) return -1;
    tf links[MAX_LINKS], shapeT[MAX_LINKS];
    fk_links(r, q, finger_open, links);
    shape_transforms(r, links, shapeT);
    for (int i = 0; i < r->S * (sc->P + 1); ++i) out[i] = INFINITY;
    scene_min_distance(r, sc, shapeT, NULL, r->S, qd, flags, out);
    return 0;
}

int orc_self_distances(const orc_robot* r, const double* q, int dof, double finger_open, double qd, uint32_t flags,
                       double* out) {
    if (dof != r->dof || r->L > MAX_LINKS) return -1;
    tf links[MAX_LINKS];
    fk_links(r, q, finger_open, links);
    self_min_distance(r, links, qd, flags, out);
    return 0;
}

static void poses_range(int64_t b, int64_t e, void* p) {
    job_ctx* c = (job_ctx*)p;
    for (int64_t i = b; i < e; ++i) {
        tf shapeT[MAX_LINKS];
        const tf* P = (const tf*)(c->poses + 12 * i);
        int finite = 1;
        for (int k = 0; k < 12; ++k) finite &= isfinite(c->poses[12 * i + k]) ? 1 : 0;
        if (!finite) {   /* no geometry: rejected */
            if (c->bits) c->bits[i] = 1;
            if (c->dist) c->dist[i] = c->qd;
            continue;
        }
        for (int k = 0; k < c->ns; ++k) shapeT[c->shapes[k]] = tf_mul(P, &c->rel[c->shapes[k]]);
        double d = scene_min_distance(c->r, c->sc, shapeT, c->shapes, c->ns, c->qd, c->flags, NULL);
        if (c->bits) c->bits[i] = (uint8_t)(d < c->threshold);
        if (c->dist) c->dist[i] = d < c->qd ? d : c->qd;
    }
}

static void edges_range(int64_t b, int64_t e_, void* p) {
    job_ctx* c = (job_ctx*)p;
    for (int64_t e = b; e < e_; ++e) {
        double q[MAX_LINKS];
        uint8_t any = 0;
        double dmin = c->qd;
        for (int k = 0; k < c->substeps; ++k) {
            double t = c->substeps > 1 ? (double)k / (double)(c->substeps - 1) : 0.0;
            for (int j = 0; j < c->dof; ++j)
                q[j] = c->q[e * c->dof + j] + t * (c->qb[e * c->dof + j] - c->q[e * c->dof + j]);
            uint8_t bit;
            double d;
            check_one(c->r, c->sc, q, c->flags, c->finger_open, c->threshold, c->qd, &bit, &d, NULL);
            any |= bit;
            if (d < dmin) dmin = d;
        }
        if (c->bits) c->bits[e] = any;
        if (c->dist) c->dist[e] = dmin;
    }
}

int orc_check_poses(const orc_robot* r, const orc_scene* sc, const double* poses, int64_t n, int ee_link,
                    const int* use_links, int n_use, double finger_open, double threshold, double qd, uint32_t flags,
                    uint8_t* bits, double* dist, int n_threads) {
    if (r->L > MAX_LINKS || r->S > MAX_LINKS) return -1;
    /* rigid offsets of the gripper shapes relative to the end-effector link (independent of the arm joints) */
    double q0[MAX_LINKS] = {0};
    tf links[MAX_LINKS], rel[MAX_LINKS];
    int shapes[MAX_LINKS], ns = 0;
    fk_links(r, q0, finger_open, links);
    tf inv_ee = tf_inv(&links[ee_link]);
    for (int k = 0; k < n_use; ++k)
        for (int s = 0; s < r->S; ++s)
            if (r->shape_link[s] == use_links[k]) {
                rel[s] = tf_mul(&inv_ee, &links[use_links[k]]);
                shapes[ns++] = s;
            }
    job_ctx c = {r, sc, NULL, NULL, poses, 0, 0, flags, finger_open, threshold, qd, bits, dist, NULL, rel, shapes, ns};
    parallel_for(n, 64, n_threads, poses_range, &c);
    return 0;
}

int orc_check_edges(const orc_robot* r, const orc_scene* sc, const double* qa, const double* qb, int64_t m, int dof,
                    int substeps, uint32_t flags, double finger_open, double threshold, double qd, uint8_t* bits,
                    double* dist, int n_threads) {
    if (dof != r->dof || r->L > MAX_LINKS || r->S > MAX_LINKS || substeps < 1) return -1;
    job_ctx c = {r, sc, qa, qb, NULL, dof, substeps, flags, finger_open, threshold, qd, bits, dist, NULL, NULL, NULL, 0};
    parallel_for(m, 16, n_threads, edges_range, &c);
    return 0;
}
