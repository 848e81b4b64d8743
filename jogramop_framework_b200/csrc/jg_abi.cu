// jg_abi.cu -- host side of libjogramop_b200.so: handle management, table upload, pass orchestration.
// Implements include/jogramop_b200.h.  Plain CUDA runtime, no torch types anywhere in this library.
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cfloat>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/jogramop_b200.h"
#include "jg_handles.h"
#include "jg_kernels.cuh"

using namespace jg;

// ----------------------------------------------------------------------------------------------- errors
static thread_local std::string g_err;
int jg_fail(int code, const char* fmt, ...);
static int fail(int code, const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_err = buf;
    return code;
}
#define CUDA_TRY(x)                                                                              \
    do {                                                                                         \
        cudaError_t e_ = (x);                                                                    \
        if (e_ != cudaSuccess) return fail(JG_ERR_CUDA, "%s: %s", #x, cudaGetErrorString(e_));   \
    } while (0)

int jg_fail(int code, const char* fmt, ...) {   // for the other translation units of the library
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_err = buf;
    return code;
}
extern "C" const char* jg_last_error(void) { return g_err.c_str(); }
extern "C" int jg_abi_version(void) { return JG_ABI_VERSION; }
extern "C" void jg_default_params(jg_params* p) {
    if (!p) return;
    p->threshold = -0.001f;
    p->query_distance = 0.01f;
    p->finger_open = 0.04f;
    p->chunk = 0;
}
extern "C" int jg_device_count(void) {
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) return fail(JG_ERR_NODEVICE, "cudaGetDeviceCount: %s", cudaGetErrorString(e));
    return n;
}

// ----------------------------------------------------------------------------------------------- handles
extern "C" jg_robot* jg_robot_create(int L, const int32_t* joint_type, const int32_t* joint_parent,
                                     const double* joint_origin, const double* joint_axis, const int32_t* joint_qidx,
                                     const double* q_limits, int dof, int S, const int32_t* shape_link,
                                     const double* shape_margin, const int32_t* shape_off, const double* core,
                                     const int32_t* plane_off, const double* plane_v, const double* plane_margin,
                                     int K, const int32_t* self_pairs) {
    if (L <= 0 || L > MAX_LINKS || S < 0 || S > MAX_SHAPES || dof <= 0 || dof > MAX_DOF || !joint_type ||
        !joint_parent || !joint_origin || !joint_axis || !joint_qidx || !q_limits ||
        (S && (!shape_link || !shape_off || !shape_margin || !core || !plane_off || !plane_v || !plane_margin)) ||
        K < 0 || (K && !self_pairs)) {
        fail(JG_ERR_INVALID, "jg_robot_create: bad sizes (L=%d<=%d, S=%d<=%d, dof=%d<=%d) or NULL table", L, MAX_LINKS, S,
             MAX_SHAPES, dof, MAX_DOF);
        return nullptr;
    }
    for (int i = 0; i < L; ++i) {
        if (joint_parent[i] >= i || joint_parent[i] < -1) {
            fail(JG_ERR_INVALID, "jg_robot_create: joint %d has parent %d (parents must precede children)", i, joint_parent[i]);
            return nullptr;
        }
        if (joint_qidx[i] >= dof || joint_qidx[i] < -2) {
            fail(JG_ERR_INVALID, "jg_robot_create: joint %d q index %d out of range", i, joint_qidx[i]);
            return nullptr;
        }
    }
    for (int s = 0; s < S; ++s) {
        if (shape_link[s] < -1 || shape_link[s] >= L || (s && shape_link[s] < shape_link[s - 1])) {
            fail(JG_ERR_INVALID, "jg_robot_create: shapes must be sorted by link id in [-1,%d)", L);
            return nullptr;
        }
        if (shape_off[s + 1] - shape_off[s] <= 0 || shape_off[s + 1] - shape_off[s] >= 65536) {
            fail(JG_ERR_INVALID, "jg_robot_create: shape %d has an unsupported vertex count", s);
            return nullptr;
        }
    }
    jg_robot* r = new jg_robot;
    r->L = L; r->S = S; r->dof = dof; r->K = K;
    r->joint_type.assign(joint_type, joint_type + L);
    r->joint_parent.assign(joint_parent, joint_parent + L);
    r->joint_qidx.assign(joint_qidx, joint_qidx + L);
    r->joint_origin.assign(joint_origin, joint_origin + 12 * L);
    r->joint_axis.assign(joint_axis, joint_axis + 3 * L);
    r->q_limits.assign(q_limits, q_limits + 2 * dof);
    r->shape_link.assign(shape_link, shape_link + S);
    r->shape_margin.assign(shape_margin, shape_margin + S);
    r->shape_off.assign(shape_off, shape_off + S + 1);
    r->core.assign(core, core + 3 * (S ? shape_off[S] : 0));
    r->plane_off.assign(plane_off, plane_off + S + 1);
    r->plane_v.assign(plane_v, plane_v + 3 * (S ? plane_off[S] : 0));
    r->plane_margin.assign(plane_margin, plane_margin + S);
    r->self_pairs.assign(self_pairs, self_pairs + 2 * K);
    return r;
}
extern "C" void jg_robot_free(jg_robot* r) { delete r; }

extern "C" jg_scene* jg_scene_create(int P, const int32_t* off, const double* v, const double* margin,
                                     const double* plane) {
    if (P < 0 || (P && (!off || !v || !margin))) {
        fail(JG_ERR_INVALID, "jg_scene_create: bad arguments");
        return nullptr;
    }
    for (int p = 0; p < P; ++p)
        if (off[p + 1] - off[p] <= 0 || off[p + 1] - off[p] >= 65536) {
            fail(JG_ERR_INVALID, "jg_scene_create: piece %d has an unsupported vertex count", p);
            return nullptr;
        }
    jg_scene* s = new jg_scene;
    s->P = P;
    if (P) {
        s->off.assign(off, off + P + 1);
        s->v.assign(v, v + 3 * off[P]);
        s->margin.assign(margin, margin + P);
    } else {
        s->off.assign(1, 0);
    }
    s->has_plane = plane != nullptr;
    if (plane) memcpy(s->plane, plane, sizeof(double) * 4);
    return s;
}
extern "C" jg_scene* jg_scene_create_mesh(int n_verts, const double* verts, int n_tris, const int32_t* tris, double margin,
                                          const double* plane) {
    if (n_verts <= 0 || n_tris <= 0 || !verts || !tris || !(margin >= 0)) {
        fail(JG_ERR_INVALID, "jg_scene_create_mesh: bad arguments");
        return nullptr;
    }
    if (n_tris >= 65536) {
        fail(JG_ERR_INVALID, "jg_scene_create_mesh: at most 65535 triangles are supported (got %d)", n_tris);
        return nullptr;
    }
    for (int i = 0; i < 3 * n_tris; ++i)
        if (tris[i] < 0 || tris[i] >= n_verts) {
            fail(JG_ERR_INVALID, "jg_scene_create_mesh: triangle %d references vertex %d of %d", i / 3, tris[i], n_verts);
            return nullptr;
        }
    jg_scene* s = new jg_scene;
    s->is_mesh = true;
    s->off.assign(1, 0);
    s->mesh_v.resize(3 * (size_t)n_verts);
    for (size_t i = 0; i < s->mesh_v.size(); ++i) s->mesh_v[i] = (float)verts[i];
    s->mesh_t.assign(tris, tris + 3 * (size_t)n_tris);
    s->mesh_margin = margin;
    s->has_plane = plane != nullptr;
    if (plane) memcpy(s->plane, plane, sizeof(double) * 4);
    return s;
}
extern "C" void jg_scene_free(jg_scene* s) { delete s; }

// ----------------------------------------------------------------------------------------------- engine
struct jg_engine {
    int device = 0;
    jg_params prm;
    int L = 0, S = 0, dof = 0, P = 0, K = 0;
    int n_keys = 0, n_scene_keys = 0, n_verts = 0;
    bool has_plane = false;
    float plane[4] = {0, 0, 1, 0};
    std::vector<int> shape_link;          // host copy, for pose mode
    std::vector<int> link_first_shape;    // [L+1] first shape of link (index link+1) or -1
    RobotTab h_robot;
    // device tables
    RobotTab* d_robot = nullptr;
    PieceRec* d_pieces = nullptr;
    SelfRec* d_selfs = nullptr;
    KeyRec* d_keys = nullptr;
    float4* d_verts = nullptr;
    // scratch arena
    int cap = 0;                          // configurations per pass the arena currently holds (grows on demand up to cap_max)
    int cap_max = 0;                      // largest pass: jg_params.chunk, or what fits the memory budget (<= 2^20)
    float4 *q_r0 = nullptr, *q_r1 = nullptr, *q_r2 = nullptr;
    int* q_cfg = nullptr;
    int* epa_slot = nullptr;              // [n_keys*cap]
    uint4* epa_ids = nullptr;             // [n_keys*cap]
    int2* ovf_list = nullptr;             // [cap]
    uint32_t* ovf_poly = nullptr;         // [ovf_poly_cap][JG_OVF_WORDS]
    int ovf_poly_cap = 0;
    float* epa_prio = nullptr;            // [n_keys*cap]
    float* q_prio = nullptr;              // [n_keys*cap]
    float max_msum = 0.f, min_msum = 0.f;
    cudaEvent_t ev_order = nullptr;      // orders a call on a new stream after the previous call's work (shared scratch arena)
    bool mesh = false;
    float mesh_margin = 0.f;
    Lbvh bvh;
    int* d_err = nullptr;
    int mesh_pass = 0;
    float4* mesh_T = nullptr;            // [mesh_cfg_cap * S * 3] shape transforms of a mesh-mode pass (queries of k_mesh_traverse)
    int mesh_cfg_cap = 0;                // configurations (substep-expanded) a mesh-mode pass may hold
    int n_stash = 0;                     // link frames the broad phase stashes for the self pairs                   // configurations per pass in mesh mode (adaptive)
    float* d_qrest = nullptr;
    int* act_list = nullptr;              // [n_keys*cap]
    int2* act_meta = nullptr;             // [n_keys*cap]
    struct Pending { bool busy = false; int64_t n = 0; uint32_t* bits = nullptr; float* dist = nullptr; uint8_t* reason = nullptr;
                     bool pb = false, pd = false, pr = false; } pending[2];   // jg_check_host_submit tickets (ticket & 1)
    int64_t next_ticket = 0;
    unsigned long long *champ_scene = nullptr, *champ_self = nullptr;   // [cap]
    int* counts = nullptr;                // [3*n_keys + 16], layout in jg_kernels.cuh PassArgs
    int *enc_scene = nullptr, *enc_self = nullptr;
    uint32_t* viol = nullptr;
    int* d_link_slot = nullptr;           // FK: [L]
    Tf* d_rel = nullptr;                  // poses: [MAX_SHAPES]
    int* d_use_shapes = nullptr;
    // host staging for jg_check_host
    cudaStream_t s_in = nullptr, s_run = nullptr, s_out = nullptr;
    float* stage_q[2] = {nullptr, nullptr};      // device
    uint32_t* stage_bits[2] = {nullptr, nullptr};
    float* stage_dist[2] = {nullptr, nullptr};
    uint8_t* stage_reason[2] = {nullptr, nullptr};
    float* pin_q[2] = {nullptr, nullptr};        // pinned host
    uint32_t* pin_bits[2] = {nullptr, nullptr};
    float* pin_dist[2] = {nullptr, nullptr};
    uint8_t* pin_reason[2] = {nullptr, nullptr};
    cudaEvent_t ev_in[2] = {nullptr, nullptr}, ev_run[2] = {nullptr, nullptr}, ev_out[2] = {nullptr, nullptr};
    bool staging_ready = false;          // streams + events exist
    int staging_cap = 0;                 // pass size the staging buffers were allocated for
    // stats
    jg_stats stats{};
    unsigned long long* d_stats = nullptr;       // [8]
    cudaStream_t last_stream = nullptr;
    bool have_last_stream = false;
    // lanes (jg_engine_set_lanes): a call of more than one pass hands its passes round-robin to this engine and to
    // lanes - 1 sibling engines (own tables, own arena, created on first use), each on its own stream, so that one pass's
    // kernels run under the tails of the others
    int lanes = 1;
    std::vector<jg_engine*> sib;
    std::vector<cudaStream_t> lane_stream;
    std::vector<cudaEvent_t> lane_ev;
    cudaEvent_t ev_fork = nullptr;
    jg_robot robot_copy;
    jg_scene scene_copy;
    bool has_scene = false;
    bool accumulate_stats = false;       // a lane's second, third ... piece of one call: keep counting
    int last_call_lanes = 0;             // lanes the last call used (jg_get_stats adds the siblings' counters)
    int sm_count = 148;
    int np_blocks = 0, epa_blocks = 0, epa_threads = 0;
    size_t smem_narrow = 0, smem_epa = 0, smem_epa_big = 0;
    // optional per-kernel timing (jg_profile_enable)
    bool profiling = false;
    std::vector<cudaEvent_t> prof_ev;    // event pool
    struct ProfSection { int cat; cudaEvent_t a, b; };
    std::vector<ProfSection> prof_sections;
    size_t prof_used = 0;
    double prof_ms[4] = {0, 0, 0, 0};
    int64_t prof_n[4] = {0, 0, 0, 0};
};

static cudaEvent_t prof_event(jg_engine* e) {
    if (e->prof_used == e->prof_ev.size()) {
        cudaEvent_t ev;
        cudaEventCreate(&ev);
        e->prof_ev.push_back(ev);
    }
    return e->prof_ev[e->prof_used++];
}
// timed section of category `cat` (0 broad phase, 1 GJK stages, 2 penetration stages, 3 finalize): events on the
// launching stream; sections are (cat, start, end) triples resolved in prof_collect
static void prof_begin(jg_engine* e, int cat, cudaStream_t st) {
    if (!e->profiling) return;
    cudaEvent_t ev = prof_event(e);
    cudaEventRecord(ev, st);
    e->prof_sections.push_back({cat, ev, nullptr});
}
static void prof_end(jg_engine* e, cudaStream_t st) {
    if (!e->profiling || e->prof_sections.empty()) return;
    cudaEvent_t ev = prof_event(e);
    cudaEventRecord(ev, st);
    e->prof_sections.back().b = ev;
}

static void prof_collect(jg_engine* e) {
    bool counted[4] = {false, false, false, false};
    int last_pass_cat = -1;
    for (const auto& sec : e->prof_sections) {
        if (!sec.b) continue;
        cudaEventSynchronize(sec.b);
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, sec.a, sec.b) == cudaSuccess) e->prof_ms[sec.cat] += ms;
        // launches = number of passes in which the category ran (a new pass starts with category 0)
        if (sec.cat == 0 || last_pass_cat < 0) { for (bool& c : counted) c = false; }
        if (!counted[sec.cat]) { e->prof_n[sec.cat] += 1; counted[sec.cat] = true; }
        last_pass_cat = sec.cat;
    }
    e->prof_sections.clear();
    e->prof_used = 0;
}

template <class T>
static void free_dev(T*& p) { cudaFree(p); p = nullptr; }
template <class T>
static void free_pin(T*& p) { cudaFreeHost(p); p = nullptr; }

// the per-pass scratch arena (everything sized by e->cap)
static void arena_free(jg_engine* e) {
    free_dev(e->q_r0); free_dev(e->q_r1); free_dev(e->q_r2); free_dev(e->q_cfg); free_dev(e->epa_slot); free_dev(e->epa_ids);
    free_dev(e->ovf_list); free_dev(e->ovf_poly); free_dev(e->epa_prio); free_dev(e->q_prio); free_dev(e->act_list);
    free_dev(e->act_meta); free_dev(e->champ_scene); free_dev(e->champ_self); free_dev(e->enc_scene); free_dev(e->enc_self);
    free_dev(e->viol); free_dev(e->mesh_T);
    e->cap = 0;
}
// staging buffers of the host path (sized by e->cap as well)
static void staging_free(jg_engine* e) {
    for (int i = 0; i < 2; ++i) {
        free_dev(e->stage_q[i]); free_dev(e->stage_bits[i]); free_dev(e->stage_dist[i]); free_dev(e->stage_reason[i]);
        free_pin(e->pin_q[i]); free_pin(e->pin_bits[i]); free_pin(e->pin_dist[i]); free_pin(e->pin_reason[i]);
    }
    e->staging_cap = 0;
}

static void engine_release(jg_engine* e) {
    if (!e) return;
    for (jg_engine* sb : e->sib) engine_release(sb);
    e->sib.clear();
    cudaSetDevice(e->device);
    for (cudaStream_t ls : e->lane_stream) cudaStreamDestroy(ls);
    for (cudaEvent_t ev : e->lane_ev) cudaEventDestroy(ev);
    if (e->ev_fork) cudaEventDestroy(e->ev_fork);
    cudaFree(e->d_robot); cudaFree(e->d_pieces); cudaFree(e->d_selfs); cudaFree(e->d_keys); cudaFree(e->d_verts);
    arena_free(e);
    cudaFree(e->bvh.nodes); cudaFree(e->bvh.tri_v); cudaFree(e->bvh.tri_id); cudaFree(e->d_err); cudaFree(e->d_qrest); cudaFree(e->counts);
    cudaFree(e->d_link_slot); cudaFree(e->d_rel);
    cudaFree(e->d_use_shapes); cudaFree(e->d_stats);
    staging_free(e);
    for (int i = 0; i < 2; ++i) {
        if (e->ev_in[i]) cudaEventDestroy(e->ev_in[i]);
        if (e->ev_run[i]) cudaEventDestroy(e->ev_run[i]);
        if (e->ev_out[i]) cudaEventDestroy(e->ev_out[i]);
    }
    if (e->s_in) cudaStreamDestroy(e->s_in);
    if (e->s_run) cudaStreamDestroy(e->s_run);
    if (e->s_out) cudaStreamDestroy(e->s_out);
    for (cudaEvent_t ev : e->prof_ev) cudaEventDestroy(ev);
    if (e->ev_order) cudaEventDestroy(e->ev_order);
    delete e;
}

static int build_tables(jg_engine* e, const jg_robot* r, const jg_scene* sc) {
    RobotTab& R = e->h_robot;
    memset(&R, 0, sizeof R);
    R.L = r->L; R.S = r->S; R.dof = r->dof; R.n_self = r->K;
    for (int j = 0; j < r->dof; ++j) {
        R.qlo[j] = (float)r->q_limits[2 * j];
        R.qhi[j] = (float)r->q_limits[2 * j + 1];
        // float32 rounding must not reject the float32 image of a limit itself
        if ((double)R.qlo[j] > r->q_limits[2 * j]) R.qlo[j] = nextafterf(R.qlo[j], -INFINITY);
        if ((double)R.qhi[j] < r->q_limits[2 * j + 1]) R.qhi[j] = nextafterf(R.qhi[j], INFINITY);
    }
    // which links must be kept for later children (one register slot: enough for chains with one branch point)
    int saved = -2;
    std::vector<int> need_save(r->L, 0);
    for (int i = 0; i < r->L; ++i) {
        const int p = r->joint_parent[i];
        if (p >= 0 && p != i - 1) need_save[p] = 1;
    }
    for (int i = 0; i < r->L; ++i) {
        JointRec& J = R.joints[i];
        for (int k = 0; k < 12; ++k) J.o[k] = (float)r->joint_origin[12 * i + k];
        double ax = r->joint_axis[3 * i], ay = r->joint_axis[3 * i + 1], az = r->joint_axis[3 * i + 2];
        double nn = std::sqrt(ax * ax + ay * ay + az * az);
        if (r->joint_type[i] != JG_JOINT_FIXED && !(nn > 0))
            return fail(JG_ERR_INVALID, "joint %d: zero axis on a movable joint", i);
        if (nn > 0) { ax /= nn; ay /= nn; az /= nn; }
        J.axis[0] = (float)ax; J.axis[1] = (float)ay; J.axis[2] = (float)az;
        J.type = r->joint_type[i];
        J.qidx = r->joint_qidx[i];
        {
            const float I12[12] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0};
            bool ident = true;
            for (int k = 0; k < 12; ++k) ident = ident && J.o[k] == I12[k];
            J.fast = (ident ? 1 : 0) | ((J.type == JG_JOINT_REVOLUTE && J.axis[0] == 0.f && J.axis[1] == 0.f && J.axis[2] == 1.f) ? 2 : 0);
        }
        const int p = r->joint_parent[i];
        if (p == -1) J.parent_sel = 2;
        else if (p == i - 1) J.parent_sel = 0;
        else if (p == saved) J.parent_sel = 1;
        else return fail(JG_ERR_INVALID, "joint %d: unsupported tree topology (more than one live branch point)", i);
        J.save = need_save[i];
        if (J.save) saved = i;
        J.shape_first = 0; J.shape_count = 0;
    }
    e->link_first_shape.assign(r->L + 1, -1);
    e->shape_link = r->shape_link;
    // every vertex set is padded to a multiple of 4 with copies of its last vertex (support_scan works on groups of 4)
    std::vector<float4> verts;
    verts.reserve(4096);
    auto pad4 = [&verts]() { while (verts.size() % 4) verts.push_back(verts.back()); };
    float epa_box_c[MAX_SHAPES][3], epa_box_h[MAX_SHAPES][3];
    for (int s = 0; s < r->S; ++s) {
        ShapeRec& Sh = R.shapes[s];
        double lo[3] = {1e300, 1e300, 1e300}, hi[3] = {-1e300, -1e300, -1e300};
        for (int i = r->shape_off[s]; i < r->shape_off[s + 1]; ++i)
            for (int k = 0; k < 3; ++k) {
                lo[k] = std::min(lo[k], r->core[3 * i + k]);
                hi[k] = std::max(hi[k], r->core[3 * i + k]);
            }
        for (int k = 0; k < 3; ++k) {
            Sh.c[k] = (float)(0.5 * (lo[k] + hi[k]));
            Sh.h[k] = (float)(0.5 * (hi[k] - lo[k]) + 1e-6);
        }
        Sh.margin = (float)r->shape_margin[s];
        Sh.plane_margin = (float)r->plane_margin[s];
        Sh.voff = (int)verts.size();
        Sh.vn = r->shape_off[s + 1] - r->shape_off[s];
        for (int i = r->shape_off[s]; i < r->shape_off[s + 1]; ++i)
            verts.push_back(make_float4((float)r->core[3 * i], (float)r->core[3 * i + 1], (float)r->core[3 * i + 2], 0.f));
        pad4();
        Sh.link = r->shape_link[s];
        {   // exactly the 8 corners of its local box?  (exact box-vs-triangle tests in the boolean mesh mode)
            int seen = 0;
            bool box = Sh.vn == 8;
            for (int i = r->shape_off[s]; box && i < r->shape_off[s + 1]; ++i) {
                int code = 0;
                for (int k = 0; k < 3; ++k) {
                    const double v = r->core[3 * i + k];
                    if (std::fabs(v - lo[k]) <= 1e-12) code |= 0;
                    else if (std::fabs(v - hi[k]) <= 1e-12) code |= 1 << k;
                    else box = false;
                }
                seen |= 1 << code;
            }
            Sh.is_box = (box && seen == 0xFF) ? 1 : 0;
        }
        // plane / EPA vertices: share the core range when identical (hulls), else append (boxes)
        const int pn = r->plane_off[s + 1] - r->plane_off[s];
        bool same = pn == Sh.vn;
        for (int i = 0; same && i < 3 * pn; ++i)
            same = r->plane_v[3 * r->plane_off[s] + i] == r->core[3 * r->shape_off[s] + i];
        if (same) { Sh.poff = Sh.voff; Sh.pn = Sh.vn; }
        else {
            Sh.poff = (int)verts.size(); Sh.pn = pn;
            for (int i = r->plane_off[s]; i < r->plane_off[s + 1]; ++i)
                verts.push_back(make_float4((float)r->plane_v[3 * i], (float)r->plane_v[3 * i + 1],
                                            (float)r->plane_v[3 * i + 2], 0.f));
            pad4();
        }
        {   // AABB of the penetration geometry (box: exact corners), used for the penetration upper bound
            double plo[3] = {1e300, 1e300, 1e300}, phi[3] = {-1e300, -1e300, -1e300};
            for (int i = r->plane_off[s]; i < r->plane_off[s + 1]; ++i)
                for (int k = 0; k < 3; ++k) {
                    plo[k] = std::min(plo[k], r->plane_v[3 * i + k]);
                    phi[k] = std::max(phi[k], r->plane_v[3 * i + k]);
                }
            for (int k = 0; k < 3; ++k) {
                epa_box_c[s][k] = (float)(0.5 * (plo[k] + phi[k]));
                epa_box_h[s][k] = (float)(0.5 * (phi[k] - plo[k]) + 1e-6);
            }
        }
        const int link = r->shape_link[s];
        if (e->link_first_shape[link + 1] < 0) e->link_first_shape[link + 1] = s;
        if (link < 0) {
            if (R.base_shape_count == 0) R.base_shape_first = s;
            R.base_shape_count++;
        } else {
            if (R.joints[link].shape_count == 0) R.joints[link].shape_first = s;
            R.joints[link].shape_count++;
        }
    }
    // pieces: centred vertices (mesh mode: the whole triangle soup is "piece 0" of the key numbering)
    e->mesh = sc && sc->is_mesh;
    e->mesh_margin = e->mesh ? (float)sc->mesh_margin : 0.f;
    const int P = e->mesh ? 0 : (sc ? sc->P : 0);
    e->P = e->mesh ? 1 : P;
    std::vector<PieceRec> pieces(P);
    for (int p = 0; p < P; ++p) {
        PieceRec& Pc = pieces[p];
        memset(&Pc, 0, sizeof Pc);
        double lo[3] = {1e300, 1e300, 1e300}, hi[3] = {-1e300, -1e300, -1e300};
        for (int i = sc->off[p]; i < sc->off[p + 1]; ++i)
            for (int k = 0; k < 3; ++k) {
                lo[k] = std::min(lo[k], sc->v[3 * i + k]);
                hi[k] = std::max(hi[k], sc->v[3 * i + k]);
            }
        for (int k = 0; k < 3; ++k) {
            Pc.c[k] = (float)(0.5 * (lo[k] + hi[k]));
            Pc.h[k] = (float)(0.5 * (hi[k] - lo[k]) + 1e-6);
        }
        Pc.margin = (float)sc->margin[p];
        Pc.voff = (int)verts.size();
        Pc.vn = sc->off[p + 1] - sc->off[p];
        for (int i = sc->off[p]; i < sc->off[p + 1]; ++i)   // subtract the (float) centre in double, then round
            verts.push_back(make_float4((float)(sc->v[3 * i] - (double)Pc.c[0]), (float)(sc->v[3 * i + 1] - (double)Pc.c[1]),
                                        (float)(sc->v[3 * i + 2] - (double)Pc.c[2]), 0.f));
        pad4();
    }
    e->has_plane = sc && sc->has_plane;
    if (e->has_plane)
        for (int k = 0; k < 4; ++k) e->plane[k] = (float)sc->plane[k];
    // keys
    std::vector<KeyRec> keys;
    for (int s = 0; s < r->S; ++s) {
        if (e->mesh)
            keys.push_back(KeyRec{R.shapes[s].voff, R.shapes[s].vn, 0, 0, R.shapes[s].margin + e->mesh_margin, KEY_MESH, 0, 0, 0, 0,
                                  0.f, 0, {0, 0, 0}, {0, 0, 0}, {0, 0, 0}, {0, 0, 0}});
        for (int p = 0; p < P; ++p)
        {
            KeyRec k{R.shapes[s].voff, R.shapes[s].vn, pieces[p].voff, pieces[p].vn,
                     R.shapes[s].margin + pieces[p].margin, KEY_HULL, R.shapes[s].poff, R.shapes[s].pn,
                     pieces[p].voff, pieces[p].vn, R.shapes[s].plane_margin + pieces[p].margin, 0, {0, 0, 0}, {0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
            for (int c = 0; c < 3; ++c) {
                k.a_c[c] = epa_box_c[s][c]; k.a_h[c] = epa_box_h[s][c];
                k.b_c[c] = 0.f; k.b_h[c] = pieces[p].h[c];     // piece vertices are centred on their AABB centre
            }
            keys.push_back(k);
        }
        keys.push_back(KeyRec{R.shapes[s].poff, R.shapes[s].pn, 0, 0, R.shapes[s].plane_margin, KEY_PLANE, 0, 0, 0, 0, 0.f, 0,
                              {0, 0, 0}, {0, 0, 0}, {0, 0, 0}, {0, 0, 0}});
    }
    e->n_scene_keys = (int)keys.size();
    // self pairs sorted by their second link: the broad phase tests (la, lb) when the chain reaches lb, with la's frame
    // taken from a shared-memory stash (only the pairs' first links are stashed)
    std::vector<SelfRec> selfs;
    std::vector<int> order(r->K);
    for (int k = 0; k < r->K; ++k) order[k] = k;
    std::stable_sort(order.begin(), order.end(), [&](int x, int y) { return r->self_pairs[2 * x + 1] < r->self_pairs[2 * y + 1]; });
    for (int i = 0; i < r->L; ++i) { R.joints[i].self_first = 0; R.joints[i].self_count = 0; R.joints[i].stash_slot = -1; }
    R.n_stash = 0;
    for (int kk = 0; kk < r->K; ++kk) {
        const int k = order[kk];
        const int la = r->self_pairs[2 * k], lb = r->self_pairs[2 * k + 1];
        if (la < 0 || lb < 0 || la >= r->L || lb >= r->L || la >= lb)
            return fail(JG_ERR_INVALID, "self pair %d: link ids (%d,%d) must satisfy 0 <= first < second < %d", k, la, lb, r->L);
        // a link without a collision shape yields no closest points in pyBullet: the pair can never collide
        if (e->link_first_shape[la + 1] < 0 || e->link_first_shape[lb + 1] < 0) continue;
        if (R.joints[la].stash_slot < 0) R.joints[la].stash_slot = R.n_stash++;
        // getClosestPoints(linkA, linkB) covers every collision shape of both links: one key per shape combination
        for (int sa = 0; sa < r->S; ++sa) {
            if (r->shape_link[sa] != la) continue;
            for (int sb = 0; sb < r->S; ++sb) {
                if (r->shape_link[sb] != lb) continue;
                if (R.joints[lb].self_count == 0) R.joints[lb].self_first = (int)selfs.size();
                R.joints[lb].self_count++;
                selfs.push_back(SelfRec{sa, sb, la, lb});
                KeyRec kr{R.shapes[sb].voff, R.shapes[sb].vn, R.shapes[sa].voff, R.shapes[sa].vn,
                          R.shapes[sa].margin + R.shapes[sb].margin, KEY_SELF, R.shapes[sb].poff, R.shapes[sb].pn,
                          R.shapes[sa].poff, R.shapes[sa].pn, R.shapes[sa].plane_margin + R.shapes[sb].plane_margin, 0,
                          {0, 0, 0}, {0, 0, 0}, {0, 0, 0}, {0, 0, 0}};
                for (int c = 0; c < 3; ++c) {
                    kr.a_c[c] = epa_box_c[sb][c]; kr.a_h[c] = epa_box_h[sb][c];
                    kr.b_c[c] = epa_box_c[sa][c]; kr.b_h[c] = epa_box_h[sa][c];
                }
                keys.push_back(kr);
            }
        }
    }
    R.n_self = (int)selfs.size();
    e->n_stash = R.n_stash;
    e->n_keys = (int)keys.size();
    e->n_verts = (int)verts.size();
    e->min_msum = FLT_MAX;
    for (const KeyRec& k : keys) {
        e->max_msum = std::max(e->max_msum, k.msum);
        if (k.kind != KEY_PLANE) e->min_msum = std::min(e->min_msum, k.msum);
    }
    // champion codes pack (key << 20 | queue slot) into 32 bits: 11 bits of key, 20 bits of slot (cap <= 2^20)
    if (e->n_keys >= 2048) return fail(JG_ERR_INVALID, "too many (shape, piece) keys (%d >= 2048)", e->n_keys);
    e->K = (int)selfs.size();

    CUDA_TRY(cudaMalloc(&e->d_robot, sizeof(RobotTab)));
    CUDA_TRY(cudaMemcpy(e->d_robot, &R, sizeof(RobotTab), cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMalloc(&e->d_pieces, sizeof(PieceRec) * std::max(P, 1)));
    if (P) CUDA_TRY(cudaMemcpy(e->d_pieces, pieces.data(), sizeof(PieceRec) * P, cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMalloc(&e->d_selfs, sizeof(SelfRec) * std::max<size_t>(selfs.size(), 1)));
    if (!selfs.empty()) CUDA_TRY(cudaMemcpy(e->d_selfs, selfs.data(), sizeof(SelfRec) * selfs.size(), cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMalloc(&e->d_keys, sizeof(KeyRec) * std::max(e->n_keys, 1)));
    if (e->n_keys) CUDA_TRY(cudaMemcpy(e->d_keys, keys.data(), sizeof(KeyRec) * e->n_keys, cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMalloc(&e->d_verts, sizeof(float4) * std::max(e->n_verts, 1)));
    if (e->n_verts) CUDA_TRY(cudaMemcpy(e->d_verts, verts.data(), sizeof(float4) * e->n_verts, cudaMemcpyHostToDevice));
    return JG_OK;
}


// GPU build of the LBVH (set-up time, once per scene)
static int build_lbvh(jg_engine* e, const jg_scene* sc) {
    const int nv = (int)sc->mesh_v.size() / 3, n = (int)sc->mesh_t.size() / 3;
    float lo[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, hi[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
    for (int i = 0; i < nv; ++i)
        for (int k = 0; k < 3; ++k) {
            lo[k] = std::min(lo[k], sc->mesh_v[3 * i + k]);
            hi[k] = std::max(hi[k], sc->mesh_v[3 * i + k]);
        }
    float *d_v = nullptr;
    int *d_t = nullptr, *d_vals = nullptr, *d_vals2 = nullptr, *d_parent = nullptr, *d_visits = nullptr;
    unsigned long long *d_keys = nullptr, *d_keys2 = nullptr;
    void* d_tmp = nullptr;
    size_t tmp_bytes = 0;
    Lbvh& B = e->bvh;
    B.n_tris = n;
    CUDA_TRY(cudaMalloc(&d_v, sizeof(float) * 3 * nv));
    CUDA_TRY(cudaMalloc(&d_t, sizeof(int) * 3 * n));
    CUDA_TRY(cudaMemcpy(d_v, sc->mesh_v.data(), sizeof(float) * 3 * nv, cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMemcpy(d_t, sc->mesh_t.data(), sizeof(int) * 3 * n, cudaMemcpyHostToDevice));
    CUDA_TRY(cudaMalloc(&d_keys, sizeof(unsigned long long) * n));
    CUDA_TRY(cudaMalloc(&d_keys2, sizeof(unsigned long long) * n));
    CUDA_TRY(cudaMalloc(&d_vals, sizeof(int) * n));
    CUDA_TRY(cudaMalloc(&d_vals2, sizeof(int) * n));
    CUDA_TRY(cudaMalloc(&d_parent, sizeof(int) * (2 * n)));
    CUDA_TRY(cudaMalloc(&d_visits, sizeof(int) * std::max(n, 1)));
    CUDA_TRY(cudaMemset(d_visits, 0, sizeof(int) * std::max(n, 1)));
    CUDA_TRY(cudaMalloc(&B.nodes, sizeof(BvhNode) * std::max(n - 1, 1)));
    CUDA_TRY(cudaMalloc(&B.tri_v, sizeof(float4) * 3 * n));
    CUDA_TRY(cudaMalloc(&B.tri_id, sizeof(int) * n));
    const float3 flo = make_float3(lo[0], lo[1], lo[2]);
    const float3 inv = make_float3(1.f / std::max(hi[0] - lo[0], 1e-9f), 1.f / std::max(hi[1] - lo[1], 1e-9f),
                                   1.f / std::max(hi[2] - lo[2], 1e-9f));
    const int blocks = (n + 127) / 128;
    k_lbvh_morton<<<blocks, 128>>>(d_v, d_t, n, flo, inv, d_keys, d_vals);
    CUDA_TRY(cub::DeviceRadixSort::SortPairs(nullptr, tmp_bytes, d_keys, d_keys2, d_vals, d_vals2, n));
    CUDA_TRY(cudaMalloc(&d_tmp, std::max<size_t>(tmp_bytes, 16)));
    CUDA_TRY(cub::DeviceRadixSort::SortPairs(d_tmp, tmp_bytes, d_keys, d_keys2, d_vals, d_vals2, n));
    k_lbvh_gather<<<blocks, 128>>>(d_v, d_t, d_vals2, n, B.tri_v, B.tri_id);
    if (n > 1) {
        k_lbvh_hierarchy<<<blocks, 128>>>(d_keys2, n, B.nodes, d_parent);
        k_lbvh_refit<<<blocks, 128>>>(B.tri_v, n, B.nodes, d_parent, d_visits);
    }
    CUDA_TRY(cudaDeviceSynchronize());
    CUDA_TRY(cudaGetLastError());
    cudaFree(d_v); cudaFree(d_t); cudaFree(d_keys); cudaFree(d_keys2); cudaFree(d_vals); cudaFree(d_vals2);
    cudaFree(d_parent); cudaFree(d_visits); cudaFree(d_tmp);
    return JG_OK;
}

constexpr int BP_BLOCK = 128;   // broad phase: threads (= configurations) per block
constexpr int NP_BLOCK = 256;   // narrow phase
constexpr int EPA_BIG_BLOCK = 512;  // overflow pass of the penetration depth: one pair per warp, 16 warps per block

static void launch_broadphase(const jg_engine* e, const PassArgs& a, int grid, size_t smem, cudaStream_t st) {
    if (e->mesh) k_broadphase<BP_BLOCK, true><<<grid, BP_BLOCK, smem, st>>>(a);
    else k_broadphase<BP_BLOCK, false><<<grid, BP_BLOCK, smem, st>>>(a);
}

#define JG_COUNTS_INTS(n_keys) (7 * (n_keys) + 8)   /* layout: jg_kernels.cuh cnt_*() */

static size_t bp_smem(const jg_engine* e, uint32_t flags) {
    size_t s = sizeof(RobotTab) + sizeof(PieceRec) * e->P + sizeof(float) * BP_BLOCK * e->dof;
    if (flags & JG_SELF) s += sizeof(float) * 12 * e->n_stash * BP_BLOCK;
    return s;
}

static int arena_alloc(jg_engine* e, int cap) {
    e->cap = cap;
    e->ovf_poly_cap = std::min(std::max(cap / 4, 1), 65536);   // hand-over records (496 B each); later overflows restart
    const size_t qn = (size_t)std::max(e->n_keys, 1) * cap;
    bool okm = cudaMalloc(&e->q_r0, sizeof(float4) * qn) == cudaSuccess && cudaMalloc(&e->q_r1, sizeof(float4) * qn) == cudaSuccess &&
               cudaMalloc(&e->q_r2, sizeof(float4) * qn) == cudaSuccess && cudaMalloc(&e->q_cfg, sizeof(int) * qn) == cudaSuccess &&
               cudaMalloc(&e->epa_slot, sizeof(int) * qn) == cudaSuccess && cudaMalloc(&e->epa_ids, sizeof(uint4) * qn) == cudaSuccess &&
               cudaMalloc(&e->ovf_list, sizeof(int2) * 4 * (size_t)cap) == cudaSuccess &&
               cudaMalloc(&e->ovf_poly, sizeof(uint32_t) * JG_OVF_WORDS * (size_t)e->ovf_poly_cap) == cudaSuccess &&
               cudaMalloc(&e->epa_prio, sizeof(float) * qn) == cudaSuccess && cudaMalloc(&e->q_prio, sizeof(float) * qn) == cudaSuccess &&
               cudaMalloc(&e->act_list, sizeof(int) * qn) == cudaSuccess && cudaMalloc(&e->act_meta, sizeof(int2) * qn) == cudaSuccess &&
               cudaMalloc(&e->champ_scene, sizeof(unsigned long long) * cap) == cudaSuccess &&
               cudaMalloc(&e->champ_self, sizeof(unsigned long long) * cap) == cudaSuccess &&
               cudaMalloc(&e->enc_scene, sizeof(int) * cap) == cudaSuccess && cudaMalloc(&e->enc_self, sizeof(int) * cap) == cudaSuccess &&
               cudaMalloc(&e->viol, sizeof(uint32_t) * (cap / 32 + 1)) == cudaSuccess;
    if (okm && e->mesh) {
        e->mesh_cfg_cap = std::min(cap, 1 << 19);    // 48 B per (configuration, shape): 300 MB at 12 shapes
        okm = cudaMalloc(&e->mesh_T, sizeof(float4) * 3 * (size_t)e->mesh_cfg_cap * std::max(e->S, 1)) == cudaSuccess;
    }
    if (!okm) {
        const int rc = fail(JG_ERR_NOMEM, "scratch allocation failed (%zu queue entries): %s", qn, cudaGetErrorString(cudaGetLastError()));
        arena_free(e);
        return rc;
    }
    return JG_OK;
}

static int drain_pending(jg_engine* e);

// make the arena hold passes of min(want, cap_max) configurations (power-of-two growth; everything in flight is
// finished first: growth is a rare set-up event, not part of a steady-state call)
static int ensure_arena(jg_engine* e, int64_t want) {
    const int need = (int)std::min<int64_t>(e->cap_max, std::max<int64_t>(want, 1));
    if (need <= e->cap) return JG_OK;
    int cap = std::max(e->cap, 1024);
    while (cap < need) cap <<= 1;
    cap = std::min(cap, e->cap_max);
    int rc = drain_pending(e);
    if (rc) return rc;
    CUDA_TRY(cudaDeviceSynchronize());
    arena_free(e);      // (the host path's staging buffers are separate: ensure_staging grows them when a host call needs it)
    e->mesh_pass = 0;
    return arena_alloc(e, cap);
}

extern "C" jg_engine* jg_engine_create(const jg_robot* r, const jg_scene* sc, int device, const jg_params* prm) {
    if (!r) { fail(JG_ERR_INVALID, "jg_engine_create: robot is NULL"); return nullptr; }
    int ndev = 0;
    cudaError_t ce = cudaGetDeviceCount(&ndev);
    if (ce != cudaSuccess || ndev <= 0) {
        fail(JG_ERR_NODEVICE, "no CUDA device available (%s); this engine has no CPU fallback",
             ce == cudaSuccess ? "device count 0" : cudaGetErrorString(ce));
        return nullptr;
    }
    if (device < 0 || device >= ndev) { fail(JG_ERR_INVALID, "device %d out of range [0,%d)", device, ndev); return nullptr; }
    if (cudaSetDevice(device) != cudaSuccess) { fail(JG_ERR_CUDA, "cudaSetDevice(%d) failed", device); return nullptr; }
    jg_engine* e = new jg_engine;
    e->device = device;
    if (prm) e->prm = *prm; else jg_default_params(&e->prm);
    e->L = r->L; e->S = r->S; e->dof = r->dof;
    e->robot_copy = *r;                   // (a few hundred KB: what sibling engines are built from)
    if (sc) { e->scene_copy = *sc; e->has_scene = true; }
    if (build_tables(e, r, sc) != JG_OK) { engine_release(e); return nullptr; }
    if (e->mesh && build_lbvh(e, sc) != JG_OK) { engine_release(e); return nullptr; }
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) == cudaSuccess) e->sm_count = prop.multiProcessorCount;
    e->smem_narrow = sizeof(float4) * e->n_verts + sizeof(KeyRec) * e->n_keys + sizeof(int) * (e->n_keys + 2) + 64;
    const size_t smem_max = 227 * 1024;
    if (e->smem_narrow > smem_max || bp_smem(e, JG_SELF) > smem_max) {
        fail(JG_ERR_INVALID, "scene too large for the shared-memory resident path (%zu B of vertices/keys)", e->smem_narrow);
        engine_release(e);
        return nullptr;
    }
    bool ok = cudaFuncSetAttribute(k_narrowphase<NP_BLOCK>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_max) == cudaSuccess;
    ok = ok && cudaFuncSetAttribute(k_penetration<416>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_max) == cudaSuccess;
    ok = ok && cudaFuncSetAttribute(k_penetration<384>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_max) == cudaSuccess;
    ok = ok && cudaFuncSetAttribute(k_penetration<352>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_max) == cudaSuccess;
    ok = ok && cudaFuncSetAttribute(k_penetration<320>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_max) == cudaSuccess;
    ok = ok && cudaFuncSetAttribute(k_penetration<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_max) == cudaSuccess;
    ok = ok && cudaFuncSetAttribute(k_penetration<192>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_max) == cudaSuccess;
    ok = ok && cudaFuncSetAttribute(k_penetration<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_max) == cudaSuccess;
    ok = ok && cudaFuncSetAttribute(k_penetration<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_max) == cudaSuccess;
    ok = ok && cudaFuncSetAttribute(k_penetration_big<EPA_BIG_BLOCK>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_max) == cudaSuccess;
    ok = ok && cudaFuncSetAttribute(k_broadphase<BP_BLOCK, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_max) == cudaSuccess;
    ok = ok && cudaFuncSetAttribute(k_broadphase<BP_BLOCK, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_max) == cudaSuccess;
    ok = ok && cudaFuncSetAttribute(k_broadphase_poses<BP_BLOCK, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_max) == cudaSuccess;
    ok = ok && cudaFuncSetAttribute(k_broadphase_poses<BP_BLOCK, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_max) == cudaSuccess;
    if (!ok) {
        fail(JG_ERR_CUDA, "cudaFuncSetAttribute failed: %s (is this an sm_100 device?)", cudaGetErrorString(cudaGetLastError()));
        engine_release(e);
        return nullptr;
    }
    // persistent grids: one resident wave of each narrow-phase kernel
    {
        int occ = 0;
        if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_narrowphase<NP_BLOCK>, NP_BLOCK, e->smem_narrow) == cudaSuccess && occ > 0)
            e->np_blocks = e->sm_count * occ;
        else
            e->np_blocks = e->sm_count * 2;
        // penetration pass: one block per SM, as many threads as shared-memory polytopes fit beside the vertex table
        const size_t tables = sizeof(float4) * e->n_verts + sizeof(int) * ((e->n_keys + 2 + 3) & ~3);
        const size_t per_thread = sizeof(uint32_t) * EpaCompact<1>::WORDS;
        e->epa_threads = 0;
        int want = 416;
        if (const char* ev = getenv("JG_EPA_THREADS")) want = atoi(ev);   // tuning knob (experiments only)
        for (int t : {416, 384, 352, 320, 256, 192, 128, 64})
            if (t <= want && tables + per_thread * t + 64 <= smem_max) { e->epa_threads = t; break; }
        if (!e->epa_threads) {
            fail(JG_ERR_INVALID, "scene too large for the shared-memory penetration pass (%zu B of tables)", tables);
            engine_release(e);
            return nullptr;
        }
        e->smem_epa = tables + per_thread * e->epa_threads + 64;
        e->smem_epa_big = sizeof(float4) * e->n_verts + sizeof(uint32_t) * EpaPolyBig::WORDS * (EPA_BIG_BLOCK / 32) + 64;
        e->epa_blocks = e->sm_count;
    }
    // scratch arena: per-key queues of `cap` entries (92 B each).  Default: the largest power of two <= 2^20
    // configurations per pass whose arena stays within 16 GB and a quarter of the free device memory.
    int cap = e->prm.chunk;
    if (cap > (1 << 20)) {   // queue slots are packed into 20 bits of the champion codes (queue_push / champ_publish / k_select)
        fail(JG_ERR_INVALID, "jg_params.chunk = %d exceeds the maximum pass size 2^20 = 1048576 (larger batches are cut into passes)", cap);
        engine_release(e);
        return nullptr;
    }
    if (cap <= 0) {
        size_t free_b = 0, total_b = 0;
        cudaMemGetInfo(&free_b, &total_b);
        const size_t budget = std::min<size_t>((size_t)16 << 30, free_b / 4);
        const size_t per_cfg = (size_t)std::max(e->n_keys, 1) * 92 + 48;
        cap = 1 << 20;
        while (cap > 4096 && (size_t)cap * per_cfg > budget) cap >>= 1;
    }
    cap = std::min((cap + 1023) / 1024 * 1024, 1 << 20);
    static_assert((1 << 20) - 1 == 0xFFFFF, "slot field of the champion code");
    e->cap_max = cap;
    bool okm = cudaMalloc(&e->counts, sizeof(int) * JG_COUNTS_INTS(e->n_keys)) == cudaSuccess &&
               cudaMalloc(&e->d_link_slot, sizeof(int) * MAX_LINKS) == cudaSuccess &&
               cudaMalloc(&e->d_rel, sizeof(Tf) * MAX_SHAPES) == cudaSuccess &&
               cudaMalloc(&e->d_use_shapes, sizeof(int) * MAX_SHAPES) == cudaSuccess &&
               cudaMalloc(&e->d_stats, sizeof(unsigned long long) * 8) == cudaSuccess && cudaMalloc(&e->d_err, sizeof(int)) == cudaSuccess && cudaMalloc(&e->d_qrest, sizeof(float) * MAX_DOF) == cudaSuccess;
    if (!okm) {
        fail(JG_ERR_NOMEM, "table allocation failed: %s", cudaGetErrorString(cudaGetLastError()));
        engine_release(e);
        return nullptr;
    }
    // an explicit chunk is allocated at once; the automatic size starts small (single-shot queries of the drop-in API
    // never need more) and grows with the batches that actually arrive (ensure_arena)
    if (arena_alloc(e, e->prm.chunk > 0 ? e->cap_max : std::min(e->cap_max, 4096)) != JG_OK) {
        engine_release(e);
        return nullptr;
    }
    cudaMemset(e->d_stats, 0, sizeof(unsigned long long) * 8);
    cudaMemset(e->d_err, 0, sizeof(int));
    return e;
}
extern "C" void jg_engine_free(jg_engine* e) { engine_release(e); }
extern "C" int jg_engine_device(const jg_engine* e) { return e ? e->device : JG_ERR_INVALID; }
extern "C" int64_t jg_engine_pass_size(const jg_engine* e) { return e ? (int64_t)e->cap_max : (int64_t)JG_ERR_INVALID; }

// ----------------------------------------------------------------------------------------------- passes
// want_dist = false: the caller only wants verdict bits.  The verdict is "some contact distance < threshold", so the
// query distance shrinks to the threshold itself (pairs farther apart cannot change a bit: the broad phase drops
// them and GJK stops as soon as a separating axis proves it) and penetration depth is never needed.  Same bits.
// The shortcut is exact only while every overlap is a collision: overlapping cores report exactly -(margin sum), so the
// threshold must lie above -(smallest margin sum); for a lower threshold (a free parameter of links_in_collision /
// are_in_collision) the penetration stages stay on and only the distance output is dropped.
static void fill_args(const jg_engine* e, PassArgs& a, uint32_t& flags, bool want_dist = true) {
    memset(&a, 0, sizeof a);
    if (!want_dist && (e->mesh || e->n_keys == 0 || e->prm.threshold > -e->min_msum + 1e-6f)) flags |= JG_NO_EPA;
    a.dof = e->dof; a.substeps = 1; a.flags = flags;
    a.finger_open = e->prm.finger_open; a.threshold = e->prm.threshold;
    a.qd = want_dist ? e->prm.query_distance : std::min(e->prm.query_distance, e->prm.threshold + 1e-6f);
    for (int k = 0; k < 4; ++k) a.plane[k] = e->plane[k];
    a.has_plane = e->has_plane ? 1 : 0;
    a.robot = e->d_robot; a.pieces = e->d_pieces; a.P = e->P; a.selfs = e->d_selfs; a.keys = e->d_keys;
    a.n_keys = e->n_keys; a.n_scene_keys = e->n_scene_keys; a.verts = e->d_verts; a.n_verts = e->n_verts;
    a.q_r0 = e->q_r0; a.q_r1 = e->q_r1; a.q_r2 = e->q_r2; a.q_cfg = e->q_cfg; a.epa_slot = e->epa_slot; a.epa_ids = e->epa_ids; a.ovf_list = e->ovf_list; a.ovf_cap = 4 * e->cap; a.ovf_poly = e->ovf_poly; a.ovf_poly_cap = e->ovf_poly_cap; a.epa_prio = e->epa_prio; a.q_prio = e->q_prio; a.max_msum = e->max_msum; a.stats = e->d_stats;
    a.mesh = e->mesh ? 1 : 0; a.bvh_nodes = e->bvh.nodes; a.bvh_tri = e->bvh.tri_v; a.bvh_n = e->bvh.n_tris;
    a.mesh_margin = e->mesh_margin; a.mesh_T = e->mesh_T; a.mesh_n_use = e->S; a.q_tri = e->epa_slot; a.err_flag = e->d_err; a.act_list = e->act_list; a.act_meta = e->act_meta; a.champ_scene = e->champ_scene; a.champ_self = e->champ_self; a.cap = e->cap; a.counts = e->counts;
    a.enc_scene = e->enc_scene; a.enc_self = e->enc_self; a.viol = e->viol;
}

// Launch a kernel of a pass so that it may start while its predecessor in the stream is still draining (programmatic
// dependent launch: the kernel waits for the predecessor's results at its JG_WAIT_FOR_PREDECESSOR()).  Off while the
// per-section profiling events sit between the kernels, or with JG_NO_PDL=1 in the environment.
template <class... KArgs, class... Args>
static void launch_dep(const jg_engine* e, void (*kernel)(KArgs...), int grid, int block, size_t smem, cudaStream_t st, Args... args) {
    static const bool disabled = getenv("JG_NO_PDL") != nullptr;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)grid);
    cfg.blockDim = dim3((unsigned)block);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = (disabled || e->profiling) ? 0 : 1;
    cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}

// narrow phase + finalize of the pass described by `a` (queues already filled)
static int run_narrow_and_finalize(jg_engine* e, const PassArgs& a, int n_out, uint32_t flags, uint32_t* bits, float* dist,
                                   uint8_t* reason, cudaStream_t st) {
    prof_end(e, st);   // broad phase
    // Stages (each works on lists built by k_select):  GJK on every configuration's champion pair (+ plane entries),
    // penetration depth of the champions that overlap, GJK on the pairs that can still matter, their penetration
    // depth, then the rare large polytopes of both penetration stages.
    const bool epa = !(flags & JG_NO_EPA);
    const size_t sel_smem = sizeof(int) * (2 * e->n_keys + 2);
    auto launch_penetration = [&](const PassArgs& a) {
#ifdef JG_TRACE
        {
            static int trace_slot = 0;
            const int slot = trace_slot++ & 1;
            cudaMemcpyToSymbolAsync(g_jg_trace_slot, &slot, sizeof(int), 0, cudaMemcpyHostToDevice, st);
        }
#endif
        switch (e->epa_threads) {
            case 416: launch_dep(e, k_penetration<416>, e->epa_blocks, 416, e->smem_epa, st, a); break;
            case 384: launch_dep(e, k_penetration<384>, e->epa_blocks, 384, e->smem_epa, st, a); break;
            case 352: launch_dep(e, k_penetration<352>, e->epa_blocks, 352, e->smem_epa, st, a); break;
            case 320: launch_dep(e, k_penetration<320>, e->epa_blocks, 320, e->smem_epa, st, a); break;
            case 256: launch_dep(e, k_penetration<256>, e->epa_blocks, 256, e->smem_epa, st, a); break;
            case 192: launch_dep(e, k_penetration<192>, e->epa_blocks, 192, e->smem_epa, st, a); break;
            case 128: launch_dep(e, k_penetration<128>, e->epa_blocks, 128, e->smem_epa, st, a); break;
            default: launch_dep(e, k_penetration<64>, e->epa_blocks, 64, e->smem_epa, st, a); break;
        }
    };
    // The stage lists live in two sets: every k_select builds the next stage's lists in one set and clears the set of
    // the stage that has just finished (and books its entry count into stats[4] GJK pairs / stats[5] polytopes).
    int set = 0, n_sel = 0, prev_slot = -1;
    PassArgs as = a;
    if (e->n_keys > 0) {
        for (int round = 0; round < 2; ++round) {
            prof_begin(e, 1, st);
            // the broad phase already wrote the stage-1 lists itself (set 0, thread-local champions) unless the champions
            // were found with atomics (edge substeps, mesh queues)
            const bool listed = round == 0 && a.substeps <= 1 && !a.mesh;
            if (!listed) {
                if (prev_slot >= 0) set ^= 1;
                as.stage_set = set;
                launch_dep(e, k_select<256>, e->sm_count * 4, 256, sel_smem, st, as, (int)(round == 0 ? SEL_GJK_CHAMP : SEL_GJK_REST), n_sel++,
                           prev_slot, (int)(round == 1 && epa ? 1 : 0));
                e->stats.kernel_launches += 1;
            }
            as.stage_set = set;
            launch_dep(e, k_narrowphase<NP_BLOCK>, e->np_blocks, NP_BLOCK, e->smem_narrow, st, as);
            prev_slot = 4;
            prof_end(e, st);
            e->stats.kernel_launches += 1;
            if (epa) {
                prof_begin(e, 2, st);
                set ^= 1;
                as.stage_set = set;
                launch_dep(e, k_select<256>, e->sm_count * 4, 256, sel_smem, st, as, (int)SEL_EPA_NEW, n_sel++, prev_slot, 0);
                launch_penetration(as);
                prev_slot = 5;
                prof_end(e, st);
                e->stats.kernel_launches += 2;
            }
        }
        if (epa) {
            prof_begin(e, 2, st);
            launch_dep(e, k_penetration_big<EPA_BIG_BLOCK>, e->sm_count * 2, EPA_BIG_BLOCK, e->smem_epa_big, st, a);
            prof_end(e, st);
            e->stats.kernel_launches += 1;
        }
    }
    prof_begin(e, 3, st);
    // one extra block of k_finalize folds this pass's queue counters into the engine statistics (and books the last
    // stage, which no k_select followed)
    StatArgs sa = {nullptr, 0, 0, 0, nullptr, 0, -1};
    if (e->n_keys > 0) sa = StatArgs{e->counts, e->n_scene_keys, e->n_keys, e->P, e->d_stats, set, prev_slot};
    launch_dep(e, k_finalize, (n_out + 255) / 256 + 1, 256, 0, st, (const int*)e->enc_scene, (const int*)e->enc_self, (const uint32_t*)e->viol,
               n_out, a.qd, a.threshold, flags, bits, dist, reason, sa);
    e->stats.kernel_launches += 1;
    prof_end(e, st);
    CUDA_TRY(cudaGetLastError());
    return JG_OK;
}

// All entry points share ONE scratch arena (queues, counters, per-configuration slots).  Calls on the same stream are
// ordered by the stream; a call on a DIFFERENT stream than the previous one (torch stream after the library's own
// host-path stream, or the other way round) first waits for everything the previous call enqueued.
static int begin_call(jg_engine* e, cudaStream_t st) {
    CUDA_TRY(cudaSetDevice(e->device));
    // a stream that is being captured into a CUDA graph cannot wait on uncaptured work: the caller synchronises before
    // the capture (CollisionEngine.capture_check does) and orders the replays against other calls itself
    cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
    const bool capturing = cudaStreamIsCapturing(st, &cs) == cudaSuccess && cs == cudaStreamCaptureStatusActive;
    if (!capturing) {
        if (e->have_last_stream && st != e->last_stream) {
            if (!e->ev_order) CUDA_TRY(cudaEventCreateWithFlags(&e->ev_order, cudaEventDisableTiming));
            if (cudaEventRecord(e->ev_order, e->last_stream) == cudaSuccess) CUDA_TRY(cudaStreamWaitEvent(st, e->ev_order, 0));
            else cudaGetLastError();   // the previous stream no longer exists: its work has completed
        }
        e->have_last_stream = true;
        e->last_stream = st;
    }
    e->last_call_lanes = 0;
    if (!e->accumulate_stats) {
        e->stats = jg_stats{};
        CUDA_TRY(cudaMemsetAsync(e->d_stats, 0, sizeof(unsigned long long) * 8, st));
    }
    return JG_OK;
}

// ---- lanes
extern "C" int jg_engine_set_lanes(jg_engine* e, int lanes) {
    if (!e) return fail(JG_ERR_INVALID, "jg_engine_set_lanes: engine is NULL");
    if (lanes < 1 || lanes > 8) return fail(JG_ERR_INVALID, "jg_engine_set_lanes: %d lanes out of range [1,8]", lanes);
    CUDA_TRY(cudaSetDevice(e->device));
    while ((int)e->sib.size() > lanes - 1) {   // fewer lanes: give the siblings' arenas back
        CUDA_TRY(cudaDeviceSynchronize());
        engine_release(e->sib.back());
        e->sib.pop_back();
    }
    e->lanes = lanes;
    return JG_OK;
}
extern "C" int jg_engine_lanes(const jg_engine* e) { return e ? e->lanes : JG_ERR_INVALID; }

static int ensure_lanes(jg_engine* e) {
    while ((int)e->sib.size() < e->lanes - 1) {
        jg_engine* sb = jg_engine_create(&e->robot_copy, e->has_scene ? &e->scene_copy : nullptr, e->device, &e->prm);
        if (!sb) return JG_ERR_CUDA;   // jg_last_error() holds the reason
        e->sib.push_back(sb);
    }
    while ((int)e->lane_stream.size() < e->lanes) {
        cudaStream_t ls;
        cudaEvent_t ev;
        CUDA_TRY(cudaStreamCreateWithFlags(&ls, cudaStreamNonBlocking));
        e->lane_stream.push_back(ls);
        CUDA_TRY(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
        e->lane_ev.push_back(ev);
    }
    if (!e->ev_fork) CUDA_TRY(cudaEventCreateWithFlags(&e->ev_fork, cudaEventDisableTiming));
    return JG_OK;
}

// may this call be cut over the lanes?  (not while profiling: sections of overlapping passes would time each other; not
// while its stream is being captured; not from a sibling)
static bool use_lanes(const jg_engine* e, int64_t n, int64_t unit, cudaStream_t st) {
    if (e->lanes < 2 || n <= unit || e->profiling) return false;
    cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
    return !(cudaStreamIsCapturing(st, &cs) == cudaSuccess && cs == cudaStreamCaptureStatusActive);
}

// n units in the fewest equal pieces of at most `unit` (multiples of 32: the packed words of the pieces stay aligned),
// piece i on lane i % lanes: fork from the caller's stream, join back into it.  piece(engine, offset, count, stream).
template <class F>
static int run_lanes(jg_engine* e, int64_t n, int64_t unit, cudaStream_t st, F piece) {
    int rc = ensure_lanes(e);
    if (rc) return rc;
    unit = std::max<int64_t>(32, unit / 32 * 32);
    const int64_t k = (n + unit - 1) / unit;
    const int64_t p = std::max<int64_t>(32, std::min<int64_t>(unit, ((n + k - 1) / k + 31) / 32 * 32));
    const int L = (int)std::min<int64_t>(e->lanes, (n + p - 1) / p);
    CUDA_TRY(cudaEventRecord(e->ev_fork, st));
    for (int l = 0; l < L; ++l) CUDA_TRY(cudaStreamWaitEvent(e->lane_stream[l], e->ev_fork, 0));
    int64_t i = 0;
    for (int64_t off = 0; off < n && rc == JG_OK; off += p, ++i) {
        const int l = (int)(i % L);
        jg_engine* le = l == 0 ? e : e->sib[l - 1];
        le->accumulate_stats = i >= L;
        rc = piece(le, off, std::min<int64_t>(p, n - off), e->lane_stream[l]);
        le->accumulate_stats = false;
    }
    for (int l = 0; l < L; ++l) {   // join (also after an error: nothing may still run on the lanes when the caller reads `st`)
        if (cudaEventRecord(e->lane_ev[l], e->lane_stream[l]) == cudaSuccess) cudaStreamWaitEvent(st, e->lane_ev[l], 0);
    }
    e->last_call_lanes = L;
    return rc;
}

// Mesh mode: a (pose, link shape) may queue many triangle candidates, so a pass takes fewer configurations than the
// per-key queue capacity; the pass size adapts.  Called when a pass reported a queue overflow: halve the pass, or -- at
// the smallest pass -- grow the arena (while it is below cap_max).  > 0: redo the pass.
static int mesh_overflow_retry(jg_engine* e, cudaStream_t st, const char* who, int min_pass = 256) {
    CUDA_TRY(cudaMemsetAsync(e->d_err, 0, sizeof(int), st));
    if (e->mesh_pass > min_pass) {
        e->mesh_pass = std::max(min_pass, e->mesh_pass / 2 / 32 * 32);
        return 1;
    }
    if (e->cap < e->cap_max) {
        const int rc = ensure_arena(e, std::min<int64_t>(e->cap_max, (int64_t)e->cap * 4));
        if (rc) return rc;
        e->mesh_pass = std::max(min_pass, std::min(e->cap / 2, e->mesh_cfg_cap) / 32 * 32);
        return 1;
    }
    return fail(JG_ERR_NOMEM, "%s: mesh mode: a triangle candidate queue overflowed even at %d configurations per pass "
                "(%d entries per link shape, the maximum): simplify the obstacle mesh", who, min_pass, e->cap);
}
static void mesh_first_pass(jg_engine* e) {
    if (e->mesh && e->mesh_pass <= 0) e->mesh_pass = std::max(256, std::min(e->cap / 2, e->mesh_cfg_cap) / 32 * 32);
    if (e->mesh) e->mesh_pass = std::min(e->mesh_pass, e->mesh_cfg_cap / 32 * 32);
}

constexpr int MESH_GROUP = 8;   // lanes that walk the LBVH together for one (configuration, shape) query
// the triangle candidates of a mesh-mode pass (the broad phase has put the shape transforms down): one group of lanes per
// query.  use_shapes == NULL: every shape of the robot (k_broadphase), else the gripper's shape list (k_broadphase_poses).
// Boolean fast path: verdict only, zero margins, every queried shape is exactly its box -> exact box-vs-triangle tests
// in the traversal itself, nothing queued.  Returns 1 when the pass must be redone (queue overflow), < 0 on error.
static int mesh_candidates(jg_engine* e, PassArgs& a, int n_cfg, int n_use, const int* d_use, const std::vector<int>& use_host,
                           uint32_t flags, cudaStream_t st, const char* who, int min_pass = 256) {
    if (!(flags & JG_SCENE)) return 0;
    bool boolean = (flags & JG_NO_EPA) && a.qd <= 1e-5f && e->mesh_margin == 0.f;
    for (int k = 0; boolean && k < n_use; ++k) {
        const ShapeRec& Sh = e->h_robot.shapes[use_host.empty() ? k : use_host[k]];
        boolean = Sh.is_box && Sh.margin == 0.f;
    }
    a.mesh_n_use = n_use;
    const int n_q = n_cfg * n_use;
    const int per_block = 256 / MESH_GROUP;
    const size_t smem = sizeof(int) * per_block * 32 * MESH_GROUP;
    if (n_q > 0) {
        if (boolean) k_mesh_traverse<MESH_GROUP, true><<<(n_q + per_block - 1) / per_block, 256, smem, st>>>(a, n_q, d_use);
        else k_mesh_traverse<MESH_GROUP, false><<<(n_q + per_block - 1) / per_block, 256, smem, st>>>(a, n_q, d_use);
        e->stats.kernel_launches += 1;
    }
    int err = 0;
    prof_end(e, st);
    CUDA_TRY(cudaMemcpyAsync(&err, e->d_err, sizeof(int), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    if (err & 2) {
        cudaMemsetAsync(e->d_err, 0, sizeof(int), st);
        return fail(JG_ERR_UNSUPPORTED, "%s: mesh mode: the LBVH traversal stack of a query overflowed (%d entries): the tree over this "
                    "mesh is too deep / too wide for the shared-memory stack; no result was produced", who, 32 * MESH_GROUP);
    }
    if (err) {
        const int rc = mesh_overflow_retry(e, st, who, min_pass);
        return rc < 0 ? rc : 1;
    }
    prof_begin(e, 0, st);
    return 0;
}

extern "C" int jg_check(jg_engine* e, const float* q, int64_t n, int dof, uint32_t flags, uint32_t* bits, float* min_dist,
                        uint8_t* reason, void* stream) {
    if (!e) return fail(JG_ERR_INVALID, "jg_check: engine is NULL");
    if (dof != e->dof) return fail(JG_ERR_INVALID, "jg_check: dof %d does not match the robot's %d", dof, e->dof);
    if (n < 0 || (n && (!q || !bits))) return fail(JG_ERR_INVALID, "jg_check: NULL q/bits");
    cudaStream_t st = (cudaStream_t)stream;
    CUDA_TRY(cudaSetDevice(e->device));
    if (use_lanes(e, n, e->cap_max, st))
        return run_lanes(e, n, e->cap_max, st, [&](jg_engine* le, int64_t off, int64_t m, cudaStream_t ls) {
            return jg_check(le, q + off * dof, m, dof, flags, bits + off / 32, min_dist ? min_dist + off : nullptr,
                            reason ? reason + off : nullptr, ls);
        });
    int rc = ensure_arena(e, e->mesh ? n * 16 : n);
    if (rc) return rc;
    rc = begin_call(e, st);
    if (rc) return rc;
    e->stats.configs += n;
    mesh_first_pass(e);
    for (int64_t off = 0; off < n;) {
        const int64_t pass = e->mesh ? e->mesh_pass : e->cap;
        const int m = (int)std::min<int64_t>(pass, n - off);
        PassArgs a;
        fill_args(e, a, flags, min_dist != nullptr);
        a.q = q + off * dof; a.n = m;
        CUDA_TRY(cudaMemsetAsync(e->counts, 0, sizeof(int) * JG_COUNTS_INTS(e->n_keys), st));
        CUDA_TRY(cudaMemsetAsync(e->viol, 0, sizeof(uint32_t) * (e->cap / 32 + 1), st));
        prof_begin(e, 0, st);
        launch_broadphase(e, a, (m + BP_BLOCK - 1) / BP_BLOCK, bp_smem(e, flags), st);
        e->stats.kernel_launches += 1;
        if (e->mesh) {   // triangle candidates: group-cooperative LBVH traversal; a full queue means: redo the pass smaller
            rc = mesh_candidates(e, a, m, e->S, nullptr, {}, flags, st, "jg_check");
            if (rc < 0) return rc;
            if (rc > 0) continue;
        }
        rc = run_narrow_and_finalize(e, a, m, flags, bits + off / 32, min_dist ? min_dist + off : nullptr,
                                     reason ? reason + off : nullptr, st);
        if (rc) return rc;
        off += m;
    }
    return JG_OK;
}

extern "C" int jg_check_edges(jg_engine* e, const float* qa, const float* qb, int64_t m, int dof, int substeps,
                              uint32_t flags, uint32_t* bits, float* min_dist, void* stream) {
    if (!e) return fail(JG_ERR_INVALID, "jg_check_edges: engine is NULL");
    if (dof != e->dof) return fail(JG_ERR_INVALID, "jg_check_edges: dof %d does not match the robot's %d", dof, e->dof);
    if (substeps < 2 || substeps > e->cap_max / 32) return fail(JG_ERR_INVALID, "jg_check_edges: substeps %d out of range [2,%d]", substeps, e->cap_max / 32);
    if (m < 0 || (m && (!qa || !qb || !bits))) return fail(JG_ERR_INVALID, "jg_check_edges: NULL qa/qb/bits");
    cudaStream_t st = (cudaStream_t)stream;
    CUDA_TRY(cudaSetDevice(e->device));
    if (use_lanes(e, m, (e->cap_max / substeps) / 32 * 32, st))
        return run_lanes(e, m, (e->cap_max / substeps) / 32 * 32, st, [&](jg_engine* le, int64_t off, int64_t k, cudaStream_t ls) {
            return jg_check_edges(le, qa + off * dof, qb + off * dof, k, dof, substeps, flags, bits + off / 32,
                                  min_dist ? min_dist + off : nullptr, ls);
        });
    int rc = ensure_arena(e, std::max<int64_t>(m, 32) * substeps * (e->mesh ? 16 : 1));
    if (rc) return rc;
    if (e->mesh && 32 * substeps > e->mesh_cfg_cap)
        return fail(JG_ERR_INVALID, "jg_check_edges: mesh mode takes at most %d substeps per edge", e->mesh_cfg_cap / 32);
    rc = begin_call(e, st);
    if (rc) return rc;
    e->stats.configs += m * substeps;
    mesh_first_pass(e);
    for (int64_t off = 0; off < m;) {
        // packed words stay aligned across passes: a pass holds a multiple of 32 edges (mesh mode: adaptive pass size)
        const int pass_cfgs = e->mesh ? std::max(e->mesh_pass, 32 * substeps) : e->cap;
        const int edges_per_pass = (pass_cfgs / substeps) / 32 * 32;
        const int me = (int)std::min<int64_t>(edges_per_pass, m - off);
        PassArgs a;
        fill_args(e, a, flags, min_dist != nullptr);
        a.q = qa + off * dof; a.qb = qb + off * dof; a.n = me * substeps; a.substeps = substeps;
        CUDA_TRY(cudaMemsetAsync(e->counts, 0, sizeof(int) * JG_COUNTS_INTS(e->n_keys), st));
        CUDA_TRY(cudaMemsetAsync(e->viol, 0, sizeof(uint32_t) * (e->cap / 32 + 1), st));
        prof_begin(e, 0, st);
        k_init_scratch<<<(me + 255) / 256, 256, 0, st>>>(e->enc_scene, e->enc_self, e->champ_scene, e->champ_self, me);
        launch_broadphase(e, a, (a.n + BP_BLOCK - 1) / BP_BLOCK, bp_smem(e, flags), st);
        e->stats.kernel_launches += 2;
        if (e->mesh) {
            rc = mesh_candidates(e, a, a.n, e->S, nullptr, {}, flags, st, "jg_check_edges", 32 * substeps);
            if (rc < 0) return rc;
            if (rc > 0) continue;
        }
        rc = run_narrow_and_finalize(e, a, me, flags, bits + off / 32, min_dist ? min_dist + off : nullptr, nullptr, st);
        if (rc) return rc;
        off += me;
    }
    return JG_OK;
}

// host-side double FK (set-up only: rigid offsets of the gripper shapes for pose mode)
static void host_fk(const RobotTab& R, float finger_open, double out[][12]) {
    for (int i = 0; i < R.L; ++i) {
        const JointRec& J = R.joints[i];
        double O[12], M[12] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0};
        for (int k = 0; k < 12; ++k) O[k] = J.o[k];
        const double val = J.qidx == -2 ? finger_open : 0.0;
        if (J.type == 1) {
            const double x = J.axis[0], y = J.axis[1], z = J.axis[2], c = std::cos(val), s = std::sin(val), t = 1 - c;
            M[0] = t * x * x + c; M[1] = t * x * y - s * z; M[2] = t * x * z + s * y;
            M[4] = t * x * y + s * z; M[5] = t * y * y + c; M[6] = t * y * z - s * x;
            M[8] = t * x * z - s * y; M[9] = t * y * z + s * x; M[10] = t * z * z + c;
        } else if (J.type == 2) {
            M[3] = J.axis[0] * val; M[7] = J.axis[1] * val; M[11] = J.axis[2] * val;
        }
        auto mul = [](const double* A, const double* B, double* C) {
            for (int r = 0; r < 3; ++r)
                for (int c = 0; c < 4; ++c)
                    C[4 * r + c] = A[4 * r] * B[c] + A[4 * r + 1] * B[4 + c] + A[4 * r + 2] * B[8 + c] + (c == 3 ? A[4 * r + 3] : 0.0);
        };
        double loc[12];
        mul(O, M, loc);
        int parent = -1;
        if (J.parent_sel == 0) parent = i - 1;
        else if (J.parent_sel == 1) { for (int k = i - 1; k >= 0; --k) if (R.joints[k].save) { parent = k; break; } }
        if (parent >= 0) mul(out[parent], loc, out[i]);
        else memcpy(out[i], loc, sizeof loc);
    }
}

extern "C" int jg_check_poses(jg_engine* e, const float* poses, int64_t n, int fmt, int ee_link, const int32_t* links,
                              int n_links, uint32_t flags, uint32_t* bits, float* min_dist, void* stream) {
    if (!e) return fail(JG_ERR_INVALID, "jg_check_poses: engine is NULL");
    if (fmt != JG_POSE_MAT34 && fmt != JG_POSE_POS_QUAT_XYZW) return fail(JG_ERR_INVALID, "jg_check_poses: unknown pose format %d", fmt);
    if (ee_link < 0 || ee_link >= e->L || !links || n_links <= 0) return fail(JG_ERR_INVALID, "jg_check_poses: bad link arguments");
    if (n < 0 || (n && (!poses || !bits))) return fail(JG_ERR_INVALID, "jg_check_poses: NULL poses/bits");
    flags &= (JG_SCENE | JG_PLANE | JG_NO_EPA);
    cudaStream_t st = (cudaStream_t)stream;
    CUDA_TRY(cudaSetDevice(e->device));
    if (use_lanes(e, n, e->cap_max, st)) {
        const int stride = fmt == JG_POSE_MAT34 ? 12 : 7;
        return run_lanes(e, n, e->cap_max, st, [&](jg_engine* le, int64_t off, int64_t k, cudaStream_t ls) {
            return jg_check_poses(le, poses + off * stride, k, fmt, ee_link, links, n_links, flags, bits + off / 32,
                                  min_dist ? min_dist + off : nullptr, ls);
        });
    }
    int rc = ensure_arena(e, e->mesh ? n * 16 : n);
    if (rc) return rc;
    rc = begin_call(e, st);
    if (rc) return rc;
    e->stats.configs += n;
    // rigid offsets rel_s = inv(T_ee) * T_link(s)
    double fk[MAX_LINKS][12];
    host_fk(e->h_robot, e->prm.finger_open, fk);
    std::vector<Tf> rel;
    std::vector<int> use;
    for (int k = 0; k < n_links; ++k) {
        if (links[k] < 0 || links[k] >= e->L) return fail(JG_ERR_INVALID, "jg_check_poses: link id %d out of range", links[k]);
        for (int s = 0; s < e->S; ++s) {
            if (e->shape_link[s] != links[k]) continue;
            const double* A = fk[ee_link];
            const double* B = fk[links[k]];
            double C[12];
            for (int r = 0; r < 3; ++r) {
                for (int c = 0; c < 3; ++c) C[4 * r + c] = A[r] * B[c] + A[4 + r] * B[4 + c] + A[8 + r] * B[8 + c];
                C[4 * r + 3] = A[r] * (B[3] - A[3]) + A[4 + r] * (B[7] - A[7]) + A[8 + r] * (B[11] - A[11]);
            }
            Tf T;
            T.r00 = (float)C[0]; T.r01 = (float)C[1]; T.r02 = (float)C[2]; T.tx = (float)C[3];
            T.r10 = (float)C[4]; T.r11 = (float)C[5]; T.r12 = (float)C[6]; T.ty = (float)C[7];
            T.r20 = (float)C[8]; T.r21 = (float)C[9]; T.r22 = (float)C[10]; T.tz = (float)C[11];
            rel.push_back(T);
            use.push_back(s);
        }
    }
    if (use.empty() || (int)use.size() > MAX_SHAPES) return fail(JG_ERR_INVALID, "jg_check_poses: the listed links carry no collision shapes");
    CUDA_TRY(cudaMemcpyAsync(e->d_rel, rel.data(), sizeof(Tf) * rel.size(), cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaMemcpyAsync(e->d_use_shapes, use.data(), sizeof(int) * use.size(), cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaStreamSynchronize(st));   // rel/use are stack-lifetime host vectors
    const int stride = fmt == JG_POSE_MAT34 ? 12 : 7;
    mesh_first_pass(e);
    for (int64_t off = 0; off < n;) {
        const int64_t pass = e->mesh ? e->mesh_pass : e->cap;   // mesh mode: adaptive pass size, see mesh_overflow_retry
        const int m = (int)std::min<int64_t>(pass, n - off);
        PassArgs a;
        fill_args(e, a, flags, min_dist != nullptr);
        a.n = m;
        a.mesh_n_use = (int)use.size();
        CUDA_TRY(cudaMemsetAsync(e->counts, 0, sizeof(int) * JG_COUNTS_INTS(e->n_keys), st));
        CUDA_TRY(cudaMemsetAsync(e->viol, 0, sizeof(uint32_t) * (e->cap / 32 + 1), st));
        const size_t smem = sizeof(RobotTab) + sizeof(PieceRec) * e->P;
        prof_begin(e, 0, st);
        if (e->mesh)
            k_broadphase_poses<BP_BLOCK, true><<<(m + BP_BLOCK - 1) / BP_BLOCK, BP_BLOCK, smem, st>>>(
                a, poses + off * stride, fmt, e->d_rel, e->d_use_shapes, (int)use.size());
        else
            k_broadphase_poses<BP_BLOCK, false><<<(m + BP_BLOCK - 1) / BP_BLOCK, BP_BLOCK, smem, st>>>(
                a, poses + off * stride, fmt, e->d_rel, e->d_use_shapes, (int)use.size());
        e->stats.kernel_launches += 1;
        if (e->mesh) {
            rc = mesh_candidates(e, a, m, (int)use.size(), e->d_use_shapes, use, flags, st, "jg_check_poses");
            if (rc < 0) return rc;
            if (rc > 0) continue;
        }
        rc = run_narrow_and_finalize(e, a, m, flags, bits + off / 32, min_dist ? min_dist + off : nullptr, nullptr, st);
        if (rc) return rc;
        off += m;
    }
    return JG_OK;
}

extern "C" int jg_fk(jg_engine* e, const float* q, int64_t n, int dof, const int32_t* link_ids, int n_link_ids,
                     const double* base_pose, float* out, void* stream) {
    if (!e) return fail(JG_ERR_INVALID, "jg_fk: engine is NULL");
    if (dof != e->dof) return fail(JG_ERR_INVALID, "jg_fk: dof %d does not match the robot's %d", dof, e->dof);
    if (n < 0 || n > 2000000000LL || (n && (!q || !out)) || !link_ids || n_link_ids <= 0 || n_link_ids > MAX_LINKS)
        return fail(JG_ERR_INVALID, "jg_fk: bad arguments");
    cudaStream_t st = (cudaStream_t)stream;
    CUDA_TRY(cudaSetDevice(e->device));
    int slot[MAX_LINKS];
    for (int i = 0; i < MAX_LINKS; ++i) slot[i] = -1;
    for (int k = 0; k < n_link_ids; ++k) {
        if (link_ids[k] < 0 || link_ids[k] >= e->L) return fail(JG_ERR_INVALID, "jg_fk: link id %d out of range", link_ids[k]);
        if (slot[link_ids[k]] >= 0) return fail(JG_ERR_INVALID, "jg_fk: link id %d listed twice", link_ids[k]);
        slot[link_ids[k]] = k;
    }
    CUDA_TRY(cudaMemcpyAsync(e->d_link_slot, slot, sizeof slot, cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    Tf B = {1, 0, 0, 0, 1, 0, 0, 0, 1, 0, 0, 0};
    if (base_pose) {
        B.r00 = (float)base_pose[0]; B.r01 = (float)base_pose[1]; B.r02 = (float)base_pose[2]; B.tx = (float)base_pose[3];
        B.r10 = (float)base_pose[4]; B.r11 = (float)base_pose[5]; B.r12 = (float)base_pose[6]; B.ty = (float)base_pose[7];
        B.r20 = (float)base_pose[8]; B.r21 = (float)base_pose[9]; B.r22 = (float)base_pose[10]; B.tz = (float)base_pose[11];
    }
    if (n == 0) return JG_OK;
    const size_t smem = sizeof(RobotTab) + sizeof(float) * BP_BLOCK * e->dof;
    k_fk<BP_BLOCK><<<(int)((n + BP_BLOCK - 1) / BP_BLOCK), BP_BLOCK, smem, st>>>(e->d_robot, q, (int)n, dof, e->prm.finger_open,
                                                                              e->d_link_slot, n_link_ids, B, out);
    CUDA_TRY(cudaGetLastError());
    return JG_OK;
}

// ----------------------------------------------------------------------------------------------- host path
static int ensure_staging(jg_engine* e) {
    if (!e->staging_ready) {
        CUDA_TRY(cudaStreamCreateWithFlags(&e->s_in, cudaStreamNonBlocking));
        CUDA_TRY(cudaStreamCreateWithFlags(&e->s_run, cudaStreamNonBlocking));
        CUDA_TRY(cudaStreamCreateWithFlags(&e->s_out, cudaStreamNonBlocking));
        for (int i = 0; i < 2; ++i) {
            CUDA_TRY(cudaEventCreateWithFlags(&e->ev_in[i], cudaEventDisableTiming));
            CUDA_TRY(cudaEventCreateWithFlags(&e->ev_run[i], cudaEventDisableTiming));
            CUDA_TRY(cudaEventCreateWithFlags(&e->ev_out[i], cudaEventDisableTiming));
        }
        e->staging_ready = true;
    }
    if (e->staging_cap >= e->cap) return JG_OK;
    {
        const int rc = drain_pending(e);      // a ticket in flight still uses the old buffers
        if (rc) return rc;
        CUDA_TRY(cudaDeviceSynchronize());
    }
    staging_free(e);
    for (int i = 0; i < 2; ++i) {
        CUDA_TRY(cudaMalloc(&e->stage_q[i], sizeof(float) * e->cap * e->dof));
        CUDA_TRY(cudaMalloc(&e->stage_bits[i], sizeof(uint32_t) * (e->cap / 32)));
        CUDA_TRY(cudaMalloc(&e->stage_dist[i], sizeof(float) * e->cap));
        CUDA_TRY(cudaMalloc(&e->stage_reason[i], e->cap));
        CUDA_TRY(cudaMallocHost(&e->pin_q[i], sizeof(float) * e->cap * e->dof));
        CUDA_TRY(cudaMallocHost(&e->pin_bits[i], sizeof(uint32_t) * (e->cap / 32)));
        CUDA_TRY(cudaMallocHost(&e->pin_dist[i], sizeof(float) * e->cap));
        CUDA_TRY(cudaMallocHost(&e->pin_reason[i], e->cap));
    }
    e->staging_cap = e->cap;
    return JG_OK;
}

static bool is_pinned(const void* p) {
    if (!p) return false;
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, p) != cudaSuccess) { cudaGetLastError(); return false; }
    return at.type == cudaMemoryTypeHost;
}

// One pass of the host path into staging buffer set `b`: sliced upload overlapped with the broad phase of the same pass,
// staged narrow phase, download on the third stream; ev_out[b] fires when the results are in host memory.
static int host_pass(jg_engine* e, int b, const float* q, int64_t off, int m, int dof, uint32_t flags, uint32_t* bits,
                     float* dist, uint8_t* reason, bool pq, bool pb, bool pd, bool pr) {
    cudaStream_t st = e->s_run;
    constexpr int SLICES = 4;
    int rc;
    PassArgs a;
    uint32_t fl = flags;
    fill_args(e, a, fl, dist != nullptr);
    a.q = e->stage_q[b]; a.n = m;
    CUDA_TRY(cudaMemsetAsync(e->counts, 0, sizeof(int) * JG_COUNTS_INTS(e->n_keys), st));
    CUDA_TRY(cudaMemsetAsync(e->viol, 0, sizeof(uint32_t) * (e->cap / 32 + 1), st));
    const int slice = std::max(1024, ((m + SLICES - 1) / SLICES + 1023) / 1024 * 1024);
    prof_begin(e, 0, st);
    for (int c0 = 0; c0 < m; c0 += slice) {
        const int c1 = std::min(m, c0 + slice);
        const float* src = q + (off + c0) * dof;
        if (!pq) {
            memcpy(e->pin_q[b] + (size_t)c0 * dof, src, sizeof(float) * (c1 - c0) * dof);
            src = e->pin_q[b] + (size_t)c0 * dof;
        }
        CUDA_TRY(cudaMemcpyAsync(e->stage_q[b] + (size_t)c0 * dof, src, sizeof(float) * (c1 - c0) * dof, cudaMemcpyHostToDevice, e->s_in));
        CUDA_TRY(cudaEventRecord(e->ev_in[b], e->s_in));
        CUDA_TRY(cudaStreamWaitEvent(st, e->ev_in[b], 0));
        PassArgs as = a;
        as.cfg_base = c0;
        as.n = c1;     // configurations [c0, c1) of the pass
        launch_broadphase(e, as, (c1 - c0 + BP_BLOCK - 1) / BP_BLOCK, bp_smem(e, fl), st);
        e->stats.kernel_launches += 1;
    }
    rc = run_narrow_and_finalize(e, a, m, fl, e->stage_bits[b], dist ? e->stage_dist[b] : nullptr,
                                 reason ? e->stage_reason[b] : nullptr, st);
    if (rc) return rc;
    CUDA_TRY(cudaEventRecord(e->ev_run[b], st));
    CUDA_TRY(cudaStreamWaitEvent(e->s_out, e->ev_run[b], 0));
    CUDA_TRY(cudaMemcpyAsync(pb ? bits + off / 32 : e->pin_bits[b], e->stage_bits[b], sizeof(uint32_t) * ((m + 31) / 32),
                             cudaMemcpyDeviceToHost, e->s_out));
    if (dist) CUDA_TRY(cudaMemcpyAsync(pd ? dist + off : e->pin_dist[b], e->stage_dist[b], sizeof(float) * m, cudaMemcpyDeviceToHost, e->s_out));
    if (reason) CUDA_TRY(cudaMemcpyAsync(pr ? reason + off : e->pin_reason[b], e->stage_reason[b], m, cudaMemcpyDeviceToHost, e->s_out));
    CUDA_TRY(cudaEventRecord(e->ev_out[b], e->s_out));
    return JG_OK;
}

extern "C" int jg_check_host(jg_engine* e, const float* q, int64_t n, int dof, uint32_t flags, uint32_t* bits, float* dist,
                             uint8_t* reason) {
    if (!e) return fail(JG_ERR_INVALID, "jg_check_host: engine is NULL");
    if (dof != e->dof) return fail(JG_ERR_INVALID, "jg_check_host: dof %d does not match the robot's %d", dof, e->dof);
    if (n < 0 || (n && (!q || !bits))) return fail(JG_ERR_INVALID, "jg_check_host: NULL q/bits");
    CUDA_TRY(cudaSetDevice(e->device));
    if (e->pending[0].busy || e->pending[1].busy)
        return fail(JG_ERR_INVALID, "jg_check_host: batches submitted with jg_check_host_submit are still in flight");
    int rc = ensure_arena(e, n);
    if (rc) return rc;
    rc = ensure_staging(e);
    if (rc) return rc;
    const bool pq = is_pinned(q), pb = is_pinned(bits), pd = dist && is_pinned(dist), pr = reason && is_pinned(reason);
    if (n == 0) return JG_OK;
    if (e->mesh) {
        // mesh mode: the device entry point checks its queue-overflow flag per pass (a host round trip), so the host
        // path is a staged copy around it: `cap` configurations at a time through the pinned staging buffers
        const int chunk = e->staging_cap;      // (the arena may grow inside jg_check; the staging buffers do not)
        for (int64_t off = 0; off < n; off += chunk) {
            const int m = (int)std::min<int64_t>(chunk, n - off);
            const float* src = q + off * dof;
            if (!pq) { memcpy(e->pin_q[0], src, sizeof(float) * m * dof); src = e->pin_q[0]; }
            CUDA_TRY(cudaMemcpyAsync(e->stage_q[0], src, sizeof(float) * m * dof, cudaMemcpyHostToDevice, e->s_run));
            rc = jg_check(e, e->stage_q[0], m, dof, flags, e->stage_bits[0], dist ? e->stage_dist[0] : nullptr,
                          reason ? e->stage_reason[0] : nullptr, e->s_run);
            if (rc) return rc;
            CUDA_TRY(cudaMemcpyAsync(pb ? bits + off / 32 : e->pin_bits[0], e->stage_bits[0], sizeof(uint32_t) * ((m + 31) / 32),
                                     cudaMemcpyDeviceToHost, e->s_run));
            if (dist) CUDA_TRY(cudaMemcpyAsync(pd ? dist + off : e->pin_dist[0], e->stage_dist[0], sizeof(float) * m, cudaMemcpyDeviceToHost, e->s_run));
            if (reason) CUDA_TRY(cudaMemcpyAsync(pr ? reason + off : e->pin_reason[0], e->stage_reason[0], m, cudaMemcpyDeviceToHost, e->s_run));
            CUDA_TRY(cudaStreamSynchronize(e->s_run));
            if (!pb) memcpy(bits + off / 32, e->pin_bits[0], sizeof(uint32_t) * ((m + 31) / 32));
            if (dist && !pd) memcpy(dist + off, e->pin_dist[0], sizeof(float) * m);
            if (reason && !pr) memcpy(reason + off, e->pin_reason[0], m);
        }
        return JG_OK;
    }
    // One engine pass per `cap` configurations.  Within a pass the upload is cut into slices: the broad phase of slice
    // i runs while slice i+1 is still crossing PCIe (the broad phase is per-configuration, so it can be launched
    // slice by slice into the same queues); the staged narrow phase then runs once over the whole pass.  Results go
    // back on a third stream while the next pass uploads.  Pageable host memory is staged through pinned buffers.
    cudaStream_t st = e->s_run;
    rc = begin_call(e, st);
    if (rc) return rc;
    e->stats.configs = n;
    const int64_t n_pass = (n + e->cap - 1) / e->cap;
    for (int64_t i = 0; i < n_pass; ++i) {
        const int b = (int)(i & 1);
        const int64_t off = i * e->cap;
        const int m = (int)std::min<int64_t>(e->cap, n - off);
        if (i >= 2) {   // buffer b was used by pass i-2: wait for its download, drain it into pageable user memory
            CUDA_TRY(cudaEventSynchronize(e->ev_out[b]));
            const int64_t o2 = (i - 2) * e->cap;
            if (!pb) memcpy(bits + o2 / 32, e->pin_bits[b], sizeof(uint32_t) * (e->cap / 32));
            if (dist && !pd) memcpy(dist + o2, e->pin_dist[b], sizeof(float) * e->cap);
            if (reason && !pr) memcpy(reason + o2, e->pin_reason[b], e->cap);
        }
        rc = host_pass(e, b, q, off, m, dof, flags, bits, dist, reason, pq, pb, pd, pr);
        if (rc) return rc;
    }
    CUDA_TRY(cudaStreamSynchronize(e->s_out));
    for (int64_t i = std::max<int64_t>(0, n_pass - 2); i < n_pass; ++i) {
        const int b = (int)(i & 1);
        const int64_t off = i * e->cap;
        const int m = (int)std::min<int64_t>(e->cap, n - off);
        if (!pb) memcpy(bits + off / 32, e->pin_bits[b], sizeof(uint32_t) * ((m + 31) / 32));
        if (dist && !pd) memcpy(dist + off, e->pin_dist[b], sizeof(float) * m);
        if (reason && !pr) memcpy(reason + off, e->pin_reason[b], m);
    }
    e->last_stream = st;
    return JG_OK;
}

// Asynchronous form of the host path for callers that stream batches: up to TWO batches in flight, so the upload of
// batch k+1 and the download of batch k overlap the kernels of the other one (a synchronous call can only overlap
// within its own batch).  n <= pass capacity; tickets alternate between the two staging buffer sets.
extern "C" int jg_check_host_submit(jg_engine* e, const float* q, int64_t n, int dof, uint32_t flags, uint32_t* bits, float* dist,
                                    uint8_t* reason, int64_t* ticket) {
    if (!e || !ticket) return fail(JG_ERR_INVALID, "jg_check_host_submit: NULL engine / ticket");
    if (dof != e->dof) return fail(JG_ERR_INVALID, "jg_check_host_submit: dof %d does not match the robot's %d", dof, e->dof);
    if (n <= 0 || !q || !bits) return fail(JG_ERR_INVALID, "jg_check_host_submit: empty batch or NULL q/bits");
    if (e->mesh) {
        // mesh-mode passes check their queue-overflow flag on the host, so there is nothing to overlap: the batch is done
        // when this returns and its ticket is already complete (jg_check_host_wait returns at once)
        if (e->pending[0].busy || e->pending[1].busy) return fail(JG_ERR_INVALID, "jg_check_host_submit: wait for the tickets in flight first");
        const int rc0 = jg_check_host(e, q, n, dof, flags, bits, dist, reason);
        if (rc0) return rc0;
        *ticket = e->next_ticket++;
        return JG_OK;
    }
    if (n > e->cap_max) return fail(JG_ERR_INVALID, "jg_check_host_submit: %lld configurations exceed the pass capacity %d (use jg_check_host)",
                                    (long long)n, e->cap_max);
    CUDA_TRY(cudaSetDevice(e->device));
    int rc = ensure_arena(e, n);
    if (rc) return rc;
    rc = ensure_staging(e);
    if (rc) return rc;
    const int b = (int)(e->next_ticket & 1);
    if (e->pending[b].busy) return fail(JG_ERR_INVALID, "jg_check_host_submit: two batches are already in flight (wait for ticket %lld first)",
                                        (long long)(e->next_ticket - 2));
    rc = begin_call(e, e->s_run);
    if (rc) return rc;
    e->stats.configs = n;
    jg_engine::Pending& P = e->pending[b];
    P.n = n; P.bits = bits; P.dist = dist; P.reason = reason;
    P.pb = is_pinned(bits); P.pd = dist && is_pinned(dist); P.pr = reason && is_pinned(reason);
    rc = host_pass(e, b, q, 0, (int)n, dof, flags, bits, dist, reason, is_pinned(q), P.pb, P.pd, P.pr);
    if (rc) return rc;
    P.busy = true;
    *ticket = e->next_ticket++;
    e->last_stream = e->s_run;
    return JG_OK;
}

extern "C" int jg_check_host_wait(jg_engine* e, int64_t ticket) {
    if (!e) return fail(JG_ERR_INVALID, "jg_check_host_wait: engine is NULL");
    if (ticket < 0 || ticket >= e->next_ticket) return fail(JG_ERR_INVALID, "jg_check_host_wait: unknown ticket %lld", (long long)ticket);
    const int b = (int)(ticket & 1);
    jg_engine::Pending& P = e->pending[b];
    if (!P.busy || ticket + 2 < e->next_ticket) return JG_OK;   // already collected
    CUDA_TRY(cudaSetDevice(e->device));
    CUDA_TRY(cudaEventSynchronize(e->ev_out[b]));
    if (!P.pb) memcpy(P.bits, e->pin_bits[b], sizeof(uint32_t) * ((P.n + 31) / 32));
    if (P.dist && !P.pd) memcpy(P.dist, e->pin_dist[b], sizeof(float) * P.n);
    if (P.reason && !P.pr) memcpy(P.reason, e->pin_reason[b], P.n);
    P.busy = false;
    return JG_OK;
}

static int drain_pending(jg_engine* e) {
    for (int64_t t = std::max<int64_t>(0, e->next_ticket - 2); t < e->next_ticket; ++t) {
        const int rc = jg_check_host_wait(e, t);
        if (rc) return rc;
    }
    return JG_OK;
}

extern "C" int jg_get_stats(jg_engine* e, jg_stats* out) {
    if (!e || !out) return fail(JG_ERR_INVALID, "jg_get_stats: NULL argument");
    CUDA_TRY(cudaSetDevice(e->device));
    CUDA_TRY(cudaStreamSynchronize(e->last_stream));
    unsigned long long h[8];
    CUDA_TRY(cudaMemcpy(h, e->d_stats, sizeof h, cudaMemcpyDeviceToHost));
    *out = e->stats;
    out->pairs = (int64_t)h[0];
    out->plane_pairs = (int64_t)h[1];
    out->self_pairs = (int64_t)h[2];
    out->epa_pairs = (int64_t)h[3];
    out->gjk_run = (int64_t)h[4];
    out->epa_run = (int64_t)h[5];
    out->overflow_run = (int64_t)h[6];
    for (int l = 1; l < e->last_call_lanes; ++l) {   // a call that was cut over the lanes: add the siblings' share
        jg_stats sb;
        e->sib[l - 1]->last_call_lanes = 0;
        const int rc = jg_get_stats(e->sib[l - 1], &sb);
        if (rc) return rc;
        out->configs += sb.configs; out->pairs += sb.pairs; out->plane_pairs += sb.plane_pairs; out->self_pairs += sb.self_pairs;
        out->epa_pairs += sb.epa_pairs; out->gjk_run += sb.gjk_run; out->epa_run += sb.epa_run;
        out->kernel_launches += sb.kernel_launches; out->overflow_run += sb.overflow_run;
    }
    return JG_OK;
}

extern "C" int jg_profile_enable(jg_engine* e, int enable) {
    if (!e) return fail(JG_ERR_INVALID, "jg_profile_enable: engine is NULL");
    CUDA_TRY(cudaSetDevice(e->device));
    if (e->profiling) prof_collect(e);
    e->profiling = enable != 0;
    if (enable) {
        for (int k = 0; k < 4; ++k) { e->prof_ms[k] = 0; e->prof_n[k] = 0; }
        e->prof_used = 0;
    }
    return JG_OK;
}

extern "C" int jg_get_profile(jg_engine* e, double ms[4], int64_t launches[4]) {
    if (!e || !ms || !launches) return fail(JG_ERR_INVALID, "jg_get_profile: NULL argument");
    CUDA_TRY(cudaSetDevice(e->device));
    prof_collect(e);
    for (int k = 0; k < 4; ++k) { ms[k] = e->prof_ms[k]; launches[k] = e->prof_n[k]; }
    return JG_OK;
}

// 8 independent FMA chains per thread, 4096 FMAs each
__global__ void __launch_bounds__(256) k_fma_peak(float* out, float a, float b, int iters) {
    float x0 = threadIdx.x, x1 = x0 + 1.f, x2 = x0 + 2.f, x3 = x0 + 3.f, x4 = x0 + 4.f, x5 = x0 + 5.f, x6 = x0 + 6.f, x7 = x0 + 7.f;
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int k = 0; k < 16; ++k) {
            x0 = fmaf(x0, a, b); x1 = fmaf(x1, a, b); x2 = fmaf(x2, a, b); x3 = fmaf(x3, a, b);
            x4 = fmaf(x4, a, b); x5 = fmaf(x5, a, b); x6 = fmaf(x6, a, b); x7 = fmaf(x7, a, b);
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = x0 + x1 + x2 + x3 + x4 + x5 + x6 + x7;
}

extern "C" int jg_work_counters(int device, uint64_t out[4]) {
#ifdef JG_COUNT_WORK
    if (!out) return fail(JG_ERR_INVALID, "jg_work_counters: NULL output");
    CUDA_TRY(cudaSetDevice(device));
    CUDA_TRY(cudaDeviceSynchronize());
    unsigned long long h[JG_WORK_N];
    CUDA_TRY(cudaMemcpyFromSymbol(h, g_jg_work, sizeof(h)));
    for (int i = 0; i < 4; ++i) out[i] = h[i];
    return JG_OK;
#else
    (void)device; (void)out;
    return fail(JG_ERR_UNSUPPORTED, "jg_work_counters: this library was built without -DJG_COUNT_WORK");
#endif
}

// JG_TRACE build only: the per-warp timeline of the last two penetration-kernel launches: out[2][4096][4]
extern "C" int jg_trace_read(unsigned long long* out) {
#ifdef JG_TRACE
    CUDA_TRY(cudaDeviceSynchronize());
    CUDA_TRY(cudaMemcpyFromSymbol(out, g_jg_trace, sizeof(unsigned long long) * 2 * 4096 * 8));
    return JG_OK;
#else
    (void)out;
    return fail(JG_ERR_UNSUPPORTED, "jg_trace_read: this library was built without -DJG_TRACE");
#endif
}

extern "C" int jg_work_counters_reset(int device) {
#ifdef JG_COUNT_WORK
    CUDA_TRY(cudaSetDevice(device));
    CUDA_TRY(cudaDeviceSynchronize());
    unsigned long long h[JG_WORK_N] = {0ull, 0ull, 0ull, 0ull};
    CUDA_TRY(cudaMemcpyToSymbol(g_jg_work, h, sizeof(h)));
    return JG_OK;
#else
    (void)device;
    return fail(JG_ERR_UNSUPPORTED, "jg_work_counters_reset: this library was built without -DJG_COUNT_WORK");
#endif
}

extern "C" int jg_measure_fp32_peak(int device, int reps, double* tflops) {
    if (!tflops) return fail(JG_ERR_INVALID, "jg_measure_fp32_peak: NULL output");
    CUDA_TRY(cudaSetDevice(device));
    cudaDeviceProp prop;
    CUDA_TRY(cudaGetDeviceProperties(&prop, device));
    const int blocks = prop.multiProcessorCount * 8, threads = 256, iters = 2048;
    float* out = nullptr;
    CUDA_TRY(cudaMalloc(&out, sizeof(float) * blocks * threads));
    cudaEvent_t a, b;
    CUDA_TRY(cudaEventCreate(&a));
    CUDA_TRY(cudaEventCreate(&b));
    double best = 0;
    for (int r = 0; r < std::max(reps, 1) + 1; ++r) {
        CUDA_TRY(cudaEventRecord(a));
        k_fma_peak<<<blocks, threads>>>(out, 0.999f, 0.001f, iters);
        CUDA_TRY(cudaEventRecord(b));
        CUDA_TRY(cudaEventSynchronize(b));
        float ms = 0;
        CUDA_TRY(cudaEventElapsedTime(&ms, a, b));
        const double fl = 2.0 * 8 * 16 * (double)iters * blocks * threads;
        if (r > 0) best = std::max(best, fl / (ms * 1e-3) / 1e12);
    }
    cudaEventDestroy(a); cudaEventDestroy(b); cudaFree(out);
    *tflops = best;
    return JG_OK;
}

extern "C" int jg_ik(jg_engine* e, const float* targets, int64_t n, int ee_link, const float* q_init, const double* q_rest,
                     int iters, float tol, float damping, float null_gain, int position_only, float* q_out, float* err_out,
                     void* stream) {
    if (!e) return fail(JG_ERR_INVALID, "jg_ik: engine is NULL");
    if (n < 0 || n > 2000000000LL || (n && (!targets || !q_out || !err_out)) || !q_rest)
        return fail(JG_ERR_INVALID, "jg_ik: NULL argument");
    if (ee_link < 0 || ee_link >= e->L) return fail(JG_ERR_INVALID, "jg_ik: end-effector link %d out of range", ee_link);
    if (iters < 0 || !(tol > 0.f) || !(damping > 0.f)) return fail(JG_ERR_INVALID, "jg_ik: bad iteration parameters");
    cudaStream_t st = (cudaStream_t)stream;
    CUDA_TRY(cudaSetDevice(e->device));
    float rest[MAX_DOF] = {0};
    for (int j = 0; j < e->dof; ++j) rest[j] = (float)q_rest[j];
    CUDA_TRY(cudaMemcpyAsync(e->d_qrest, rest, sizeof rest, cudaMemcpyHostToDevice, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    if (n == 0) return JG_OK;
    IkArgs a;
    a.robot = e->d_robot; a.targets = targets; a.q_init = q_init; a.q_rest = e->d_qrest; a.n = (int)n; a.dof = e->dof;
    a.ee_link = ee_link; a.iters = iters; a.pos_only = position_only ? 1 : 0; a.tol = tol; a.damping = damping; a.null_gain = null_gain;
    a.finger_open = e->prm.finger_open; a.q_out = q_out; a.err_out = err_out;
    k_ik<128><<<(int)((n + 127) / 128), 128, sizeof(RobotTab), st>>>(a);
    CUDA_TRY(cudaGetLastError());
    return JG_OK;
}
