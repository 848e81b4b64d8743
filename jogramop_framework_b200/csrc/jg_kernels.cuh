// jg_kernels.cuh -- device tables and the kernels of one collision-check pass.
//
//   k_broadphase       one thread per configuration: forward kinematics down the URDF chain (link transform in
//                      registers), world box of every link shape vs obstacle-piece boxes / plane / self pairs;
//                      survivors go to per-(shape,piece) queues in HBM: all keys of a shape are tested first, then
//                      lanes 0..P reserve the slots of their keys side by side (one atomic round trip per shape).
//   k_select           builds the per-key index lists of a narrow-phase stage (champions / pairs that can still matter).
//   k_narrowphase      GJK.  A warp works on ONE key at a time, so every lane runs on the same two vertex sets
//                      (shared-memory broadcast reads), each on its own configuration; entries come from the key's
//                      cursor through refill_loop (lane refill, prefetched reservations).  Result: atomicMin of the
//                      contact distance into the configuration's slot; overlapping cores go to the penetration queue.
//   k_penetration      expanding-polytope depth, one compact polytope per thread in shared memory, same scheduling.
//   k_penetration_big  the few polytopes that outgrow the compact storage: one per warp, continued from a hand-over
//                      record.
//   k_finalize         per configuration: clamp at the query distance, threshold, ballot-pack the verdict bits.
//
// HBM layout of a pass over C configurations and K keys (key = shape*(P+1)+piece, plane = piece P, then the
// self-collision pairs): see PassArgs below and DESIGN.md section 4.
#pragma once
#include "jg_epa.cuh"
#include "jg_lbvh.cuh"

// JG_DEBUG_BOUNDS build (libjogramop_b200_bounds.so): every index into a queue, a stage list or the overflow records is
// checked on the device before it is used; a violation prints what and where and traps (the launch fails, the test
// fails).  compute-sanitizer is not available on the GPU pool, so the GPU suite is run once per round against this
// build instead (profiles/r02_bounds_run.txt).  The shipped library compiles the checks away.
#if defined(JG_DEBUG_BOUNDS) && !defined(JG_HOST_EMU)
#define JG_BOUNDS(cond, what)                                                                                       \
    do {                                                                                                            \
        if (!(cond)) {                                                                                              \
            printf("JG_DEBUG_BOUNDS: %s violated at %s:%d (block %d thread %d)\n", what, __FILE__, __LINE__, (int)blockIdx.x, \
                   (int)threadIdx.x);                                                                                \
            __trap();                                                                                               \
        }                                                                                                           \
    } while (0)
#else
#define JG_BOUNDS(cond, what) ((void)0)
#endif

// Programmatic dependent launch (sm_90+): the kernels of a pass are launched with the programmatic-stream-serialization
// attribute (jg_abi.cu launch_dep), so a kernel may START while its predecessor is still draining; it loads what does not
// depend on the predecessor (vertex table, key table -> shared memory), then waits here for the predecessor's results.
// Persistent kernels (one resident wave) also announce early that their successor may be scheduled: its blocks then take
// over each SM as soon as this kernel's block on it has exited -- the successor's prologue hides under this kernel's tail.
// Without the launch attribute both calls are no-ops.
#if defined(__CUDA_ARCH__) && !defined(JG_HOST_EMU)
#define JG_WAIT_FOR_PREDECESSOR() cudaGridDependencySynchronize()
#define JG_SUCCESSOR_MAY_START() cudaTriggerProgrammaticLaunchCompletion()
#else
#define JG_WAIT_FOR_PREDECESSOR() ((void)0)
#define JG_SUCCESSOR_MAY_START() ((void)0)
#endif

namespace jg {

// JG_TRACE build (experiments only, never shipped): per-warp timeline of the penetration kernel -- when each warp started and
// ended, how many keys it visited and how many polytopes it started -- to see where the SMs idle at the end of a stage.
#ifdef JG_TRACE
__device__ unsigned long long g_jg_trace[2][4096][8];
__device__ int g_jg_trace_slot;
__device__ __forceinline__ unsigned long long jg_now() { unsigned long long t; asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t)); return t; }
#endif

#ifndef JG_BP_OCC
#define JG_BP_OCC 7   /* resident blocks per SM the broad phase is compiled for (register budget) */
#endif
constexpr int MAX_LINKS = 24;
constexpr int MAX_SHAPES = 24;
constexpr int MAX_DOF = 16;

struct JointRec {
    float o[12];      // joint origin (parent link frame -> joint frame), row-major 3x4
    float axis[3];
    int type;         // JG_JOINT_*
    int qidx;         // >=0 slot of q, -1 fixed, -2 finger
    int parent_sel;   // 0 previous link, 1 saved link, 2 base
    int save;         // keep this link's transform for later children
    int shape_first, shape_count;
    int self_first, self_count;   // self-collision pairs whose SECOND link is this one (sorted by second link)
    int stash_slot;               // >= 0: this link is the first link of some pair: keep its frame in shared memory
    int fast;                     // bit 0: the origin is the identity; bit 1: revolute about +z of the joint frame
};

struct ShapeRec {
    float c[3], h[3];  // local AABB of the core vertices
    float margin, plane_margin;
    int voff, vn;      // core vertices
    int poff, pn;      // plane / EPA vertices
    int link;
    int is_box;        // the core vertices are exactly the 8 corners of the local box (c, h): exact box tests may be used
    int pad[2];
};

struct RobotTab {
    int L, S, dof, base_shape_first, base_shape_count, n_self, n_stash, pad1;
    float qlo[MAX_DOF], qhi[MAX_DOF];
    JointRec joints[MAX_LINKS];
    ShapeRec shapes[MAX_SHAPES];
};

static_assert(sizeof(RobotTab) % 16 == 0, "the piece table follows the robot table in shared memory and is read in 16-byte words");

struct __align__(16) PieceRec {
    float c[3], h[3];  // AABB centre / half extents (robot frame); vertices are stored relative to c
    float margin;
    int voff, vn;
    int pad[3];
};
static_assert(sizeof(PieceRec) == 48, "PieceRec is read as three 16-byte words");

struct SelfRec {
    int sa, sb;        // shape ids (sa = shape of the first link: the static "B" side of the GJK)
    int la, lb;        // link ids
};

enum { KEY_HULL = 0, KEY_PLANE = 1, KEY_SELF = 2, KEY_MESH = 3 };
struct KeyRec {
    int a_off, a_n;    // moving vertex set (robot shape), GJK core
    int b_off, b_n;    // static vertex set (piece / first link shape), GJK core
    float msum;        // margins subtracted from the core distance
    int kind;
    int ea_off, ea_n;  // penetration-depth geometry (box: exact corners instead of the shrunk core)
    int eb_off, eb_n;
    float emsum;       // margins added to the penetration depth
    int pad;
    float a_c[3], a_h[3];   // AABB (own frame) of the moving penetration geometry
    float b_c[3], b_h[3];   // AABB (own frame) of the static penetration geometry
};

struct PassArgs {
    // inputs
    const float* q;    // [n,dof] (or qa for edges)
    const float* qb;   // edges: end configurations, else NULL
    int n;             // configurations in this pass (substep-expanded for edges)
    int cfg_base;      // first configuration of THIS launch (the broad phase of a pass may be split into several
                       // launches so that the host path can overlap it with the upload of the next slice)
    int dof;
    int substeps;      // 1, or substeps per edge
    int stage_set;     // which of the two stage-list sets (cnt_act / cnt_cur) this launch works on
    uint32_t flags;
    float finger_open, qd, threshold;
    float plane[4];
    int has_plane;
    // tables
    const RobotTab* robot;
    const PieceRec* pieces;
    int P;
    const SelfRec* selfs;
    const KeyRec* keys;
    int n_keys, n_scene_keys;
    const float4* verts;
    int n_verts;
    // queues
    float4 *q_r0, *q_r1, *q_r2;
    int* q_cfg;
    int cap;           // queue capacity per key (= max configs per pass)
    int* counts;       // 5*n_keys + 8 ints (JG_COUNTS_INTS): pair-queue counts | penetration-queue counts | entries selected for the
                       // current stage | penetration entries already expanded | gjk range ctr, penetration range ctr,
                       // overflow-list length, select range ctr
    float* q_prio;     // [n_keys*cap] upper bound of the pair's total penetration (AABB overlap + margins)
    int* epa_slot;     // [n_keys*cap] queue slots whose cores overlap (penetration-depth pass)
    uint4* epa_ids;    // [n_keys*cap] their GJK end simplices (vertex ids, 0xFFFFFFFF = unused)
    int2* ovf_list;    // [ovf_cap] (key, index in the penetration queue) of polytopes that outgrew shared memory
    int ovf_cap;       // = 4 * cap: a pass may hand more polytopes to the overflow kernel than it has configurations
    uint32_t* ovf_poly;  // [ovf_poly_cap][JG_OVF_WORDS] the compact polytope of the first ovf_poly_cap of them: the overflow
    int ovf_poly_cap;    // kernel goes on from there instead of repeating the 14 expansions that filled it
    float* epa_prio;   // [n_keys*cap] the same bound for the entries of the penetration queue
    int* act_list;     // [n_keys*cap] per key: queue indices selected for the current stage (k_select)
    int2* act_meta;    // [n_keys*cap] beside act_list (hull / triangle keys): (pair-queue slot, configuration) of the entry,
                       // so that a refilling lane does not chase list -> slot -> configuration through three loads
    unsigned long long *champ_scene, *champ_self;   // [cap] per configuration: (prio bits << 32 | key << 20 | slot) of the
                                                    // queued pair with the largest upper bound ("champion")
    float max_msum;    // largest margin sum over the keys (JG_NO_EPA pruning)
    unsigned long long* stats;   // [8] device counters (jg_get_stats)
    // mesh mode: obstacle triangle soup behind an LBVH instead of convex pieces
    int mesh;
    const BvhNode* bvh_nodes;
    const float4* bvh_tri;     // [3 * bvh_n] triangle vertices, Morton order
    int bvh_n;
    float mesh_margin;
    int* q_tri;                // [n_keys*cap] triangle slot of a mesh queue entry (aliases epa_slot: no EPA on triangles)
    float4* mesh_T;            // [mesh pass][shapes used][3] rows of every shape's transform: the queries of k_mesh_traverse
    int mesh_n_use;            // shapes per configuration in mesh_T
    int* err_flag;             // set when a mesh queue overflowed
    // per-configuration scratch
    int* enc_scene;    // encoded min contact distance vs scene+plane
    int* enc_self;
    uint32_t* viol;    // limit violation bits
};

__device__ __forceinline__ int* cnt_pair(const PassArgs& a) { return a.counts; }
__device__ __forceinline__ int* cnt_epa(const PassArgs& a) { return a.counts + a.n_keys; }
// Stage lists come in TWO sets (a.stage_set selects the one a kernel works on): k_select builds the lists of the next
// stage in one set and, in passing, clears the set of the stage that has just finished -- no reset kernel in between.
//   counts layout (JG_COUNTS_INTS = 7 K + 8 ints): pair | epa | act[0] | done | misc(8) | cur[0] | act[1] | cur[1]
__device__ __forceinline__ int* cnt_act_of(const PassArgs& a, int set) { return a.counts + (set ? 5 * a.n_keys + 8 : 2 * a.n_keys); }
__device__ __forceinline__ int* cnt_cur_of(const PassArgs& a, int set) { return a.counts + (set ? 6 * a.n_keys + 8 : 4 * a.n_keys + 8); }
__device__ __forceinline__ int* cnt_act(const PassArgs& a) { return cnt_act_of(a, a.stage_set); }
__device__ __forceinline__ int* cnt_done(const PassArgs& a) { return a.counts + 3 * a.n_keys; }
enum { CTR_OVF = 2, CTR_SEL = 3 /* .. CTR_SEL + 3: one range counter per k_select launch of a pass */ };
// hand-over record of an overflowing compact polytope: its words, then lb, ub, it, nv, nf (padded to 124)
#define JG_OVF_WORDS 124
__device__ __forceinline__ int* cnt_misc(const PassArgs& a) { return a.counts + 4 * a.n_keys; }
// per-key cursors of the running narrow-phase stage (entries of the key's stage list handed out so far)
__device__ __forceinline__ int* cnt_cur(const PassArgs& a) { return cnt_cur_of(a, a.stage_set); }

__device__ __forceinline__ Tf joint_local(const JointRec& J, float val) {
    Tf O;
    O.r00 = J.o[0]; O.r01 = J.o[1]; O.r02 = J.o[2];  O.tx = J.o[3];
    O.r10 = J.o[4]; O.r11 = J.o[5]; O.r12 = J.o[6];  O.ty = J.o[7];
    O.r20 = J.o[8]; O.r21 = J.o[9]; O.r22 = J.o[10]; O.tz = J.o[11];
    if (J.type == 1) {  // revolute: Rodrigues about the unit axis
        float s, c;
        sincosf(val, &s, &c);
        const float x = J.axis[0], y = J.axis[1], z = J.axis[2], t = 1.f - c;
        Tf M;
        M.r00 = fmaf(t * x, x, c);      M.r01 = fmaf(t * x, y, -s * z); M.r02 = fmaf(t * x, z, s * y);
        M.r10 = fmaf(t * x, y, s * z);  M.r11 = fmaf(t * y, y, c);      M.r12 = fmaf(t * y, z, -s * x);
        M.r20 = fmaf(t * x, z, -s * y); M.r21 = fmaf(t * y, z, s * x);  M.r22 = fmaf(t * z, z, c);
        M.tx = M.ty = M.tz = 0.f;
        return tf_mul(O, M);
    }
    if (J.type == 2) {  // prismatic
        const float ax = J.axis[0] * val, ay = J.axis[1] * val, az = J.axis[2] * val;
        O.tx = fmaf(O.r00, ax, fmaf(O.r01, ay, fmaf(O.r02, az, O.tx)));
        O.ty = fmaf(O.r10, ax, fmaf(O.r11, ay, fmaf(O.r12, az, O.ty)));
        O.tz = fmaf(O.r20, ax, fmaf(O.r21, ay, fmaf(O.r22, az, O.tz)));
    }
    return O;
}

// parent frame x joint origin x joint motion, specialised for what the chains of this repo are made of: identity origins
// (the platform joints) and revolutes about the joint frame's z axis (every Panda joint): the z rotation only mixes the
// first two columns of (parent x origin).  Anything else takes the generic path.
__device__ __forceinline__ Tf chain_step(const Tf& par, const JointRec& J, float val) {
    if (J.type == 1 && !(J.fast & 2)) return tf_mul(par, joint_local(J, val));
    Tf A = par;
    if (!(J.fast & 1)) {
        Tf O;
        O.r00 = J.o[0]; O.r01 = J.o[1]; O.r02 = J.o[2];  O.tx = J.o[3];
        O.r10 = J.o[4]; O.r11 = J.o[5]; O.r12 = J.o[6];  O.ty = J.o[7];
        O.r20 = J.o[8]; O.r21 = J.o[9]; O.r22 = J.o[10]; O.tz = J.o[11];
        A = tf_mul(par, O);
    }
    if (J.type == 1) {
        float s, c;
        sincosf(val, &s, &c);
        const float a00 = A.r00, a10 = A.r10, a20 = A.r20;
        A.r00 = fmaf(c, a00, s * A.r01);  A.r01 = fmaf(c, A.r01, -s * a00);
        A.r10 = fmaf(c, a10, s * A.r11);  A.r11 = fmaf(c, A.r11, -s * a10);
        A.r20 = fmaf(c, a20, s * A.r21);  A.r21 = fmaf(c, A.r21, -s * a20);
    } else if (J.type == 2) {
        const float ax = J.axis[0] * val, ay = J.axis[1] * val, az = J.axis[2] * val;
        A.tx = fmaf(A.r00, ax, fmaf(A.r01, ay, fmaf(A.r02, az, A.tx)));
        A.ty = fmaf(A.r10, ax, fmaf(A.r11, ay, fmaf(A.r12, az, A.ty)));
        A.tz = fmaf(A.r20, ax, fmaf(A.r21, ay, fmaf(A.r22, az, A.tz)));
    }
    return A;
}

// A configuration's champion = its queued pair with the largest penetration bound.  All pairs of a configuration are
// produced by ONE thread, so the thread keeps the running champion in a register (packed prio bits << 32 | key << 20 |
// slot, compared as an integer) and publishes it once at the end (k_broadphase epilogue): no atomicMax per pair and no
// selection kernel for the first stage.  Edge substeps of one edge are spread over threads: they keep the atomic path.
struct Champ {
    unsigned long long scene, self;
};

// warp-aggregated append of the lanes with `survive` to queue `key`.  prio = upper bound of the pair's total
// penetration.  `local` != nullptr: track the champion in the caller's register; else `champ` != nullptr: atomicMax.
// stage1 = true (plane entries): every entry also goes straight into the key's stage-1 list.
__device__ __forceinline__ void queue_push(const PassArgs& a, bool survive, int key, const Tf& T, int cfg, float prio,
                                           unsigned long long* champ, unsigned long long* local, bool stage1) {
    const unsigned mask = __ballot_sync(0xFFFFFFFFu, survive);
    if (mask == 0) return;
    const int lane = threadIdx.x & 31;
    const int leader = __ffs(mask) - 1;
    int base = 0;
    if (lane == leader) {
        base = atomicAdd(a.counts + key, __popc(mask));
        if (stage1) atomicAdd(cnt_act(a) + key, __popc(mask));   // same slots, same count: the list is the identity
    }
    base = __shfl_sync(0xFFFFFFFFu, base, leader);
    if (survive) {
        const int idx = base + __popc(mask & ((1u << lane) - 1u));
        JG_BOUNDS(idx >= 0 && idx < a.cap && key >= 0 && key < a.n_keys, "pair-queue slot (queue_push)");
        const size_t slot = (size_t)key * a.cap + idx;
        a.q_r0[slot] = make_float4(T.r00, T.r01, T.r02, T.tx);
        a.q_r1[slot] = make_float4(T.r10, T.r11, T.r12, T.ty);
        a.q_r2[slot] = make_float4(T.r20, T.r21, T.r22, T.tz);
        a.q_cfg[slot] = cfg;
        a.q_prio[slot] = prio;
        if (stage1) a.act_list[slot] = idx;
        const unsigned long long code = ((unsigned long long)__float_as_uint(prio) << 32) | (unsigned)((key << 20) | idx);
        if (local) { if (code > *local) *local = code; }
        else if (champ) atomicMax(champ + cfg, code);
    }
}

// publish a thread-local champion: remember it for the later stages and put it on its key's stage-1 list
// (lanes of the warp that share a key reserve together)
__device__ __forceinline__ void champ_publish(const PassArgs& a, unsigned long long code, unsigned long long* champ, int cfg,
                                              bool valid) {
    const bool have = valid && code != 0ull;
    if (valid) champ[cfg] = code;
    const unsigned act = __ballot_sync(0xFFFFFFFFu, have);
    if (!have) return;
    const int key = (int)((unsigned)(code & 0xFFFFFFFFull) >> 20), idx = (int)(code & 0xFFFFFull);
    const unsigned peers = __match_any_sync(act, key);
    const int lane = threadIdx.x & 31, leader = __ffs(peers) - 1;
    int base = 0;
    if (lane == leader) base = atomicAdd(cnt_act(a) + key, __popc(peers));
    base = __shfl_sync(peers, base, leader);
    JG_BOUNDS(key >= 0 && key < a.n_keys && base + __popc(peers) <= a.cap && idx < a.cap, "stage-1 list slot (champ_publish)");
    const size_t li = (size_t)key * a.cap + base + __popc(peers & ((1u << lane) - 1u));
    a.act_list[li] = idx;
    a.act_meta[li] = make_int2(idx, cfg);
}

// AABB of the shape (local centre c / half h) under T against a box (centre bc, half bh): smallest signed overlap over
// the three axes (negative = separated along that axis by that much)
__device__ __forceinline__ float aabb_overlap(const Tf& T, const float* c, const float* h, const float* bc,
                                              const float* bh) {
    const float cx = fmaf(T.r00, c[0], fmaf(T.r01, c[1], fmaf(T.r02, c[2], T.tx)));
    const float cy = fmaf(T.r10, c[0], fmaf(T.r11, c[1], fmaf(T.r12, c[2], T.ty)));
    const float cz = fmaf(T.r20, c[0], fmaf(T.r21, c[1], fmaf(T.r22, c[2], T.tz)));
    const float hx = fmaf(fabsf(T.r00), h[0], fmaf(fabsf(T.r01), h[1], fabsf(T.r02) * h[2]));
    const float hy = fmaf(fabsf(T.r10), h[0], fmaf(fabsf(T.r11), h[1], fabsf(T.r12) * h[2]));
    const float hz = fmaf(fabsf(T.r20), h[0], fmaf(fabsf(T.r21), h[1], fabsf(T.r22) * h[2]));
    return fminf(fminf(hx + bh[0] - fabsf(cx - bc[0]), hy + bh[1] - fabsf(cy - bc[1])), hz + bh[2] - fabsf(cz - bc[2]));
}
// bound of the total penetration from the box overlap: the core boxes may be up to 1 mm (embedded box margin)
// smaller than the penetration geometry on each side
__device__ __forceinline__ float prio_of(float ov, float msum) { return fmaxf(ov, 0.f) + msum + 0.0021f; }

// world-frame bounding box of a shape placed by T: computed once per shape, tested against every piece
struct WorldBox {
    float cx, cy, cz, hx, hy, hz;
};
__device__ __forceinline__ WorldBox world_box(const Tf& T, const ShapeRec& Sh) {
    WorldBox w;
    w.cx = fmaf(T.r00, Sh.c[0], fmaf(T.r01, Sh.c[1], fmaf(T.r02, Sh.c[2], T.tx)));
    w.cy = fmaf(T.r10, Sh.c[0], fmaf(T.r11, Sh.c[1], fmaf(T.r12, Sh.c[2], T.ty)));
    w.cz = fmaf(T.r20, Sh.c[0], fmaf(T.r21, Sh.c[1], fmaf(T.r22, Sh.c[2], T.tz)));
    w.hx = fmaf(fabsf(T.r00), Sh.h[0], fmaf(fabsf(T.r01), Sh.h[1], fabsf(T.r02) * Sh.h[2]));
    w.hy = fmaf(fabsf(T.r10), Sh.h[0], fmaf(fabsf(T.r11), Sh.h[1], fabsf(T.r12) * Sh.h[2]));
    w.hz = fmaf(fabsf(T.r20), Sh.h[0], fmaf(fabsf(T.r21), Sh.h[1], fabsf(T.r22) * Sh.h[2]));
    return w;
}
// (shape, piece) candidate test: world boxes first, then a second look along the shape's own box axes (its local box is
// tight, the world box of a rotated link is not), which culls false candidates and tightens the penetration bound `ov`
// that ranks the pairs
__device__ __forceinline__ bool piece_test(const PassArgs& a, const ShapeRec& Sh, const PieceRec& Pc, const Tf& T,
                                           const WorldBox& w, float& ov) {
    const float msum = Sh.margin + Pc.margin;
    const float dx = w.cx - Pc.c[0], dy = w.cy - Pc.c[1], dz = w.cz - Pc.c[2];
    ov = fminf(fminf(w.hx + Pc.h[0] - fabsf(dx), w.hy + Pc.h[1] - fabsf(dy)), w.hz + Pc.h[2] - fabsf(dz));
    if (!(ov + a.qd + msum >= 0.f)) return false;
    const float o0 = Sh.h[0] + fmaf(fabsf(T.r00), Pc.h[0], fmaf(fabsf(T.r10), Pc.h[1], fabsf(T.r20) * Pc.h[2])) -
                     fabsf(fmaf(T.r00, dx, fmaf(T.r10, dy, T.r20 * dz)));
    const float o1 = Sh.h[1] + fmaf(fabsf(T.r01), Pc.h[0], fmaf(fabsf(T.r11), Pc.h[1], fabsf(T.r21) * Pc.h[2])) -
                     fabsf(fmaf(T.r01, dx, fmaf(T.r11, dy, T.r21 * dz)));
    const float o2 = Sh.h[2] + fmaf(fabsf(T.r02), Pc.h[0], fmaf(fabsf(T.r12), Pc.h[1], fabsf(T.r22) * Pc.h[2])) -
                     fabsf(fmaf(T.r02, dx, fmaf(T.r12, dy, T.r22 * dz)));
    ov = fminf(ov, fminf(o0, fminf(o1, o2)));
    return ov + a.qd + msum >= 0.f;
}
// plane candidate test (conservative): lowest point of the shape's bounding box above the plane
__device__ __forceinline__ bool plane_test(const PassArgs& a, const ShapeRec& Sh, const Tf& T, const WorldBox& w) {
    const V3 nl = tf_rot_t(T, mk(a.plane[0], a.plane[1], a.plane[2]));
    const float ext = fmaf(fabsf(nl.x), Sh.h[0], fmaf(fabsf(nl.y), Sh.h[1], fabsf(nl.z) * Sh.h[2]));
    const float low = fmaf(a.plane[0], w.cx, fmaf(a.plane[1], w.cy, fmaf(a.plane[2], w.cz, a.plane[3]))) - ext;
    // the plane vertices may exceed the core AABB by the embedded box margin: 2 mm slack
    return low - 0.002f - Sh.plane_margin <= a.qd;
}

// scene + plane candidates of one shape placed by T (uniform control flow across the warp)
// MESH is a compile-time switch: the triangle path's candidate lists must not cost the hull path registers
template <bool MESH>
__device__ __forceinline__ void shape_vs_scene(const PassArgs& a, const ShapeRec& Sh, int s, const PieceRec* pieces,
                                               const Tf& T, bool valid, int out_idx, Champ* local, int qslot) {
    if (MESH && (a.flags & 1u)) {
        // mesh mode: the triangle candidates of this shape are found by k_mesh_traverse (a group of lanes per query walks
        // the LBVH together); here the query -- the shape's transform -- is only put down
        if (qslot >= 0) {   // every in-range query is written (a NaN transform included: the traversal skips it)
            float4* q = a.mesh_T + (size_t)qslot * 3;
            q[0] = make_float4(T.r00, T.r01, T.r02, T.tx);
            q[1] = make_float4(T.r10, T.r11, T.r12, T.ty);
            q[2] = make_float4(T.r20, T.r21, T.r22, T.tz);
        }
    }
    // Hull pieces + plane.  The slot reservation is an atomic with return (~1 us round trip under load); issued per
    // (shape, piece) it was a third of this kernel's time.  So: first test ALL keys of the shape (bit k of `hm` = this
    // lane hits key k; lane k keeps the warp's ballot for key k), then lanes 0..P reserve their key's slots side by
    // side -- one round trip per shape -- and only then are the entries written.
    const int lane = threadIdx.x & 31;
    const int key0 = s * (a.P + 1);
    const bool want_plane = (a.flags & 2u) && a.has_plane;
    const int k_lo = (!MESH && (a.flags & 1u)) ? 0 : a.P, k_hi = want_plane ? a.P + 1 : a.P;
    const WorldBox wb = world_box(T, Sh);
    for (int c0 = k_lo; c0 < k_hi; c0 += 32) {   // chunks of 32 keys (one chunk unless a scene has > 31 pieces)
        const int c1 = min(c0 + 32, k_hi);
        // (1) every lane tests ITS shape against all keys of the chunk on its own: the cheap first look (world boxes) in
        // a straight loop without votes, then the second look (shape's own box axes) only on the few survivors.  Bit k of
        // `hm` = this lane hits key c0 + k.
        unsigned hm = 0u;
        int k_a = -1, k_b = -1;      // the lane's first two hits keep their box overlap (the bound that ranks the pair)
        float ov_a = 0.f, ov_b = 0.f;
        if (valid) {
            const int pe = min(c1, a.P);
            const float slack = a.qd + Sh.margin;
            for (int k = c0; k < pe; ++k) {
                // box + margin of the piece in two 16-byte broadcast reads: (c0 c1 c2 h0) (h1 h2 margin .)
                const float4 pa = reinterpret_cast<const float4*>(pieces + k)[0], pb = reinterpret_cast<const float4*>(pieces + k)[1];
                const float ov1 = fminf(fminf(wb.hx + pa.w - fabsf(wb.cx - pa.x), wb.hy + pb.x - fabsf(wb.cy - pa.y)),
                                        wb.hz + pb.y - fabsf(wb.cz - pa.z));
                if (ov1 + slack + pb.z >= 0.f) hm |= 1u << (k - c0);
            }
            for (unsigned h2 = hm; h2 != 0u; h2 &= h2 - 1u) {
                const int b = __ffs(h2) - 1;
                float ov;
                if (!piece_test(a, Sh, pieces[c0 + b], T, wb, ov)) hm &= ~(1u << b);
                else if (k_a < 0) { k_a = c0 + b; ov_a = ov; }
                else if (k_b < 0) { k_b = c0 + b; ov_b = ov; }
            }
            if (c1 > a.P && plane_test(a, Sh, T, wb)) hm |= 1u << (a.P - c0);
        }
        // (2) one vote per key that ANY lane hit (lane k keeps the warp's ballot of key k) ...
        unsigned mymask = 0u;
        for (unsigned un = __reduce_or_sync(0xFFFFFFFFu, hm); un != 0u; un &= un - 1u) {
            const int b = __ffs(un) - 1;
            const unsigned m = __ballot_sync(0xFFFFFFFFu, (hm >> b) & 1u);
            if (lane == b) mymask = m;
        }
        // (3) ... lanes 0..P reserve the slots of their keys side by side: ONE atomic round trip per shape
        int mybase = 0;
        if (mymask != 0u) {
            const int k = c0 + lane;
            mybase = atomicAdd(a.counts + key0 + k, __popc(mymask));
            // plane entries go straight onto the key's stage-1 list: same slots, same count (the list is the identity)
            if (k == a.P && local != nullptr) atomicAdd(cnt_act(a) + key0 + k, __popc(mymask));
        }
        // (4) every lane writes ITS OWN hits (lanes work on different keys side by side; the key's ballot and base come
        // from the lane that reserved it by a shuffle with a per-lane source)
        const int rounds = __reduce_max_sync(0xFFFFFFFFu, __popc(hm));
        unsigned mine = hm;
        for (int it = 0; it < rounds; ++it) {
            const bool have = mine != 0u;
            const int b = have ? __ffs(mine) - 1 : 0;
            mine &= mine - 1u;
            const unsigned m = __shfl_sync(0xFFFFFFFFu, mymask, b);
            const int base = __shfl_sync(0xFFFFFFFFu, mybase, b);
            if (!have) continue;
            const int k = c0 + b;
            const int idx = base + __popc(m & ((1u << lane) - 1u));
            JG_BOUNDS(idx >= 0 && idx < a.cap && key0 + k < a.n_scene_keys && ((m >> lane) & 1u), "pair-queue slot (broad phase)");
            const size_t slot = (size_t)(key0 + k) * a.cap + idx;
            float prio = 0.f, ox = 0.f, oy = 0.f, oz = 0.f;
            if (k < a.P) {
                const PieceRec& Pc = pieces[k];
                float ov = k == k_a ? ov_a : ov_b;
                if (k != k_a && k != k_b) piece_test(a, Sh, Pc, T, wb, ov);   // third and later hits of one shape: rare
                prio = prio_of(ov, Sh.margin + Pc.margin);
                ox = Pc.c[0]; oy = Pc.c[1]; oz = Pc.c[2];   // vertices of the piece are stored relative to its AABB centre
            }
            a.q_r0[slot] = make_float4(T.r00, T.r01, T.r02, T.tx - ox);
            a.q_r1[slot] = make_float4(T.r10, T.r11, T.r12, T.ty - oy);
            a.q_r2[slot] = make_float4(T.r20, T.r21, T.r22, T.tz - oz);
            a.q_cfg[slot] = out_idx;
            a.q_prio[slot] = prio;
            if (k < a.P) {
                const unsigned long long code = ((unsigned long long)__float_as_uint(prio) << 32) | (unsigned)(((key0 + k) << 20) | idx);
                if (local) { if (code > local->scene) local->scene = code; }
                else atomicMax(a.champ_scene + out_idx, code);
            } else if (local != nullptr) {
                a.act_list[slot] = idx;
            }
        }
    }
}

template <int BLOCK, bool MESH>
__global__ void __launch_bounds__(BLOCK, JG_BP_OCC) k_broadphase(const PassArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    RobotTab* R = reinterpret_cast<RobotTab*>(smem_raw);
    PieceRec* pieces = reinterpret_cast<PieceRec*>(smem_raw + sizeof(RobotTab));
    float* qs = reinterpret_cast<float*>(smem_raw + sizeof(RobotTab) + sizeof(PieceRec) * a.P);
    float* stash = qs + BLOCK * a.dof;   // [n_stash][12][BLOCK], only with JG_SELF: frames of the pairs' first links
    const int tid = threadIdx.x;
    {
        const int* src = reinterpret_cast<const int*>(a.robot);
        int* dst = reinterpret_cast<int*>(R);
        for (int i = tid; i < (int)(sizeof(RobotTab) / 4); i += BLOCK) dst[i] = src[i];
        const int* ps = reinterpret_cast<const int*>(a.pieces);
        int* pd = reinterpret_cast<int*>(pieces);
        for (int i = tid; i < (int)(sizeof(PieceRec) / 4) * a.P; i += BLOCK) pd[i] = ps[i];
    }
    const int cfg0 = a.cfg_base + blockIdx.x * BLOCK;
    const int cfg = cfg0 + tid;
    const bool valid = cfg < a.n;
    const int dof = a.dof;
    if (a.substeps <= 1) {   // coalesced load of the block's [BLOCK,dof] tile, transposed into shared memory
        const int total = min(BLOCK, a.n - cfg0) * dof;
        const float* src = a.q + (size_t)cfg0 * dof;
        for (int i = tid; i < total; i += BLOCK) {
            const int c = i / dof, j = i - c * dof;
            qs[j * BLOCK + c] = src[i];
        }
    } else if (valid) {      // edge substep: q = qa + k/(substeps-1) (qb - qa)
        const int e = cfg / a.substeps, k = cfg - e * a.substeps;
        const float t = (float)k / (float)(a.substeps - 1);
        for (int j = 0; j < dof; ++j) {
            const float qa = a.q[(size_t)e * dof + j], qb = a.qb[(size_t)e * dof + j];
            qs[j * BLOCK + tid] = fmaf(t, qb - qa, qa);
        }
    }
    __syncthreads();
    const int out_idx = a.substeps <= 1 ? cfg : cfg / a.substeps;
    bool finite = true;

    if (valid) {
        if (a.substeps <= 1) {
            a.enc_scene[out_idx] = JG_ENC_INF;
            a.enc_self[out_idx] = JG_ENC_INF;
            a.champ_scene[out_idx] = 0ull;
            a.champ_self[out_idx] = 0ull;
        }
        {   // joint limits (JG_LIMITS); a NaN / infinite joint value is rejected whatever the flags say: it has no pose
            const bool limits = (a.flags & 8u) != 0;
            bool bad = false;
            for (int j = 0; j < dof; ++j) {
                const float v = qs[j * BLOCK + tid];
                const bool fin = fabsf(v) <= FLT_MAX;
                finite = finite && fin;
                bad |= !fin || (limits && !(R->qlo[j] <= v && v <= R->qhi[j]));
            }
            if (bad) atomicOr(a.viol + (out_idx >> 5), 1u << (out_idx & 31));
        }
    }
    // ... and takes no part in the geometry (no pairs, distance stays at the query distance)
    const bool geo = valid && finite;
    const bool want_self = (a.flags & 4u) != 0;
    const bool want_scene = (a.flags & 3u) != 0;
    // one thread owns all pairs of its configuration: champion in registers (not for edge substeps / mesh queues)
    Champ champ;
    champ.scene = champ.self = 0ull;
    Champ* lc = (a.substeps <= 1 && !MESH) ? &champ : nullptr;

    // base-link shapes (link -1)
    Tf cur = tf_identity(), saved = tf_identity();
    if (want_scene)
        for (int s = R->base_shape_first; s < R->base_shape_first + R->base_shape_count; ++s)
            shape_vs_scene<MESH>(a, R->shapes[s], s, pieces, cur, geo, out_idx, lc, valid ? (cfg - a.cfg_base) * a.mesh_n_use + s : -1);

    for (int i = 0; i < R->L; ++i) {
        const JointRec& J = R->joints[i];
        float val = 0.f;
        if (J.qidx >= 0) val = qs[J.qidx * BLOCK + tid];
        else if (J.qidx == -2) val = a.finger_open;
        if (J.parent_sel == 1) cur = chain_step(saved, J, val);
        else if (J.parent_sel == 2) cur = joint_local(J, val);
        else cur = chain_step(cur, J, val);
        if (J.save) saved = cur;
        if (want_self) {
            // pairs (la, i): la's frame was stashed when the chain passed it; the pair is tested in la's frame
            for (int k = J.self_first; k < J.self_first + J.self_count; ++k) {
                const SelfRec sr = a.selfs[k];
                const float* st = stash + (size_t)R->joints[sr.la].stash_slot * 12 * BLOCK + tid;
                Tf Ta;
                Ta.r00 = st[0 * BLOCK]; Ta.r01 = st[1 * BLOCK]; Ta.r02 = st[2 * BLOCK];  Ta.tx = st[3 * BLOCK];
                Ta.r10 = st[4 * BLOCK]; Ta.r11 = st[5 * BLOCK]; Ta.r12 = st[6 * BLOCK];  Ta.ty = st[7 * BLOCK];
                Ta.r20 = st[8 * BLOCK]; Ta.r21 = st[9 * BLOCK]; Ta.r22 = st[10 * BLOCK]; Ta.tz = st[11 * BLOCK];
                const Tf rel = tf_inv_mul(Ta, cur);   // shape of link i expressed in the frame of link la
                const ShapeRec& Sa = R->shapes[sr.sa];
                const ShapeRec& Sb = R->shapes[sr.sb];
                const float msum = Sa.margin + Sb.margin;
                const float ov = aabb_overlap(rel, Sb.c, Sb.h, Sa.c, Sa.h);
                const bool hit = geo && ov + a.qd + msum >= 0.f;
                queue_push(a, hit, a.n_scene_keys + k, rel, out_idx, prio_of(ov, msum), a.champ_self, lc ? &lc->self : nullptr, false);
            }
            if (J.stash_slot >= 0) {
                float* st = stash + (size_t)J.stash_slot * 12 * BLOCK + tid;
                st[0 * BLOCK] = cur.r00; st[1 * BLOCK] = cur.r01; st[2 * BLOCK] = cur.r02;  st[3 * BLOCK] = cur.tx;
                st[4 * BLOCK] = cur.r10; st[5 * BLOCK] = cur.r11; st[6 * BLOCK] = cur.r12;  st[7 * BLOCK] = cur.ty;
                st[8 * BLOCK] = cur.r20; st[9 * BLOCK] = cur.r21; st[10 * BLOCK] = cur.r22; st[11 * BLOCK] = cur.tz;
            }
        }
        if (want_scene)
            for (int s = J.shape_first; s < J.shape_first + J.shape_count; ++s)
                shape_vs_scene<MESH>(a, R->shapes[s], s, pieces, cur, geo, out_idx, lc, valid ? (cfg - a.cfg_base) * a.mesh_n_use + s : -1);
    }
    if (lc) {
        if (want_scene) champ_publish(a, champ.scene, a.champ_scene, out_idx, valid);
        if (want_self) champ_publish(a, champ.self, a.champ_self, out_idx, valid);
    }
}

// ------------------------------------------------------------------------------------------- mesh mode: candidates
// One GROUP of G lanes per query (configuration, link shape): the group walks the LBVH together.  Its frontier lives in
// a shared-memory stack; every step the group pops up to G nodes (one per lane: coalesced 32-byte node reads), each lane
// tests its node's box against the query's oriented box and the survivors' children go back on the stack (ballot +
// prefix inside the group) or, if they are leaves, through the triangle tests.  All groups of a warp step in lockstep
// (no per-lane divergent walk, no local-memory stack; a full stack raises the error flag instead of dropping a child).
//   BOOLEAN = false: triangles that survive the separating-axis culls are queued for the hull-vs-triangle GJK stage
//                    (slots reserved once per step and key by the lanes that emit together);
//   BOOLEAN = true : the shape IS its box (ShapeRec::is_box) and only "does it touch a surface" is asked (zero margins,
//                    verdict only): exact box-vs-triangle test at the leaves, first hit ends the query, nothing queued.
template <int G, bool BOOLEAN>
__global__ void __launch_bounds__(256) k_mesh_traverse(const PassArgs a, int n_q, const int* __restrict__ use_shapes) {
    constexpr int STK = 32 * G;
    extern __shared__ int stack_all[];
    const int lane = threadIdx.x & 31, gl = lane & (G - 1), shift = lane & ~(G - 1);
    const unsigned gfull = G == 32 ? 0xFFFFFFFFu : ((1u << G) - 1u);
    const unsigned lower = (1u << gl) - 1u, lane_lt = (1u << lane) - 1u;
    int* stack = stack_all + (threadIdx.x / G) * STK;
    const int q = blockIdx.x * (256 / G) + threadIdx.x / G;
    bool active = q < n_q;
    const int n_use = a.mesh_n_use;
    const int cfg_l = active ? q / n_use : 0, k = active ? q - cfg_l * n_use : 0;
    const int s = use_shapes ? use_shapes[k] : k;
    const int out_idx = a.substeps <= 1 ? cfg_l : cfg_l / a.substeps;
    const ShapeRec Sh = a.robot->shapes[s];
    const int key = s * (a.P + 1);
    const float msum = Sh.margin + a.mesh_margin;
    if (active && a.substeps <= 1 && cfg_l >= a.n) active = false;
    if (active && a.substeps > 1 && cfg_l >= a.n) active = false;
    // this configuration is already known to touch (another shape of it, or an earlier query): nothing left to learn
    if (BOOLEAN && active && dec_float(a.enc_scene[out_idx]) <= -msum) active = false;
    Tf T = tf_identity();
    if (active) {
        const float4 r0 = a.mesh_T[(size_t)q * 3], r1 = a.mesh_T[(size_t)q * 3 + 1], r2 = a.mesh_T[(size_t)q * 3 + 2];
        T.r00 = r0.x; T.r01 = r0.y; T.r02 = r0.z; T.tx = r0.w;
        T.r10 = r1.x; T.r11 = r1.y; T.r12 = r1.z; T.ty = r1.w;
        T.r20 = r2.x; T.r21 = r2.y; T.r22 = r2.z; T.tz = r2.w;
        // a NaN / infinite configuration has no geometry (its verdict bit is set by the broad phase): it must not walk
        // the tree -- every box test is "not separated" for NaN
        const float chk = fabsf(T.r00) + fabsf(T.r01) + fabsf(T.r02) + fabsf(T.r10) + fabsf(T.r11) + fabsf(T.r12) + fabsf(T.r20) +
                          fabsf(T.r21) + fabsf(T.r22) + fabsf(T.tx) + fabsf(T.ty) + fabsf(T.tz);
        if (!(chk <= FLT_MAX)) active = false;
    }
    const float infl = a.qd + msum + (BOOLEAN ? 0.f : 1e-5f);
    ObbQuery Q;
    Q.c = tf_apply(T, mk(Sh.c[0], Sh.c[1], Sh.c[2]));
    Q.a0 = mk(T.r00, T.r10, T.r20); Q.a1 = mk(T.r01, T.r11, T.r21); Q.a2 = mk(T.r02, T.r12, T.r22);
    Q.h = mk(Sh.h[0] + infl, Sh.h[1] + infl, Sh.h[2] + infl);
    Q.e = mk(fmaf(fabsf(T.r00), Sh.h[0], fmaf(fabsf(T.r01), Sh.h[1], fabsf(T.r02) * Sh.h[2])) + infl,
             fmaf(fabsf(T.r10), Sh.h[0], fmaf(fabsf(T.r11), Sh.h[1], fabsf(T.r12) * Sh.h[2])) + infl,
             fmaf(fabsf(T.r20), Sh.h[0], fmaf(fabsf(T.r21), Sh.h[1], fabsf(T.r22) * Sh.h[2])) + infl);
    int sp = 0;
    int single = -1;      // a one-triangle mesh has no internal node
    if (active) {
        if (a.bvh_n == 1) single = 0;
        else if (a.bvh_n > 1) { if (gl == 0) stack[0] = 0; sp = 1; }
    }
    bool hit = false;
    unsigned long long best = 0ull;
    bool first = true;
    while (__any_sync(0xFFFFFFFFu, sp > 0 || (first && single >= 0))) {
        // ---- pop up to G nodes, test them, put their internal children back
        const int take = min(sp, G);
        __syncwarp();
        int node = -1;
        if (gl < take) node = stack[sp - 1 - gl];
        sp -= take;
        __syncwarp();
        int chL = 0, chR = 0;
        bool pass = false;
        if (node >= 0) {
            const float4* np = reinterpret_cast<const float4*>(a.bvh_nodes + node);
            const float4 n0 = np[0], n1 = np[1];     // lo.xyz hi.x | hi.yz left right
            const float lo[3] = {n0.x, n0.y, n0.z}, hi[3] = {n0.w, n1.x, n1.y};
            pass = obb_vs_aabb(Q, lo, hi);
            chL = __float_as_int(n1.z);
            chR = __float_as_int(n1.w);
        }
        const bool pL = pass && chL >= 0, pR = pass && chR >= 0;
        const unsigned bL = (__ballot_sync(0xFFFFFFFFu, pL) >> shift) & gfull, bR = (__ballot_sync(0xFFFFFFFFu, pR) >> shift) & gfull;
        const int nL = __popc(bL), nR = __popc(bR);
        if (sp + nL + nR > STK) {        // group-uniform: never drop a child silently
            if (gl == 0) atomicOr(a.err_flag, 2);
            sp = 0;
            active = false;
        } else {
            if (pL) stack[sp + __popc(bL & lower)] = chL;
            if (pR) stack[sp + nL + __popc(bR & lower)] = chR;
            sp += nL + nR;
        }
        // ---- leaf children: up to two triangles per lane
#pragma unroll
        for (int c = 0; c < 2; ++c) {
            int slot = -1;
            if (c == 0) { if (pass && chL < 0) slot = ~chL; else if (first && single >= 0 && gl == 0) slot = single; }
            else if (pass && chR < 0) slot = ~chR;
            if (!active) slot = -1;
            bool cand = false;
            float cov = 0.f;
            if (slot >= 0) {
                const float4 v0 = a.bvh_tri[3 * slot], v1 = a.bvh_tri[3 * slot + 1], v2 = a.bvh_tri[3 * slot + 2];
                const V3 p0 = mk(v0.x, v0.y, v0.z), p1 = mk(v1.x, v1.y, v1.z), p2 = mk(v2.x, v2.y, v2.z);
                if (BOOLEAN) {
                    JG_WORK(JG_WORK_TRI_TESTS, 1u);
                    cand = box_tri_overlap(Q, p0, p1, p2);
                } else {
                    // the triangle's own box against the query's world box (also the bound that ranks the candidate),
                    // then the triangle's plane and the three axes of the shape's box
                    const float tlo[3] = {fminf(p0.x, fminf(p1.x, p2.x)), fminf(p0.y, fminf(p1.y, p2.y)), fminf(p0.z, fminf(p1.z, p2.z))};
                    const float thi[3] = {fmaxf(p0.x, fmaxf(p1.x, p2.x)), fmaxf(p0.y, fmaxf(p1.y, p2.y)), fmaxf(p0.z, fmaxf(p1.z, p2.z))};
                    cov = fminf(fminf(fminf(Q.c.x + Q.e.x, thi[0]) - fmaxf(Q.c.x - Q.e.x, tlo[0]),
                                      fminf(Q.c.y + Q.e.y, thi[1]) - fmaxf(Q.c.y - Q.e.y, tlo[1])),
                                fminf(Q.c.z + Q.e.z, thi[2]) - fmaxf(Q.c.z - Q.e.z, tlo[2]));
                    cand = cov >= 0.f;
                    if (cand) {
                        const V3 nt = cross(p1 - p0, p2 - p0);
                        const float l2 = norm2(nt);
                        if (l2 > 1e-24f) {
                            const float r = fabsf(dot(nt, Q.a0)) * Sh.h[0] + fabsf(dot(nt, Q.a1)) * Sh.h[1] + fabsf(dot(nt, Q.a2)) * Sh.h[2];
                            if (fabsf(dot(nt, Q.c - p0)) > r + infl * sqrtf(l2)) cand = false;
                        }
                    }
                    if (cand) {
                        const V3 d0 = p0 - Q.c, d1 = p1 - Q.c, d2 = p2 - Q.c;
                        const float t0 = dot(Q.a0, d0), t1 = dot(Q.a0, d1), t2 = dot(Q.a0, d2);
                        if (fminf(t0, fminf(t1, t2)) > Q.h.x || fmaxf(t0, fmaxf(t1, t2)) < -Q.h.x) cand = false;
                        const float u0 = dot(Q.a1, d0), u1 = dot(Q.a1, d1), u2 = dot(Q.a1, d2);
                        if (fminf(u0, fminf(u1, u2)) > Q.h.y || fmaxf(u0, fmaxf(u1, u2)) < -Q.h.y) cand = false;
                        const float w0 = dot(Q.a2, d0), w1 = dot(Q.a2, d1), w2 = dot(Q.a2, d2);
                        if (fminf(w0, fminf(w1, w2)) > Q.h.z || fmaxf(w0, fmaxf(w1, w2)) < -Q.h.z) cand = false;
                    }
                }
            }
            if (BOOLEAN) {
                const unsigned hb = (__ballot_sync(0xFFFFFFFFu, cand) >> shift) & gfull;
                if (hb) { hit = true; sp = 0; active = false; }       // first hit ends the query (group-uniform)
            } else {
                const unsigned cm = __ballot_sync(0xFFFFFFFFu, cand);
                if (cand) {     // lanes of one key reserve their slots together
                    const unsigned peers = __match_any_sync(cm, key);
                    const int leader = __ffs(peers) - 1;
                    int base = 0;
                    if (lane == leader) base = atomicAdd(a.counts + key, __popc(peers));
                    base = __shfl_sync(peers, base, leader);
                    const int idx = base + __popc(peers & lane_lt);
                    if (idx < a.cap) {
                        const size_t qs = (size_t)key * a.cap + idx;
                        a.q_r0[qs] = make_float4(T.r00, T.r01, T.r02, T.tx);
                        a.q_r1[qs] = make_float4(T.r10, T.r11, T.r12, T.ty);
                        a.q_r2[qs] = make_float4(T.r20, T.r21, T.r22, T.tz);
                        a.q_cfg[qs] = out_idx;
                        a.q_tri[qs] = slot;
                        // a triangle pair never reports less than -(margins): the bound only ranks the candidates
                        const float prio = prio_of(cov - infl, msum);
                        a.q_prio[qs] = prio;
                        const unsigned long long code = ((unsigned long long)__float_as_uint(prio) << 32) | (unsigned)((key << 20) | idx);
                        if (code > best) best = code;
                    } else {
                        atomicOr(a.err_flag, 1);     // queue full: the host redoes the pass with fewer configurations
                    }
                }
            }
        }
        first = false;
    }
    if (BOOLEAN) {
        if (hit && gl == 0) atomicMin(a.enc_scene + out_idx, enc_float(-msum));
    } else {
        // the group's best candidate becomes (a candidate for) the configuration's champion
#pragma unroll
        for (int o = G / 2; o; o >>= 1) {
            const unsigned long long other = __shfl_xor_sync(0xFFFFFFFFu, best, o);
            if (other > best) best = other;
        }
        if (gl == 0 && best) atomicMax(a.champ_scene + out_idx, best);
    }
}

// ------------------------------------------------------------------------------------------------ narrow phase
// Work distribution shared by the GJK and the penetration kernels: PER-KEY CURSORS with LANE REFILL.
// Every lane of a warp must work on the same key (same two vertex sets -> broadcast reads), pairs differ wildly in
// cost, and a stage has only a handful of pairs per lane -- so the unit of scheduling is the single entry:
//   * a warp starts on the key at its proportional position in the stage's work and stays on it; it reserves 32
//     entries at a time from the key's cursor (atomicAdd) and hands them to lanes as they go idle, so the warp-uniform
//     part of an iteration (the vertex scans) runs with all lanes busy;
//   * the next reservation is issued the moment the current one has been handed out and is only looked at when it is
//     needed -- at least one iteration later -- so the round trip of the atomic is never waited for;
//   * a key whose cursor passed its count is exhausted: the lanes drain (the only tail, once per key and warp) and
//     the warp moves to the next key, in circular order, that still has entries.
// sched_prefix: exclusive scan of the per-key entry counts into shared memory (warp 0 of the block).
__device__ __forceinline__ void sched_prefix(const int* __restrict__ counts, int n_keys, int* prefix, int lane) {
    int carry = 0;
    for (int base = 0; base < n_keys; base += 32) {
        const int k = base + lane;
        const int t = k < n_keys ? counts[k] : 0;
        int incl = t;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int up = __shfl_up_sync(0xFFFFFFFFu, incl, o);
            if (lane >= o) incl += up;
        }
        if (k < n_keys) prefix[k] = carry + incl - t;
        carry += __shfl_sync(0xFFFFFFFFu, incl, 31);
    }
    if (lane == 0) prefix[n_keys] = carry;
}

// range scheduler of the selection kernel (uniform cost per entry): a key's span is cut into ranges of G entries that
// warps pull from one global counter.  called by warp 0 of the block; counts = per-key entry counts
__device__ __forceinline__ void sched_build(const int* __restrict__ counts, int n_keys, int n_warps_total, int* prefix,
                                            int* g_out, int lane) {
    int tot = 0;
    for (int k = lane; k < n_keys; k += 32) tot += counts[k];
#pragma unroll
    for (int o = 16; o; o >>= 1) tot += __shfl_xor_sync(0xFFFFFFFFu, tot, o);
    int G = (tot / (n_warps_total * 4) + 31) & ~31;
    G = max(64, min(G, 2048));
    int carry = 0;
    for (int base = 0; base < n_keys; base += 32) {
        const int k = base + lane;
        const int t = k < n_keys ? (counts[k] + G - 1) / G : 0;
        int incl = t;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int up = __shfl_up_sync(0xFFFFFFFFu, incl, o);
            if (lane >= o) incl += up;
        }
        if (k < n_keys) prefix[k] = carry + incl - t;
        carry += __shfl_sync(0xFFFFFFFFu, incl, 31);
    }
    if (lane == 0) { prefix[n_keys] = carry; *g_out = G; }
}

__device__ __forceinline__ int sched_find_key(const int* prefix, int n_keys, int r) {
    int lo = 0, hi = n_keys - 1;
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (prefix[mid] <= r) lo = mid; else hi = mid - 1;
    }
    return lo;
}

__device__ __forceinline__ int cursor_grab(int* cur, int n, int lane) {
    int b = 0;
    if (lane == 0) b = atomicAdd(cur, n);
    return __shfl_sync(0xFFFFFFFFu, b, 0);
}

// first key at or after `from` (circular) whose cursor is still below its count, or -1: lanes probe 32 keys at a time
__device__ __forceinline__ int sched_next_key(const int* __restrict__ counts, const int* cur, int n_keys, int from, int lane) {
    for (int j = 0; j < n_keys; j += 32) {
        int k = from + j + lane;
        if (k >= n_keys) k -= n_keys;
        const bool has = (j + lane < n_keys) && (*(const volatile int*)(cur + k) < counts[k]);
        const unsigned m = __ballot_sync(0xFFFFFFFFu, has);
        if (m) {
            int kk = from + j + (__ffs(m) - 1);
            if (kk >= n_keys) kk -= n_keys;
            return kk;
        }
    }
    return -1;
}

// The refill loop of one key.  idle() / start(list entry, pair-queue slot, configuration) act on the calling lane; step()
// runs one iteration and is called by ALL lanes (converged) while any lane is busy.  Three reservations are in the
// pipeline: `cur` is being handed out (its list entries + metadata sit in registers, lane l <-> entry cbase + l, and
// reach the refilling lane by shuffle), `nxt` has its metadata loads in flight, `pf` is the atomic in flight.
template <class Idle, class Start, class Step>
__device__ __forceinline__ void refill_loop(int* cur, int cnt, int lane, const int* __restrict__ list,
                                            const int2* __restrict__ meta, Idle idle, Start start, Step step) {
    const unsigned lt_mask = (1u << lane) - 1u;
    int next = 0, end = 0, cbase = 0, cj = 0;
    int2 cm = make_int2(0, 0);
    int nbase = -1, nj = 0;
    int2 nm = make_int2(0, 0);
    int pf = cursor_grab(cur, 32, lane);
    bool pf_pending = true, exhausted = false;
    for (;;) {
        const unsigned need = __ballot_sync(0xFFFFFFFFu, idle());
        if (need && next >= end) {
            if (nbase >= 0) {   // promote the loaded interval
                cbase = next = nbase; end = min(nbase + 32, cnt); cj = nj; cm = nm; nbase = -1;
            }
            if (pf_pending) {   // its atomic was issued a whole interval ago
                pf_pending = false;
                if (pf < cnt) {
                    nbase = pf;
                    if (pf + lane < cnt) { nj = list[pf + lane]; nm = meta[pf + lane]; }
                } else exhausted = true;
            }
            if (!exhausted) {
                pf = cursor_grab(cur, 32, lane);
                pf_pending = true;
            }
            if (next >= end && nbase >= 0) {   // start-up / starved: use the interval just loaded (waits for its loads)
                cbase = next = nbase; end = min(nbase + 32, cnt); cj = nj; cm = nm; nbase = -1;
            }
        }
        const int avail = end - next;
        if (need && avail > 0) {
            const int rank = __popc(need & lt_mask);
            const int src = (next - cbase + rank) & 31;
            const int j = __shfl_sync(0xFFFFFFFFu, cj, src);
            const int slot = __shfl_sync(0xFFFFFFFFu, cm.x, src), cfg = __shfl_sync(0xFFFFFFFFu, cm.y, src);
            if (idle() && rank < avail) start(j, slot, cfg);
            next += min(__popc(need), avail);
        }
        if (!__any_sync(0xFFFFFFFFu, !idle())) {
            if (exhausted && next >= end && nbase < 0) break;   // drained
            continue;
        }
        step();
    }
}

// queue entry -> transform (the configuration index comes with the list metadata)
__device__ __forceinline__ Tf load_tf(const PassArgs& a, size_t slot) {
    const float4 r0 = a.q_r0[slot], r1 = a.q_r1[slot], r2 = a.q_r2[slot];
    Tf T;
    T.r00 = r0.x; T.r01 = r0.y; T.r02 = r0.z; T.tx = r0.w;
    T.r10 = r1.x; T.r11 = r1.y; T.r12 = r1.z; T.ty = r1.w;
    T.r20 = r2.x; T.r21 = r2.y; T.r22 = r2.z; T.tz = r2.w;
    return T;
}

__device__ __forceinline__ Tf load_entry(const PassArgs& a, size_t slot, int& cfg) {
    const float4 r0 = a.q_r0[slot], r1 = a.q_r1[slot], r2 = a.q_r2[slot];
    cfg = a.q_cfg[slot];
    Tf T;
    T.r00 = r0.x; T.r01 = r0.y; T.r02 = r0.z; T.tx = r0.w;
    T.r10 = r1.x; T.r11 = r1.y; T.r12 = r1.z; T.ty = r1.w;
    T.r20 = r2.x; T.r21 = r2.y; T.r22 = r2.z; T.tz = r2.w;
    return T;
}

template <int BLOCK>
__global__ void __launch_bounds__(BLOCK) k_narrowphase(const PassArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float4* verts = reinterpret_cast<float4*>(smem_raw);
    KeyRec* keys = reinterpret_cast<KeyRec*>(smem_raw + sizeof(float4) * a.n_verts);
    int* prefix = reinterpret_cast<int*>(keys + a.n_keys);   // [n_keys+1], then G
    const int tid = threadIdx.x, lane = tid & 31;
    JG_SUCCESSOR_MAY_START();
    for (int i = tid; i < a.n_verts; i += BLOCK) verts[i] = a.verts[i];
    {
        const int* ks = reinterpret_cast<const int*>(a.keys);
        int* kd = reinterpret_cast<int*>(keys);
        for (int i = tid; i < (int)(sizeof(KeyRec) / 4) * a.n_keys; i += BLOCK) kd[i] = ks[i];
    }
    JG_WAIT_FOR_PREDECESSOR();         // the tables above do not depend on it; the stage lists and queues do
    const int* acounts = cnt_act(a);   // entries selected for this stage (k_select)
    if (tid < 32) sched_prefix(acounts, a.n_keys, prefix, lane);
    __syncthreads();
    const int total = prefix[a.n_keys];
    if (total == 0) return;
    int* cur = cnt_cur(a);
    const bool want_epa = !(a.flags & 16u);
    const int n_warps = gridDim.x * (BLOCK / 32), gw = blockIdx.x * (BLOCK / 32) + (tid >> 5);
    const unsigned lt_mask = (1u << lane) - 1u;

    // start at the key that holds this warp's proportional share of the stage, then go round the keys
    int key = sched_find_key(prefix, a.n_keys, (int)((long long)gw * total / n_warps));
    for (;;) {
        key = sched_next_key(acounts, cur, a.n_keys, key, lane);
        if (key < 0) break;
        const KeyRec K = keys[key];
        const int cnt = acounts[key];
        const size_t kbase = (size_t)key * a.cap;
        int* enc = K.kind == KEY_SELF ? a.enc_self : a.enc_scene;

        if (K.kind == KEY_PLANE) {   // uniform cost per entry: plain strided loop over chunks of 256
            for (;;) {
                const int begin = cursor_grab(cur + key, 256, lane);
                if (begin >= cnt) break;
                const int end = min(begin + 256, cnt);
                for (int i = begin + lane; i < end; i += 32) {
                    int cfg;
                    const Tf T = load_entry(a, kbase + a.act_list[kbase + i], cfg);
                    const float d = plane_distance(verts + K.a_off, K.a_n, T, mk(a.plane[0], a.plane[1], a.plane[2]),
                                                   a.plane[3]) - K.msum;
                    if (d <= a.qd) atomicMin(enc + cfg, enc_float(d));
                }
            }
            continue;
        }

        const float4* A = verts + K.a_off;
        const float4* B = verts + K.b_off;
        const float dmax = fmaxf(a.qd + K.msum, 0.f);   // cores apart at all => contact distance > -msum
        bool active = false;
        GjkState S;
        Tf T = tf_identity();
        int cfg = 0, off = 0;
        float best_total = -FLT_MAX, dmax_l = dmax;
        if (K.kind == KEY_MESH) {   // link hull vs one triangle per lane (registers); same lane refill, no penetration pass
            TriSupport tri;
            tri.p0 = tri.p1 = tri.p2 = mk(0.f, 0.f, 0.f);
            refill_loop(cur + key, cnt, lane, a.act_list + kbase, a.act_meta + kbase, [&]() { return !active; },
                [&](int, int slot, int cfg_in) {
                    off = slot;
                    cfg = cfg_in;
                    T = load_tf(a, kbase + off);
                    const int t3 = 3 * a.q_tri[kbase + off];
                    const float4 v0 = a.bvh_tri[t3], v1 = a.bvh_tri[t3 + 1], v2 = a.bvh_tri[t3 + 2];
                    // work in a frame centred on the triangle (fp32 precision, like the centred pieces)
                    const V3 c = mk((v0.x + v1.x + v2.x) * (1.f / 3.f), (v0.y + v1.y + v2.y) * (1.f / 3.f),
                                    (v0.z + v1.z + v2.z) * (1.f / 3.f));
                    tri.p0 = mk(v0.x, v0.y, v0.z) - c; tri.p1 = mk(v1.x, v1.y, v1.z) - c; tri.p2 = mk(v2.x, v2.y, v2.z) - c;
                    T.tx -= c.x; T.ty -= c.y; T.tz -= c.z;
                    gjk_init(S, mk(T.tx, T.ty, T.tz));
                    active = true;
                },
                [&]() {
                    if (!active) return;
                    const int st = gjk_step(S, A, K.a_n, tri, T, dmax);
                    if (st != JG_GJK_RUNNING) {
                        active = false;
                        if (st == JG_GJK_SEPARATED) {
                            const float d = sqrtf(S.vv) - K.msum;
                            if (d <= a.qd) atomicMin(enc + cfg, enc_float(d));
                        } else if (st == JG_GJK_PENETRATING) {
                            atomicMin(enc + cfg, enc_float(-K.msum));
                        }
                    }
                });
            continue;
        }
        refill_loop(cur + key, cnt, lane, a.act_list + kbase, a.act_meta + kbase, [&]() { return !active; },
            [&](int, int slot, int c) {
                off = slot;
                cfg = c;
                JG_BOUNDS(slot >= 0 && slot < a.cap && c >= 0 && c < a.cap, "stage list entry (GJK refill)");
                const int e0 = enc[cfg];   // what is already proven for this configuration
                T = load_tf(a, kbase + off);
                gjk_init(S, mk(T.tx, T.ty, T.tz));
                active = true;
                best_total = -dec_float(e0);
                // only a contact distance below the proven one can change the minimum: GJK may stop as soon as a
                // separating axis shows the cores are further apart than that
                dmax_l = fminf(dmax, fmaxf(K.msum - best_total, 0.f));
            },
            [&]() {
                // ---- one GJK iteration on every busy lane
                bool deep = false;
                if (active) {
                    int st = gjk_step(S, A, K.a_n, SetSupport{B, K.b_n}, T, dmax_l);
                    // exact early out: another pair of this configuration already penetrates deeper than this pair possibly
                    // can (hmin = smallest support value seen = upper bound of this pair's depth if it overlaps at all)
                    if (st == JG_GJK_RUNNING && best_total > 0.f && fmaxf(S.hmin, 0.f) + fmaxf(K.emsum, K.msum) + 0.004f <= best_total) {
                        active = false;
                        st = JG_GJK_RUNNING;
                    } else
                    if (st != JG_GJK_RUNNING) {
                        active = false;
                        if (st == JG_GJK_SEPARATED) {
                            const float d = sqrtf(S.vv) - K.msum;
                            if (d <= a.qd) atomicMin(enc + cfg, enc_float(d));
                        } else if (st == JG_GJK_PENETRATING) {
                            if (want_epa) deep = true;
                            else atomicMin(enc + cfg, enc_float(-K.msum));
                        }
                    }
                }
                if (want_epa) {   // cores overlap: hand the pair (and its end simplex) to the penetration-depth pass
                    const unsigned m = __ballot_sync(0xFFFFFFFFu, deep);
                    if (m) {
                        const int leader = __ffs(m) - 1;
                        int base = 0;
                        if (lane == leader) base = atomicAdd(cnt_epa(a) + key, __popc(m));
                        base = __shfl_sync(0xFFFFFFFFu, base, leader);
                        if (deep) {
                            const int jj = base + __popc(m & lt_mask);
                            JG_BOUNDS(jj >= 0 && jj < a.cap && off >= 0 && off < a.cap, "penetration-queue slot");
                            const size_t j = kbase + jj;
                            // bound of the total penetration: box overlap from the broad phase, tightened by the smallest
                            // support value seen during GJK (+4 mm: margins of the penetration geometry vs the cores)
                            a.epa_prio[j] = fminf(a.q_prio[kbase + off], fmaxf(S.hmin, 0.f) + fmaxf(K.emsum, K.msum) + 0.004f);
                            a.epa_slot[j] = off;
                            a.epa_ids[j] = make_uint4(S.S.n > 0 ? S.S.ia : 0xFFFFFFFFu, S.S.n > 1 ? S.S.ib : 0xFFFFFFFFu,
                                                      S.S.n > 2 ? S.S.ic : 0xFFFFFFFFu, S.S.n > 3 ? S.S.id : 0xFFFFFFFFu);
                        }
                    }
                }
            });
    }
}

// ------------------------------------------------------------------------------------------- penetration depth
// Pairs whose cores overlap.  Start simplex = the GJK end simplex handed over by the narrow phase (rebuilt from its
// vertex ids); shapes whose penetration geometry differs from the GJK core (boxes) rerun GJK on that geometry first.
__device__ __forceinline__ V3 md_point(const float4* __restrict__ A, const float4* __restrict__ B, const Tf& T,
                                       uint32_t id) {
    const float4 pa = A[id >> 16], pb = B[id & 0xFFFFu];
    return tf_apply(T, mk(pa.x, pa.y, pa.z)) - mk(pb.x, pb.y, pb.z);
}

__device__ __forceinline__ Simplex simplex_from_ids(const float4* __restrict__ A, const float4* __restrict__ B,
                                                    const Tf& T, uint4 ids) {
    Simplex sx;
    sx.a = sx.b = sx.c = sx.d = mk(0.f, 0.f, 0.f);
    sx.ia = ids.x; sx.ib = ids.y; sx.ic = ids.z; sx.id = ids.w;
    sx.n = 0;
    if (ids.x != 0xFFFFFFFFu) { sx.a = md_point(A, B, T, ids.x); sx.n = 1; }
    if (ids.y != 0xFFFFFFFFu) { sx.b = md_point(A, B, T, ids.y); sx.n = 2; }
    if (ids.z != 0xFFFFFFFFu) { sx.c = md_point(A, B, T, ids.z); sx.n = 3; }
    if (ids.w != 0xFFFFFFFFu) { sx.d = md_point(A, B, T, ids.w); sx.n = 4; }
    return sx;
}

// Stage selection.  Every stage of a pass (GJK on the champions, penetration depth of the champions, GJK on the
// rest, penetration depth of the rest) works on per-key lists of queue indices built here:
//   SEL_GJK_CHAMP  pair queue: the champion of every configuration (largest penetration bound) + all plane entries
//   SEL_GJK_REST   pair queue: every other pair whose bound still exceeds the configuration's proven penetration
//                  (all of them while no penetration is proven: then the minimum distance needs every pair)
//   SEL_EPA_NEW    penetration queue: entries appended since the last penetration stage, same bound test
// Exact: a skipped pair provably cannot change the configuration's minimum.
enum { SEL_GJK_CHAMP = 0, SEL_GJK_REST = 1, SEL_EPA_NEW = 2 };

template <int BLOCK>
__global__ void __launch_bounds__(BLOCK) k_select(const PassArgs a, int mode, int sel_index, int prev_stat_slot, int mark_done) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    int* prefix = reinterpret_cast<int*>(smem_raw);   // [n_keys + 2]
    int* span = prefix + a.n_keys + 2;                // [n_keys] entries to look at per key
    const int tid = threadIdx.x, lane = tid & 31;
    JG_SUCCESSOR_MAY_START();
    JG_WAIT_FOR_PREDECESSOR();
    const bool epa = mode == SEL_EPA_NEW;
    const int* qcount = epa ? cnt_epa(a) : cnt_pair(a);
    const int* qstart = cnt_done(a);
    for (int k = tid; k < a.n_keys; k += BLOCK) span[k] = min(qcount[k], a.cap) - (epa ? qstart[k] : 0);
    __syncthreads();
    if (tid < 32) sched_build(span, a.n_keys, gridDim.x * (BLOCK / 32), prefix, prefix + a.n_keys + 1, lane);
    __syncthreads();
    const int total = prefix[a.n_keys], G = prefix[a.n_keys + 1];
    int* range_counter = cnt_misc(a) + CTR_SEL + sel_index;
    int* acounts = cnt_act(a);
    if (blockIdx.x == 0) {
        // the stage that has just finished used the OTHER set: account for it (stats[prev_stat_slot] += entries it ran),
        // clear it for the stage after this one, and -- after a penetration stage -- remember how far the penetration
        // queue has been expanded.  Nothing else in this launch touches that set.
        int* oact = cnt_act_of(a, a.stage_set ^ 1);
        int* ocur = cnt_cur_of(a, a.stage_set ^ 1);
        int mine = 0;
        for (int k = tid; k < a.n_keys; k += BLOCK) {
            mine += oact[k];
            oact[k] = 0;
            ocur[k] = 0;
            if (mark_done) cnt_done(a)[k] = cnt_epa(a)[k];
        }
        if (prev_stat_slot >= 0) {
#pragma unroll
            for (int o = 16; o; o >>= 1) mine += __shfl_xor_sync(0xFFFFFFFFu, mine, o);
            if (lane == 0 && mine) atomicAdd(a.stats + prev_stat_slot, (unsigned long long)mine);
        }
    }
    const unsigned lt_mask = (1u << lane) - 1u;
    const bool no_epa = (a.flags & 16u) != 0;
    for (;;) {
        int r = 0;
        if (lane == 0) r = atomicAdd(range_counter, 1);
        r = __shfl_sync(0xFFFFFFFFu, r, 0);
        if (r >= total) break;
        const int key = sched_find_key(prefix, a.n_keys, r);
        const int kind = a.keys[key].kind;
        const int start = epa ? qstart[key] : 0;
        const int begin = start + (r - prefix[key]) * G;
        const int end = min(begin + G, min(qcount[key], a.cap));
        const size_t kbase = (size_t)key * a.cap;
        const unsigned long long* champ = kind == KEY_SELF ? a.champ_self : a.champ_scene;
        const int* enc = kind == KEY_SELF ? a.enc_self : a.enc_scene;
        // four entries per lane and step: their dependent load chains (entry -> configuration -> champion / proven
        // distance) are independent of each other, which is what this latency-bound kernel needs
        constexpr int U = 4;
        for (int j0 = begin; j0 < end; j0 += 32 * U) {
            int jj[U], slot[U], cfg[U];
            bool sel[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                jj[u] = j0 + 32 * u + lane;
                slot[u] = jj[u];
                cfg[u] = 0;
                sel[u] = false;
                if (jj[u] < end && kind != KEY_PLANE) {
                    if (kind != KEY_MESH && epa) slot[u] = a.epa_slot[kbase + jj[u]];
                }
            }
#pragma unroll
            for (int u = 0; u < U; ++u)
                if (jj[u] < end && kind != KEY_PLANE) cfg[u] = a.q_cfg[kbase + slot[u]];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                if (jj[u] >= end) continue;
                const int j = jj[u];
                if (kind == KEY_PLANE) {
                    sel[u] = mode == SEL_GJK_CHAMP;
                } else if (kind == KEY_MESH) {
                    // a triangle pair reports at least -(margins): once the configuration is known to be that deep
                    // (by another triangle or by the plane) no other triangle can lower its minimum
                    const bool is_champ = (unsigned)(champ[cfg[u]] & 0xFFFFFFFFull) == (unsigned)((key << 20) | j);
                    if (mode == SEL_GJK_CHAMP) sel[u] = is_champ;
                    else if (mode == SEL_GJK_REST) sel[u] = !is_champ && -dec_float(enc[cfg[u]]) < a.keys[key].msum - 1e-6f;
                } else {
                    if (mode == SEL_GJK_CHAMP) {
                        sel[u] = (unsigned)(champ[cfg[u]] & 0xFFFFFFFFull) == (unsigned)((key << 20) | slot[u]);
                    } else {
                        float best = -dec_float(enc[cfg[u]]);       // proven total penetration (negative: none yet)
                        if (no_epa && best >= a.max_msum - 1e-6f) best = FLT_MAX;   // clamped mode: any overlap settles it
                        sel[u] = (epa ? a.epa_prio[kbase + j] : a.q_prio[kbase + slot[u]]) > best;
                        if (mode == SEL_GJK_REST)
                            sel[u] = sel[u] && (unsigned)(champ[cfg[u]] & 0xFFFFFFFFull) != (unsigned)((key << 20) | slot[u]);
                    }
                }
            }
            unsigned m[U];
            int total_sel = 0;
#pragma unroll
            for (int u = 0; u < U; ++u) {
                m[u] = __ballot_sync(0xFFFFFFFFu, sel[u]);
                total_sel += __popc(m[u]);
            }
            if (total_sel) {
                int base = 0;
                if (lane == 0) base = atomicAdd(acounts + key, total_sel);
                base = __shfl_sync(0xFFFFFFFFu, base, 0);
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    if (sel[u]) {
                        JG_BOUNDS(base + __popc(m[u] & lt_mask) < a.cap && jj[u] < a.cap && slot[u] >= 0 && slot[u] < a.cap &&
                                      cfg[u] >= 0 && cfg[u] < a.cap, "stage list slot (k_select)");
                        const size_t li = kbase + base + __popc(m[u] & lt_mask);
                        a.act_list[li] = jj[u];
                        if (kind != KEY_PLANE) a.act_meta[li] = make_int2(slot[u], cfg[u]);
                    }
                    base += __popc(m[u]);
                }
            }
        }
    }
}

template <int BLOCK>
__global__ void __launch_bounds__(BLOCK) k_penetration(const PassArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float4* verts = reinterpret_cast<float4*>(smem_raw);
    int* prefix = reinterpret_cast<int*>(smem_raw + sizeof(float4) * a.n_verts);   // [n_keys + 2]
    uint32_t* poly_words = reinterpret_cast<uint32_t*>(prefix + ((a.n_keys + 2 + 3) & ~3));
    const int tid = threadIdx.x, lane = tid & 31;
    const int* ecounts = cnt_act(a);   // entries selected for this stage (k_select)
    JG_SUCCESSOR_MAY_START();
    for (int i = tid; i < a.n_verts; i += BLOCK) verts[i] = a.verts[i];
    JG_WAIT_FOR_PREDECESSOR();
    if (tid < 32) sched_prefix(ecounts, a.n_keys, prefix, lane);
    __syncthreads();
    const int total = prefix[a.n_keys];
    if (total == 0) return;
    int* cur = cnt_cur(a);
    int* ovf_counter = cnt_misc(a) + CTR_OVF;
    const int n_warps = gridDim.x * (BLOCK / 32), gw = blockIdx.x * (BLOCK / 32) + (tid >> 5);

    EpaStateC<BLOCK> X;   // compact polytope, interleaved across the block in shared memory (jg_epa.cuh)
    X.poly.w = poly_words + tid;
    X.poly.nv = X.poly.nf = 0;
    X.lb = 0.f; X.ub = FLT_MAX; X.it = 0;

#ifdef JG_TRACE
    const unsigned long long tr_t0 = jg_now();
    int tr_keys = 0, tr_started = 0, tr_steps = 0, tr_steps_at_last = 0, tr_key0 = -1;
    unsigned long long tr_last = 0;
#endif
    int key = sched_find_key(prefix, a.n_keys, (int)((long long)gw * total / n_warps));
    for (;;) {
        key = sched_next_key(ecounts, cur, a.n_keys, key, lane);
        if (key < 0) break;
#ifdef JG_TRACE
        ++tr_keys;
        if (tr_key0 < 0) tr_key0 = key;
#endif
        const KeyRec K = a.keys[key];
        const size_t kbase = (size_t)key * a.cap;
        int* enc = K.kind == KEY_SELF ? a.enc_self : a.enc_scene;
        const float4* A = verts + K.ea_off;
        const float4* B = verts + K.eb_off;
        const bool regjk = (K.ea_off != K.a_off) || (K.eb_off != K.b_off);   // warp-uniform

        int phase = 0;   // 0 idle, 1 GJK on the penetration geometry, 2 expanding polytope
        GjkState S;
        Tf T = tf_identity();
        int cfg = 0, myj = 0;
        float best_total = -FLT_MAX;
        refill_loop(cur + key, ecounts[key], lane, a.act_list + kbase, a.act_meta + kbase, [&]() { return phase == 0; },
            [&](int jj, int slot, int c) {
                myj = jj;
                cfg = c;
#ifdef JG_TRACE
                ++tr_started;
                tr_last = jg_now();
                tr_steps_at_last = tr_steps;
#endif
                JG_BOUNDS(jj >= 0 && jj < a.cap && slot >= 0 && slot < a.cap && c >= 0 && c < a.cap, "stage list entry (penetration refill)");
                const size_t j = kbase + myj;
                const int e0 = enc[cfg];
                T = load_tf(a, kbase + slot);
                best_total = -dec_float(e0);
                if (regjk) {
                    gjk_init(S, mk(T.tx, T.ty, T.tz));
                    phase = 1;
                } else {
                    phase = 2;
                    const ThreadSupport sup = {A, K.ea_n, B, K.eb_n, T};
                    if (!epa_begin(X, sup, simplex_from_ids(A, B, T, a.epa_ids[j]))) {
                        atomicMin(enc + cfg, enc_float(fminf(-K.emsum, -K.msum)));
                        phase = 0;
                    }
                }
            },
            [&]() {
#ifdef JG_TRACE
                ++tr_steps;
#endif
                if (phase == 1) {
                    const int st = gjk_step(S, A, K.ea_n, SetSupport{B, K.eb_n}, T, 1e30f);
                    if (st == JG_GJK_PENETRATING) {
                        phase = 2;
                        const ThreadSupport sup = {A, K.ea_n, B, K.eb_n, T};
                        if (!epa_begin(X, sup, S.S)) {
                            atomicMin(enc + cfg, enc_float(fminf(-K.emsum, -K.msum)));
                            phase = 0;
                        }
                    } else if (st != JG_GJK_RUNNING) {   // the penetration geometry does not overlap after all
                        atomicMin(enc + cfg, enc_float(fminf(sqrtf(S.vv) - K.emsum, -K.msum)));
                        phase = 0;
                    }
                } else if (phase == 2) {
                    const ThreadSupport sup = {A, K.ea_n, B, K.eb_n, T};
                    const int st = (a.flags & 0x40000000u) ? (int)JG_EPA_OVERFLOW : epa_step(X, sup);   // test hook
                    if (st == JG_EPA_RUNNING && X.ub + K.emsum + 1e-5f <= best_total) {
                        phase = 0;   // exact early out: this pair's depth is bounded below what its configuration already has
                    } else
                    if (st == JG_EPA_DONE) {
                        // the cores overlap: never report less penetration than the margins
                        atomicMin(enc + cfg, enc_float(fminf(-(epa_result(X) + K.emsum), -K.msum)));
                        phase = 0;
                    } else if (st == JG_EPA_OVERFLOW) {   // rare large polytope: redo with per-thread storage
                        const int o = atomicAdd(ovf_counter, 1);
                        if (o < a.ovf_cap) {
                            a.ovf_list[o] = make_int2(key, myj);
                            if (o < a.ovf_poly_cap) {   // hand the polytope over (nothing was modified by the failed step)
                                uint32_t* rec = a.ovf_poly + (size_t)o * JG_OVF_WORDS;
                                for (int k = 0; k < EpaCompact<BLOCK>::WORDS; ++k) rec[k] = X.poly.at(k);
                                rec[EpaCompact<BLOCK>::WORDS + 0] = __float_as_uint(X.lb);
                                rec[EpaCompact<BLOCK>::WORDS + 1] = __float_as_uint(X.ub);
                                rec[EpaCompact<BLOCK>::WORDS + 2] = (uint32_t)X.it;
                                rec[EpaCompact<BLOCK>::WORDS + 3] = (uint32_t)X.poly.nv;
                                rec[EpaCompact<BLOCK>::WORDS + 4] = (uint32_t)X.poly.nf;
                            }
                        } else atomicMin(enc + cfg, enc_float(fminf(-(epa_result(X) + K.emsum), -K.msum)));
                        phase = 0;
                    }
                }
            });
    }
#ifdef JG_TRACE
    {
        int started = tr_started, sal = tr_steps_at_last;
        unsigned long long last = tr_last;
        for (int o = 16; o; o >>= 1) {
            started += __shfl_xor_sync(0xFFFFFFFFu, started, o);
            sal = max(sal, __shfl_xor_sync(0xFFFFFFFFu, sal, o));
            const unsigned long long ol = __shfl_xor_sync(0xFFFFFFFFu, last, o);
            last = ol > last ? ol : last;
        }
        if (lane == 0 && gw < 4096) {
            unsigned long long* r = g_jg_trace[g_jg_trace_slot & 1][gw];
            r[0] = tr_t0; r[1] = jg_now(); r[2] = (unsigned long long)tr_keys; r[3] = (unsigned long long)started;
            r[4] = last; r[5] = (unsigned long long)tr_key0; r[6] = (unsigned long long)tr_steps; r[7] = (unsigned long long)(tr_steps - sal);
        }
    }
#endif
}

// Overflow pass: the few pairs whose polytope outgrew the fast storage class.  ONE PAIR PER WARP: the lanes split
// the vertex scans (lane-strided + shuffle arg-max, identical result on every lane) and then all execute the same
// polytope update on a large polytope held in shared memory (identical values written by every lane).  This keeps
// the latency of a 40-expansion query in the microseconds instead of a single thread walking 260 vertices serially.
struct WarpSupport {
    const float4* A; int nA;
    const float4* B; int nB;
    Tf T;
    int lane;
    __device__ __forceinline__ int scan(const float4* __restrict__ v, int n, V3 d) const {
        float best = -FLT_MAX;
        int bi = 0x7FFFFFFF;
        JG_WORK(JG_WORK_DOTS, (unsigned)((n - lane + 31) / 32));
        for (int i = lane; i < n; i += 32) {
            const float s = dot4(v[i], d);
            if (s > best) { best = s; bi = i; }
        }
#pragma unroll
        for (int o = 16; o; o >>= 1) {
            const float ob = __shfl_xor_sync(0xFFFFFFFFu, best, o);
            const int oi = __shfl_xor_sync(0xFFFFFFFFu, bi, o);
            if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
        }
        return bi;
    }
    __device__ __forceinline__ V3 operator()(V3 d, uint32_t& id) const {
        const int ia = scan(A, nA, tf_rot_t(T, d));
        const int ib = scan(B, nB, -d);
        const float4 pa = A[ia], pb = B[ib];
        id = ((uint32_t)ia << 16) | (uint32_t)ib;
        return tf_apply(T, mk(pa.x, pa.y, pa.z)) - mk(pb.x, pb.y, pb.z);
    }
    __device__ __forceinline__ unsigned work_unit() const { return lane == 0 ? 1u : 0u; }   // one expansion per WARP
};

template <int BLOCK>
__global__ void __launch_bounds__(BLOCK) k_penetration_big(const PassArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float4* verts = reinterpret_cast<float4*>(smem_raw);
    uint32_t* poly_words = reinterpret_cast<uint32_t*>(smem_raw + sizeof(float4) * a.n_verts);
    JG_SUCCESSOR_MAY_START();
    JG_WAIT_FOR_PREDECESSOR();
    const int n = min(cnt_misc(a)[CTR_OVF], a.ovf_cap);
    if (n == 0) return;
    for (int i = threadIdx.x; i < a.n_verts; i += BLOCK) verts[i] = a.verts[i];
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    EpaStateT<EpaPolyBig> X;
    X.poly.w = poly_words + (size_t)warp * EpaPolyBig::WORDS;
    for (int i = blockIdx.x * (BLOCK / 32) + warp; i < n; i += gridDim.x * (BLOCK / 32)) {
        const int2 e = a.ovf_list[i];
        JG_BOUNDS(e.x >= 0 && e.x < a.n_keys && e.y >= 0 && e.y < a.cap, "overflow list entry");
        const KeyRec K = a.keys[e.x];
        const size_t kbase = (size_t)e.x * a.cap;
        const size_t j = kbase + e.y;
        int cfg;
        const Tf T = load_entry(a, kbase + a.epa_slot[j], cfg);
        const WarpSupport sup = {verts + K.ea_off, K.ea_n, verts + K.eb_off, K.eb_n, T, lane};
        Simplex sx;
        bool pen = true;
        float d = 0.f;
        if (i < a.ovf_poly_cap) {
            // continue the compact polytope: lane l converts vertex l and face l into the large storage class (the
            // neighbour's edge index, which the compact form leaves out, is found from the shared vertex)
            typedef EpaCompact<1> C;
            const uint32_t* rec = a.ovf_poly + (size_t)i * JG_OVF_WORDS;
            const int nv = (int)rec[C::WORDS + 3], nf = (int)rec[C::WORDS + 4];
            __syncwarp();
            if (lane < nv)   // vertex ids are not kept by the compact form: unique dummies (the bracket test ends the run)
                X.poly.setP(lane, mk(__uint_as_float(rec[3 * lane]), __uint_as_float(rec[3 * lane + 1]), __uint_as_float(rec[3 * lane + 2])),
                            0xFFFFFF00u + (uint32_t)lane);
            if (lane < nf) {
                const uint32_t fx = rec[C::oFx + lane];
                uint32_t fv = 0u, fa = 0u;
                float fd = FLT_MAX;
                if ((fx >> 30) & 1u) {
                    fv = (uint32_t)cx_vert(fx, 0) | ((uint32_t)cx_vert(fx, 1) << 8) | ((uint32_t)cx_vert(fx, 2) << 16) | (1u << 24);
                    fd = __uint_as_float(rec[C::oFd + lane]);
#pragma unroll
                    for (int k = 0; k < 3; ++k) {
                        const int g = cx_nbr(fx, k), vend = cx_vert(fx, k == 2 ? 0 : k + 1);
                        const uint32_t fxg = rec[C::oFx + g];
                        const int eg = cx_vert(fxg, 0) == vend ? 0 : (cx_vert(fxg, 1) == vend ? 1 : 2);
                        fa |= ((uint32_t)g | ((uint32_t)eg << 8)) << (10 * k);
                    }
                }
                X.poly.fv(lane) = fv;
                X.poly.fa(lane) = fa;
                X.poly.set_fd(lane, fd);
            }
            X.poly.nv = nv;
            X.poly.nf = nf;
            X.lb = __uint_as_float(rec[C::WORDS + 0]);
            X.ub = __uint_as_float(rec[C::WORDS + 1]);
            X.it = (int)rec[C::WORDS + 2];
            __syncwarp();
            while (epa_step(X, sup) == JG_EPA_RUNNING) { __syncwarp(); }
            d = -(epa_result(X) + K.emsum);
            __syncwarp();
            if (lane == 0) atomicMin((K.kind == KEY_SELF ? a.enc_self : a.enc_scene) + cfg, enc_float(fminf(d, -K.msum)));
            continue;
        }
        if ((K.ea_off != K.a_off) || (K.eb_off != K.b_off)) {   // box: every lane reruns the (identical) GJK
            float dist;
            int it;
            pen = gjk_distance(sup.A, sup.nA, sup.B, sup.nB, T, mk(T.tx, T.ty, T.tz), 1e30f, dist, it, sx) == JG_GJK_PENETRATING;
            d = dist - K.emsum;
        } else {
            sx = simplex_from_ids(sup.A, sup.B, T, a.epa_ids[j]);
        }
        if (pen) {
            d = -K.emsum;
            __syncwarp();
            if (epa_begin(X, sup, sx)) {
                __syncwarp();
                while (epa_step(X, sup) == JG_EPA_RUNNING) { __syncwarp(); }
                d = -(epa_result(X) + K.emsum);
            }
        }
        __syncwarp();
        if (lane == 0) atomicMin((K.kind == KEY_SELF ? a.enc_self : a.enc_scene) + cfg, enc_float(fminf(d, -K.msum)));
    }
}

// ------------------------------------------------------------------------------------------------ finalize
// statistics of the pass: queue counters -> the engine's running totals (jg_get_stats)
struct StatArgs {
    const int* counts;      // NULL: nothing to account
    int n_scene_keys, n_keys, P;
    unsigned long long* stats;   // [0] hull pairs, [1] plane, [2] self, [3] overlapping, [4] GJK run, [5] polytopes, [6] overflows
    int last_set, last_slot;     // stage-list set / statistics slot of the pass's last stage (no k_select followed it)
};
__device__ __forceinline__ void accumulate_stats(const StatArgs& t, int lane) {
    const int n_keys = t.n_keys;
    unsigned long long h = 0, pl = 0, sf = 0, ep = 0, last = 0;
    const int* lact = t.counts + (t.last_set ? 5 * n_keys + 8 : 2 * n_keys);
    for (int k = lane; k < n_keys; k += 32) {
        const int c = t.counts[k];
        if (k >= t.n_scene_keys) sf += c;
        else if (k % (t.P + 1) == t.P) pl += c;
        else h += c;
        ep += t.counts[n_keys + k];
        last += lact[k];
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) {
        h += __shfl_xor_sync(0xFFFFFFFFu, h, o);
        pl += __shfl_xor_sync(0xFFFFFFFFu, pl, o);
        sf += __shfl_xor_sync(0xFFFFFFFFu, sf, o);
        ep += __shfl_xor_sync(0xFFFFFFFFu, ep, o);
        last += __shfl_xor_sync(0xFFFFFFFFu, last, o);
    }
    if (lane == 0) {
        if (t.last_slot >= 0) t.stats[t.last_slot] += last;
        t.stats[0] += h; t.stats[1] += pl; t.stats[2] += sf; t.stats[3] += ep;
        t.stats[6] += (unsigned long long)t.counts[4 * n_keys + CTR_OVF];   // polytopes handed to the overflow kernel
    }
}

// grid = ceil(n / 256) blocks for the configurations + ONE block that books the statistics
__global__ void k_finalize(const int* __restrict__ enc_scene, const int* __restrict__ enc_self,
                           const uint32_t* __restrict__ viol, int n, float qd, float threshold, uint32_t flags,
                           uint32_t* __restrict__ bits, float* __restrict__ dist, uint8_t* __restrict__ reason, const StatArgs st) {
    JG_WAIT_FOR_PREDECESSOR();
    if (blockIdx.x == gridDim.x - 1) {
        if (st.counts != nullptr && threadIdx.x < 32) accumulate_stats(st, threadIdx.x);
        return;
    }
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const bool valid = i < n;
    float ds = FLT_MAX, dself = FLT_MAX;
    bool lim = false;
    if (valid) {
        ds = dec_float(enc_scene[i]);
        dself = dec_float(enc_self[i]);
        lim = (viol[i >> 5] >> (i & 31)) & 1u;    // outside the joint limits (JG_LIMITS) or not a finite configuration / pose
    }
    const bool cself = dself < threshold, cscene = ds < threshold;
    const bool bit = valid && (lim || cself || cscene);
    const unsigned w = __ballot_sync(0xFFFFFFFFu, bit);
    if (valid) {
        if ((threadIdx.x & 31) == 0) bits[i >> 5] = w;
        if (dist) dist[i] = fminf(fminf(ds, dself), qd);
        if (reason) reason[i] = lim ? 1 : (cself ? 2 : (cscene ? 3 : 0));
    }
}

__global__ void k_init_scratch(int* enc_scene, int* enc_self, unsigned long long* champ_scene,
                               unsigned long long* champ_self, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) {
        enc_scene[i] = JG_ENC_INF;
        enc_self[i] = JG_ENC_INF;
        champ_scene[i] = 0ull;
        champ_self[i] = 0ull;
    }
}

// ------------------------------------------------------------------------------------------------ gripper poses
// Free-floating gripper: one thread per pose of the end-effector link; the listed shapes ride on it rigidly.
template <int BLOCK, bool MESH>
__global__ void __launch_bounds__(BLOCK) k_broadphase_poses(const PassArgs a, const float* __restrict__ poses, int fmt,
                                                            const Tf* __restrict__ rel, const int* __restrict__ use_shapes,
                                                            int n_use) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    RobotTab* R = reinterpret_cast<RobotTab*>(smem_raw);
    PieceRec* pieces = reinterpret_cast<PieceRec*>(smem_raw + sizeof(RobotTab));
    const int tid = threadIdx.x;
    {
        const int* src = reinterpret_cast<const int*>(a.robot);
        int* dst = reinterpret_cast<int*>(R);
        for (int i = tid; i < (int)(sizeof(RobotTab) / 4); i += BLOCK) dst[i] = src[i];
        const int* ps = reinterpret_cast<const int*>(a.pieces);
        int* pd = reinterpret_cast<int*>(pieces);
        for (int i = tid; i < (int)(sizeof(PieceRec) / 4) * a.P; i += BLOCK) pd[i] = ps[i];
    }
    __syncthreads();
    const int cfg = blockIdx.x * BLOCK + tid;
    const bool valid = cfg < a.n;
    bool geo = valid;
    Tf Pz = tf_identity();
    if (valid) {
        a.enc_scene[cfg] = JG_ENC_INF;
        a.enc_self[cfg] = JG_ENC_INF;
        a.champ_scene[cfg] = 0ull;
        a.champ_self[cfg] = 0ull;
        if (fmt == 0) {
            const float4* p4 = reinterpret_cast<const float4*>(poses + (size_t)cfg * 12);
            const float4 r0 = p4[0], r1 = p4[1], r2 = p4[2];
            Pz.r00 = r0.x; Pz.r01 = r0.y; Pz.r02 = r0.z; Pz.tx = r0.w;
            Pz.r10 = r1.x; Pz.r11 = r1.y; Pz.r12 = r1.z; Pz.ty = r1.w;
            Pz.r20 = r2.x; Pz.r21 = r2.y; Pz.r22 = r2.z; Pz.tz = r2.w;
        } else {
            const float* p = poses + (size_t)cfg * 7;
            float x = p[3], y = p[4], z = p[5], w = p[6];
            const float inv = rsqrtf(fmaf(x, x, fmaf(y, y, fmaf(z, z, w * w))));
            x *= inv; y *= inv; z *= inv; w *= inv;
            Pz.r00 = 1.f - 2.f * (y * y + z * z); Pz.r01 = 2.f * (x * y - z * w);       Pz.r02 = 2.f * (x * z + y * w);
            Pz.r10 = 2.f * (x * y + z * w);       Pz.r11 = 1.f - 2.f * (x * x + z * z); Pz.r12 = 2.f * (y * z - x * w);
            Pz.r20 = 2.f * (x * z - y * w);       Pz.r21 = 2.f * (y * z + x * w);       Pz.r22 = 1.f - 2.f * (x * x + y * y);
            Pz.tx = p[0]; Pz.ty = p[1]; Pz.tz = p[2];
        }
        const float chk = fabsf(Pz.r00) + fabsf(Pz.r01) + fabsf(Pz.r02) + fabsf(Pz.r10) + fabsf(Pz.r11) + fabsf(Pz.r12) + fabsf(Pz.r20) +
                          fabsf(Pz.r21) + fabsf(Pz.r22) + fabsf(Pz.tx) + fabsf(Pz.ty) + fabsf(Pz.tz);
        if (!(chk <= FLT_MAX)) {   // NaN / infinite pose: rejected, no geometry
            atomicOr(a.viol + (cfg >> 5), 1u << (cfg & 31));
            geo = false;
        }
    }
    Champ champ;
    champ.scene = champ.self = 0ull;
    Champ* lc = MESH ? nullptr : &champ;   // mesh queues find their champions with atomics (see k_broadphase)
    for (int k = 0; k < n_use; ++k) {
        const int s = use_shapes[k];
        const Tf T = tf_mul(Pz, rel[k]);
        shape_vs_scene<MESH>(a, R->shapes[s], s, pieces, T, geo, cfg, lc, valid ? cfg * n_use + k : -1);
    }
    if (lc) champ_publish(a, champ.scene, a.champ_scene, cfg, valid);
}

// ------------------------------------------------------------------------------------------------ FK only
template <int BLOCK>
__global__ void __launch_bounds__(BLOCK) k_fk(const RobotTab* __restrict__ robot, const float* __restrict__ q, int n,
                                              int dof, float finger_open, const int* __restrict__ link_slot /*[L]*/,
                                              int n_out, Tf base, float* __restrict__ out) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    RobotTab* R = reinterpret_cast<RobotTab*>(smem_raw);
    float* qs = reinterpret_cast<float*>(smem_raw + sizeof(RobotTab));
    const int tid = threadIdx.x;
    {
        const int* src = reinterpret_cast<const int*>(robot);
        int* dst = reinterpret_cast<int*>(R);
        for (int i = tid; i < (int)(sizeof(RobotTab) / 4); i += BLOCK) dst[i] = src[i];
    }
    const int cfg0 = blockIdx.x * BLOCK, cfg = cfg0 + tid;
    const int total = min(BLOCK, n - cfg0) * dof;
    for (int i = tid; i < total; i += BLOCK) {
        const int c = i / dof, j = i - c * dof;
        qs[j * BLOCK + c] = q[(size_t)cfg0 * dof + i];
    }
    __syncthreads();
    if (cfg >= n) return;
    Tf cur = tf_identity(), saved = tf_identity();
    for (int i = 0; i < R->L; ++i) {
        const JointRec& J = R->joints[i];
        float val = 0.f;
        if (J.qidx >= 0) val = qs[J.qidx * BLOCK + tid];
        else if (J.qidx == -2) val = finger_open;
        if (J.parent_sel == 1) cur = chain_step(saved, J, val);
        else if (J.parent_sel == 2) cur = joint_local(J, val);
        else cur = chain_step(cur, J, val);
        if (J.save) saved = cur;
        const int slot = link_slot[i];
        if (slot >= 0) {
            const Tf W = tf_mul(base, cur);
            float* o = out + ((size_t)cfg * n_out + slot) * 12;
            o[0] = W.r00; o[1] = W.r01; o[2] = W.r02;  o[3] = W.tx;
            o[4] = W.r10; o[5] = W.r11; o[6] = W.r12;  o[7] = W.ty;
            o[8] = W.r20; o[9] = W.r21; o[10] = W.r22; o[11] = W.tz;
        }
    }
}

}  // namespace jg

// ================================================================================================ inverse kinematics
// "Next" row (f1): replaces pybullet.calculateInverseKinematics (simulation.py:251-310).  One thread per target pose of
// the end-effector link: damped least squares on the geometric Jacobian with a null-space pull towards the rest
// configuration (the reference passes lowerLimits / upperLimits / jointRanges / restPoses = home_conf), joint-limit
// clamping, up to `iters` iterations (reference: 100) or until the residual drops below `tol` (reference: 1e-3).
namespace jg {

struct IkArgs {
    const RobotTab* robot;
    const float* targets;   // [n,12] pose of the end-effector link, robot base frame
    const float* q_init;    // [n,dof] or NULL (rest pose)
    const float* q_rest;    // [dof] device
    int n, dof, ee_link, iters, pos_only;
    float tol, damping, null_gain, finger_open;
    float* q_out;           // [n,dof]
    float* err_out;         // [n] position error [m] + orientation error [deg]/1000 (simulation.py:299-306)
};

// solve (M) y = b for a symmetric positive definite 6x6 (Cholesky, in place)
__device__ __forceinline__ void spd6_factor(float (&M)[6][6]) {
#pragma unroll
    for (int j = 0; j < 6; ++j) {
        float d = M[j][j];
#pragma unroll
        for (int k = 0; k < 6; ++k) if (k < j) d -= M[j][k] * M[j][k];
        d = sqrtf(fmaxf(d, 1e-20f));
        M[j][j] = d;
#pragma unroll
        for (int i = 0; i < 6; ++i) if (i > j) {
            float v = M[i][j];
#pragma unroll
            for (int k = 0; k < 6; ++k) if (k < j) v -= M[i][k] * M[j][k];
            M[i][j] = v / d;
        }
    }
}
__device__ __forceinline__ void spd6_solve(const float (&L)[6][6], float (&b)[6]) {
#pragma unroll
    for (int i = 0; i < 6; ++i) {
        float v = b[i];
#pragma unroll
        for (int k = 0; k < 6; ++k) if (k < i) v -= L[i][k] * b[k];
        b[i] = v / L[i][i];
    }
#pragma unroll
    for (int i = 5; i >= 0; --i) {
        float v = b[i];
#pragma unroll
        for (int k = 0; k < 6; ++k) if (k > i) v -= L[k][i] * b[k];
        b[i] = v / L[i][i];
    }
}

template <int BLOCK>
__global__ void __launch_bounds__(BLOCK) k_ik(const IkArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    RobotTab* R = reinterpret_cast<RobotTab*>(smem_raw);
    {
        const int* src = reinterpret_cast<const int*>(a.robot);
        int* dst = reinterpret_cast<int*>(R);
        for (int i = threadIdx.x; i < (int)(sizeof(RobotTab) / 4); i += BLOCK) dst[i] = src[i];
    }
    __syncthreads();
    const int idx = blockIdx.x * BLOCK + threadIdx.x;
    if (idx >= a.n) return;
    const int dof = a.dof;
    float q[MAX_DOF];
    for (int j = 0; j < MAX_DOF; ++j) q[j] = j < dof ? (a.q_init ? a.q_init[(size_t)idx * dof + j] : a.q_rest[j]) : 0.f;
    Tf Tt;
    {
        const float* t = a.targets + (size_t)idx * 12;
        Tt.r00 = t[0]; Tt.r01 = t[1]; Tt.r02 = t[2]; Tt.tx = t[3];
        Tt.r10 = t[4]; Tt.r11 = t[5]; Tt.r12 = t[6]; Tt.ty = t[7];
        Tt.r20 = t[8]; Tt.r21 = t[9]; Tt.r22 = t[10]; Tt.tz = t[11];
    }
    float err = FLT_MAX;
    for (int it = 0; it <= a.iters; ++it) {
        // ---- FK with the world axis / origin of every driven joint
        V3 jax[MAX_DOF], jor[MAX_DOF];
        int jtype[MAX_DOF];
        for (int j = 0; j < MAX_DOF; ++j) { jax[j] = mk(0.f, 0.f, 0.f); jor[j] = mk(0.f, 0.f, 0.f); jtype[j] = 0; }
        Tf cur = tf_identity(), saved = tf_identity(), ee = tf_identity();
        // which joints lie on the path to the end effector: all driven joints of the Panda do (serial arm)
        for (int i = 0; i < R->L; ++i) {
            const JointRec& J = R->joints[i];
            float val = 0.f;
            if (J.qidx >= 0) val = q[J.qidx];
            else if (J.qidx == -2) val = a.finger_open;
            const Tf par = J.parent_sel == 1 ? saved : (J.parent_sel == 2 ? tf_identity() : cur);
            if (J.qidx >= 0) {
                JointRec J0 = J;
                J0.type = 0;
                const Tf frame = tf_mul(par, joint_local(J0, 0.f));   // joint frame before its own motion
                jax[J.qidx] = mk(fmaf(frame.r00, J.axis[0], fmaf(frame.r01, J.axis[1], frame.r02 * J.axis[2])),
                                 fmaf(frame.r10, J.axis[0], fmaf(frame.r11, J.axis[1], frame.r12 * J.axis[2])),
                                 fmaf(frame.r20, J.axis[0], fmaf(frame.r21, J.axis[1], frame.r22 * J.axis[2])));
                jor[J.qidx] = mk(frame.tx, frame.ty, frame.tz);
                jtype[J.qidx] = J.type;
            }
            cur = tf_mul(par, joint_local(J, val));
            if (J.save) saved = cur;
            if (i == a.ee_link) ee = cur;
        }
        // ---- task-space error: position + rotation vector of R_target * R_ee^T
        float e6[6];
        e6[0] = Tt.tx - ee.tx; e6[1] = Tt.ty - ee.ty; e6[2] = Tt.tz - ee.tz;
        {
            const float r00 = Tt.r00 * ee.r00 + Tt.r01 * ee.r01 + Tt.r02 * ee.r02, r01 = Tt.r00 * ee.r10 + Tt.r01 * ee.r11 + Tt.r02 * ee.r12,
                        r02 = Tt.r00 * ee.r20 + Tt.r01 * ee.r21 + Tt.r02 * ee.r22;
            const float r10 = Tt.r10 * ee.r00 + Tt.r11 * ee.r01 + Tt.r12 * ee.r02, r11 = Tt.r10 * ee.r10 + Tt.r11 * ee.r11 + Tt.r12 * ee.r12,
                        r12 = Tt.r10 * ee.r20 + Tt.r11 * ee.r21 + Tt.r12 * ee.r22;
            const float r20 = Tt.r20 * ee.r00 + Tt.r21 * ee.r01 + Tt.r22 * ee.r02, r21 = Tt.r20 * ee.r10 + Tt.r21 * ee.r11 + Tt.r22 * ee.r12,
                        r22 = Tt.r20 * ee.r20 + Tt.r21 * ee.r21 + Tt.r22 * ee.r22;
            const V3 sk = mk(0.5f * (r21 - r12), 0.5f * (r02 - r20), 0.5f * (r10 - r01));   // sin(angle) * axis
            const float sn = sqrtf(norm2(sk)), cs = 0.5f * (r00 + r11 + r22 - 1.f);
            const float ang = atan2f(sn, cs);
            V3 w;
            if (sn > 1e-6f) w = sk * (ang / sn);
            else if (cs > 0.f) w = sk;                       // ~identity
            else {                                           // ~pi: axis from the diagonal
                const V3 ax = mk(sqrtf(fmaxf(0.5f * (r00 + 1.f), 0.f)), sqrtf(fmaxf(0.5f * (r11 + 1.f), 0.f)),
                                 sqrtf(fmaxf(0.5f * (r22 + 1.f), 0.f)));
                w = ax * ang;
            }
            e6[3] = w.x; e6[4] = w.y; e6[5] = w.z;
            if (a.pos_only) { e6[3] = 0.f; e6[4] = 0.f; e6[5] = 0.f; }
        }
        const float perr = sqrtf(e6[0] * e6[0] + e6[1] * e6[1] + e6[2] * e6[2]);
        const float aerr = sqrtf(e6[3] * e6[3] + e6[4] * e6[4] + e6[5] * e6[5]);
        err = perr + aerr * (57.29577951f / 1000.f);
        if (it == a.iters || (perr < a.tol && aerr < a.tol)) break;
        // ---- Jacobian columns, (J J^T + lambda^2 I)
        float Jc[MAX_DOF][6];
        for (int j = 0; j < MAX_DOF; ++j) {
            if (j < dof && jtype[j] == 1) {
                const V3 lin = cross(jax[j], mk(ee.tx, ee.ty, ee.tz) - jor[j]);
                Jc[j][0] = lin.x; Jc[j][1] = lin.y; Jc[j][2] = lin.z;
                Jc[j][3] = a.pos_only ? 0.f : jax[j].x; Jc[j][4] = a.pos_only ? 0.f : jax[j].y; Jc[j][5] = a.pos_only ? 0.f : jax[j].z;
            } else if (j < dof && jtype[j] == 2) {
                Jc[j][0] = jax[j].x; Jc[j][1] = jax[j].y; Jc[j][2] = jax[j].z; Jc[j][3] = 0.f; Jc[j][4] = 0.f; Jc[j][5] = 0.f;
            } else {
                for (int r = 0; r < 6; ++r) Jc[j][r] = 0.f;
            }
        }
        // active set: a joint that sits on a limit and that the task gradient J^T e pushes further out cannot move; with
        // its column left in, the damped step keeps spending itself against the clamp (the commonest way a start ends
        // short of a reachable pose).  Take it out of this iteration's Jacobian.
        for (int j = 0; j < dof; ++j) {
            float g = 0.f;
#pragma unroll
            for (int r = 0; r < 6; ++r) g = fmaf(Jc[j][r], e6[r], g);
            if ((q[j] <= R->qlo[j] && g < 0.f) || (q[j] >= R->qhi[j] && g > 0.f)) {
#pragma unroll
                for (int r = 0; r < 6; ++r) Jc[j][r] = 0.f;
            }
        }
        float M[6][6];
#pragma unroll
        for (int r = 0; r < 6; ++r)
#pragma unroll
            for (int c = 0; c < 6; ++c) {
                float v = r == c ? a.damping * a.damping : 0.f;
                for (int j = 0; j < dof; ++j) v = fmaf(Jc[j][r], Jc[j][c], v);
                M[r][c] = v;
            }
        spd6_factor(M);
        // primary task
        float y[6];
#pragma unroll
        for (int r = 0; r < 6; ++r) y[r] = e6[r];
        spd6_solve(M, y);
        // null-space pull towards the rest configuration: z - J^T (J J^T + l^2 I)^-1 J z
        float z[MAX_DOF], jz[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        for (int j = 0; j < dof; ++j) {
            z[j] = a.null_gain * (a.q_rest[j] - q[j]);
#pragma unroll
            for (int r = 0; r < 6; ++r) jz[r] = fmaf(Jc[j][r], z[j], jz[r]);
        }
        spd6_solve(M, jz);
        float dq[MAX_DOF], big = 0.f;
        for (int j = 0; j < dof; ++j) {
            float v = z[j];
#pragma unroll
            for (int r = 0; r < 6; ++r) v += Jc[j][r] * (y[r] - jz[r]);
            dq[j] = v;
            big = fmaxf(big, fabsf(v));
        }
        const float scale = big > 0.3f ? 0.3f / big : 1.f;   // trust region: at most 0.3 rad / 0.3 m per iteration
        for (int j = 0; j < dof; ++j) q[j] = fminf(fmaxf(fmaf(scale, dq[j], q[j]), R->qlo[j]), R->qhi[j]);
    }
    for (int j = 0; j < dof; ++j) a.q_out[(size_t)idx * dof + j] = q[j];
    a.err_out[idx] = err;
}

}  // namespace jg
