/*
 * jogramop_b200.h -- C-ABI of the B200-native batched collision-query engine.
 *
 * The reference (mrudorfer/jogramop-framework) has no FFI of its own: its hot path crosses from Python
 * into the pybullet C extension.  Each entry point below names the reference call (file:line under the
 * reference checkout) -- and the pybullet C-API call behind it -- that it replaces.  The Python classes
 * in jogramop_framework_b200/{simulation,scenario}.py bind these symbols with ctypes (INTEGRATION.md
 * shows the stub a reference maintainer would add).
 *
 * Conventions
 *   - extern "C", plain pointers and sizes, no torch / C++ types.  All functions that return int
 *     return 0 on success and a negative JG_ERR_* code on failure; jg_last_error() gives the
 *     thread-local message.  No exception crosses the boundary.
 *   - Handles (jg_robot / jg_scene / jg_engine) are created and destroyed in pairs by the library and
 *     are immutable after creation: batched calls on one engine are stream-ordered and may be issued
 *     from several host threads as long as each uses its own stream and output buffers ... EXCEPT that
 *     an engine owns ONE scratch arena, so concurrent calls on the same engine must be serialised by
 *     the caller (use one engine per stream).
 *   - Geometry is passed in double precision host arrays (the engine converts to fp32);
 *     configurations / results of the `jg_check*` / `jg_fk` entry points are DEVICE pointers (fp32),
 *     `stream` is a cudaStream_t (NULL = legacy default stream).  The `*_host` variants take HOST
 *     pointers and run the pipelined H2D -> kernels -> D2H path themselves.
 *   - Transforms are row-major 3x4 (R | t).
 *   - Frames: everything is expressed in the ROBOT base frame (scene pieces are
 *     inv(robot_pose) @ pose @ vertices, create_cpp_export.py:60-61).
 */
#ifndef JOGRAMOP_B200_H
#define JOGRAMOP_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define JG_ABI_VERSION 2

typedef struct jg_robot jg_robot;
typedef struct jg_scene jg_scene;
typedef struct jg_engine jg_engine;

enum {
    JG_OK = 0,
    JG_ERR_INVALID = -1,   /* bad argument (NULL, dof mismatch, unsupported topology ...) */
    JG_ERR_CUDA = -2,      /* a CUDA runtime call failed */
    JG_ERR_NOMEM = -3,
    JG_ERR_NODEVICE = -4,  /* no usable CUDA device: there is NO CPU fallback */
    JG_ERR_UNSUPPORTED = -5 /* entry point not available in this build of the library */
};

/* what jg_check tests (scenario.py:109-119 order: limits, self-collision, scene collision) */
enum {
    JG_SCENE = 1,   /* robot shapes vs obstacle / object pieces   (simulation.py:97-121 body_in_collision) */
    JG_PLANE = 2,   /* robot shapes vs ground plane               (burg plane.urdf in _env_bodies)         */
    JG_SELF = 4,    /* 42 ordered link pairs                      (simulation.py:439-464 in_self_collision) */
    JG_LIMITS = 8,  /* lo <= q <= hi                              (scenario.py:90-91 in_joint_limits)       */
    JG_NO_EPA = 16, /* skip the penetration-depth pass: overlapping cores report -(marginA+marginB)         */
    JG_DEBUG_FORCE_OVERFLOW = 0x40000000 /* test hook: send every penetrating pair through the overflow kernel */
};

enum { JG_JOINT_FIXED = 0, JG_JOINT_REVOLUTE = 1, JG_JOINT_PRISMATIC = 2 };
enum { JG_POSE_MAT34 = 0, JG_POSE_POS_QUAT_XYZW = 1 };

typedef struct {
    float threshold;       /* contactDistance below this = collision; default -0.001 (simulation.py:82)     */
    float query_distance;  /* getClosestPoints distance; default 0.01 (simulation.py:86)                    */
    float finger_open;     /* finger joint value; default 0.04 (simulation.py:172, reset_gripper 312-316)   */
    int32_t chunk;         /* configurations per internal pass; 0 = default                                 */
} jg_params;

void jg_default_params(jg_params* p);
const char* jg_last_error(void);
int jg_abi_version(void);
/* number of CUDA devices visible, or a negative JG_ERR_* */
int jg_device_count(void);

/* ---- robot: replaces burg SimulatorBase.load_robot -> pybullet.loadURDF (simulation.py:137-144) ----
 * Link i is the child of joint i (pybullet numbering), parents must precede children.
 * joint_origin [L,12]; joint_axis [L,3]; joint_qidx [L]: slot of the configuration vector driving the
 * joint, -1 = fixed at 0, -2 = finger (driven by params.finger_open); q_limits [dof,2].
 * Shapes (sorted by link; link -1 = base): convex hulls given by their vertices in the LINK frame.
 * core_verts are the GJK cores (box: shrunk by its margin), plane_verts the vertices used against the
 * plane and for penetration depth (box: exact corners); shape_margin / shape_plane_margin the margins
 * subtracted in either case.  self_pairs [K,2] are LINK ids. */
jg_robot* jg_robot_create(int n_links, const int32_t* joint_type, const int32_t* joint_parent,
                          const double* joint_origin, const double* joint_axis, const int32_t* joint_qidx,
                          const double* q_limits, int dof,
                          int n_shapes, const int32_t* shape_link, const double* shape_margin,
                          const int32_t* shape_vert_offset, const double* core_verts,
                          const int32_t* shape_plane_offset, const double* plane_verts,
                          const double* shape_plane_margin,
                          int n_self_pairs, const int32_t* self_pairs);
void jg_robot_free(jg_robot*);

/* ---- scene: replaces burg SimulatorBase.add_scene -> pybullet.loadURDF per object (scenario.py:64) ----
 * Convex pieces (hull vertices, robot frame) + optional ground plane n.p + d = 0 (plane[4], NULL = none). */
jg_scene* jg_scene_create(int n_pieces, const int32_t* piece_vert_offset, const double* piece_verts,
                          const double* piece_margin, const double* plane);
/* ---- scene as ONE triangle mesh ("mesh mode"): the merged export/obstacles.obj of create_cpp_export.py:40-69 ----
 * verts [n_verts,3] (robot frame), tris [n_tris,3] vertex indices.  The engine builds an LBVH over the triangles on
 * the GPU and reports the distance link hull <-> triangle (triangle margin `margin`, default 0.001); overlapping
 * pairs report -(margins) (no penetration depth on a surface).  Verdict / query-distance semantics as for pieces. */
jg_scene* jg_scene_create_mesh(int n_verts, const double* verts, int n_tris, const int32_t* tris, double margin,
                               const double* plane);
void jg_scene_free(jg_scene*);

/* ---- file-level loaders: the same handles straight from the reference's own files, for callers without Python
 *      (the sister C++ planners, README.md:52-56).  Replace pybullet.loadURDF (simulation.py:137-144) and
 *      burg Scene.from_yaml + SimulatorBase.add_scene (scenario.py:24,64); semantics of SURVEY.md Appendix A/C/E:
 *      one convex hull (margin 0.001) per OBJ `o`/`g` shape, a file without groups is ONE shape, <box> embeds its
 *      margin, link i = child of joint i (depth first, file order), pieces in the ROBOT frame
 *      inv(robot_pose) @ pose @ v with objects first, then bg_objects (create_cpp_export.py:36,41,60-61), plane z_world = 0.
 * jg_robot_load: `urdf_path` = robots/franka_panda/{mobile_panda,panda}.urdf; mesh filenames are relative to `mesh_root`
 *   (NULL = the URDF's directory).  Configuration vector = the movable joints in link order except joints whose name
 *   contains "finger" (these follow params.finger_open): [base_y, base_x, j1..j7] for the mobile Panda
 *   (simulation.py:159).  Self-collision pairs = FrankaRobot.in_self_collision's (simulation.py:439-464).
 * jg_scene_load: `scene_yaml` = scenarios/<id>/scene.yaml (its object_library_fn is followed);  robot_pose = row-major
 *   4x4 world pose of the robot base, NULL = Scenario.default_robot_pose() (scenario.py:69-80).  Objects with
 *   scale != 1 are refused (JG_ERR_UNSUPPORTED, create_cpp_export.py:33-34).
 * jg_scene_load_ex: ungrouped_obj_policy for OBJ files without groups: JG_OBJ_SINGLE_HULL (pyBullet's reading) or
 *   JG_OBJ_CONNECTED_COMPONENTS (one hull per connected component: SURVEY quirk Q1, the stale gate_0_54_vhacd.obj).
 * jg_scene_load_mesh: the scene as one triangle mesh (mesh mode): JG_MESH_VHACD = the merged vhacd files (what
 *   create_cpp_export.py writes as export/obstacles.obj), JG_MESH_VISUAL = the library's mesh_fn meshes (what burg's
 *   check_collisions tests against, create_graspset.py:54-59). */
enum { JG_OBJ_SINGLE_HULL = 0, JG_OBJ_CONNECTED_COMPONENTS = 1 };
enum { JG_MESH_VHACD = 0, JG_MESH_VISUAL = 1 };
jg_robot* jg_robot_load(const char* urdf_path, const char* mesh_root);
jg_scene* jg_scene_load(const char* scene_yaml, const double robot_pose[16]);
jg_scene* jg_scene_load_ex(const char* scene_yaml, const double robot_pose[16], int ungrouped_obj_policy);
jg_scene* jg_scene_load_mesh(const char* scene_yaml, const double robot_pose[16], int mesh_source, double margin);

/* read-only views of a handle's tables (valid until the handle is freed): what a C caller needs to size its buffers,
 * and what the tests compare with the Python ingestion */
typedef struct {
    int n_links, n_shapes, dof, n_self_pairs;
    int ee_link;                     /* end-effector link (panda_grasptarget: 14 / 11) or -1 when unknown */
    const int32_t *joint_type, *joint_parent, *joint_qidx;
    const double *joint_origin, *joint_axis, *q_limits;
    const int32_t* shape_link;
    const double* shape_margin;
    const int32_t* shape_vert_offset;
    const double* core_verts;
    const int32_t* shape_plane_offset;
    const double* plane_verts;
    const double* shape_plane_margin;
    const int32_t* self_pairs;
} jg_robot_desc;
typedef struct {
    int n_pieces;
    const int32_t* piece_vert_offset;
    const double* piece_verts;
    const double* piece_margin;
    const int32_t* piece_body;       /* body index of every piece (loaded scenes; NULL otherwise) */
    int n_bodies, n_target_bodies;   /* bodies 0 .. n_target_bodies-1 are the scene's `objects` (grasp targets) */
    int has_plane;
    double plane[4];
    int is_mesh, n_mesh_verts, n_mesh_tris;
    const float* mesh_verts;
    const int32_t* mesh_tris;
} jg_scene_desc;
int jg_robot_describe(const jg_robot*, jg_robot_desc* out);
int jg_robot_link_index(const jg_robot*, const char* link_name);   /* link id, or a negative JG_ERR_* */
int jg_scene_describe(const jg_scene*, jg_scene_desc* out);

/* ---- engine: uploads hulls / pieces to HBM on `device`, owns the scratch arena ---- */
jg_engine* jg_engine_create(const jg_robot*, const jg_scene* /* may be NULL: FK / self-collision only */,
                            int device, const jg_params* /* NULL = defaults */);
void jg_engine_free(jg_engine*);
int jg_engine_device(const jg_engine*);
/* Configurations one pass of this engine holds at most (jg_params.chunk, or what the memory budget allowed; <= 2^20).
 * Larger calls are processed pass after pass on the engine's one scratch arena; a caller that wants the passes of a large
 * call side by side cuts it at multiples of this size (and of 32) over several engines on several streams. */
int64_t jg_engine_pass_size(const jg_engine*);
/* Lanes (default 1): with lanes > 1 a jg_check / jg_check_edges / jg_check_poses call of MORE than one pass hands its passes
 * round-robin to `lanes` scratch arenas (this engine's and lanes - 1 sibling engines', created on first use, each as large as
 * this engine's) on internal streams, forked from and joined back into the caller's stream: the kernels of one pass run
 * under the tails of the others (1 M edges x 64 substeps: +40 % at 4 lanes).  Same results, same stream semantics; calls of
 * one pass, calls made while profiling and calls on a stream under CUDA-graph capture are not cut.  1 <= lanes <= 8.
 * Mesh-mode engines accept lanes too but gain nothing on the pose path (the LBVH traversal has no persistent-kernel tails to
 * fill: measured 5.4e8 vs 5.3e8 poses/s), so the Python layer keeps them at 1. */
int jg_engine_set_lanes(jg_engine*, int lanes);
int jg_engine_lanes(const jg_engine*);

/* ---- forward kinematics: replaces reset_arm_joints + getLinkState (simulation.py:206-212, 223-249) ----
 * q [n,dof] device fp32; link_ids [L_out] HOST ints; out [n,L_out,12] device fp32: link frames (URDF link
 * frame = getLinkState items 4-5) pre-multiplied by base_pose (HOST 3x4 doubles, NULL = identity). */
int jg_fk(jg_engine*, const float* q, int64_t n, int dof, const int32_t* link_ids, int n_link_ids,
          const double* base_pose, float* out, void* stream);

/* ---- collision check: replaces reset_arm_joints + in_collision / in_self_collision
 *      (simulation.py:206-212, 439-472 -> getClosestPoints, point[8] < threshold) ----
 * q [n,dof] device fp32.  bits: ceil(n/32) uint32 words, bit i%32 of word i/32 = configuration i collides
 * (or violates limits if JG_LIMITS).  min_dist [n] (may be NULL): min contactDistance over everything
 * checked, clamped to query_distance.  With min_dist == NULL only verdicts are computed: the engine then prunes
 * everything that cannot flip a bit (same bits, several times faster).  reason [n] uint8 (may be NULL): 0 ok, 1 limits, 2 self, 3 scene.
 * A configuration (edge end point, pose) with a NaN or infinite component has no geometry: its bit is set and its reason
 * is 1 whatever `flags` says; its distance is the query distance. */
int jg_check(jg_engine*, const float* q, int64_t n, int dof, uint32_t flags, uint32_t* bits, float* min_dist,
             uint8_t* reason, void* stream);

/* ---- edge validation (new batched API; north_star "planner edge validation") ----
 * qa,qb [m,dof]; substep k of `substeps` is qa + k/(substeps-1) (qb-qa); edge collides iff any substep does;
 * min_dist [m] = min over substeps. */
int jg_check_edges(jg_engine*, const float* qa, const float* qb, int64_t m, int dof, int substeps, uint32_t flags,
                   uint32_t* bits, float* min_dist, void* stream);

/* ---- free-floating gripper (create_graspset.py:53-60 check_collisions) ----
 * poses: pose of link `ee_link` in the robot frame, fmt JG_POSE_MAT34 [n,12] or JG_POSE_POS_QUAT_XYZW [n,7];
 * the shapes of `links` (HOST ints) move rigidly with it (offsets from FK with params.finger_open).  On an engine whose
 * scene was made with jg_scene_create_mesh the gripper is tested against the triangle surfaces (trimesh / fcl style;
 * with triangle margin 0, min_dist <= -(hull margin) means the hull itself touches a surface). */
int jg_check_poses(jg_engine*, const float* poses, int64_t n, int fmt, int ee_link, const int32_t* links,
                   int n_links, uint32_t flags, uint32_t* bits, float* min_dist, void* stream);

/* ---- inverse kinematics ("next" row: replaces pybullet.calculateInverseKinematics, simulation.py:251-310) ----
 * targets [n,12]: pose of link `ee_link` in the robot frame (device fp32); q_init [n,dof] or NULL (start from
 * q_rest); q_rest [dof] HOST doubles = null-space rest pose (the reference passes home_conf).  Damped least squares
 * with null-space pull and joint-limit clamping, at most `iters` iterations (reference 100), residual `tol`
 * (reference 1e-3).  q_out [n,dof]; err_out [n] = position error [m] + orientation error [deg]/1000, the quantity the
 * reference thresholds at 0.015 (simulation.py:299-306). */
int jg_ik(jg_engine*, const float* targets, int64_t n, int ee_link, const float* q_init, const double* q_rest, int iters,
          float tol, float damping, float null_gain, int position_only, float* q_out, float* err_out, void* stream);

/* ---- host-buffer convenience (the end-to-end path a drop-in caller uses): pinned staging, chunked
 *      H2D -> kernels -> D2H overlapped on two streams; blocks until the results are in host memory ---- */
int jg_check_host(jg_engine*, const float* q_host, int64_t n, int dof, uint32_t flags, uint32_t* bits_host,
                  float* min_dist_host, uint8_t* reason_host);

/* Asynchronous form for callers that stream batches (a planner validating batch after batch): up to TWO batches in flight,
 * so the upload of batch k+1 and the download of batch k overlap the kernels of the other batch.  n <= the pass
 * capacity (2^20 configurations unless jg_params.chunk says otherwise); buffers must stay valid until the wait returns.
 * jg_check_host_submit returns at once with a ticket; jg_check_host_wait blocks until that batch's results are in the
 * host buffers.  A third submit before the oldest ticket was waited for is refused (JG_ERR_INVALID). */
int jg_check_host_submit(jg_engine*, const float* q_host, int64_t n, int dof, uint32_t flags, uint32_t* bits_host,
                         float* min_dist_host, uint8_t* reason_host, int64_t* ticket);
int jg_check_host_wait(jg_engine*, int64_t ticket);

/* ---- instrumentation ---- */
typedef struct {
    int64_t configs;          /* configurations processed by the last jg_check* call                         */
    int64_t pairs;            /* (shape, piece) pairs that survived the broad phase                          */
    int64_t plane_pairs;      /* (shape, plane) candidates                                                   */
    int64_t self_pairs;       /* self-collision candidates                                                   */
    int64_t epa_pairs;        /* pairs whose cores overlapped (queued for the penetration pass)              */
    int64_t gjk_run;          /* pair-queue entries the GJK stages really processed (incl. plane entries)    */
    int64_t epa_run;          /* polytopes the penetration stages really expanded (after bound pruning)      */
    int64_t kernel_launches;  /* kernels launched by the last call                                           */
    int64_t overflow_run;     /* polytopes that outgrew the compact storage and were redone by the overflow kernel */
} jg_stats;
/* synchronises the engine's last stream and reads the device counters */
int jg_get_stats(jg_engine*, jg_stats*);

/* Per-kernel timing: when enabled, every pass brackets its kernels with CUDA events on the launching stream;
 * jg_get_profile synchronises and returns the accumulated milliseconds and launch counts since the last
 * jg_profile_enable(e, 1).  ms[0..3] / launches[0..3] = broad phase, narrow phase (GJK), penetration depth (EPA),
 * finalize. */
int jg_profile_enable(jg_engine*, int enable);
int jg_get_profile(jg_engine*, double ms[4], int64_t launches[4]);

/* FP32 FMA micro-benchmark (roofline denominator of the narrow phase): dependent-chain-free FFMA loop on every
 * SM of `device`; returns the best of `reps` runs in TFLOP/s (2 flops per FMA). */
int jg_measure_fp32_peak(int device, int reps, double* tflops);

/* Work counters of the instrumented build (libjogramop_b200_count.so, compiled with -DJG_COUNT_WORK): executed support
 * dot products, GJK iterations, polytope expansions, reserved -- accumulated on `device` since the last reset.  Used by
 * bench.py to audit the nominal flops-per-check of the roofline with a counted figure.  The
 * shipped library does not count: both calls return JG_ERR_UNSUPPORTED there. */
int jg_work_counters(int device, uint64_t out[4]);
int jg_work_counters_reset(int device);

#ifdef __cplusplus
}
#endif
#endif
