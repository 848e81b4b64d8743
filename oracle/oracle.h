/*
 * oracle.h -- CPU restatement (double precision, plain C) of the jogramop collision query.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the shipped product path; only
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load it.
 *
 * What it restates (the arithmetic lives in un-vendored, unpinned third-party code -- pybullet
 * (Bullet 3.x, double precision) reached through burg_toolkit@dev -- so the algorithm is restated
 * from its published behaviour and anchored on the reference's own call sites and golden files):
 *   simulation.py:206-212   FrankaRobot.reset_arm_joints      -> orc_fk (joint vector -> link frames)
 *   simulation.py:223-249   FrankaRobot.forward_kinematics    -> orc_fk
 *   simulation.py:82-95     GraspingSimulator.links_in_collision (getClosestPoints(..., 0.01), point[8] < -0.001)
 *   simulation.py:97-121    GraspingSimulator.body_in_collision  (plane + bg objects + target object)
 *   simulation.py:439-464   FrankaRobot.in_self_collision (42 ordered link pairs with the platform)
 *   simulation.py:466-472   FrankaRobot.in_collision
 *   scenario.py:90-91       in_joint_limits
 * Bullet semantics: every OBJ shape is a convex hull with collision margin
 * 0.001; a <box> embeds its margin (core = box shrunk by 0.001); contactDistance = GJK distance of
 * the cores - marginA - marginB, or -(EPA depth of the margin-inflated shapes) once the cores
 * overlap; hull-vs-plane distance = lowest vertex - margin; points are only reported up to the query
 * distance (0.01); verdict = any contactDistance < threshold (-0.001).
 *
 * PARITY PIN: pybullet cannot be installed in this image, so the oracle is pinned by the golden
 * artefacts the reference ships (tests/golden/export_golden.npz: 89 accepted IK rows = FK known
 * answers + collision known-negatives, grasps.csv frames, obstacles.obj merge) and by independent
 * numerical cross-checks (scipy NNLS / qhull on the Minkowski difference).  Round 2 added two pins
 * from data the reference ships: the 20 x 200 grasps of scenarios/<id>/grasps.npy, which all passed
 * the reference's gripper filter (create_graspset.py:53-60) -- the restated filter (zero-margin
 * gripper boxes vs the triangles of the mesh_fn meshes, orc_check_poses on triangle pieces) accepts
 * 4 000 of 4 000 and every change of the recovered gripper geometry starts rejecting them
 * (tests/test_graspset.py) -- and the row counts of export/grasp_IK_solutions.csv (IK outcomes,
 * tests/test_export_and_ik.py).  There are still no golden positives and no golden distances for the
 * robot path: its distance parity is "oracle-pinned", see DESIGN.md section 3.
 */
#ifndef JG_ORACLE_H
#define JG_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct orc_robot orc_robot;
typedef struct orc_scene orc_scene;

enum { ORC_SCENE = 1, ORC_PLANE = 2, ORC_SELF = 4, ORC_LIMITS = 8, ORC_NO_EPA = 16 };

/* joint_origin: [L,12] row-major 3x4.  joint_qidx[L]: index into the configuration vector, -1 fixed
 * (value 0), -2 finger (value = finger_open).  shape arrays as in ingest.RobotModel. */
orc_robot* orc_robot_create(int n_links, const int* joint_type, const int* joint_parent,
                            const double* joint_origin, const double* joint_axis, const int* joint_qidx,
                            const double* q_limits /*[dof,2]*/, int dof,
                            int n_shapes, const int* shape_link, const double* shape_margin,
                            const int* shape_vert_offset, const double* core_verts,
                            const int* shape_plane_offset, const double* plane_verts,
                            const double* shape_plane_margin,
                            int n_self_pairs, const int* self_pairs /*[K,2] link ids*/);
void orc_robot_free(orc_robot*);

orc_scene* orc_scene_create(int n_pieces, const int* piece_vert_offset, const double* piece_verts,
                            const double* piece_margin, const double* plane /*[4] or NULL*/);
void orc_scene_free(orc_scene*);

/* link frames in the robot base frame: out[n, L, 12] (row-major 3x4) */
int orc_fk(const orc_robot*, const double* q, int64_t n, int dof, double finger_open, double* out);

/* one check per configuration: bits[n] (0/1), dist[n] = min contactDistance clamped to query_distance,
 * reason[n] (optional): 0 ok, 1 limits, 2 self, 3 scene (first failing test in scenario.py:109-119 order). */
int orc_check(const orc_robot*, const orc_scene*, const double* q, int64_t n, int dof, uint32_t flags,
              double finger_open, double threshold, double query_distance,
              uint8_t* bits, double* dist, uint8_t* reason, int n_threads);

/* per (shape, piece|plane) contact distances for ONE configuration: out[S, P+1] (plane last), values
 * beyond the query distance are reported as +inf. */
int orc_pair_distances(const orc_robot*, const orc_scene*, const double* q, int dof, double finger_open,
                       double query_distance, uint32_t flags, double* out);
/* per self pair contact distances for ONE configuration: out[K] */
int orc_self_distances(const orc_robot*, const double* q, int dof, double finger_open,
                       double query_distance, uint32_t flags, double* out);

/* free-floating gripper: poses[n,12] = pose of link `ee_link` (3x4 row-major, robot frame); the shapes of
 * links in `links[n_links_used]` are placed rigidly relative to it using FK(q_ref). */
int orc_check_poses(const orc_robot*, const orc_scene*, const double* poses, int64_t n, int ee_link,
                    const int* links, int n_links_used, double finger_open, double threshold,
                    double query_distance, uint32_t flags, uint8_t* bits, double* dist, int n_threads);

/* edges: qa,qb [m,dof], `substeps` interpolated configurations q_k = qa + k/(substeps-1) (qb-qa) */
int orc_check_edges(const orc_robot*, const orc_scene*, const double* qa, const double* qb, int64_t m, int dof,
                    int substeps, uint32_t flags, double finger_open, double threshold, double query_distance,
                    uint8_t* bits, double* dist, int n_threads);

/* raw geometric kernels (for cross-checks): vertices [n,3], transforms 3x4 row-major (NULL = identity) */
double orc_gjk_distance(const double* va, int na, const double* Ta, const double* vb, int nb, const double* Tb,
                        double dmax, int* iterations);
/* penetration depth of two overlapping polytopes (EPA); returns < 0 if they do not overlap */
double orc_epa_depth(const double* va, int na, const double* Ta, const double* vb, int nb, const double* Tb,
                     int* iterations);

#ifdef __cplusplus
}
#endif
#endif
