"""Grasp-set creation for a scenario's target object: the front-end of ``create_graspset.py``.

The reference (``create_graspset.py:27-70``) samples 2000 antipodal grasp candidates on the target object with burg's
``AntipodalGraspSampler(n_orientations=7, only_grasp_from_above=True, no_contact_below_z=None)``, rejects the ones
whose gripper collides with the scene or the plane (``check_collisions(..., with_plane=True)``: trimesh / fcl), keeps a
random 200 and writes ``grasps.npy`` (``[m,4,4]`` world poses in burg's grasp convention).

burg_toolkit is not vendored in the reference and not installable here, so the sampler below RESTATES the published
antipodal scheme (parity unpinned; what is pinned are its properties, ``tests/test_graspset.py``):
  * a contact point p1 is drawn uniformly on the surface (area-weighted triangle, uniform barycentric);
  * a closing direction is drawn inside the friction cone (half angle atan(mu)) around the inward normal at p1, the
    ray from p1 along it leaves the object at p2; the pair is kept when the direction also lies inside the friction cone
    around the outward normal at p2 (antipodal), and 0 < |p2 - p1| <= max gripper width;
  * the grasp frame sits at the midpoint, x = closing direction, and ``n_orientations`` approach directions are spread
    evenly around x; ``only_grasp_from_above`` keeps those whose z axis (which points from the object towards the
    gripper: ``FrankaRobot.tf_grasp2ee`` maps it to -z of the end effector) has a non-negative world z component
    (checked on the reference's own ``grasps.npy``: every annotation has z_z >= 0, some exactly 0).
The sampling is host-side set-up work (2000 rays against a few hundred triangles).  The hot part -- the collision
filter -- is the batched gripper-pose kernel (``GraspingSimulator.gripper_in_collision_batch`` -> ``jg_check_poses``).
"""
import os

import numpy as np


def _ray_mesh_exit(orig, dirs, tri_a, e1, e2, eps=1e-9):
    """Farthest hit of each ray (t > eps) -> (t [m], triangle index [m], -1 = none): where a finger closing along the ray
    from outside meets the object (across cavities of a non-convex target too).  Moeller-Trumbore, rays x triangles."""
    m = len(orig)
    t_best = np.full(m, -1.0)
    idx = np.full(m, -1, dtype=np.int64)
    for r0 in range(0, m, 256):
        o, d = orig[r0:r0 + 256], dirs[r0:r0 + 256]
        p = np.cross(d[:, None, :], e2[None])
        det = np.einsum('rtk,tk->rt', p, e1)
        ok = np.abs(det) > 1e-14
        inv = np.where(ok, 1.0 / np.where(ok, det, 1.0), 0.0)
        s = o[:, None, :] - tri_a[None]
        u = np.einsum('rtk,rtk->rt', s, p) * inv
        qv = np.cross(s, e1[None])
        v = np.einsum('rk,rtk->rt', d, qv) * inv
        t = np.einsum('rtk,tk->rt', qv, e2) * inv
        hit = ok & (u >= -1e-9) & (v >= -1e-9) & (u + v <= 1 + 1e-9) & (t > eps)
        t = np.where(hit, t, -1.0)
        j = t.argmax(1)
        tb = t[np.arange(len(j)), j]
        t_best[r0:r0 + 256] = tb
        idx[r0:r0 + 256] = np.where(tb > 0, j, -1)
    return t_best, idx


class AntipodalGraspSampler:
    """``burg.sampling.AntipodalGraspSampler`` as used by ``create_graspset.py:40-50`` (same constructor / ``sample``
    arguments; ``sample`` takes the target's triangle mesh instead of a burg ObjectInstance)."""

    def __init__(self, n_orientations=7, only_grasp_from_above=True, no_contact_below_z=None, mu=0.25, seed=None):
        self.n_orientations = int(n_orientations)
        self.only_grasp_from_above = bool(only_grasp_from_above)
        self.no_contact_below_z = no_contact_below_z
        self.mu = float(mu)
        self.rng = np.random.default_rng(seed)

    def sample(self, verts, tris, n=10, max_gripper_width=0.08, max_rounds=200):
        """-> (poses [m,4,4] in the frame of ``verts``, contacts [m,2,3]), m <= n (== n unless the object offers no
        antipodal pairs)."""
        verts = np.asarray(verts, dtype=np.float64)
        tris = np.asarray(tris, dtype=np.int64)
        a, b, c = verts[tris[:, 0]], verts[tris[:, 1]], verts[tris[:, 2]]
        e1, e2 = b - a, c - a
        nrm = np.cross(e1, e2)
        area2 = np.linalg.norm(nrm, axis=1)
        keep = area2 > 1e-14
        a, e1, e2, nrm, area2 = a[keep], e1[keep], e2[keep], nrm[keep], area2[keep]
        nrm = nrm / area2[:, None]
        # outward orientation: signed volume must be positive (flip a consistently inward-wound mesh)
        if np.einsum('tk,tk->', a + (e1 + e2) / 3.0 - verts.mean(0), nrm * area2[:, None]) < 0:
            nrm = -nrm
        cdf = np.cumsum(area2) / area2.sum()
        cone = np.arctan(self.mu)
        poses, contacts = [], []
        for _ in range(max_rounds):
            if len(poses) >= n:
                break
            m = max(4 * n, 256)
            t = np.searchsorted(cdf, self.rng.random(m)).clip(0, len(cdf) - 1)
            r1, r2 = np.sqrt(self.rng.random(m)), self.rng.random(m)
            p1 = a[t] + (r1 * (1 - r2))[:, None] * e1[t] + (r1 * r2)[:, None] * e2[t]
            n1 = nrm[t]
            # direction inside the friction cone around -n1 (uniform in angle up to the cone half angle)
            ang = cone * np.sqrt(self.rng.random(m))
            phi = 2 * np.pi * self.rng.random(m)
            ref = np.where(np.abs(n1[:, :1]) < 0.9, np.array([[1.0, 0, 0]]), np.array([[0, 1.0, 0]]))
            u = np.cross(n1, ref)
            u /= np.linalg.norm(u, axis=1, keepdims=True)
            w = np.cross(n1, u)
            d = -n1 * np.cos(ang)[:, None] + (u * np.cos(phi)[:, None] + w * np.sin(phi)[:, None]) * np.sin(ang)[:, None]
            tt, hit = _ray_mesh_exit(p1 - 1e-7 * n1, d, a, e1, e2, eps=1e-6)
            ok = hit >= 0
            n2 = nrm[np.where(ok, hit, 0)]
            p2 = p1 + d * tt[:, None]
            width = np.linalg.norm(p2 - p1, axis=1)
            ok &= (np.einsum('ij,ij->i', d, n2) >= np.cos(cone) - 1e-12) & (width > 1e-4) & (width <= max_gripper_width)
            if self.no_contact_below_z is not None:
                ok &= (p1[:, 2] >= self.no_contact_below_z) & (p2[:, 2] >= self.no_contact_below_z)
            for i in np.nonzero(ok)[0]:
                x = (p2[i] - p1[i]) / width[i]
                centre = 0.5 * (p1[i] + p2[i])
                ref_i = np.array([0.0, 0, 1.0]) if abs(x[2]) < 0.95 else np.array([1.0, 0, 0])
                y0 = np.cross(ref_i, x)
                y0 /= np.linalg.norm(y0)
                z0 = np.cross(x, y0)
                off = self.rng.random() * 2 * np.pi / self.n_orientations
                for k in range(self.n_orientations):
                    th = off + 2 * np.pi * k / self.n_orientations
                    z = z0 * np.cos(th) + y0 * np.sin(th)
                    if self.only_grasp_from_above and z[2] < 0:   # horizontal approaches stay (the reference's sets have them)
                        continue
                    T = np.eye(4)
                    T[:3, 0], T[:3, 1], T[:3, 2], T[:3, 3] = x, np.cross(z, x), z, centre
                    poses.append(T)
                    contacts.append(np.stack([p1[i], p2[i]]))
                    if len(poses) >= n:
                        break
                if len(poses) >= n:
                    break
        if not poses:
            return np.zeros((0, 4, 4)), np.zeros((0, 2, 3))
        return np.stack(poses), np.stack(contacts)


def _box(lo, hi):
    lo, hi = np.asarray(lo, float), np.asarray(hi, float)
    return np.array([[x, y, z] for x in (lo[0], hi[0]) for y in (lo[1], hi[1]) for z in (lo[2], hi[2])])


def burg_gripper_model(opening_width=0.08, finger_length=0.05, finger_thickness=0.003, stem_length=0.0):
    """burg's simplified two-finger gripper (``burg.gripper.TwoFingerGripperVisualisation(opening_width=0.08)``,
    create_graspset.py:53) as a RobotModel of convex boxes in the GRASP frame (x = closing direction, z = from the finger
    tips back towards the gripper base; ``FrankaRobot.tf_grasp2ee`` maps it to the end-effector frame):

        finger 1   x in [-w/2 - t, -w/2]   y in [-t/2, t/2]   z in [0, l]
        finger 2   x in [ w/2,  w/2 + t]   y in [-t/2, t/2]   z in [0, l]
        bridge     x in [-w/2 - t, w/2 + t], y in [-t/2, t/2], z in [l, l + t]
        stem       (optional) x, y in [-t/2, t/2], z in [l + t, l + t + stem_length]

    burg_toolkit is not vendored, so the mesh itself is not in the reference tree; the layout (two bar fingers outside
    the opening joined by a bar) and the numbers w = 0.08 (given by the call site), l = 0.05, t = 0.003 were RECOVERED
    from the reference's own 20 x 200 kept grasps (``scenarios/*/grasps.npy``, 4 000 poses that passed burg's filter;
    tools/fit_burg_gripper.py): in the grasp frame the scene surfaces leave exactly this region empty -- the finger
    lines x = +-0.04 are never crossed for z in [0, 0.0535] but are reached down to z = -0.0001, the segment between
    them is crossed in 30 grasps at z = 0.048 and in none for z in [0.050, 0.062], and the void is < 6 mm wide in x and
    < 4 mm in y.  Nothing above z = 0.077 on the centre line can belong to the gripper (68 kept grasps have target
    surface there), so a stem, if burg's mesh has one, is shorter than 24 mm: ``stem_length`` defaults to 0.
    All margins are zero: the model is meant for the boolean surface-mesh mode (trimesh / fcl semantics)."""
    from . import ingest
    w, l, t = float(opening_width), float(finger_length), float(finger_thickness)
    parts = [_box([-w / 2 - t, -t / 2, 0], [-w / 2, t / 2, l]), _box([w / 2, -t / 2, 0], [w / 2 + t, t / 2, l]),
             _box([-w / 2 - t, -t / 2, l], [w / 2 + t, t / 2, l + t])]
    if stem_length > 0:
        parts.append(_box([-t / 2, -t / 2, l + t], [t / 2, t / 2, l + t + stem_length]))
    S = len(parts)
    off = (np.arange(S + 1) * 8).astype(np.int32)
    verts = np.concatenate(parts)
    return ingest.RobotModel(
        name='burg_two_finger_gripper', link_names=['grasp_frame'], joint_names=['grasp_frame_joint'],
        joint_type=np.array([ingest.JOINT_PRISMATIC], np.int32), joint_parent=np.array([-1], np.int32),
        joint_origin=np.eye(4)[None].copy(), joint_axis=np.array([[0.0, 0.0, 1.0]]), joint_limits=np.zeros((1, 2)),
        shape_link=np.zeros(S, np.int32), shape_type=np.full(S, ingest.SHAPE_HULL, np.int32), shape_margin=np.zeros(S),
        shape_vert_offset=off, core_verts=verts, shape_plane_offset=off.copy(), plane_verts=verts.copy(),
        shape_plane_margin=np.zeros(S), arm_joint_ids=[0], finger_joint_ids=[], end_effector_id=0,
        self_pairs=np.zeros((0, 2), np.int32))


def target_mesh_world(scenario):
    """Triangle mesh (verts [V,3], tris [F,3]) of the scenario's target object in the WORLD frame."""
    sc = scenario._scene_model
    tgt = np.nonzero(np.asarray(sc.body_is_target))[0]
    sel = np.isin(sc.tri_body, tgt)
    tris = sc.tris[sel]
    used, inv = np.unique(tris.reshape(-1), return_inverse=True)
    v = sc.tri_verts[used] @ sc.robot_pose[:3, :3].T + sc.robot_pose[:3, 3]
    return v, inv.reshape(-1, 3)


def check_collisions(grasp_poses, scenario, open_scale=1.0, surface_mesh=False, gripper='burg', sim=None):
    """``grasp_sampler.check_collisions(candidates, scene, gripper.mesh, with_plane=True)`` (create_graspset.py:53-60):
    0 = free, 1 = the gripper at that grasp collides with a scene body or the plane.  One batched GPU call.

    ``gripper='burg'`` (default, the reference's filter): burg's two-finger gripper (``burg_gripper_model``) against
    the surfaces of the scene's ``mesh_fn`` meshes, boolean (``GraspingSimulator.burg_gripper_collision_batch``).  All
    4 000 grasps the reference ships pass it (tests/test_graspset.py).
    ``gripper='panda'``: the real Panda hand + finger hulls placed through ``FrankaRobot.tf_grasp2ee`` -- stricter (the
    fingers are 2 cm wide and reach deeper): against the convex pieces with Bullet margins, or with
    ``surface_mesh=True`` against the triangle surfaces of export/obstacles.obj without margins."""
    if gripper == 'burg':
        if sim is None:
            sim = scenario.get_robot_and_sim()[1]
        return sim.burg_gripper_collision_batch(grasp_poses).cpu().numpy().astype(np.int32)
    if gripper != 'panda':
        raise ValueError("gripper must be 'burg' or 'panda'")
    from .simulation import FrankaRobot
    if sim is None:
        sim = scenario.get_robot_and_sim()[1]
    ee = np.asarray(grasp_poses, dtype=np.float64).reshape(-1, 4, 4) @ FrankaRobot.tf_grasp2ee
    if surface_mesh:   # boolean mesh-surface mode: no margins, obstacle interiors are free
        hit = sim.gripper_mesh_collision_batch(ee, open_scale=open_scale)
    else:
        hit, _ = sim.gripper_in_collision_batch(ee, open_scale=open_scale, want_dist=False)
    return hit.cpu().numpy().astype(np.int32)


def create_grasps(scenario, n_candidates=2000, keep=200, gripper_width=0.08, seed=None, save_fn=None, gripper='burg'):
    """``create_graspset.create_grasps`` (27-70) for a loaded ``Scenario``: sample, filter, keep a random subset,
    optionally save ``grasps.npy``.  Returns (grasps [m,4,4] world, number of candidates, number collision-free)."""
    sampler = AntipodalGraspSampler(n_orientations=7, only_grasp_from_above=True, no_contact_below_z=None, seed=seed)
    verts, tris = target_mesh_world(scenario)
    candidates, _ = sampler.sample(verts, tris, n=n_candidates, max_gripper_width=gripper_width)
    collisions = check_collisions(candidates, scenario, gripper=gripper)
    grasps = candidates[collisions == 0]
    n_free = len(grasps)
    if len(grasps) > keep:
        grasps = grasps[sampler.rng.choice(len(grasps), keep, replace=False)]
    if save_fn is not None:
        os.makedirs(os.path.dirname(os.path.abspath(save_fn)), exist_ok=True)
        np.save(save_fn, grasps)
    return grasps, len(candidates), n_free
