#!/usr/bin/env python
"""How the geometry of burg's two-finger gripper (``burg.gripper.TwoFingerGripperVisualisation(opening_width=0.08)``,
create_graspset.py:53 -- burg_toolkit is not vendored in the reference and not installable here) was RECOVERED from the
reference's own data: the 20 x 200 grasps in ``scenarios/*/grasps.npy`` all passed burg's mesh-vs-mesh filter, so in the
grasp frame the scene surfaces (``mesh_fn`` meshes of every object + the plane) must leave the gripper's volume empty.

    python tools/fit_burg_gripper.py [/root/reference] > profiles/r02_burg_gripper_fit.txt

Part 1 probes that empty region with line segments (where do scene surfaces cross the lines x = +-w/2 and the bar between
them?).  Part 2 sweeps the parameters of ``graspset.burg_gripper_model`` with the CPU oracle's zero-margin box-vs-triangle
test and counts how many of the 4 000 golden grasps each variant rejects: the recovered geometry is the only one with 0.
Builder-run (reads the reference tree); the resulting model is pinned by tests/test_graspset.py on baked data.
"""
import copy
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import oracle  # noqa: E402
from jogramop_framework_b200 import ingest as I  # noqa: E402
from jogramop_framework_b200.graspset import burg_gripper_model  # noqa: E402

IDS = [11, 12, 13, 14, 15, 21, 22, 23, 24, 25, 31, 32, 33, 34, 35, 41, 42, 43, 44, 45]


def visual_scene(ref, sid):
    scene, _ = I.load_scene_yaml(f'{ref}/scenarios/{sid:03d}/scene.yaml')
    bodies = [(*I.read_obj_mesh(o.mesh_fn), o.pose, o.is_target) for o in list(scene.objects) + list(scene.bg_objects)]
    return I.merge_meshes(bodies, I.default_robot_pose())


def seg_hits(p0, p1, tri):
    """segments p0->p1 [m,3] vs triangles [t,3,3] -> parameter of every crossing in [0,1], NaN elsewhere [m,t]"""
    d = p1 - p0
    a, e1, e2 = tri[:, 0], tri[:, 1] - tri[:, 0], tri[:, 2] - tri[:, 0]
    res = np.full((len(p0), len(tri)), np.nan)
    for r0 in range(0, len(p0), 64):
        o, dd = p0[r0:r0 + 64], d[r0:r0 + 64]
        pv = np.cross(dd[:, None, :], e2[None])
        det = np.einsum('rtk,tk->rt', pv, e1)
        ok = np.abs(det) > 1e-16
        inv = np.where(ok, 1 / np.where(ok, det, 1), 0)
        s = o[:, None, :] - a[None]
        u = np.einsum('rtk,rtk->rt', s, pv) * inv
        qv = np.cross(s, e1[None])
        v = np.einsum('rk,rtk->rt', dd, qv) * inv
        t = np.einsum('rtk,tk->rt', qv, e2) * inv
        res[r0:r0 + 64] = np.where(ok & (u >= 0) & (v >= 0) & (u + v <= 1) & (t >= 0) & (t <= 1), t, np.nan)
    return res


def main():
    ref = sys.argv[1] if len(sys.argv) > 1 else '/root/reference'
    scenes, grasps = {}, {}
    for sid in IDS:
        sc = visual_scene(ref, sid)
        scenes[sid] = sc
        grasps[sid] = np.linalg.inv(sc.robot_pose) @ np.load(f'{ref}/scenarios/{sid:03d}/grasps.npy').astype(float)
    plane = np.array([[[-5, -5, -0.05], [5, -5, -0.05], [5, 5, -0.05]], [[-5, -5, -0.05], [5, 5, -0.05], [-5, 5, -0.05]]], float)

    print('# part 1: where do scene surfaces cross lines of the grasp frame (all 4 000 kept grasps)?')
    z0, z1 = -0.05, 0.30
    for name, xy in (('finger line x=-0.040', (-0.040, 0)), ('finger line x=+0.040', (0.040, 0)), ('centre line x=0', (0, 0))):
        zs = []
        for sid in IDS:
            tri = np.concatenate([scenes[sid].tri_verts[scenes[sid].tris], plane])
            G = grasps[sid]
            p0 = G[:, :3, :3] @ np.array([*xy, z0]) + G[:, :3, 3]
            p1 = G[:, :3, :3] @ np.array([*xy, z1]) + G[:, :3, 3]
            t = seg_hits(p0, p1, tri)
            zs.append((z0 + (z1 - z0) * t)[~np.isnan(t)])
        zs = np.sort(np.concatenate(zs))
        lo, hi = zs[zs < 0.03], zs[zs > 0.0495]
        print(f'{name}: {len(zs)} crossings; highest below z=0.03: {lo[-3:].round(4).tolist()}; '
              f'lowest above z=0.0495: {hi[:3].round(4).tolist()}')
    print('bar between the finger lines at height z: number of grasps in which a scene surface crosses it')
    for zb in (0.040, 0.044, 0.046, 0.048, 0.049, 0.050, 0.052, 0.056, 0.060, 0.064, 0.072, 0.080):
        c = 0
        for sid in IDS:
            tri = np.concatenate([scenes[sid].tri_verts[scenes[sid].tris], plane])
            G = grasps[sid]
            p0 = G[:, :3, :3] @ np.array([-0.04, 0, zb]) + G[:, :3, 3]
            p1 = G[:, :3, :3] @ np.array([0.04, 0, zb]) + G[:, :3, 3]
            c += int((~np.isnan(seg_hits(p0, p1, tri))).any(1).sum())
        print(f'  z = {zb:.3f}: {c}')

    print('# part 2: golden grasps (of 4 000) rejected by variants of burg_gripper_model (oracle, boolean, zero margins)')
    orc = {}

    def rejected(shift=0.0, **kw):
        n = 0
        for sid in IDS:
            sc = scenes[sid]
            if sid not in orc:
                m = copy.copy(sc)
                m.piece_verts = sc.tri_verts[sc.tris].reshape(-1, 3)
                m.piece_vert_offset = (np.arange(len(sc.tris) + 1) * 3).astype(np.int32)
                m.piece_body, m.piece_margin = sc.tri_body, np.zeros(len(sc.tris))
                orc[sid] = m
            o = oracle.Oracle(burg_gripper_model(**kw), orc[sid])
            G = grasps[sid].copy()
            G[:, :3, 3] += shift * G[:, :3, 2]
            _, d = o.check_poses(G, links=[0], ee_link=0, flags=oracle.SCENE | oracle.PLANE | oracle.NO_EPA, threshold=1e-9)
            n += int((d <= 1e-9).sum())
        return n
    print('recovered model (w 0.08, l 0.05, t 0.003):', rejected())
    for name, vals in (('finger_length', [0.046, 0.048, 0.049, 0.051, 0.052, 0.054, 0.058, 0.065]),
                       ('finger_thickness', [0.001, 0.002, 0.004, 0.006, 0.010]),
                       ('opening_width', [0.070, 0.076, 0.078, 0.079, 0.081, 0.082]),
                       ('stem_length', [0.01, 0.02, 0.03, 0.05])):
        print(f'{name}:', [(v, rejected(**{name: v})) for v in vals])
    print('whole gripper shifted along its z axis [m]:',
          [(z, rejected(shift=z)) for z in (0.002, 0.001, -0.0005, -0.001, -0.002, -0.005, -0.01, -0.02)])


if __name__ == '__main__':
    main()
