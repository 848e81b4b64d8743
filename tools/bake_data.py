#!/usr/bin/env python
"""Bake the reference's data files into the compact arrays this repo ships.

The GPU box has no ``/root/reference``; the engine therefore ships *derived* geometry
(convex-hull vertices of the robot links, obstacle pieces in the robot frame, the merged
triangle mesh, the 200 grasps per scenario) as small ``.npz`` files, and the reference's
golden artefacts as test fixtures.  This script is the committed
recipe that generated them:

    python tools/bake_data.py [/root/reference]

Outputs
  jogramop_framework_b200/data/robot_{mobile_panda,panda}.npz   RobotModel
  jogramop_framework_b200/data/robot_mobile_panda_simplified.npz  RobotModel of mobile_panda_fingersSmallMesh.urdf
  jogramop_framework_b200/data/scene_<id>.npz                   SceneModel (+ grasps_world)
  jogramop_framework_b200/data/scene_045_singlehull.npz         045 with the stale one-piece gate (quirk Q1)
  jogramop_framework_b200/data/gate_0_54_pieces.npz             recovered 2-piece gate decomposition
  jogramop_framework_b200/data/visual_meshes.npz                the library's ``mesh_fn`` meshes of every object type used
                                                                by a scenario (what burg's check_collisions tests against)
  tests/golden/export_golden.npz                                 export/*.csv, obstacles.obj triangles, per scenario
"""
import os
import sys
import warnings

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(__file__), '..'))
from jogramop_framework_b200 import ingest as I  # noqa: E402

SCENARIO_IDS = [11, 12, 13, 14, 15, 21, 22, 23, 24, 25, 31, 32, 33, 34, 35, 41, 42, 43, 44, 45]  # util.py:10


def recover_gate_054(ref):
    """Quirk Q1: recover the gate_0_54 decomposition used at export time from
    scenarios/045/export/obstacles.obj (triangles after the banana, before the pedestal)."""
    scene, lib = I.load_scene_yaml(os.path.join(ref, 'scenarios/045/scene.yaml'))
    order = list(scene.objects) + list(scene.bg_objects)
    ntri = [sum(len(f) for _, f in I.read_obj_shapes(o.vhacd_fn)) for o in order]
    (v, f), = I.read_obj_shapes(os.path.join(ref, 'scenarios/045/export/obstacles.obj'))
    gi = [o.object_type for o in order].index('gate_0_54')
    n_before = sum(ntri[:gi])
    n_after = sum(ntri[gi + 1:])
    gate_tris = f[n_before:len(f) - n_after]
    used, inv = np.unique(gate_tris.reshape(-1), return_inverse=True)
    gv, gf = v[used], inv.reshape(-1, 3)
    comps = I.split_connected_components(gv, gf, tol=1e-6)   # weld in the file's own 6-digit coordinates
    T = np.linalg.inv(np.linalg.inv(I.default_robot_pose()) @ order[gi].pose)
    return [c[0] @ T[:3, :3].T + T[:3, 3] for c in comps], [c[1] for c in comps]


def main():
    ref = sys.argv[1] if len(sys.argv) > 1 else '/root/reference'
    here = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..')
    data = os.path.join(here, 'jogramop_framework_b200', 'data')
    gold = os.path.join(here, 'tests', 'golden')
    os.makedirs(data, exist_ok=True)
    os.makedirs(gold, exist_ok=True)

    # mobile_panda_fingersSmallMesh.urdf: the simplified geometry profile of the sister planners (boxes around the links,
    # meshes/collisionSmall/*BBox.obj); pyBullet's checks in the reference use the full meshes, this one is optional
    for urdf, out in (('mobile_panda.urdf', 'robot_mobile_panda.npz'), ('panda.urdf', 'robot_panda.npz'),
                      ('mobile_panda_fingersSmallMesh.urdf', 'robot_mobile_panda_simplified.npz')):
        r = I.load_robot_urdf(os.path.join(ref, 'robots/franka_panda', urdf))
        r.save(os.path.join(data, out))
        print(out, 'links', r.n_links, 'shapes', r.n_shapes, 'hull verts', len(r.core_verts))

    gate_v, gate_f = recover_gate_054(ref)
    print('gate_0_54 recovered pieces:', [len(v) for v in gate_v], [len(f) for f in gate_f])
    np.savez_compressed(os.path.join(data, 'gate_0_54_pieces.npz'),
                        n=np.array(len(gate_v)), **{f'v{i}': v for i, v in enumerate(gate_v)},
                        **{f'f{i}': f for i, f in enumerate(gate_f)})

    golden = {}
    pose = I.default_robot_pose()
    visual = {}
    for sid in SCENARIO_IDS:
        sdir = os.path.join(ref, 'scenarios', f'{sid:03d}')
        scene, _ = I.load_scene_yaml(os.path.join(sdir, 'scene.yaml'))
        grasps = np.load(os.path.join(sdir, 'grasps.npy'), allow_pickle=True).astype(np.float32)
        for o in list(scene.objects) + list(scene.bg_objects):
            if o.object_type not in visual:
                v, f = I.read_obj_mesh(o.mesh_fn)
                used, inv = np.unique(f.reshape(-1), return_inverse=True)     # drop unreferenced vertices
                visual[o.object_type] = (v[used].astype(np.float32), inv.reshape(-1, 3).astype(np.int32))
        variants = [(f'scene_{sid:03d}.npz', {'gate_0_54': list(zip(gate_v, gate_f))})]
        if sid == 45:
            variants.append((f'scene_{sid:03d}_singlehull.npz', None))
        for fn, ov in variants:
            m = I.build_scene_model(scene, pose, name=f'{sid:03d}', piece_overrides=ov)
            d = m.to_dict()
            d['grasps_world'] = grasps
            d['object_types'] = np.array([o.object_type for o in scene.objects])
            d['bg_object_types'] = np.array([o.object_type for o in scene.bg_objects])
            d['object_poses'] = np.array([o.pose for o in scene.objects]).reshape(-1, 4, 4)
            d['bg_object_poses'] = np.array([o.pose for o in scene.bg_objects]).reshape(-1, 4, 4)
            np.savez_compressed(os.path.join(data, fn), **d)
            print(fn, 'pieces', m.n_pieces, 'tris', len(m.tris))
        # golden artefacts
        exp = os.path.join(sdir, 'export')
        with warnings.catch_warnings():
            warnings.simplefilter('ignore')
            ik = np.loadtxt(os.path.join(exp, 'grasp_IK_solutions.csv'), delimiter=',', ndmin=2)
        golden[f'ik_{sid:03d}'] = ik.reshape(-1, 9)
        golden[f'grasps_csv_{sid:03d}'] = np.loadtxt(os.path.join(exp, 'grasps.csv'), delimiter=',')
        golden[f'start_conf_{sid:03d}'] = np.loadtxt(os.path.join(exp, 'robot_start_conf.csv'), delimiter=',')
        (v, f), = I.read_obj_shapes(os.path.join(exp, 'obstacles.obj'))
        golden[f'obstacles_tris_{sid:03d}'] = v[f].astype(np.float64)
    np.savez_compressed(os.path.join(gold, 'export_golden.npz'), **golden)
    np.savez_compressed(os.path.join(data, 'visual_meshes.npz'), types=np.array(sorted(visual)),
                        **{f'v_{t}': visual[t][0] for t in visual}, **{f'f_{t}': visual[t][1] for t in visual})
    print('visual meshes:', {t: (len(v), len(f)) for t, (v, f) in visual.items()})
    print('golden IK rows:', sum(len(golden[f'ik_{s:03d}']) for s in SCENARIO_IDS))


if __name__ == '__main__':
    main()
