"""Export writers of ``create_cpp_export.export_for_cpp`` (create_cpp_export.py:23-82): the files the sister C++
planners consume, written to ``<out_dir>/{obstacles.obj, grasps.csv, robot_start_conf.csv, grasp_IK_solutions.csv}``
in the robot frame.  No open3d / burg_toolkit.

Formats: obstacles.obj = Open3D-style header, ``v x y z`` with 6 significant digits, ``f a b c``
(objects first, then bg_objects, vertices not shared across source meshes); grasps.csv = 200 x 16 row-major 4x4
(``np.savetxt`` default ``%.18e``); robot_start_conf.csv = 9 x 1; grasp_IK_solutions.csv = n x 9, possibly empty.
"""
import os

import numpy as np

HOME_CONF_WITH_PLATFORM = [0.0, 0.0, -0.017792060227770554, -0.7601235411041661, 0.019782607023391807, -2.342050140544315,
                           0.029840531355804868, 1.5411935298621688, 0.7534486589746342]   # create_cpp_export.py:75-77


def write_obstacles_obj(fn, scene_model, name='obstacles'):
    """The merged obstacle mesh in the robot frame (create_cpp_export.py:40-69), Open3D writer layout."""
    v, f = scene_model.tri_verts, scene_model.tris
    with open(fn, 'w') as fh:
        fh.write('# Created by Open3D \n')
        fh.write(f'# object name: {name}\n')
        fh.write(f'# number of vertices: {len(v)}\n')
        fh.write(f'# number of triangles: {len(f)}\n')
        for p in v:
            fh.write('v %g %g %g\n' % (p[0], p[1], p[2]))
        for t in f:
            fh.write('f %d %d %d\n' % (t[0] + 1, t[1] + 1, t[2] + 1))


def export_for_cpp(scenario, out_dir=None, ik_solutions=None, with_ik=True):
    """create_cpp_export.py:23-82.  ``ik_solutions``: precomputed [n,9] rows; else ``scenario.get_ik_solutions()`` is
    run when ``with_ik`` (needs the GPU).  Returns the output directory."""
    if out_dir is None:
        out_dir = os.path.join(os.path.dirname(scenario.scene_fn), 'export')
    os.makedirs(out_dir, exist_ok=True)
    model = scenario._scene_model.retarget(scenario.robot_pose)
    write_obstacles_obj(os.path.join(out_dir, 'obstacles.obj'), model)
    tf = np.linalg.inv(scenario.robot_pose)
    grasps = (tf @ scenario.gs.poses).reshape(len(scenario.gs), -1)
    np.savetxt(os.path.join(out_dir, 'grasps.csv'), grasps, delimiter=',')
    np.savetxt(os.path.join(out_dir, 'robot_start_conf.csv'), np.array(HOME_CONF_WITH_PLATFORM), delimiter=',')
    if ik_solutions is None and with_ik:
        ik_solutions = scenario.get_ik_solutions()
    if ik_solutions is not None:
        np.savetxt(os.path.join(out_dir, 'grasp_IK_solutions.csv'), np.asarray(ik_solutions).reshape(-1, 9), delimiter=',')
    return out_dir
