"""Golden artefacts G1/G2/G5: frames, obstacle merge, start configuration."""
import numpy as np
import pytest

from jogramop_framework_b200 import ingest
from ref_numpy import SCENARIO_IDS

TF_GRASP2EE = np.array([[0, 1, 0, 0], [1, 0, 0, 0], [0, 0, -1, 0], [0, 0, 0, 1.0]])  # simulation.py:125-130


@pytest.mark.parametrize('sid', SCENARIO_IDS)
def test_g2_obstacle_merge_matches_export(sid, golden, scene_loader):
    """create_cpp_export.py:40-61: objects then bg_objects, inv(robot_pose) @ pose @ vhacd, same triangle order."""
    m = scene_loader(sid)
    ref = golden[f'obstacles_tris_{sid:03d}']
    mine = m.tri_verts[m.tris]
    assert mine.shape == ref.shape
    assert np.abs(mine - ref).max() < 1e-5   # obstacles.obj is written with 6 significant digits


def test_q1_stale_gate_file_differs(golden, scene_loader):
    """Quirk Q1: the current gate_0_54_vhacd.obj (one shape) does not reproduce scenarios/045/export/obstacles.obj."""
    assert scene_loader(45, '_singlehull').tris.shape[0] != golden['obstacles_tris_045'].shape[0]
    assert scene_loader(45).n_pieces == scene_loader(45, '_singlehull').n_pieces + 1


@pytest.mark.parametrize('sid', SCENARIO_IDS)
def test_g1_grasp_frames(sid, golden):
    """grasps.csv = inv(robot_pose) @ grasps.npy @ tf_grasp2ee (scenario.py:31, create_cpp_export.py:60-72)."""
    import os
    from conftest import ROOT
    z = np.load(os.path.join(ROOT, 'jogramop_framework_b200', 'data', f'scene_{sid:03d}.npz'))
    g = z['grasps_world'].astype(np.float64)
    assert g.shape == (200, 4, 4)
    mine = np.linalg.inv(ingest.default_robot_pose()) @ (g @ TF_GRASP2EE)
    assert np.abs(mine.reshape(200, 16) - golden[f'grasps_csv_{sid:03d}']).max() < 1e-6


def test_g5_start_conf(golden):
    home = [0.0, 0.0, -0.017792060227770554, -0.7601235411041661, 0.019782607023391807, -2.342050140544315,
            0.029840531355804868, 1.5411935298621688, 0.7534486589746342]      # simulation.py:166-169
    for sid in SCENARIO_IDS:
        assert np.allclose(golden[f'start_conf_{sid:03d}'], home, atol=0, rtol=0)


def test_robot_model_tables(robot):
    """15 links, 12 shapes, hull sizes, joint limits, 42 self-collision pairs."""
    assert robot.n_links == 15 and robot.n_shapes == 12
    assert robot.link_names[14] == 'panda_grasptarget' and robot.link_names[1] == 'mobile_platform'
    assert list(np.diff(robot.shape_vert_offset)) == [8, 102, 152, 152, 152, 152, 152, 260, 102, 102, 18, 18]
    assert robot.arm_joint_ids == [0, 1, 3, 4, 5, 6, 7, 8, 9] and robot.finger_joint_ids == [12, 13]
    lim = robot.joint_limits[robot.arm_joint_ids]
    assert np.allclose(lim[:3], [[-1, 1], [-1, 1], [-2.9671, 2.9671]])
    assert np.allclose(lim[5], [-3.1416, 0.0]) and np.allclose(lim[7], [-0.0873, 3.8223])
    assert robot.self_pairs.shape == (42, 2)
    assert (2, 4) in map(tuple, robot.self_pairs) and (8, 13) in map(tuple, robot.self_pairs)
    assert not any(10 in p or 1 in p or 14 in p for p in map(tuple, robot.self_pairs))
    # the platform box embeds its 1 mm margin (btBoxShape): core = half extents - margin
    assert np.allclose(np.abs(robot.shape_verts(0)).max(0), [0.199, 0.199, 0.009])
    assert np.allclose(np.abs(robot.shape_plane_verts(0)).max(0), [0.2, 0.2, 0.01])


def test_obj_reader_dialects(tmp_path):
    """OBJ syntax variants (Appendix A/E): f a//n, f a/b/c, polygons, negative indices, o/g splitting, no groups."""
    p = tmp_path / 'a.obj'
    p.write_text('# c\nmtllib x\nv  0 0 0\nv 1 0 0\nv 0 1 0\nv 0 0 1\nvn 0 0 1\ng one\nf 1//1 2//1 3//1\n'
                 'o two\nusemtl m\ns off\nf 1/1/1 2/1/1 4/1/1 3/1/1\nl 1 2\nf -4 -3 -1 \n')
    shapes = ingest.read_obj_shapes(str(p))
    assert len(shapes) == 2
    assert shapes[0][1].shape == (1, 3) and shapes[1][1].shape == (3, 3)
    p2 = tmp_path / 'b.obj'
    p2.write_text('v 0 0 0\nv 1 0 0\nv 0 1 0\nv 0 0 1\nf 1 2 3\nf 1 2 4\n')
    assert len(ingest.read_obj_shapes(str(p2))) == 1
