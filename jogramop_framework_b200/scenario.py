"""Drop-in replacement for the reference's ``scenario.py`` (``Scenario``, scenario.py:12-129).

Loads ``scenarios/<id>/scene.yaml`` + ``grasps.npy`` from a jogramop-framework checkout (``JOGRAMOP_ROOT`` or the
cwd) or, when none is present, the baked copies shipped under ``jogramop_framework_b200/data`` (tools/bake_data.py).
No burg_toolkit / open3d / pybullet.
"""
import os

import numpy as np

from . import ingest, simulation, util
from .simulation import FrankaRobot, GraspingSimulator

_DATA = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'data')


class GraspSet:
    """Minimal stand-in for ``burg.GraspSet``: ``poses`` [n,4,4], ``len()``, ``from_poses``, ``transform``."""

    def __init__(self, poses):
        self.poses = np.asarray(poses).reshape(-1, 4, 4)

    @classmethod
    def from_poses(cls, poses):
        return cls(poses)

    def __len__(self):
        return len(self.poses)

    def transform(self, tf):
        self.poses = np.asarray(tf) @ self.poses
        return self


class Scenario:
    """
    Describes a scenario, including:
    - scene with objects and obstacles
    - set of grasp candidates
    - simulator
    - robot
    """

    def __init__(self, scenario_id, robot_pose=None, with_platform=True, gate_fix=True, device=0):
        self.id = scenario_id
        self.device = device
        root = simulation.reference_root()
        self.robot_pose = np.asarray(robot_pose, dtype=np.float64) if robot_pose is not None else self.default_robot_pose()
        self.with_platform = with_platform
        if root is not None:
            self.scene_fn = os.path.join(root, util.SCENARIO_DIR, f'{scenario_id:03d}', 'scene.yaml')
            self.scene, lib = ingest.load_scene_yaml(self.scene_fn)
            overrides = None
            gate = os.path.join(_DATA, 'gate_0_54_pieces.npz')
            if gate_fix and os.path.exists(gate) and any(o.object_type == 'gate_0_54' for o in self.scene.bg_objects):
                z = np.load(gate)      # quirk Q1: the shipped gate_0_54_vhacd.obj is stale (one solid block)
                overrides = {'gate_0_54': [(z[f'v{i}'], z[f'f{i}']) for i in range(int(z['n']))]}
            self._scene_model = ingest.build_scene_model(self.scene, self.robot_pose, name=f'{scenario_id:03d}',
                                                         piece_overrides=overrides)
            grasps = np.load(os.path.join(root, util.SCENARIO_DIR, f'{scenario_id:03d}', 'grasps.npy'),
                             allow_pickle=True)
        else:
            variant = '' if gate_fix else '_singlehull'
            fn = os.path.join(_DATA, f'scene_{scenario_id:03d}{variant}.npz')
            if not os.path.exists(fn):
                fn = os.path.join(_DATA, f'scene_{scenario_id:03d}.npz')
            if not os.path.exists(fn):
                raise FileNotFoundError(f'scenario {scenario_id:03d}: neither a jogramop checkout (JOGRAMOP_ROOT) nor {fn}')
            self.scene_fn = fn
            with np.load(fn, allow_pickle=False) as z:
                d = {k: z[k] for k in z.files}
            self._scene_model = ingest.SceneModel.from_dict(d).retarget(self.robot_pose)
            grasps = d['grasps_world']
            self.scene = ingest.SceneDescription(
                objects=[ingest.SceneObject(str(t), p, None, True) for t, p in zip(d['object_types'], d['object_poses'])],
                bg_objects=[ingest.SceneObject(str(t), p, None, False)
                            for t, p in zip(d['bg_object_types'], d['bg_object_poses'])],
                scene_fn=fn)
        print(f'loaded scene with {len(self.scene.objects)} objects and {len(self.scene.bg_objects)} obstacles.')
        # convert from BURG to franka frame (scenario.py:31)
        grasps = grasps @ FrankaRobot.tf_grasp2ee
        self.gs = GraspSet.from_poses(grasps)
        print(f'loaded {len(self.gs)} grasps')
        self.select_indices = None  # for grasp set sub-selection
        self._visual_model = None

    def visual_scene_model(self):
        """The scene as burg's ``AntipodalGraspSampler.check_collisions`` sees it (create_graspset.py:54-59): the
        library's ``mesh_fn`` mesh of every object and bg object, merged, in the robot frame, plus the plane -- a
        SceneModel for mesh-mode engines (``GraspingSimulator.burg_gripper_collision_batch``)."""
        if self._visual_model is None:
            bodies = []
            baked = None
            for o in list(self.scene.objects) + list(self.scene.bg_objects):
                if getattr(o, 'mesh_fn', None):
                    v, f = ingest.read_obj_mesh(o.mesh_fn)
                else:
                    if baked is None:
                        baked = np.load(os.path.join(_DATA, 'visual_meshes.npz'), allow_pickle=False)
                    v, f = baked[f'v_{o.object_type}'], baked[f'f_{o.object_type}']
                bodies.append((v, f, o.pose, o.is_target))
            self._visual_model = ingest.merge_meshes(bodies, self.robot_pose, name=f'{self.id:03d}_visual')
        return self._visual_model

    @property
    def grasp_poses(self):
        if self.select_indices is None:
            return self.gs.poses
        return self.gs.poses[self.select_indices]

    def select_n_grasps(self, n=None, seed=None):
        # if n is None, will select all grasps again
        if n is None:
            print('erasing grasp selection. whole grasp set available again.')
            self.select_indices = None
            return
        if n > len(self.gs):
            raise ValueError('cannot select more grasps than available. n must be smaller than number of grasps')
        rng = np.random.default_rng(seed)
        self.select_indices = rng.choice(len(self.gs), n, replace=False)

    def get_robot_and_sim(self, with_gui=False, geometry='full'):
        """Get the robot and simulator handles (scenario.py:59-67).  ``geometry='simplified'`` loads the sister planners'
        box profile of the robot instead of the full collision meshes (see simulation.load_robot_model)."""
        sim = GraspingSimulator(verbose=with_gui, device=self.device)
        sim.add_scene(self._scene_model)
        sim.scene = self.scene
        sim._visual_model_fn = self.visual_scene_model
        robot = FrankaRobot(sim, self.robot_pose, with_platform=self.with_platform, geometry=geometry)
        return robot, sim

    @staticmethod
    def default_robot_pose():
        """scenario.py:69-80."""
        return ingest.default_robot_pose()

    def filter_configurations(self, candidates, robot=None):
        """The check half of ``get_ik_solutions`` (scenario.py:105-125) as ONE batched call: joint limits, then
        self-collision, then scene collision, with the reference's counters.  ``candidates``: [n,dof] (rows may be
        None-free only).  Returns (accepted [m,dof] float64, reason uint8[n], Timer)."""
        own_sim = None
        if robot is None:
            robot, own_sim = self.get_robot_and_sim()
        candidates = np.asarray(candidates, dtype=np.float64).reshape(-1, len(robot.arm_joint_ids))
        count = util.Timer()
        count.start('perform checks')
        _, _, reason = robot.check_batch(candidates.astype(np.float32), want_dist=False)
        reason = reason.cpu().numpy()
        count.stop('perform checks')
        for code, key in ((1, 'IK solution not in joint limits'), (2, 'IK solution in self-collision'),
                          (3, 'IK solution in collision'), (0, 'IK solution is OK')):
            k = int((reason == code).sum())
            if k:
                count.count(key, k)
        if own_sim is not None:
            own_sim.dismiss()      # engines of a simulator created here are released with it (robot <-> simulator is a cycle)
        return candidates[reason == 0], reason, count

    def get_ik_solutions(self):
        """scenario.py:82-129: IK for every grasp (null-space control, accepted when position error + deg/1000 <=
        0.015), then joint limits, self-collision and scene collision.  Same loop structure, state handling (the IK of
        grasp g starts from the configuration the previous grasp left the robot in) and counters as the reference."""
        robot, sim = self.get_robot_and_sim()
        joint_limits = robot.arm_joint_limits()

        def in_joint_limits(joint_conf):
            return np.all(joint_limits[:, 0] <= joint_conf) and np.all(joint_conf <= joint_limits[:, 1])

        ik_solutions = []
        count = util.Timer()
        print('calculating IK solutions for all grasps')
        for g in range(len(self.gs)):
            count.start('calculate IK')
            pos, orn = util.position_and_quaternion_from_tf(self.gs.poses[g], convention='pybullet')
            target_conf = robot.inverse_kinematics(pos, orn, null_space_control=True)
            count.stop('calculate IK')
            count.start('perform checks')
            if target_conf is not None:
                robot.reset_arm_joints(target_conf)
                if not in_joint_limits(target_conf):
                    target_conf = None
                    count.count('IK solution not in joint limits')
                elif robot.in_self_collision():
                    target_conf = None
                    count.count('IK solution in self-collision')
                elif robot.in_collision():
                    target_conf = None
                    count.count('IK solution in collision')
                else:
                    count.count('IK solution is OK')
            else:
                count.count('IK solution too far away from goal')
            count.stop('perform checks')
            if target_conf is not None:
                ik_solutions.append(target_conf)
        count.print()
        sim.dismiss()
        return np.array(ik_solutions)

    def ik_start_configurations(self, robot, restarts, seed=0):
        """Start configurations for ``restarts`` IK attempts per grasp -> [restarts * n, dof] (attempt-major).  A 9-dof
        arm has a 3-dimensional solution manifold per grasp and only a small part of it is collision-free (the hardest
        reference grasps: ~1 % of the converged attempts), so the batched variant covers it by sampling: attempt 0
        starts at home (the reference's start), odd attempts uniformly inside the joint limits, even attempts with the
        mobile base placed 0.25-0.8 m from the grasp, joint 1 turned towards it and the arm jittered around home."""
        n = len(self.gs)
        lim = robot.arm_joint_limits()
        home = np.asarray(robot.home_conf, dtype=np.float64)
        rng = np.random.default_rng(seed)
        T = np.linalg.inv(robot.pose) @ self.gs.poses            # targets in the robot frame
        starts = [np.tile(home, (n, 1))]
        for r in range(1, restarts):
            if r % 2 == 1 or not self.with_platform:
                starts.append(lim[:, 0] + (lim[:, 1] - lim[:, 0]) * rng.random((n, len(lim))))
                continue
            q = np.tile(home, (n, 1))
            rad, phi = rng.uniform(0.25, 0.8, n), rng.uniform(0, 2 * np.pi, n)
            q[:, 1] = np.clip(T[:, 0, 3] - rad * np.cos(phi), lim[1, 0], lim[1, 1])      # base_joint_x is q[1] (quirk Q2)
            q[:, 0] = np.clip(T[:, 1, 3] - rad * np.sin(phi), lim[0, 0], lim[0, 1])      # base_joint_y is q[0]
            q[:, 2] = np.arctan2(T[:, 1, 3] - q[:, 0], T[:, 0, 3] - q[:, 1]) + rng.normal(0, 0.3, n)
            q[:, 3:] += rng.normal(0, 0.4, (n, q.shape[1] - 3))
            starts.append(np.clip(q, lim[:, 0], lim[:, 1]))
        return np.concatenate(starts)

    def get_ik_solutions_batch(self, restarts=256, seed=0):
        """Batched variant: all grasps x ``restarts`` start configurations (``ik_start_configurations``) go through ONE
        IK launch and ONE filter launch (limits, self-collision, scene).  Returns (accepted [m,dof], grasp index of each
        accepted row).  With 256 restarts every scenario yields at least as many distinct grasps as the reference's
        ``grasp_IK_solutions.csv`` holds (tests/test_export_and_ik.py, profiles/r02_ik_g6_counts.txt)."""
        robot, sim = self.get_robot_and_sim()
        n = len(self.gs)
        q0 = self.ik_start_configurations(robot, restarts, seed).astype(np.float32)
        poses = np.tile(self.gs.poses, (restarts, 1, 1))
        q, err = robot.inverse_kinematics_batch(poses, q_init=q0)
        _, _, reason = robot.check_batch(q, want_dist=False)
        q, err = q.cpu().numpy().astype(np.float64), err.cpu().numpy()
        ok = (err <= 0.015) & (reason.cpu().numpy() == 0)
        idx = np.nonzero(ok)[0]
        sim.dismiss()
        return q[idx], idx % n
