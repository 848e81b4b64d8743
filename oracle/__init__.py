"""ctypes wrapper of the CPU oracle (oracle/oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, ``__graft_entry__.smoke()`` and bench.py's
``cpu_baseline`` / ``--impl reference`` legs -- never by the product package
``jogramop_framework_b200``.  See oracle/oracle.h for what is restated (reference file:line)
and for the parity-pin statement.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, '_build', 'liboracle.so')

SCENE, PLANE, SELF, LIMITS, NO_EPA = 1, 2, 4, 8, 16


def build(force=False):
    """Compile oracle.c with gcc (oracle/Makefile)."""
    src = [os.path.join(_HERE, f) for f in ('oracle.c', 'oracle.h')]
    if (not force and os.path.exists(_LIB_PATH)
            and all(os.path.getmtime(_LIB_PATH) >= os.path.getmtime(s) for s in src)):
        return _LIB_PATH
    subprocess.check_call(['make', '-C', _HERE, '-s'] + (['-B'] if force else []))
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB_PATH)
        vp, i32, i64, f64, u32 = C.c_void_p, C.c_int, C.c_int64, C.c_double, C.c_uint32
        L.orc_robot_create.restype = vp
        L.orc_robot_create.argtypes = [i32, vp, vp, vp, vp, vp, vp, i32, i32, vp, vp, vp, vp, vp, vp, vp, i32, vp]
        L.orc_robot_free.argtypes = [vp]
        L.orc_scene_create.restype = vp
        L.orc_scene_create.argtypes = [i32, vp, vp, vp, vp]
        L.orc_scene_free.argtypes = [vp]
        L.orc_fk.argtypes = [vp, vp, i64, i32, f64, vp]
        L.orc_check.argtypes = [vp, vp, vp, i64, i32, u32, f64, f64, f64, vp, vp, vp, i32]
        L.orc_pair_distances.argtypes = [vp, vp, vp, i32, f64, f64, u32, vp]
        L.orc_self_distances.argtypes = [vp, vp, i32, f64, f64, u32, vp]
        L.orc_check_poses.argtypes = [vp, vp, vp, i64, i32, vp, i32, f64, f64, f64, u32, vp, vp, i32]
        L.orc_check_edges.argtypes = [vp, vp, vp, vp, i64, i32, i32, u32, f64, f64, f64, vp, vp, i32]
        L.orc_gjk_distance.restype = f64
        L.orc_gjk_distance.argtypes = [vp, i32, vp, vp, i32, vp, f64, vp]
        L.orc_epa_depth.restype = f64
        L.orc_epa_depth.argtypes = [vp, i32, vp, vp, i32, vp, vp]
        _lib = L
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def _i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def joint_qidx(robot):
    """Map joints -> configuration-vector slots: arm joint k reads q[k], finger joints read finger_open."""
    qidx = np.full(robot.n_links, -1, np.int32)
    for k, j in enumerate(robot.arm_joint_ids):
        qidx[j] = k
    for j in robot.finger_joint_ids:
        qidx[j] = -2
    return qidx


class Oracle:
    """Double-precision CPU reference for one (robot, scene) pair."""

    def __init__(self, robot, scene=None):
        L = lib()
        self.robot, self.scene = robot, scene
        self.dof = len(robot.arm_joint_ids)
        lim = _f64(robot.joint_limits[robot.arm_joint_ids])
        self._keep = [
            _i32(robot.joint_type), _i32(robot.joint_parent), _f64(robot.joint_origin[:, :3, :]), _f64(robot.joint_axis),
            joint_qidx(robot), lim, _i32(robot.shape_link), _f64(robot.shape_margin), _i32(robot.shape_vert_offset),
            _f64(robot.core_verts), _i32(robot.shape_plane_offset), _f64(robot.plane_verts),
            _f64(robot.shape_plane_margin), _i32(robot.self_pairs)]
        k = self._keep
        self._r = L.orc_robot_create(robot.n_links, _p(k[0]), _p(k[1]), _p(k[2]), _p(k[3]), _p(k[4]), _p(k[5]),
                                     self.dof, robot.n_shapes, _p(k[6]), _p(k[7]), _p(k[8]), _p(k[9]), _p(k[10]),
                                     _p(k[11]), _p(k[12]), len(robot.self_pairs), _p(k[13]))
        self._s = None
        if scene is not None:
            ks = [_i32(scene.piece_vert_offset), _f64(scene.piece_verts), _f64(scene.piece_margin), _f64(scene.plane)]
            self._s = L.orc_scene_create(scene.n_pieces, _p(ks[0]), _p(ks[1]), _p(ks[2]),
                                         _p(ks[3]) if scene.has_plane else None)

    def __del__(self):
        try:
            L = lib()
            if getattr(self, '_s', None):
                L.orc_scene_free(self._s)
            if getattr(self, '_r', None):
                L.orc_robot_free(self._r)
        except Exception:
            pass

    def fk(self, q, finger_open=0.04):
        """All link frames in the robot base frame: [n, L, 3, 4]."""
        q = _f64(np.atleast_2d(q))
        out = np.empty((len(q), self.robot.n_links, 3, 4))
        rc = lib().orc_fk(self._r, _p(q), len(q), q.shape[1], finger_open, _p(out))
        assert rc == 0
        return out

    def check(self, q, flags=SCENE | PLANE, finger_open=0.04, threshold=-0.001, query_distance=0.01, n_threads=0,
              want_reason=False):
        q = _f64(np.atleast_2d(q))
        n = len(q)
        bits = np.zeros(n, np.uint8)
        dist = np.zeros(n)
        reason = np.zeros(n, np.uint8) if want_reason else None
        rc = lib().orc_check(self._r, self._s, _p(q), n, q.shape[1], flags, finger_open, threshold, query_distance,
                             _p(bits), _p(dist), _p(reason), n_threads)
        assert rc == 0
        return (bits.astype(bool), dist, reason) if want_reason else (bits.astype(bool), dist)

    def pair_distances(self, q, flags=SCENE | PLANE, finger_open=0.04, query_distance=0.01):
        q = _f64(q).reshape(-1)
        out = np.empty((self.robot.n_shapes, self.scene.n_pieces + 1))
        rc = lib().orc_pair_distances(self._r, self._s, _p(q), len(q), finger_open, query_distance, flags, _p(out))
        assert rc == 0
        return out

    def self_distances(self, q, flags=SELF, finger_open=0.04, query_distance=0.01):
        q = _f64(q).reshape(-1)
        out = np.empty(len(self.robot.self_pairs))
        rc = lib().orc_self_distances(self._r, _p(q), len(q), finger_open, query_distance, flags, _p(out))
        assert rc == 0
        return out

    def check_poses(self, poses, links, ee_link=None, flags=SCENE | PLANE, finger_open=0.04, threshold=-0.001,
                    query_distance=0.01, n_threads=0):
        poses = _f64(poses).reshape(-1, 4, 4)[:, :3, :] if np.asarray(poses).shape[-2:] == (4, 4) else _f64(poses)
        poses = np.ascontiguousarray(poses.reshape(-1, 12))
        n = len(poses)
        links = _i32(links)
        bits = np.zeros(n, np.uint8)
        dist = np.zeros(n)
        ee = self.robot.end_effector_id if ee_link is None else ee_link
        rc = lib().orc_check_poses(self._r, self._s, _p(poses), n, ee, _p(links), len(links), finger_open, threshold,
                                   query_distance, flags, _p(bits), _p(dist), n_threads)
        assert rc == 0
        return bits.astype(bool), dist

    def check_edges(self, qa, qb, substeps=64, flags=SCENE | PLANE, finger_open=0.04, threshold=-0.001,
                    query_distance=0.01, n_threads=0):
        qa, qb = _f64(np.atleast_2d(qa)), _f64(np.atleast_2d(qb))
        m = len(qa)
        bits = np.zeros(m, np.uint8)
        dist = np.zeros(m)
        rc = lib().orc_check_edges(self._r, self._s, _p(qa), _p(qb), m, qa.shape[1], substeps, flags, finger_open,
                                   threshold, query_distance, _p(bits), _p(dist), n_threads)
        assert rc == 0
        return bits.astype(bool), dist


def gjk_distance(va, vb, Ta=None, Tb=None, dmax=1e300):
    va, vb = _f64(va), _f64(vb)
    Ta = _f64(Ta)[:3, :] if Ta is not None else None
    Tb = _f64(Tb)[:3, :] if Tb is not None else None
    Ta = np.ascontiguousarray(Ta) if Ta is not None else None
    Tb = np.ascontiguousarray(Tb) if Tb is not None else None
    it = C.c_int(0)
    d = lib().orc_gjk_distance(_p(va), len(va), _p(Ta), _p(vb), len(vb), _p(Tb), dmax, C.byref(it))
    return d, it.value


def epa_depth(va, vb, Ta=None, Tb=None):
    va, vb = _f64(va), _f64(vb)
    Ta = np.ascontiguousarray(_f64(Ta)[:3, :]) if Ta is not None else None
    Tb = np.ascontiguousarray(_f64(Tb)[:3, :]) if Tb is not None else None
    it = C.c_int(0)
    d = lib().orc_epa_depth(_p(va), len(va), _p(Ta), _p(vb), len(vb), _p(Tb), C.byref(it))
    return d, it.value
