"""Stand-ins for the parts of the reference's ``util.py`` and of burg_toolkit's ``util`` that the hot path touches.

Reference: util.py:9-10 (SCENARIO_DIR / SCENARIO_IDS), util.py:13-49 (Timer), util.py:79-107 (quaternion helpers);
burg_toolkit.util.position_and_quaternion_from_tf / tf_from_pos_quat (pybullet convention: quaternion xyzw).
No pybullet / numpy-quaternion / burg imports.
"""
import time

import numpy as np

SCENARIO_DIR = 'scenarios'
SCENARIO_IDS = [11, 12, 13, 14, 15, 21, 22, 23, 24, 25, 31, 32, 33, 34, 35, 41, 42, 43, 44, 45]


class Timer:
    """Interface of the reference's ``util.Timer`` (util.py:13-49; used by ``Scenario.get_ik_solutions``,
    scenario.py:94-127): ``start(key)`` / ``stop(key)`` accumulate elapsed seconds per key over any number of
    intervals, ``count(key)`` counts occurrences, ``print()`` writes the summary in the reference's layout.
    Own implementation: elapsed time is kept separately from the start stamp of a running interval (monotonic
    ``perf_counter``), so ``timers`` always holds finished seconds; ``count`` takes an increment for batched callers."""

    _RULE = 33

    def __init__(self):
        self.timers = {}      # key -> accumulated seconds of finished intervals
        self.counters = {}    # key -> occurrences
        self._running = {}    # key -> perf_counter() stamp of the open interval

    def start(self, key):
        self.timers.setdefault(key, 0.0)
        self._running[key] = time.perf_counter()

    def stop(self, key):
        now = time.perf_counter()
        if key not in self._running:
            raise ValueError('attempting to stop timer that has not been started')
        self.timers[key] += now - self._running.pop(key)

    def count(self, key, n=1):
        self.counters[key] = self.counters.get(key, 0) + n

    def summary(self):
        """-> {'timings': {key: seconds}, 'counters': {key: n}} (what ``print`` shows)."""
        return {'timings': dict(self.timers), 'counters': dict(self.counters)}

    def print(self):
        for title, rows, unit, fmt in ((' TIMINGS ', self.timers, 's', '{:.2f}'), (' COUNTERS ', self.counters, 'x', '{}')):
            print(title.center(self._RULE, '*'))
            for key, val in rows.items():
                print('\t' + key + ':\t' + fmt.format(val) + unit)
        print('*' * self._RULE)


def quaternion_from_rotation_matrix(rot_mat):
    """-> [x, y, z, w] (util.py:91-93)."""
    m = np.asarray(rot_mat, dtype=np.float64)
    t = np.trace(m)
    if t > 0:
        s = np.sqrt(t + 1.0) * 2
        w, x, y, z = 0.25 * s, (m[2, 1] - m[1, 2]) / s, (m[0, 2] - m[2, 0]) / s, (m[1, 0] - m[0, 1]) / s
    elif m[0, 0] > m[1, 1] and m[0, 0] > m[2, 2]:
        s = np.sqrt(1.0 + m[0, 0] - m[1, 1] - m[2, 2]) * 2
        w, x, y, z = (m[2, 1] - m[1, 2]) / s, 0.25 * s, (m[0, 1] + m[1, 0]) / s, (m[0, 2] + m[2, 0]) / s
    elif m[1, 1] > m[2, 2]:
        s = np.sqrt(1.0 + m[1, 1] - m[0, 0] - m[2, 2]) * 2
        w, x, y, z = (m[0, 2] - m[2, 0]) / s, (m[0, 1] + m[1, 0]) / s, 0.25 * s, (m[1, 2] + m[2, 1]) / s
    else:
        s = np.sqrt(1.0 + m[2, 2] - m[0, 0] - m[1, 1]) * 2
        w, x, y, z = (m[1, 0] - m[0, 1]) / s, (m[0, 2] + m[2, 0]) / s, (m[1, 2] + m[2, 1]) / s, 0.25 * s
    return np.asarray([x, y, z, w])


def rotation_matrix_from_quaternion(q):
    x, y, z, w = np.asarray(q, dtype=np.float64) / np.linalg.norm(q)
    return np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - z * w), 2 * (x * z + y * w)],
                     [2 * (x * y + z * w), 1 - 2 * (x * x + z * z), 2 * (y * z - x * w)],
                     [2 * (x * z - y * w), 2 * (y * z + x * w), 1 - 2 * (x * x + y * y)]])


def position_and_quaternion_from_tf(tf, convention='pybullet'):
    """burg.util.position_and_quaternion_from_tf: 4x4 -> (pos[3], quat xyzw)."""
    assert convention == 'pybullet', 'only the pybullet (xyzw) convention is supported'
    tf = np.asarray(tf, dtype=np.float64)
    return tf[:3, 3].copy(), quaternion_from_rotation_matrix(tf[:3, :3])


def tf_from_pos_quat(pos=None, quat=None, convention='pybullet'):
    """burg.util.tf_from_pos_quat: (pos, quat xyzw) -> 4x4."""
    assert convention == 'pybullet', 'only the pybullet (xyzw) convention is supported'
    tf = np.eye(4)
    if pos is not None:
        tf[:3, 3] = pos
    if quat is not None:
        tf[:3, :3] = rotation_matrix_from_quaternion(quat)
    return tf


def angle_between_quaternions(q1, q2, as_degree=False):
    """util.py:79-88 (pybullet.getDifferenceQuaternion): angle of q1^-1 * q2, quaternions xyzw."""
    q1 = np.asarray(q1, dtype=np.float64) / np.linalg.norm(q1)
    q2 = np.asarray(q2, dtype=np.float64) / np.linalg.norm(q2)
    angle = 2 * np.arccos(np.clip(np.abs(np.dot(q1, q2)), 0, 1))
    return np.rad2deg(angle) if as_degree else angle


def get_translation_and_angle(pos1, orn1, pos2, orn2):
    """util.py:96-107."""
    return np.linalg.norm(np.asarray(pos2) - np.asarray(pos1)), angle_between_quaternions(orn1, orn2, as_degree=True)
