"""ctypes binding of libjogramop_b200.so (the C-ABI in include/jogramop_b200.h).

There is no CPU fallback: if the CUDA library is missing or no GPU is present the product path raises.
"""
import ctypes as C
import os

from . import build as _build

_lib = None


class JogramopError(RuntimeError):
    """Raised for every non-zero status of the C-ABI (pybullet.error -> RuntimeError)."""


def load():
    global _lib
    if _lib is not None:
        return _lib
    # JG_B200_LIB_VARIANT=count selects the instrumented twin (work counters; bench.py's flops audit runs it in a
    # subprocess).  Everything else uses the shipped library.
    path = _build.LIB_COUNT if os.environ.get('JG_B200_LIB_VARIANT') == 'count' else _build.LIB
    path = os.environ.get('JG_B200_LIB_PATH', path)   # development: an experimental build of the same sources
    if not os.path.exists(path):
        raise JogramopError(
            f'{path} is missing: build it with `python -m jogramop_framework_b200.build` '
            '(nvcc, sm_100a). The collision engine has no CPU fallback.')
    L = C.CDLL(path)
    vp, i32, i64, u32, f64 = C.c_void_p, C.c_int, C.c_int64, C.c_uint32, C.c_double
    L.jg_last_error.restype = C.c_char_p
    L.jg_abi_version.restype = i32
    L.jg_device_count.restype = i32
    L.jg_default_params.argtypes = [vp]
    L.jg_robot_create.restype = vp
    L.jg_robot_create.argtypes = [i32, vp, vp, vp, vp, vp, vp, i32, i32, vp, vp, vp, vp, vp, vp, vp, i32, vp]
    L.jg_robot_free.argtypes = [vp]
    L.jg_scene_create.restype = vp
    L.jg_scene_create.argtypes = [i32, vp, vp, vp, vp]
    L.jg_scene_create_mesh.restype = vp
    L.jg_scene_create_mesh.argtypes = [i32, vp, i32, vp, f64, vp]
    L.jg_scene_free.argtypes = [vp]
    L.jg_engine_create.restype = vp
    L.jg_engine_create.argtypes = [vp, vp, i32, vp]
    L.jg_engine_free.argtypes = [vp]
    L.jg_engine_device.argtypes = [vp]
    L.jg_fk.argtypes = [vp, vp, i64, i32, vp, i32, vp, vp, vp]
    L.jg_check.argtypes = [vp, vp, i64, i32, u32, vp, vp, vp, vp]
    L.jg_check_edges.argtypes = [vp, vp, vp, i64, i32, i32, u32, vp, vp, vp]
    L.jg_check_poses.argtypes = [vp, vp, i64, i32, i32, vp, i32, u32, vp, vp, vp]
    L.jg_check_host.argtypes = [vp, vp, i64, i32, u32, vp, vp, vp]
    L.jg_check_host_submit.argtypes = [vp, vp, i64, i32, u32, vp, vp, vp, vp]
    L.jg_check_host_wait.argtypes = [vp, i64]
    L.jg_ik.argtypes = [vp, vp, i64, i32, vp, vp, i32, C.c_float, C.c_float, C.c_float, i32, vp, vp, vp]
    L.jg_get_stats.argtypes = [vp, vp]
    L.jg_profile_enable.argtypes = [vp, i32]
    L.jg_get_profile.argtypes = [vp, vp, vp]
    L.jg_measure_fp32_peak.argtypes = [i32, i32, vp]
    L.jg_work_counters.argtypes = [i32, vp]
    L.jg_work_counters_reset.argtypes = [i32]
    _lib = L
    return L


class Params(C.Structure):
    _fields_ = [('threshold', C.c_float), ('query_distance', C.c_float), ('finger_open', C.c_float),
                ('chunk', C.c_int32)]


class Stats(C.Structure):
    _fields_ = [('configs', C.c_int64), ('pairs', C.c_int64), ('plane_pairs', C.c_int64), ('self_pairs', C.c_int64),
                ('epa_pairs', C.c_int64), ('gjk_run', C.c_int64), ('epa_run', C.c_int64), ('kernel_launches', C.c_int64),
                ('overflow_run', C.c_int64)]


def check(rc, what=''):
    if rc != 0:
        msg = load().jg_last_error()
        raise JogramopError(f'{what}: {msg.decode() if msg else rc}')


EXPORTED_SYMBOLS = [
    'jg_default_params', 'jg_last_error', 'jg_abi_version', 'jg_device_count', 'jg_robot_create', 'jg_robot_free',
    'jg_scene_create', 'jg_scene_create_mesh', 'jg_scene_free', 'jg_engine_create', 'jg_engine_free', 'jg_engine_device', 'jg_fk', 'jg_check',
    'jg_check_edges', 'jg_check_poses', 'jg_check_host', 'jg_check_host_submit', 'jg_check_host_wait', 'jg_ik', 'jg_get_stats', 'jg_profile_enable', 'jg_get_profile',
    'jg_measure_fp32_peak', 'jg_work_counters', 'jg_work_counters_reset']
