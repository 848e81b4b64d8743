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
    L.jg_robot_load.restype = vp
    L.jg_robot_load.argtypes = [C.c_char_p, C.c_char_p]
    L.jg_scene_load.restype = vp
    L.jg_scene_load.argtypes = [C.c_char_p, vp]
    L.jg_scene_load_ex.restype = vp
    L.jg_scene_load_ex.argtypes = [C.c_char_p, vp, i32]
    L.jg_scene_load_mesh.restype = vp
    L.jg_scene_load_mesh.argtypes = [C.c_char_p, vp, i32, f64]
    L.jg_robot_describe.argtypes = [vp, vp]
    L.jg_robot_link_index.argtypes = [vp, C.c_char_p]
    L.jg_scene_describe.argtypes = [vp, vp]
    L.jg_engine_create.restype = vp
    L.jg_engine_create.argtypes = [vp, vp, i32, vp]
    L.jg_engine_free.argtypes = [vp]
    L.jg_engine_device.argtypes = [vp]
    L.jg_engine_pass_size.argtypes = [vp]
    L.jg_engine_set_lanes.argtypes = [vp, C.c_int]
    L.jg_engine_lanes.argtypes = [vp]
    L.jg_engine_pass_size.restype = C.c_int64
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


class RobotDesc(C.Structure):
    _P, _D = C.POINTER(C.c_int32), C.POINTER(C.c_double)
    _fields_ = [('n_links', C.c_int), ('n_shapes', C.c_int), ('dof', C.c_int), ('n_self_pairs', C.c_int), ('ee_link', C.c_int),
                ('joint_type', _P), ('joint_parent', _P), ('joint_qidx', _P), ('joint_origin', _D), ('joint_axis', _D),
                ('q_limits', _D), ('shape_link', _P), ('shape_margin', _D), ('shape_vert_offset', _P), ('core_verts', _D),
                ('shape_plane_offset', _P), ('plane_verts', _D), ('shape_plane_margin', _D), ('self_pairs', _P)]


class SceneDesc(C.Structure):
    _P, _D = C.POINTER(C.c_int32), C.POINTER(C.c_double)
    _fields_ = [('n_pieces', C.c_int), ('piece_vert_offset', _P), ('piece_verts', _D), ('piece_margin', _D), ('piece_body', _P),
                ('n_bodies', C.c_int), ('n_target_bodies', C.c_int), ('has_plane', C.c_int), ('plane', C.c_double * 4),
                ('is_mesh', C.c_int), ('n_mesh_verts', C.c_int), ('n_mesh_tris', C.c_int),
                ('mesh_verts', C.POINTER(C.c_float)), ('mesh_tris', _P)]


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
    'jg_scene_create', 'jg_scene_create_mesh', 'jg_scene_free', 'jg_robot_load', 'jg_scene_load', 'jg_scene_load_ex', 'jg_scene_load_mesh',
    'jg_robot_describe', 'jg_robot_link_index', 'jg_scene_describe', 'jg_engine_create', 'jg_engine_free', 'jg_engine_device', 'jg_engine_pass_size', 'jg_engine_set_lanes', 'jg_engine_lanes', 'jg_fk', 'jg_check',
    'jg_check_edges', 'jg_check_poses', 'jg_check_host', 'jg_check_host_submit', 'jg_check_host_wait', 'jg_ik', 'jg_get_stats', 'jg_profile_enable', 'jg_get_profile',
    'jg_measure_fp32_peak', 'jg_work_counters', 'jg_work_counters_reset']
