"""fp32 device numerics without a GPU: the per-thread GJK / EPA device code (jg_gjk.cuh, jg_epa.cuh) is compiled for
the host with g++ (JG_HOST_EMU shim, tests/emu/gjk_emu.cpp) and compared with the double-precision oracle on real
scenario geometry.  This is what lets kernel-numerics changes be checked before any GPU time is spent."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

import oracle
from oracle import Oracle
from conftest import ROOT
from ref_numpy import uniform_configs

EMU_SRC = os.path.join(ROOT, 'tests', 'emu', 'gjk_emu.cpp')
EMU_LIB = os.path.join(ROOT, 'tests', 'emu', '_gjk_emu.so')
CSRC = os.path.join(ROOT, 'jogramop_framework_b200', 'csrc')


@pytest.fixture(scope='module')
def emu():
    deps = [EMU_SRC] + [os.path.join(CSRC, f) for f in ('jg_math.cuh', 'jg_gjk.cuh', 'jg_epa.cuh')]
    if not os.path.exists(EMU_LIB) or any(os.path.getmtime(d) > os.path.getmtime(EMU_LIB) for d in deps):
        subprocess.check_call(['g++', '-O2', '-fPIC', '-shared', '-ffp-contract=off', '-o', EMU_LIB, EMU_SRC])
    return C.CDLL(EMU_LIB)


def vp(a):
    return a.ctypes.data_as(C.c_void_p)


@pytest.mark.parametrize('sid,small', [(11, 0), (34, 0), (45, 0), (11, 1), (21, 1)])
def test_fp32_gjk_and_epa_match_oracle(sid, small, emu, robot, scene_loader):
    """small=1: the shared-memory storage class of the fast kernel (22 vertices / 40 faces) + overflow fallback."""
    emu.emu_set_small_storage(small)
    emu.emu_overflows.restype = C.c_long
    sc = scene_loader(sid)
    o = Oracle(robot, sc)
    n = 700
    q = uniform_configs(robot, n, sid)
    fk = o.fk(q.astype(np.float64))
    pd = np.stack([o.pair_distances(q[i].astype(np.float64), flags=oracle.SCENE) for i in range(n)])
    n_sep = n_pen = 0
    for s in range(robot.n_shapes):
        core = robot.shape_verts(s).astype(np.float32)
        epa_v = robot.shape_plane_verts(s).astype(np.float32)
        msum = robot.shape_margin[s] + 0.001
        em = robot.shape_plane_margin[s] + 0.001
        for p in range(sc.n_pieces):
            pv = sc.piece(p)
            c = (0.5 * (pv.min(0) + pv.max(0))).astype(np.float32)
            vb = (pv - c.astype(np.float64)).astype(np.float32)
            T = fk[:, robot.shape_link[s]].copy()
            T[:, :, 3] -= c
            T = np.ascontiguousarray(T.reshape(n, 12).astype(np.float32))
            ref = pd[:, s, p]
            out = [np.zeros(n, np.float32), np.zeros(n, np.int32), np.zeros(n, np.int32), np.zeros(n, np.float32),
                   np.zeros(n, np.int32)]
            # separated pairs: GJK on the cores with the query-distance early exit
            emu.emu_gjk_batch(vp(core), len(core), vp(vb), len(vb), vp(T), n, C.c_float(0.01 + msum), vp(out[0]),
                              vp(out[1]), vp(out[2]), None, None)
            sep = np.isfinite(ref) & (ref > -msum + 1e-6)
            assert (out[1][sep] == 0).all()
            assert np.abs(out[0][sep] - msum - ref[sep]).max(initial=0) < 2e-6
            far = ~np.isfinite(ref)
            d = out[0] - msum
            assert ((out[1][far] == 1) | (d[far] > 0.01 - 1e-5)).all()
            n_sep += int(sep.sum())
            # overlapping cores: GJK + EPA on the penetration geometry
            pen = np.isfinite(ref) & (ref < -msum - 1e-6)
            if pen.any():
                emu.emu_gjk_batch(vp(epa_v), len(epa_v), vp(vb), len(vb), vp(T), n, C.c_float(1e30), vp(out[0]),
                                  vp(out[1]), vp(out[2]), vp(out[3]), vp(out[4]))
                assert (out[1][pen] == 2).all()
                assert np.abs(-(out[3][pen] + em) - ref[pen]).max() < 2e-5
                n_pen += int(pen.sum())
    assert n_sep > 50 and n_pen > 300
    if small:
        assert emu.emu_overflows() < 0.02 * n_pen      # the overflow path is rare
    emu.emu_set_small_storage(0)


def test_fp32_plane_distance(emu, robot, scene_loader):
    sc = scene_loader(11)
    o = Oracle(robot, sc)
    n = 500
    q = uniform_configs(robot, n, 2)
    fk = o.fk(q.astype(np.float64))
    pd = np.stack([o.pair_distances(q[i].astype(np.float64), flags=oracle.PLANE, query_distance=10.0) for i in range(n)])
    plane = sc.plane.astype(np.float32)
    for s in range(robot.n_shapes):
        v = robot.shape_plane_verts(s).astype(np.float32)
        T = np.ascontiguousarray(fk[:, robot.shape_link[s]].reshape(n, 12).astype(np.float32))
        out = np.zeros(n, np.float32)
        emu.emu_plane_batch(vp(v), len(v), vp(T), n, vp(plane), vp(out))
        assert np.abs(out - robot.shape_plane_margin[s] - pd[:, s, sc.n_pieces]).max() < 2e-6
