"""The C-ABI library loads on a machine without a GPU and exports every symbol include/jogramop_b200.h declares
(no compute calls here); without a device the engine fails loudly instead of falling back to a CPU path."""
import ctypes
import os
import re

import numpy as np
import pytest

from conftest import ROOT


def declared_symbols():
    src = open(os.path.join(ROOT, 'include', 'jogramop_b200.h')).read()
    src = re.sub(r'/\*.*?\*/', '', src, flags=re.S)
    return sorted(set(re.findall(r'\b(jg_[a-z0-9_]+)\s*\(', src)))


def test_library_exports_every_declared_symbol():
    from jogramop_framework_b200 import _lib
    L = _lib.load()
    syms = declared_symbols()
    assert len(syms) >= 18
    for s in syms:
        assert hasattr(L, s), f'{s} declared in include/jogramop_b200.h but not exported'
    assert sorted(_lib.EXPORTED_SYMBOLS) == syms
    assert L.jg_abi_version() == 2


def test_default_params_and_error_channel():
    from jogramop_framework_b200 import _lib
    L = _lib.load()
    p = _lib.Params()
    L.jg_default_params(ctypes.byref(p))
    assert p.threshold == pytest.approx(-0.001) and p.query_distance == pytest.approx(0.01)
    assert p.finger_open == pytest.approx(0.04) and p.chunk == 0
    # invalid robot tables are rejected with a message, no crash
    assert L.jg_robot_create(0, None, None, None, None, None, None, 0, 0, None, None, None, None, None, None, None, 0, None) is None
    assert b'jg_robot_create' in L.jg_last_error()


def test_no_cpu_fallback(robot):
    """On a box without a CUDA device the product path must raise, not compute on the CPU."""
    import torch
    if torch.cuda.is_available():
        pytest.skip('a GPU is present')
    from jogramop_framework_b200._lib import JogramopError
    from jogramop_framework_b200.engine import CollisionEngine
    with pytest.raises(JogramopError):
        CollisionEngine(robot, None)
    from jogramop_framework_b200.scenario import Scenario
    rob, sim = Scenario(11).get_robot_and_sim()
    with pytest.raises(JogramopError):
        rob.in_collision()


def test_product_package_never_imports_the_oracle():
    """oracle/ is test infrastructure: nothing under jogramop_framework_b200/ may import or load it."""
    pkg = os.path.join(ROOT, 'jogramop_framework_b200')
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(('.py', '.cu', '.cuh', '.h')):
                txt = open(os.path.join(dirpath, f)).read()
                assert 'import oracle' not in txt and 'from oracle' not in txt and 'liboracle' not in txt, f


def test_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` (the CPU arm the driver runs beside the GPU arm): one JSON line with the same metric /
    unit as the GPU arm, impl = reference, a cpu_baseline block of kind "port" and a zero-copy e2e block."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, 'bench.py'), '--impl', 'reference', '--steps', '1', '--warmup', '1',
                          '--ref-sample', '2000'], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stderr
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line['impl'] == 'reference' and line['metric'] == 'configuration collision checks/sec' and line['unit'] == 'checks/s'
    assert line['higher_is_better'] is True and line['value'] > 0 and line['steps'] == 1
    assert line['cpu_baseline']['kind'] == 'port' and line['cpu_baseline']['cores'] >= 1 and line['cpu_baseline']['value'] == line['value']
    assert line['e2e'] == {'value': line['value'], 'unit': 'checks/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}
