"""``torch.ops.jogramop.*``: the PyTorch operator shim over the C-ABI (csrc/jg_torch.cpp, TORCH_LIBRARY(jogramop, ...)).

    from jogramop_framework_b200 import torch_ops
    ops = torch_ops.load()                                    # torch.ops.jogramop
    bits, dist, reason = ops.check(engine.handle, q_cuda, SCENE | PLANE, True, False)
    h = ops.engine_from_files('robots/franka_panda/mobile_panda.urdf', 'scenarios/011/scene.yaml', 0)

The operators take the current CUDA stream of the input's device, allocate their outputs with ATen and forward to the
``jg_*`` entry points; ``engine`` is an int64 handle (``CollisionEngine.handle`` or ``engine_from_files``).
"""
import os

import torch

from . import build as _build
from ._lib import JogramopError

_loaded = False


def load():
    global _loaded
    if not _loaded:
        if not os.path.exists(_build.LIB_TORCH):
            raise JogramopError(f'{_build.LIB_TORCH} is missing: build it with `python -m jogramop_framework_b200.build`')
        torch.ops.load_library(_build.LIB_TORCH)
        _loaded = True
    return torch.ops.jogramop
