"""Builds libjogramop_b200.so (CUDA kernels + C-ABI) in-tree with nvcc for sm_100a."""
import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
LIB = os.path.join(HERE, 'libjogramop_b200.so')
# same sources with -DJG_COUNT_WORK: counts executed support dot products / iterations (bench.py's audit of the nominal
# flops figure); never used for a timed or shipped path
LIB_COUNT = os.path.join(HERE, 'libjogramop_b200_count.so')
# same sources with -DJG_DEBUG_BOUNDS: device-side index checks that trap (the stand-in for compute-sanitizer, which the
# GPU pool does not offer); run the GPU suite against it with JG_B200_LIB_PATH=<this file>
LIB_BOUNDS = os.path.join(HERE, 'libjogramop_b200_bounds.so')
# same sources with -DJG_TRACE: per-warp timeline of the penetration kernel (tools/trace_penetration.py); built on demand only
LIB_TRACE = os.path.join(HERE, 'libjogramop_b200_trace.so')
SOURCES = ['jg_abi.cu', 'jg_ingest.cpp']
HEADERS = ['jg_handles.h', 'jg_math.cuh', 'jg_gjk.cuh', 'jg_epa.cuh', 'jg_lbvh.cuh', 'jg_kernels.cuh', os.path.join('..', '..', 'include', 'jogramop_b200.h')]
NVCC_FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-O3', '-lineinfo', '-std=c++17',
              '-Xcompiler', '-fPIC', '-shared']


def _nvcc():
    for cand in (os.environ.get('NVCC'), shutil.which('nvcc'), '/usr/local/cuda/bin/nvcc'):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError('nvcc not found: cannot build libjogramop_b200.so')


def is_stale(lib=None):
    lib = lib or LIB
    if not os.path.exists(lib):
        return True
    t = os.path.getmtime(lib)
    return any(os.path.getmtime(os.path.join(CSRC, f)) > t for f in SOURCES + HEADERS)


def build(force=False, verbose=False, counting=False):
    """Compile the library (or its instrumented twin) if it is missing or older than its sources; returns its path."""
    lib = LIB_COUNT if counting else LIB
    if not force and not is_stale(lib):
        return lib
    cmd = ([_nvcc()] + NVCC_FLAGS + (['-DJG_COUNT_WORK'] if counting else []) + (['-Xptxas', '-v'] if verbose else []) +
           ['-o', lib] + SOURCES)
    subprocess.check_call(cmd, cwd=CSRC)
    return lib


def build_trace(force=False):
    if force or is_stale(LIB_TRACE):
        subprocess.check_call([_nvcc()] + NVCC_FLAGS + ['-DJG_TRACE', '-o', LIB_TRACE] + SOURCES, cwd=CSRC)
    return LIB_TRACE


# PyTorch operator shim (TORCH_LIBRARY(jogramop, ...), csrc/jg_torch.cpp): a host-only translation unit over the C-ABI,
# compiled with g++ against the torch headers and linked to the in-tree library
LIB_TORCH = os.path.join(HERE, 'libjogramop_torch.so')


def build_torch_shim(force=False):
    src = os.path.join(CSRC, 'jg_torch.cpp')
    hdr = os.path.join(CSRC, '..', '..', 'include', 'jogramop_b200.h')
    if not force and os.path.exists(LIB_TORCH) and os.path.getmtime(LIB_TORCH) >= max(os.path.getmtime(src), os.path.getmtime(hdr)) \
            and os.path.getmtime(LIB_TORCH) >= os.path.getmtime(LIB):
        return LIB_TORCH
    import torch
    from torch.utils import cpp_extension as ce
    cuda_home = os.environ.get('CUDA_HOME', '/usr/local/cuda')
    cmd = (['g++', '-O2', '-std=c++17', '-fPIC', '-shared', '-o', LIB_TORCH, src,
            f'-D_GLIBCXX_USE_CXX11_ABI={int(torch.compiled_with_cxx11_abi())}'] +
           [f'-I{p}' for p in ce.include_paths()] + [f'-I{cuda_home}/include'] +
           [f'-L{p}' for p in ce.library_paths()] + ['-lc10', '-lc10_cuda', '-ltorch_cpu', '-ltorch',
            f'-L{HERE}', '-ljogramop_b200', '-Wl,-rpath,$ORIGIN'] + [f'-Wl,-rpath,{p}' for p in ce.library_paths()])
    subprocess.check_call(cmd, cwd=CSRC)
    return LIB_TORCH


def build_all(force=False):
    """Both libraries, compiled side by side when both are stale (each nvcc run takes about a minute)."""
    variants = ((LIB, []), (LIB_COUNT, ['-DJG_COUNT_WORK']), (LIB_BOUNDS, ['-DJG_DEBUG_BOUNDS']))
    procs = []
    for lib, defs in variants:
        if force or is_stale(lib):
            procs.append((lib, subprocess.Popen([_nvcc()] + NVCC_FLAGS + defs + ['-o', lib] + SOURCES, cwd=CSRC)))
    for lib, p in procs:
        if p.wait() != 0:
            raise RuntimeError(f'nvcc failed for {lib}')
    build_torch_shim(force=force)
    return LIB, LIB_COUNT


if __name__ == '__main__':
    print(build(force=True, verbose=True))
    print(build(force=True, counting=True))
    print(build_torch_shim(force=True))
