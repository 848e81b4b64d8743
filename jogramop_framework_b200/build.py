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
SOURCES = ['jg_abi.cu']
HEADERS = ['jg_math.cuh', 'jg_gjk.cuh', 'jg_epa.cuh', 'jg_lbvh.cuh', 'jg_kernels.cuh', os.path.join('..', '..', 'include', 'jogramop_b200.h')]
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


def build_all(force=False):
    """Both libraries, compiled side by side when both are stale (each nvcc run takes about a minute)."""
    stale = [c for c in (False, True) if force or is_stale(LIB_COUNT if c else LIB)]
    procs = []
    for c in stale:
        lib = LIB_COUNT if c else LIB
        cmd = [_nvcc()] + NVCC_FLAGS + (['-DJG_COUNT_WORK'] if c else []) + ['-o', lib] + SOURCES
        procs.append((lib, subprocess.Popen(cmd, cwd=CSRC)))
    for lib, p in procs:
        if p.wait() != 0:
            raise RuntimeError(f'nvcc failed for {lib}')
    return LIB, LIB_COUNT


if __name__ == '__main__':
    print(build(force=True, verbose=True))
    print(build(force=True, counting=True))
