"""Build liboode.so (the C-ABI library of hand-written sm_100a kernels) in-tree with nvcc.

    python -m openopt_b200.build            # or: python openopt_b200/build.py

nvcc cross-compiles for sm_100a without a GPU.  The .so is git-ignored but travels with gpurun.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
LIB = os.path.join(HERE, 'liboode.so')
SOURCES = ['oode.cu']
HEADERS = ['de_device.cuh', 'de_kernels.cuh', 'de_trial_tma.cuh', 'de_rows.cuh', 'mt19937_cpython.h',
           os.path.join('..', '..', 'include', 'oode.h')]
NVCC_FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo', '-O3', '-std=c++17', '-fmad=false',
              '-Xcompiler', '-fPIC', '-shared']


def _nvcc():
    for cand in (os.environ.get('NVCC'), '/usr/local/cuda/bin/nvcc', 'nvcc'):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    return 'nvcc'


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, s) for s in SOURCES + HEADERS] + [os.path.abspath(__file__)]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False, extra=()):
    if not force and not needs_build():
        return LIB
    cmd = [_nvcc()] + NVCC_FLAGS + list(extra) + [os.path.join(CSRC, s) for s in SOURCES] + ['-o', LIB]
    if verbose:
        print(' '.join(cmd))
    subprocess.check_call(cmd)
    return LIB


if __name__ == '__main__':
    build(force='--force' in sys.argv, verbose=True, extra=['-Xptxas', '-v'] if '-v' in sys.argv else [])
    print(LIB)
