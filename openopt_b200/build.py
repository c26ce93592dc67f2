"""Build liboode.so (the C-ABI library of hand-written sm_100a kernels) in-tree with nvcc.

    python -m openopt_b200.build [--force] [-v]           # or: python openopt_b200/build.py

nvcc cross-compiles for sm_100a without a GPU.  The objective is a template parameter of every hot kernel, so
the library is built from one translation unit per registered objective (csrc/trial_obj.cu with -DOODE_OBJ=k)
plus the host logic (csrc/oode.cu), compiled in parallel and linked into one shared object.  The .so is
git-ignored but travels with gpurun.
"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
LIB = os.path.join(HERE, 'liboode.so')
N_OBJECTIVES = 6
HEADERS = ['de_device.cuh', 'de_kernels.cuh', 'de_host_kernels.cuh', 'de_trial_tma.cuh', 'de_trial_rows.cuh', 'de_rows.cuh',
           'de_launch.h', 'mt19937_cpython.h', 'ga_kernels.cuh', 'ooga_host.cuh', os.path.join('..', '..', 'include', 'oode.h'),
           os.path.join('..', '..', 'include', 'ooga.h')]
SOURCES = ['oode.cu', 'trial_obj.cu']
NVCC_FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo', '-O3', '-std=c++17', '-fmad=false',
              '-Xcompiler', '-fPIC']


def _nvcc():
    for cand in (os.environ.get('NVCC'), '/usr/local/cuda/bin/nvcc', 'nvcc'):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    return 'nvcc'


def needs_build(lib=LIB):
    if not os.path.exists(lib):
        return True
    t = os.path.getmtime(lib)
    deps = [os.path.join(CSRC, s) for s in SOURCES + HEADERS] + [os.path.abspath(__file__)]
    return any(os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False, extra=(), lib=LIB, csrc=CSRC, objdir=None):
    """Compile `csrc` into `lib`.  extra: additional nvcc flags (e.g. -D macros of an A/B build)."""
    if not force and lib == LIB and not needs_build(lib):
        return lib
    objdir = objdir or os.path.join(HERE, 'build', os.path.basename(lib) + '.o')
    os.makedirs(objdir, exist_ok=True)
    nvcc = _nvcc()
    units = [('oode', os.path.join(csrc, 'oode.cu'), [])]
    units += [('trial_obj_%d' % k, os.path.join(csrc, 'trial_obj.cu'), ['-DOODE_OBJ=%d' % k]) for k in range(N_OBJECTIVES)]

    def compile_unit(u):
        name, src, defs = u
        obj = os.path.join(objdir, name + '.o')
        cmd = [nvcc] + NVCC_FLAGS + list(extra) + defs + ['-c', src, '-o', obj]
        if verbose:
            print(' '.join(cmd), flush=True)
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError('nvcc failed for %s:\n%s\n%s' % (name, r.stdout[-4000:], r.stderr[-8000:]))
        if verbose and (r.stdout.strip() or r.stderr.strip()):
            print(r.stdout + r.stderr, flush=True)
        return obj

    with ThreadPoolExecutor(max_workers=min(len(units), os.cpu_count() or 4)) as ex:
        objs = list(ex.map(compile_unit, units))
    cmd = [nvcc, '-shared', '-gencode', 'arch=compute_100a,code=sm_100a'] + objs + ['-o', lib]
    if verbose:
        print(' '.join(cmd), flush=True)
    subprocess.check_call(cmd)
    return lib


if __name__ == '__main__':
    build(force='--force' in sys.argv or True, verbose=True, extra=['-Xptxas', '-v'] if '-v' in sys.argv else [])
    print(LIB)
