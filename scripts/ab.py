#!/usr/bin/env python
"""Same-box A/B timing of library builds (how every kernel change is accepted: the pool's box-to-box variation
is larger than most single optimisations).

    python scripts/ab.py build NAME [--ref GITREF] [-D MACRO ...]   # -> openopt_b200/_ab/liboode_NAME.so
    python scripts/ab.py run NAME [NAME ...] [--shapes "ackley:1000:1000000 ackley:120:1000000"]

`build` compiles the csrc of a git ref (default: the working tree) with optional -D macros; `run` (on the GPU
box, e.g. `gpurun -- python scripts/ab.py run prev new`) times each build on each shape through
OODE_LIB_PATH with per-kernel CUDA events and prints one line per (build, shape).  The builds need the same
C ABI as the working tree's _lib.py."""
import argparse
import os
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
AB = os.path.join(ROOT, 'openopt_b200', '_ab')
FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo', '-O3', '-std=c++17', '-fmad=false', '-Xcompiler', '-fPIC',
         '-shared']

PROBE = r'''
import sys, os
sys.path.insert(0, %(root)r)
import numpy as np
from openopt_b200 import _lib, objectives
obj, D, NP = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
gens = 20
o = objectives.get(obj)
hw = 32.768 if obj == 'ackley' else 5.12
lb, ub = -hw * np.ones(D), hw * np.ones(D)
with _lib.DeviceDE(o.device_id, lb, ub, x0=lb + 0.25 * (ub - lb), population=NP, obj_aux=o.aux(D), max_batch=64) as dev:
    dev.init(); dev.step(3)
    dev.set_profiling(True)
    c0 = dev.counters(); dev.step(gens); c1 = dev.counters()
    name = dev.trial_kernel_name()
t = (c1['trial_kernel_ms'] - c0['trial_kernel_ms']) / gens
s = (c1['select_kernel_ms'] - c0['select_kernel_ms']) / gens
print('%%-12s %%-10s D=%%5d NP=%%8d  trial %%8.4f ms  select %%7.4f ms  %%6.1f GB/s  %%s' %% (
    os.environ.get('AB_NAME', '?'), obj, D, NP, t, s, NP * (40 * D + 16) / t / 1e6, name))
'''


def build(args):
    os.makedirs(AB, exist_ok=True)
    out = os.path.join(AB, 'liboode_%s.so' % args.name)
    if args.ref:
        tmp = tempfile.mkdtemp(prefix='oode_ab_')
        tar = subprocess.run(['git', '-C', ROOT, 'archive', args.ref, 'openopt_b200/csrc', 'include'], check=True,
                             capture_output=True).stdout
        subprocess.run(['tar', '-x', '-C', tmp], input=tar, check=True)
        src = os.path.join(tmp, 'openopt_b200', 'csrc', 'oode.cu')
    else:
        src = os.path.join(ROOT, 'openopt_b200', 'csrc', 'oode.cu')
    if os.path.exists(os.path.join(os.path.dirname(src), 'trial_obj.cu')):
        # one translation unit per objective, compiled in parallel (openopt_b200/build.py)
        sys.path.insert(0, ROOT)
        from openopt_b200 import build as b
        b.build(force=True, verbose=False, extra=['-D' + m for m in args.D], lib=out, csrc=os.path.dirname(src),
                objdir=os.path.join(AB, 'obj_' + args.name))
    else:
        cmd = ['nvcc'] + FLAGS + ['-D' + m for m in args.D] + [src, '-o', out]
        print(' '.join(cmd))
        subprocess.check_call(cmd)
    print(out)


def run(args):
    probe = os.path.join(tempfile.mkdtemp(prefix='oode_ab_'), 'probe.py')
    open(probe, 'w').write(PROBE % {'root': ROOT})
    for rep in range(args.repeat):
        for name in args.names:
            lib = os.path.join(ROOT, 'openopt_b200', 'liboode.so') if name == 'tree' else os.path.join(AB, 'liboode_%s.so' % name)
            for shape in args.shapes.split():
                obj, D, NP = shape.split(':')
                subprocess.run([sys.executable, probe, obj, D, NP], env=dict(os.environ, OODE_LIB_PATH=lib, AB_NAME=name + ''.join(' ' + e for e in args.env),
                                        **dict(e.split('=', 1) for e in args.env)), check=False)


def main():
    ap = argparse.ArgumentParser()
    sub = ap.add_subparsers(dest='cmd', required=True)
    b = sub.add_parser('build')
    b.add_argument('name')
    b.add_argument('--ref', default=None)
    b.add_argument('-D', action='append', default=[])
    r = sub.add_parser('run')
    r.add_argument('names', nargs='+', help="build names under openopt_b200/_ab/, or 'tree' for the in-tree library")
    r.add_argument('--shapes', default='ackley:1000:1000000 ackley:120:1000000 ackley:128:1000000 rastrigin:100:4194304')
    r.add_argument('--env', action='append', default=[], help='NAME=VALUE for the probe processes (e.g. OODE_TRIAL_KERNEL=rows2)')
    r.add_argument('--repeat', type=int, default=2)
    a = ap.parse_args()
    build(a) if a.cmd == 'build' else run(a)


if __name__ == '__main__':
    main()
