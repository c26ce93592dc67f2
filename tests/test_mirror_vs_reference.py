"""The host mirror (openopt_b200.GLP / NLP + the plugin) side by side with the UNMODIFIED reference on the
same problems: result fields, stop handling, evaluation counts and the printed iteration log must agree.

CPU only and only where /root/reference is mounted (the build container).  Both sides run in one
subprocess: the reference's `p.solve('de')` and the mirror's `p.solve('de', rng='cpython')` with the device
handle replaced by the oracle-backed stand-in (the kernels themselves are covered by the GPU tests)."""
import os
import subprocess
import sys

import pytest

from conftest import ROOT

REF = '/root/reference'
pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, 'openopt')), reason='reference tree not mounted')

WORKER = r'''
import sys, os, io, re, warnings, contextlib
warnings.simplefilter('ignore')
sys.path[:0] = [%(root)r, os.path.join(%(root)r, 'oracle'), os.path.join(%(root)r, 'tests')]
import numpy as np
from ref_shim import load_reference
oo, de_mod = load_reference()
import helpers
import openopt_b200 as mine
from openopt_b200 import _lib, objectives
from openopt_b200.constraints import DiagQuadratic
_lib.DeviceDE = helpers.OracleBackedDE

def text(out):            # the iteration log without the timing line
    return [l for l in out.splitlines() if 'Time Elapsed' not in l]

def both(make, solve_kw, fields=('istop', 'msg', 'stopcase', 'ff', 'isFeasible', 'rf')):
    res = []
    for lib, extra in ((oo, {}), (mine, {'rng': 'cpython'})):
        buf = io.StringIO()
        with contextlib.redirect_stdout(buf):
            p = make(lib)
            r = p.solve('de', **dict(solve_kw, **extra))
        res.append((r, text(buf.getvalue())))
    (a, ta), (b, tb) = res
    for f in fields:
        assert getattr(a, f) == getattr(b, f), (f, getattr(a, f), getattr(b, f))
    assert np.array_equal(a.xf, b.xf)
    assert list(a.iterValues.f) == list(b.iterValues.f)
    assert dict(a.evals) == dict(b.evals), (a.evals, b.evals)
    assert ta == tb, '\n'.join(ta) + '\n----\n' + '\n'.join(tb)
    return a

D = 8
lb, ub = -5.12 * np.ones(D), np.linspace(3.0, 5.12, D)
ras = objectives.Rastrigin()
# 1. the reference's own test shape (tests/rastrigin.py): default x0, default stop rules, text output on
r = both(lambda L: L.GLP(ras, lb=-5.12 * np.ones(D), ub=5.12 * np.ones(D), maxIter=1e3, maxFunEvals=1e5, iprint=10), dict(plot=0))
assert r.istop == 11
# 2. maxIter / maxFunEvals / fEnough stops, with the iteration log every 5 iterations
both(lambda L: L.GLP(ras, lb=lb, ub=ub, x0=lb + 1.0, maxIter=40, maxFunEvals=10 ** 9, maxNonSuccess=10 ** 9, iprint=5), dict(population=40))
r = both(lambda L: L.GLP(ras, lb=lb, ub=ub, x0=lb + 1.0, maxIter=10 ** 4, maxFunEvals=2000, maxNonSuccess=10 ** 9, iprint=-1), dict(population=40))
assert r.istop == -10 or 'maxFunEvals' in r.msg.lower() or r.istop < 0
both(lambda L: L.GLP(objectives.Sphere(), lb=-np.ones(5), ub=2 * np.ones(5), fEnough=1e-2, maxFunEvals=10 ** 9, maxNonSuccess=10 ** 9, iprint=1),
     dict(population=30))
# 3. goal='max', NLP with a positional x0, callbacks
both(lambda L: L.GLP(objectives.Rosenbrock(), lb=-2 * np.ones(6), ub=2 * np.ones(6), goal='max', maxIter=30, maxFunEvals=10 ** 9, iprint=10),
     dict(population=25))
both(lambda L: L.NLP(objectives.Sphere(), 0.5 * np.ones(5), lb=-np.ones(5), ub=np.ones(5), maxIter=25, maxFunEvals=10 ** 9, iprint=-1),
     dict(population=30))
calls = {}
def make_cb(L):
    calls[L] = 0
    def cb(p):
        calls[L] += 1
        return calls[L] >= 7
    return L.GLP(objectives.Sphere(), lb=-np.ones(4), ub=2 * np.ones(4), callback=cb, maxNonSuccess=10 ** 9, iprint=1)
r = both(make_cb, dict(population=20))
assert r.istop == 80
# 4. non-default strategies and general constraints (linear + registered nonlinear), with the feasibility column
both(lambda L: L.GLP(ras, lb=lb, ub=ub, x0=lb + 1.0, maxIter=30, maxFunEvals=10 ** 9, maxNonSuccess=10 ** 9, iprint=10),
     dict(population=40, baseVectorStrategy='best', searchDirectionStrategy='best', differenceFactorStrategy='constant', hndvi=2))
A = np.zeros((2, 4)); A[0, 0] = A[0, 3] = 1.0; A[1, 1] = A[1, 3] = 1.0
c = DiagQuadratic([[1, 0, 1, 0], [1.5, 1, 0, 0]], [0.15, 1.5])
both(lambda L: L.GLP(objectives.Rosenbrock(), lb=-np.ones(4), ub=np.ones(4), A=A, b=[0.15, 1.5], c=c, maxIter=50, maxFunEvals=10 ** 9,
                     maxNonSuccess=10 ** 9, iprint=10), dict(population=40))
# 5. examples/glp_Ab_c.py as written there: plot=1 without pylab (graphics are turned off with a warning) and a
#    parameter the solver does not have (mutationRate)
both(lambda L: L.GLP(objectives.Rosenbrock(), lb=-np.ones(4), ub=np.ones(4), A=A, b=[0.15, 1.5], c=c, maxIter=25, maxFunEvals=1e5,
                     maxTime=30, maxCPUTime=30), dict(mutationRate=0.15, plot=1))
print('mirror == reference')
'''


def test_mirror_matches_reference_side_by_side(tmp_path):
    script = tmp_path / 'worker.py'
    script.write_text(WORKER % {'root': ROOT})
    out = subprocess.run([sys.executable, str(script)], capture_output=True, text=True, timeout=900)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-5000:]
    assert 'mirror == reference' in out.stdout
