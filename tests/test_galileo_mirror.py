"""The galileo plugin of the host mirror side by side with the UNMODIFIED reference solver
(openopt/solvers/Standalone/galileo_oo.py behind oracle/galileo_ref_shim.py): result fields, stop handling,
evaluation counts and the printed iteration log must agree.

CPU only and only where /root/reference is mounted.  Both sides run in one subprocess; the device handle is replaced
by the oracle-backed stand-in (the kernels themselves are covered by tests/test_gpu_galileo.py)."""
import os
import subprocess
import sys

import pytest

from conftest import ROOT

REF = '/root/reference'
pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, 'openopt')), reason='reference tree not mounted')

WORKER = r'''
import sys, os, io, warnings, contextlib
warnings.simplefilter('ignore')
sys.path[:0] = [%(root)r, os.path.join(%(root)r, 'oracle'), os.path.join(%(root)r, 'tests')]
import numpy as np
from galileo_ref_shim import load_galileo
oo, g = load_galileo(150880)
import helpers
import openopt_b200 as mine
from openopt_b200 import _lib, objectives
_lib.DeviceGalileo = helpers.OracleBackedGalileo

def text(out):
    return [l for l in out.splitlines() if 'Time Elapsed' not in l]

def both(make, solve_kw, fields=('istop', 'msg', 'stopcase', 'ff', 'isFeasible', 'rf'), solver=lambda L: 'galileo'):
    res = []
    for lib, extra in ((oo, {}), (mine, {'rng': 'cpython', 'seed': 150880})):
        buf = io.StringIO()
        with contextlib.redirect_stdout(buf):
            p = make(lib)
            r = p.solve(solver(lib), **dict(solve_kw, **extra))
        res.append((r, text(buf.getvalue())))
    (a, ta), (b, tb) = res
    for f in fields:
        assert getattr(a, f) == getattr(b, f), (f, getattr(a, f), getattr(b, f))
    assert np.array_equal(a.xf, b.xf), (a.xf, b.xf)
    assert list(a.iterValues.f) == list(b.iterValues.f)
    assert dict(a.evals) == dict(b.evals), (a.evals, b.evals)
    assert ta == tb, '\n'.join(ta) + '\n----\n' + '\n'.join(tb)
    return a

D = 8
lb, ub = -5.12 * np.ones(D), np.linspace(3.0, 5.12, D)
ras = objectives.Rastrigin()
# default parameters and stop rules (GLP's maxNonSuccess fires), text output on
r = both(lambda L: L.GLP(ras, lb=lb, ub=ub, maxIter=1e3, maxFunEvals=1e5, iprint=10), dict(plot=0))
assert r.istop == 11
# maxIter, non-default rates (aliased entries), log every 5 iterations
r = both(lambda L: L.GLP(ras, lb=lb, ub=ub, maxIter=40, maxFunEvals=10 ** 9, maxNonSuccess=10 ** 9, iprint=5),
         dict(population=40, crossoverRate=0.6, mutationRate=0.1))
assert r.istop == -7
# goal='max', fEnough, maxFunEvals, a callback stop, NLP with a positional x0
both(lambda L: L.GLP(objectives.Rosenbrock(), lb=-2 * np.ones(6), ub=2 * np.ones(6), goal='max', maxIter=30, maxFunEvals=10 ** 9, iprint=10),
     dict(population=25))
r = both(lambda L: L.GLP(objectives.Sphere(), lb=-np.ones(5), ub=2 * np.ones(5), fEnough=0.3, maxFunEvals=10 ** 9, maxNonSuccess=10 ** 9, iprint=1),
         dict(population=30))
assert r.istop == 10
r = both(lambda L: L.GLP(ras, lb=lb, ub=ub, maxIter=10 ** 4, maxFunEvals=300, maxNonSuccess=10 ** 9, iprint=-1), dict(population=21))
assert r.istop == -10
calls = {}
def make_cb(L):
    calls[L] = 0
    def cb(p):
        calls[L] += 1
        return calls[L] >= 7
    return L.GLP(objectives.Sphere(), lb=-np.ones(4), ub=2 * np.ones(4), callback=cb, maxNonSuccess=10 ** 9, iprint=1)
r = both(make_cb, dict(population=20))
assert r.istop == 80
both(lambda L: L.NLP(objectives.Sphere(), 0.5 * np.ones(5), lb=-np.ones(5), ub=np.ones(5), maxIter=25, maxFunEvals=10 ** 9, iprint=-1),
     dict(population=30))
# the reference's own galileo call site (openopt/doc/solverParams.py:8-14), arguments as written there
glp1 = helpers.device_objective('glp1')
r = both(lambda L: L.GLP(glp1, lb=-np.ones(4), ub=np.ones(4), maxIter=1e3, maxCPUTime=3, maxFunEvals=1e5, fEnough=80, iprint=10),
         dict(crossoverRate=0.80, maxTime=3, population=15, mutationRate=0.15))
assert r.istop == 11
# parameters pre-bound with oosolver (kernel/nonOptMisc.py:132-160), as examples/oosolver.py does
both(lambda L: L.GLP(ras, lb=lb, ub=ub, maxIter=30, maxFunEvals=10 ** 9, maxNonSuccess=10 ** 9, iprint=-1), dict(population=18),
     solver=lambda L: L.oosolver('galileo', crossoverRate=0.7, mutationRate=0.2))
print('mirror == reference')
'''


def test_galileo_mirror_matches_reference_side_by_side(tmp_path):
    script = tmp_path / 'worker.py'
    script.write_text(WORKER % {'root': ROOT})
    out = subprocess.run([sys.executable, str(script)], capture_output=True, text=True, timeout=900)
    assert out.returncode == 0, out.stdout[-3000:] + out.stderr[-5000:]
    assert 'mirror == reference' in out.stdout
