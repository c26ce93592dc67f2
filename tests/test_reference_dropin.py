"""The drop-in claim of INTEGRATION.md, checked against the REFERENCE'S OWN kernel: the plugin stub printed in
INTEGRATION.md is written into a scratch copy of the reference tree (openopt/solvers/cuda/de_cuda_oo.py), the
reference's solver discovery finds it, and `p.solve('de_cuda', rng='cpython')` driven by the reference's
GLP / runProbSolver / ooIter gives the same result object as the stock `p.solve('de')`.

CPU only and only where /root/reference is mounted (the build container): the device handle is replaced by the
oracle-backed stand-in, so this exercises the plugin contract and the host protocol, not the kernels (those are
covered by tests/test_gpu_parity.py)."""
import os
import re
import shutil
import subprocess
import sys

import pytest

from conftest import ROOT

REF = '/root/reference'
pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, 'openopt')), reason='reference tree not mounted')

WORKER = r'''
import sys, os, warnings
warnings.simplefilter('ignore')
sys.path[:0] = [%(root)r, os.path.join(%(root)r, 'oracle'), os.path.join(%(root)r, 'tests')]
import numpy as np
from ref_shim import load_reference
oo, de_mod = load_reference(%(copy)r)
import helpers
from openopt_b200 import _lib, objectives
_lib.DeviceDE = helpers.OracleBackedDE                 # no GPU here: the oracle stands in for the device handle
lb, ub = -5.12 * np.ones(12), np.linspace(2.0, 5.12, 12)
def run(solver, **kw):
    p = oo.GLP(objectives.Rastrigin(), lb=lb, ub=ub, x0=lb + 0.25 * (ub - lb), maxIter=60, maxFunEvals=10 ** 9,
               maxNonSuccess=10 ** 9, iprint=-1)
    return p.solve(solver, population=50, **kw)
a = run('de')
b = run('de_cuda', rng='cpython')
assert b.solverInfo['name'] == 'de_cuda'
assert a.istop == b.istop == -7 and a.evals['f'] == b.evals['f'] and a.evals['iter'] == b.evals['iter']
assert a.iterValues.f == b.iterValues.f and np.array_equal(a.xf, b.xf) and a.ff == b.ff and a.isFeasible == b.isFeasible
c = run('de_cuda', rng='cpython', baseVectorStrategy='best', hndvi=2)
d = run('de', baseVectorStrategy='best', hndvi=2)
assert c.iterValues.f == d.iterValues.f and np.array_equal(c.xf, d.xf)
p = oo.GLP(lambda x: (x * x).sum(), lb=lb, ub=ub, iprint=-1)
try:
    p.solve('de_cuda')
    raise SystemExit('a plain callable was accepted')
except oo.OpenOptException as e:
    assert 'registered device objective' in str(e)
print('drop-in ok', a.ff)
'''


def test_integration_stub_runs_inside_the_reference_tree(tmp_path):
    copy = tmp_path / 'ref'
    shutil.copytree(os.path.join(REF, 'openopt'), copy / 'openopt')
    text = open(os.path.join(ROOT, 'INTEGRATION.md')).read()
    stub = re.search(r'```python\n(# openopt/solvers/cuda/de_cuda_oo\.py\n.*?)```', text, re.S).group(1)
    plug = copy / 'openopt' / 'solvers' / 'cuda'
    plug.mkdir()
    (plug / '__init__.py').write_text('')
    (plug / 'de_cuda_oo.py').write_text(stub)
    script = tmp_path / 'worker.py'
    script.write_text(WORKER % {'root': ROOT, 'copy': str(copy)})
    out = subprocess.run([sys.executable, str(script)], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-4000:]
    assert 'drop-in ok' in out.stdout


GA_WORKER = r'''
import sys, os, warnings
warnings.simplefilter('ignore')
sys.path[:0] = [%(root)r, os.path.join(%(root)r, 'oracle'), os.path.join(%(root)r, 'tests')]
import numpy as np
from galileo_ref_shim import load_galileo
oo, g = load_galileo(150880, %(copy)r)                     # the stock galileo needs its Python-2 behaviours restored
import helpers
from openopt_b200 import _lib, objectives
_lib.DeviceGalileo = helpers.OracleBackedGalileo           # no GPU here: the oracle stands in for the device handle
lb, ub = -5.12 * np.ones(9), np.linspace(2.0, 5.12, 9)
def run(solver, **kw):
    p = oo.GLP(objectives.Rastrigin(), lb=lb, ub=ub, maxIter=50, maxFunEvals=10 ** 9, maxNonSuccess=10 ** 9, iprint=-1)
    return p.solve(solver, population=24, crossoverRate=0.8, mutationRate=0.1, **kw)
a = run('galileo')
b = run('galileo_cuda', rng='cpython', seed=150880)
assert b.solverInfo['name'] == 'galileo_cuda'
assert a.istop == b.istop == -7 and a.evals['f'] == b.evals['f'] and a.evals['iter'] == b.evals['iter']
assert a.iterValues.f == b.iterValues.f and np.array_equal(a.xf, b.xf) and a.ff == b.ff and a.isFeasible == b.isFeasible
print('drop-in ok', a.ff)
'''


def test_galileo_stub_runs_inside_the_reference_tree(tmp_path):
    """The galileo stub of INTEGRATION.md inside a scratch copy of the reference tree, next to the stock galileo."""
    copy = tmp_path / 'ref'
    shutil.copytree(os.path.join(REF, 'openopt'), copy / 'openopt')
    text = open(os.path.join(ROOT, 'INTEGRATION.md')).read()
    stub = re.search(r'```python\n(# openopt/solvers/cuda/galileo_cuda_oo\.py\n.*?)```', text, re.S).group(1)
    plug = copy / 'openopt' / 'solvers' / 'cuda'
    plug.mkdir()
    (plug / '__init__.py').write_text('')
    (plug / 'galileo_cuda_oo.py').write_text(stub)
    script = tmp_path / 'worker.py'
    script.write_text(GA_WORKER % {'root': ROOT, 'copy': str(copy)})
    out = subprocess.run([sys.executable, str(script)], capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-4000:]
    assert 'drop-in ok' in out.stdout
