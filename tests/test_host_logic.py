"""CPU tests: the C-ABI library loads and exports what include/oode.h declares, fails loudly
without a GPU, reproduces CPython's random stream; the host mirror of the OpenOpt interface
(driver, stop criteria, history, evaluation counting) reproduces the reference's recorded results
when the device handle is replaced by the oracle."""
import ctypes
import os
import random
import re

import numpy as np
import pytest

import helpers
from conftest import ROOT
from openopt_b200 import GLP, NLP, OpenOptException, _lib, objectives, oosolver


def test_library_exports_every_declared_symbol():
    lib = _lib.load()
    header = open(os.path.join(ROOT, 'include', 'oode.h')).read()
    declared = set(re.findall(r'^\s*(?:int64_t|int|void|const char \*)\s*\*?(oode_\w+)\s*\(', header, re.M))
    assert declared == set(_lib.EXPORTS), declared ^ set(_lib.EXPORTS)
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.oode_abi_version() == _lib.OODE_ABI_VERSION


def test_struct_layouts_match_header(tmp_path):
    """ctypes structures against sizeof/offsetof as gcc sees include/oode.h."""
    import subprocess
    structs = {'oode_config': _lib.Config, 'oode_gen_record': _lib.GenRecord, 'oode_replay_draws': _lib.ReplayDraws,
               'oode_counters': _lib.Counters, 'oode_constraints': _lib.Constraints}
    lines = []
    for cname, cls in structs.items():
        lines.append('printf("%s %%zu\\n", sizeof(%s));' % (cname, cname))
        for fname, _ in cls._fields_:
            lines.append('printf("%s.%s %%zu\\n", offsetof(%s, %s));' % (cname, fname, cname, fname))
    src = tmp_path / 'layout.c'
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "oode.h"\nint main(void){%s return 0;}\n'
                   % ' '.join(lines))
    exe = tmp_path / 'layout'
    subprocess.check_call(['gcc', '-I', os.path.join(ROOT, 'include'), str(src), '-o', str(exe)])
    got = dict(line.split() for line in subprocess.check_output([str(exe)]).decode().splitlines())
    for cname, cls in structs.items():
        assert int(got[cname]) == ctypes.sizeof(cls), cname
        for fname, _ in cls._fields_:
            assert int(got['%s.%s' % (cname, fname)]) == getattr(cls, fname).offset, (cname, fname)


def test_c_abi_is_usable_from_plain_c(tmp_path):
    """tests/c/abi_check.c: the header compiles as strict C99 and the library links from a C host."""
    import subprocess
    exe = tmp_path / 'abi_check'
    libdir = os.path.join(ROOT, 'openopt_b200')
    _lib.load()                                   # built?
    subprocess.check_call(['gcc', '-std=c99', '-pedantic', '-Wall', '-Werror', '-I', os.path.join(ROOT, 'include'),
                           os.path.join(ROOT, 'tests', 'c', 'abi_check.c'), '-o', str(exe), '-L', libdir, '-l:liboode.so',
                           '-Wl,-rpath,' + libdir, '-lm'])
    out = subprocess.run([str(exe)], capture_output=True, text=True)
    assert out.returncode == 0, (out.returncode, out.stdout, out.stderr)
    first = random.Random()
    first.seed(150880)
    assert 'c abi ok' in out.stdout and ('first draws %d ' % first.randint(0, 99)) in out.stdout


@pytest.mark.parametrize('cr', [0.5, 0.0, 1.0, -0.5, 1.5, 0.9, 1e-10, 1 - 2.0 ** -33, 2.0 ** -32, 3 * 2.0 ** -32, 0.1, float('nan')])
def test_integer_crossover_threshold_equals_the_double_comparison(cr):
    """Philox mode decides `uc > Cr` (de_oo.py:245-250) on the raw word: w > floor(Cr * 2^32).  Against the
    double comparison on random words and on every word around the threshold."""
    t, keep_all = _lib.crossover_threshold(cr)
    rng = np.random.default_rng(1)
    edge = [0, 1, 2, 3, max(t, 1) - 1, t, min(t + 1, 2 ** 32 - 1), 2 ** 32 - 2, 2 ** 32 - 1]
    ws = np.concatenate([rng.integers(0, 2 ** 32, 100000, dtype=np.uint64), np.array(edge, dtype=np.uint64)])
    with np.errstate(invalid='ignore'):
        want = ws.astype(np.float64) * 2.0 ** -32 > cr
    assert np.array_equal(want, (ws > t) | keep_all)


def test_no_cpu_fallback():
    if _lib.device_count() > 0:
        pytest.skip('a GPU is present')
    with pytest.raises(_lib.LibraryError, match='no CPU fallback|CUDA'):
        _lib.DeviceDE(_lib.OBJ_SPHERE, -np.ones(3), np.ones(3), population=8)
    p = GLP(objectives.Sphere(), lb=-np.ones(3), ub=np.ones(3), iprint=-1)
    with pytest.raises(OpenOptException):
        p.solve('de', population=8)


def test_bad_config_is_rejected():
    lib = _lib.load()
    h = ctypes.c_void_p()
    cfg = _lib.Config()
    cfg.abi_version = 99
    assert lib.oode_create(ctypes.byref(cfg), ctypes.byref(h)) == -1
    assert b'ABI' in lib.oode_last_error()
    with pytest.raises(_lib.LibraryError, match='finite lb, ub'):
        _lib.DeviceDE(_lib.OBJ_SPHERE, np.array([-np.inf, 0.0]), np.ones(2), population=4)
    with pytest.raises(_lib.LibraryError, match='obj_aux'):
        _lib.DeviceDE(_lib.OBJ_GRIEWANK, -np.ones(2), np.ones(2), population=4)


@pytest.mark.parametrize('seed,below', [(150880, 100), (150880, 1000000), (1, 1), (2 ** 40 + 17, 37), (0, 4096)])
def test_native_cpython_stream_matches_interpreter(seed, below):
    """csrc/mt19937_cpython.h against the interpreter's own `random` (de_oo.py:16-53 draw kinds)."""
    ints, dbl = _lib.cpython_stream(seed, below, 5000, 5000)
    r = random.Random()
    r.seed(seed)
    assert [r.randint(0, below - 1) for _ in range(5000)] == ints.tolist()
    assert [r.random() for _ in range(5000)] == dbl.tolist()


# ---------------------------------------------------------------------------------------------
# host mirror against the reference's recorded runs (device handle replaced by the oracle)
# ---------------------------------------------------------------------------------------------
@pytest.fixture
def oracle_device(monkeypatch):
    monkeypatch.setattr(_lib, 'DeviceDE', helpers.OracleBackedDE)


@pytest.mark.parametrize('name', ['c1_rosenbrock_d10_np100', 'edge_sphere_d7_np13', 'edge_rosenbrock_d37_np50_max',
                                  'edge_ackley_d9_np40', 'c4_griewank_d200_np2000', 'strat_all_rastrigin_d130_np64',
                                  'strat_hndvi3_ackley_d140_np90', 'strat_directed_sphere_d20_np60', 'ex_glp1_d4_np40'])
@pytest.mark.parametrize('gens_per_call', [1, 7, 64])
def test_solve_matches_reference_record(oracle_device, name, gens_per_call):
    g, x0, _ = helpers.load_golden(name)
    kw = dict(lb=g['lb'], ub=g['ub'], maxIter=int(g['gens']) + 1, maxFunEvals=10 ** 15, maxNonSuccess=10 ** 9,
              iprint=-1)
    if x0 is not None:
        kw['x0'] = x0
    if str(g['goal']) == 'max':
        kw['goal'] = 'max'
    p = GLP(helpers.device_objective(str(g['obj'])), **kw)
    r = p.solve('de', population=int(g['NP']), rng='cpython', gensPerCall=gens_per_call,
                **helpers.solver_kwargs(helpers.golden_strategy(g)))
    assert r.istop == int(g['istop']) == -7
    assert r.evals['f'] == int(g['n_evals'])
    assert r.evals['iter'] == int(g['gens']) + 2
    assert np.array_equal(np.array(r.iterValues.f), g['iter_f'])
    assert np.array_equal(r.xf, g['xf'])
    assert r.ff == float(g['ff'])
    assert r.isFeasible and r.stopcase == 0 and r.rf == 0
    assert not hasattr(r.iterValues, 'x')
    assert r.solverInfo['name'] == 'de'


@pytest.mark.parametrize('name', helpers.CONSTRAINED_CASES)
@pytest.mark.parametrize('gens_per_call', [1, 16])
def test_constrained_solve_matches_reference_record(oracle_device, name, gens_per_call):
    """General constraints (A, b, Aeq, beq, c, h): the plugin's host protocol + the host mirror of Point /
    residuals / feasibility handling against the reference's recorded run."""
    g, x0, _ = helpers.load_golden(name)
    kw = dict(lb=g['lb'], ub=g['ub'], maxIter=int(g['gens']) + 1, maxFunEvals=10 ** 15, maxNonSuccess=10 ** 9,
              iprint=-1, **helpers.golden_constraint_kwargs(g))
    f = helpers.device_objective(str(g['obj']))
    if str(g['problem_class']) == 'NLP':
        p = NLP(f, x0, **kw)
    else:
        if x0 is not None:
            kw['x0'] = x0
        p = GLP(f, **kw)
    r = p.solve('de', population=int(g['NP']), rng='cpython', gensPerCall=gens_per_call,
                **helpers.solver_kwargs(helpers.golden_strategy(g)))
    assert r.istop == int(g['istop'])
    assert r.evals['f'] == int(g['n_evals']) and r.evals['c'] == int(g['n_evals_c']) and r.evals['h'] == int(g['n_evals_h'])
    assert np.array_equal(np.array(r.iterValues.f), g['iter_f'])
    assert np.array_equal(np.array([float(v) for v in r.iterValues.r]), g['iter_r'])
    assert np.array_equal(r.xf, g['xf'])
    assert r.ff == float(g['ff']) and r.rf == float(g['rf']) and r.isFeasible == bool(g['is_feasible'])


def test_default_stop_rules_truncate_like_the_reference(oracle_device):
    """openopt/tests/rastrigin.py: default x0 (the optimum) -> ff=0, istop=11 after 17 iterations
    (SURVEY.md section 4)."""
    p = GLP(objectives.Rastrigin(), lb=-5.12 * np.ones(10), ub=5.12 * np.ones(10), maxIter=1e3, maxFunEvals=1e5,
            maxTime=10, maxCPUTime=10, iprint=-1)
    r = p.solve('de', plot=0, rng='cpython')
    assert r.ff == 0.0 and r.istop == 11
    assert r.evals['f'] == 1720
    assert 'maxNonSuccess' in r.msg


def test_kwargs_routing_callbacks_and_oosolver(oracle_device, capsys):
    calls = []

    def cb(p):
        calls.append((p.iter, float(p.fk)))
        return len(calls) >= 5          # True -> istop 80

    p = GLP(objectives.Sphere(), lb=-np.ones(4), ub=2 * np.ones(4), iprint=1, callback=cb, maxNonSuccess=10 ** 9)
    r = p.solve(oosolver('de', population=20, rng='cpython'), bogusParam=3, maxIter=50)
    out = capsys.readouterr().out
    assert 'incorrect parameter for prob.solve(): "bogusParam"' in out
    assert ' iter   objFunVal' in out and 'istop: 80' in out
    assert r.istop == 80 and r.stopcase == 1 and len(calls) == 5
    assert p.solver.population == 20 and p.maxIter == 50
    with pytest.raises(OpenOptException, match='twice'):
        p.solve('de')


def test_plugin_rejects_what_it_cannot_run(oracle_device):
    p = GLP(lambda x: (x * x).sum(), lb=-np.ones(3), ub=np.ones(3), iprint=-1)
    with pytest.raises(OpenOptException, match='registered device objective'):
        p.solve('de')
    p = GLP(objectives.Sphere(), lb=-np.ones(3), ub=np.ones(3), iprint=-1)
    with pytest.raises(OpenOptException, match='incorrect baseVectorStrategy'):
        p.solve('de', baseVectorStrategy='bset')
    p = GLP(objectives.Sphere(), lb=-np.ones(3), ub=np.ones(3), iprint=-1)
    with pytest.raises(OpenOptException, match='hndvi'):
        p.solve('de', hndvi=0)
    p = GLP(objectives.Sphere(), lb=-np.ones(3), ub=np.ones(3), c=lambda x: x[0] - 0.5, iprint=-1)
    with pytest.raises(OpenOptException, match='registered device constraint'):
        p.solve('de')
    p = NLP(objectives.Sphere(), np.zeros(3), iprint=-1)          # infinite bounds -> implicitBounds = inf
    with pytest.raises(OpenOptException, match='finite lb, ub'):
        p.solve('de')
    p = GLP(objectives.Sphere(), lb=-np.ones(3), ub=np.ones(3), iprint=-1)
    with pytest.raises(OpenOptException, match='incorrect solver'):
        p.solve('ralg')


def test_nlp_with_box_bounds_and_fenough(oracle_device):
    p = NLP(objectives.Sphere(), 0.5 * np.ones(5), lb=-np.ones(5), ub=np.ones(5), iprint=-1, fEnough=1e-3,
            maxFunEvals=10 ** 9)
    r = p.solve('de', population=30, rng='cpython')
    assert r.istop == 10 and r.ff < 1e-3 and r.msg == 'fEnough has been reached'
