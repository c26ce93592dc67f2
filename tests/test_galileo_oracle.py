"""The galileo oracle (oracle/galileo_oracle.py) against the fixtures recorded from the unmodified reference
(tests/golden/galileo/*.npz, written by tests/golden/make_golden_galileo.py), the C ABI declared in include/ooga.h,
and properties of the native (Philox) draw schedule.  CPU only."""
import glob
import hashlib
import os
import random
import re
import subprocess

import numpy as np
import pytest

import de_oracle as orc
import galileo_oracle as gorc
from conftest import GOLDEN_DIR, ROOT
from openopt_b200 import _lib

GA_DIR = os.path.join(GOLDEN_DIR, 'galileo')
GA_CASES = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GA_DIR, '*.npz')))


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()[:16]


def oracle_for(g, stream='cpython'):
    return gorc.OracleGalileo(orc.OBJECTIVES[str(g['obj'])], g['lb'], g['ub'], population=int(g['N']),
                              crossover_rate=float(g['cr']), mutation_rate=float(g['mr']), seed=int(g['seed']),
                              stream=stream, invert=(str(g['goal']) == 'max'))


def test_fixture_set_is_complete():
    assert len(GA_CASES) == 16
    assert 'doc_solverparams_glp1_np15' in GA_CASES          # the reference's own call site, doc/solverParams.py:8-14
    assert any('cr0' in c for c in GA_CASES) and any('np1' in c for c in GA_CASES) and any('max' in c for c in GA_CASES)


@pytest.mark.parametrize('name', GA_CASES)
def test_oracle_reproduces_the_recorded_reference_run(name):
    g = np.load(os.path.join(GA_DIR, name + '.npz'))
    o = oracle_for(g)
    for k in range(int(g['iters'])):
        f, x = o.step()
        assert f == g['best_f'][k] and np.array_equal(x, g['best_x'][k]), (name, k)
        assert sha(o.pop) == str(g['pop_sha'][k]) and sha(o.fit) == str(g['fit_sha'][k]), (name, k)
    assert np.array_equal(o.pop, g['final_pop']) and np.array_equal(o.fit, g['final_fit'])
    assert o.history['n_appended'] == list(g['n_appended']) and o.history['n_aliased'] == list(g['n_aliased'])


def test_cpython_stream_consumes_the_generator_like_the_reference():
    """GACPythonStream against a direct transcription of the reference's order of consumption."""
    N, D, cr, mr = 7, 5, 0.6, 0.3
    s = gorc.GACPythonStream(99)
    r = random.Random(99)
    U = s.init_uniforms(N, D)
    assert np.array_equal(U.ravel(), [r.random() for _ in range(N * D)])
    d = s.generation_draws(1, N, D, cr, mr)
    P = (N + 1) // 2
    assert np.array_equal(d.wheel, [r.random() for _ in range(2 * P)])
    for p in range(P):
        cut = r.randint(0, D - 1) if r.random() <= cr else -1
        assert d.cut[p] == cut
    for i in range(N):
        for j in range(D):
            hit = r.random() <= mr
            assert d.mut[i, j] == hit
            if hit:
                assert d.mut_u[i, j] == r.random()


def test_philox_stream_is_counter_based():
    """Native-mode draws depend only on (seed, generation, index): any order, any subset, any population prefix."""
    s = gorc.GAPhiloxStream(7)
    a = s.generation_draws(3, 40, 11, 0.8, 0.1)
    b = s.generation_draws(3, 20, 11, 0.8, 0.1)
    assert np.array_equal(a.wheel[:20], b.wheel) and np.array_equal(a.cut[:10], b.cut)
    assert np.array_equal(a.mut[:20], b.mut) and np.array_equal(a.mut_u[:20], b.mut_u)
    c = gorc.GAPhiloxStream(7).generation_draws(4, 40, 11, 0.8, 0.1)
    assert not np.array_equal(a.wheel, c.wheel)
    assert np.all((a.wheel >= 0) & (a.wheel < 1)) and np.all((a.cut >= -1) & (a.cut < 11))
    big = s.generation_draws(1, 4000, 50, 0.7, 0.05)
    assert abs((big.cut >= 0).mean() - 0.7) < 0.05 and abs(big.mut.mean() - 0.05) < 0.005


@pytest.mark.parametrize('cr,mr,N', [(1.0, 0.05, 15), (0.5, 0.2, 16), (0.0, 0.5, 9)])
def test_oracle_invariants_in_native_mode(cr, mr, N):
    lb, ub = -5.12 * np.ones(6), np.linspace(1.0, 5.12, 6)
    o = gorc.OracleGalileo(orc.OBJECTIVES['rastrigin'], lb, ub, population=N, crossover_rate=cr, mutation_rate=mr,
                           seed=3, stream='philox')
    prev_best = np.inf
    for _ in range(30):
        f, x = o.step()
        assert np.all(np.diff(o.fit) <= 0)                          # list order = descending cached fitness
        assert np.all(o.pop >= lb) and np.all(o.pop <= ub)
        if cr >= 1.0:                                               # no in-place mutation: the best never gets worse
            assert f <= prev_best
            prev_best = f
            assert np.array_equal(-orc.OBJECTIVES['rastrigin'](o.pop), o.fit)
            assert len({r.tobytes() for r in o.pop}) == N           # replacement without duplicates
        assert f == orc.OBJECTIVES['rastrigin'](x) or cr < 1.0


@pytest.mark.parametrize('case', range(16))
def test_vectorised_step_equals_the_loop_step(case):
    """OracleGalileo.step_fast (array operations, for checks at 10^5 ... 10^6 chromosomes) against step (the line-by-line
    restatement pinned to the reference) on random problems: both draw streams, aliased entries, rates at 0 and 1."""
    rng = np.random.default_rng(500 + case)
    obj = ['sphere', 'rosenbrock', 'rastrigin', 'griewank'][case % 4]
    D = int(rng.choice([1, 2, 3, 8, 9, 33])) if obj != 'rosenbrock' else int(rng.choice([2, 3, 9]))
    N = int(rng.choice([1, 2, 3, 7, 16, 31, 64, 200]))
    kw = dict(population=N, crossover_rate=float(rng.choice([1.0, 0.8, 0.3, 0.0])), mutation_rate=float(rng.choice([0.05, 0.3, 0.0, 1.0])),
              seed=int(rng.integers(1, 2 ** 31)), stream=['philox', 'cpython'][case % 2], invert=bool(rng.random() < 0.4))
    lb, ub = -rng.random(D) * 3 - 0.5, rng.random(D) * 3 + 0.5
    a = gorc.OracleGalileo(orc.OBJECTIVES[obj], lb, ub, **kw)
    b = gorc.OracleGalileo(orc.OBJECTIVES[obj], lb, ub, **kw)
    for it in range(10):
        fa, xa = a.step()
        fb, xb = b.step_fast()
        assert fa == fb and np.array_equal(xa, xb), (case, it)
        assert np.array_equal(a.pop, b.pop) and np.array_equal(a.fit, b.fit), (case, it, obj, D, N, kw)
        assert np.array_equal(a.last_picks, b.last_picks) and np.array_equal(a.last_cut, b.last_cut)
    assert a.history['n_appended'] == b.history['n_appended'] and a.history['n_aliased'] == b.history['n_aliased']


def test_library_exports_every_symbol_of_ooga_h():
    lib = _lib.load()
    header = open(os.path.join(ROOT, 'include', 'ooga.h')).read()
    declared = set(re.findall(r'^\s*(?:int64_t|int|void)\s+\*?(ooga_\w+)\s*\(', header, re.M))
    assert declared == set(_lib.EXPORTS_GA), declared ^ set(_lib.EXPORTS_GA)
    for name in declared:
        assert hasattr(lib, name), name


def test_ooga_struct_layouts_match_header(tmp_path):
    structs = {'ooga_config': _lib.GaConfig, 'ooga_gen_record': _lib.GaRecord, 'ooga_timing': _lib.GaTiming}
    lines = []
    for cname, cls in structs.items():
        lines.append('printf("%s %%zu\\n", sizeof(%s));' % (cname, cname))
        for fname, _ in cls._fields_:
            lines.append('printf("%s.%s %%zu\\n", offsetof(%s, %s));' % (cname, fname, cname, fname))
    src = tmp_path / 'layout.c'
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "ooga.h"\nint main(void){%s return 0;}\n' % ''.join(lines))
    exe = tmp_path / 'layout'
    subprocess.check_call(['gcc', '-std=c99', '-Wall', '-Werror', '-I', os.path.join(ROOT, 'include'), str(src), '-o', str(exe)])
    out = dict(l.split() for l in subprocess.check_output([str(exe)], text=True).splitlines())
    for cname, cls in structs.items():
        assert int(out[cname]) == __import__('ctypes').sizeof(cls), cname
        for fname, _ in cls._fields_:
            assert int(out['%s.%s' % (cname, fname)]) == getattr(cls, fname).offset, (cname, fname)


def test_ooga_h_is_usable_from_plain_c(tmp_path):
    """tests/c/abi_check_ga.c: include/ooga.h compiles as strict C99 and liboode links from a C host; without a GPU the
    create call fails loudly, with one the program runs three iterations and a checkpoint round trip."""
    exe = tmp_path / 'abi_check_ga'
    libdir = os.path.join(ROOT, 'openopt_b200')
    _lib.load()
    subprocess.check_call(['gcc', '-std=c99', '-pedantic', '-Wall', '-Werror', '-I', os.path.join(ROOT, 'include'),
                           os.path.join(ROOT, 'tests', 'c', 'abi_check_ga.c'), '-o', str(exe), '-L', libdir, '-l:liboode.so',
                           '-Wl,-rpath,' + libdir, '-lm'])
    out = subprocess.run([str(exe)], capture_output=True, text=True)
    assert out.returncode == 0, (out.returncode, out.stdout, out.stderr)
    assert 'c abi (galileo) ok' in out.stdout


def test_plugin_fails_loudly_without_a_gpu():
    """No CPU fallback: on a machine without a CUDA device the solve ends in OpenOptException, not in a host GA."""
    import torch
    if torch.cuda.is_available():
        pytest.skip('a CUDA device is present')
    from openopt_b200 import GLP, OpenOptException, objectives
    p = GLP(objectives.Sphere(), lb=-np.ones(4), ub=np.ones(4), maxIter=5, iprint=-1)
    with pytest.raises(OpenOptException):
        p.solve('galileo')
    p = GLP(lambda x: (x * x).sum(), lb=-np.ones(4), ub=np.ones(4), maxIter=5, iprint=-1)
    with pytest.raises(OpenOptException, match='registered device objective'):
        p.solve('galileo')
