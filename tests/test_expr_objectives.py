"""The open objective registry (openopt_b200/expr.py): an objective written as an expression is both a host callable
(numpy, the parity definition) and generated device code.  CPU part: the host form against the formulas it restates,
the generated code, the build of the device twin and its registration in liboode (no kernel runs without a GPU).
GPU part: the device twin against the host form, and a full solve against the reference's recorded run of its own
example objective (examples/glp_1.py:4) -- the golden cases ex_glp1 / con_ex_glp_ab_c also run through every
golden-driven test of tests/test_gpu_parity.py."""
import math
import os
import sys

import numpy as np
import pytest

import de_oracle as orc
import helpers
from conftest import ROOT
from openopt_b200 import GLP, _lib, expr as E, objectives


def example_module(name):
    sys.path.insert(0, os.path.join(ROOT, 'examples'))
    return __import__(name)


def test_host_form_of_the_reference_example_is_bit_identical():
    glp_1 = example_module('glp_1')
    rs = np.random.RandomState(3)
    for _ in range(500):
        x = rs.uniform(-1, 1, 4)
        assert glp_1.f(x) == glp_1.f_ref(x) == orc.glp1(x)


def test_expression_forms_of_the_canonical_objectives():
    """Rastrigin / Ackley / Griewank / sphere written as expressions evaluate exactly like the oracle's numpy forms."""
    x = E.X
    rast = E.ExprObjective(x * x - 10 * E.cos(2 * math.pi * x) + 10)
    ack = E.ExprObjective([x * x, E.cos(2 * math.pi * x)],
                          finalize=-20.0 * E.exp(-0.2 * E.sqrt(E.S(0) / E.DIM)) - E.exp(E.S(1) / E.DIM) + 20.0 + math.e)
    gri = E.ExprObjective([x * x, ('prod', E.cos(x / E.A))], finalize=1.0 + E.S(0) / 4000.0 - E.S(1),
                          aux=lambda D: np.sqrt(np.arange(1, D + 1)))
    rs = np.random.RandomState(5)
    for D in (1, 7, 100, 257):
        X = rs.uniform(-5, 5, (9, D))
        assert np.array_equal(rast(X), orc.rastrigin(X))
        assert np.array_equal(ack(X), orc.ackley(X))
        assert np.array_equal(gri(X), orc.griewank(X))
        assert np.array_equal(rast(X[0]), orc.rastrigin(X[0]))
    assert gri.header().count('template <> struct Obj<6>') == 1 and 'NS = 1, NPR = 1' in gri.header()
    assert 'AUX = true' in gri.header() and 'AUX = false' in rast.header()


def test_expression_errors():
    with pytest.raises(ValueError):
        E.ExprObjective([('prod', E.X)])                       # a sum slot is required
    with pytest.raises(ValueError):
        E.ExprObjective([E.X, E.X, E.X])                       # at most two slots
    with pytest.raises(TypeError):
        E.X + 'a'
    with pytest.raises(ValueError):
        E.ExprObjective(E.by_column([E.X, E.X]))(np.zeros(3))  # no expression for column 2


def gpu_case_objectives(D):
    """The expression objectives of the GPU test below (also built by the CPU test, so that the compiled device twins
    travel to the GPU box with the snapshot instead of being compiled there)."""
    x = E.X
    objs = [E.ExprObjective(E.where(x > 0.5, 3.0 * x * x, (x - 1.0) * (x - 1.0)) + E.J * 0.25, name='piecewise')]
    if D == 4:
        objs.append(objectives.get('glp1'))
    else:
        objs.append(E.ExprObjective([x * x, E.cos(2 * math.pi * x)],
                                    finalize=-20.0 * E.exp(-0.2 * E.sqrt(E.S(0) / E.DIM)) - E.exp(E.S(1) / E.DIM) + 20.0 + math.e,
                                    name='ackley_expr'))
    return objs


def test_device_twin_builds_and_registers():
    """nvcc cross-compiles the generated objective into its own shared object and liboode takes it in; the id is
    stable across registrations of the same library."""
    o = objectives.get('glp1')
    lib = o.build()
    assert os.path.exists(lib) and lib.endswith('.so')
    a, b = _lib.register_objective(lib), _lib.register_objective(lib)
    assert a == b == o.device_id and a >= 6
    with pytest.raises(_lib.LibraryError):
        _lib.register_objective(os.path.join(ROOT, 'openopt_b200', 'liboode.so'))     # exports no objective
    ids = {o.device_id for D in (4, 100) for o in gpu_case_objectives(D)}
    assert len(ids) == 3 and min(ids) >= 6


@pytest.mark.gpu
@pytest.mark.parametrize('D,NP', [(4, 40), (100, 300), (1000, 64)])
def test_device_twin_matches_host_form(D, NP):
    """eval_points and a few Philox generations of expression objectives against numpy: sums of + - * terms bit-exact,
    sin / cos / exp / pow based ones within 1e-12 (every kernel form: quad for D = 4, row kernels above)."""
    rs = np.random.RandomState(D)
    objs = gpu_case_objectives(D)
    lb, ub = -1.0 * np.ones(D), np.linspace(1.0, 2.0, D)
    for k, o in enumerate(objs):
        x0 = (lb + ub) / 2.0            # what GLP passes on (kernel/GLP.py:48) and the oracle's default
        with _lib.DeviceDE(o.device_id, lb, ub, x0=x0, population=NP, obj_aux=o.aux(D), rng='philox', seed=11, max_batch=4) as dev:
            X = rs.uniform(-1, 1, (50, D))
            got, want = dev.eval_points(X), np.array([o(r) for r in X])
            if k == 0:
                assert np.array_equal(got, want)
            else:
                assert np.allclose(got, want, rtol=1e-12, atol=0)
            dev.init()
            recs = dev.step(4)
            pop, vals = dev.population()
            want = np.array([o(r) for r in pop])
            assert np.array_equal(vals, want) if k == 0 else np.allclose(vals, want, rtol=1e-12, atol=0)
            assert recs[-1].best_f == vals.min()
            ora = orc.OracleDE(o, lb, ub, population=NP, stream='philox', seed=11)
            for _ in range(4):
                ora.step()
            assert np.array_equal(pop, ora.pop)


@pytest.mark.gpu
def test_reference_examples_run_on_the_gpu_backend():
    """examples/glp_1.py and examples/glp_Ab_c.py (objective object substituted) through p.solve('de', rng='cpython'):
    the iterates of the reference's recorded runs."""
    for mod, case in (('glp_1', 'ex_glp1_d4_np40'), ('glp_Ab_c', 'con_ex_glp_ab_c_d4_np40')):
        g, x0, _ = helpers.load_golden(case)
        m = example_module(mod)
        p = m.problem(maxIter=int(g['gens']) + 1, maxFunEvals=10 ** 15, maxNonSuccess=10 ** 9, maxTime=1e9, maxCPUTime=1e9, iprint=-1)
        r = p.solve('de', rng='cpython')
        assert np.array_equal(r.xf, g['xf'])
        assert np.allclose(np.array(r.iterValues.f), g['iter_f'], rtol=1e-12, atol=0)
        assert r.istop == int(g['istop']) and r.evals['f'] == int(g['n_evals'])
