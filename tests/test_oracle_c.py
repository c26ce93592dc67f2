"""The C restatement (oracle/de_oracle.c) against the numpy oracle (pinned to the reference) and
against the golden fixtures.  Bit-exact except where glibc's exp differs from numpy's (Ackley)."""
import numpy as np
import pytest

import c_oracle
import de_oracle as orc
import helpers

RTOL = 1e-12


@pytest.mark.parametrize('name', helpers.DEFAULT_CASES)      # the C port covers the default strategies only
def test_c_oracle_matches_reference_golden(name):
    g, x0, accept = helpers.load_golden(name)
    obj, NP, gens = str(g['obj']), int(g['NP']), int(g['gens'])
    c = c_oracle.COracleDE(obj, g['lb'], g['ub'], x0=x0, population=NP, stream='cpython',
                           invert=(str(g['goal']) == 'max'))
    sign = -1.0 if str(g['goal']) == 'max' else 1.0
    bf = []
    for k in range(gens):
        c.step()
        assert np.array_equal(c.accept, accept[k]), 'accept mask, generation %d' % (k + 1)
        assert c.trial_best_idx == int(g['trial_best_idx'][k + 1])
        assert c.n_repaired == int(g['n_repaired'][k])
        bf.append(c.best_f)
    exact = obj != 'ackley'
    cmp = np.array_equal if exact else (lambda a, b: np.allclose(a, b, rtol=RTOL, atol=0))
    assert cmp(sign * np.array(bf), g['iter_f'][1:])
    assert np.array_equal(c.best_x, g['xf'])
    assert cmp(c.vals, g['final_vals'])


@pytest.mark.parametrize('obj,D,NP', [('rastrigin', 100, 300), ('rosenbrock', 33, 64), ('griewank', 200, 100),
                                      ('ackley', 1000, 50), ('sphere', 7, 20)])
def test_c_oracle_philox_matches_numpy_oracle(obj, D, NP):
    lb, ub = -5.0 * np.ones(D), np.linspace(1.0, 9.0, D)
    x0 = lb + 0.25 * (ub - lb)
    o = orc.OracleDE(orc.OBJECTIVES[obj], lb, ub, x0=x0, population=NP, stream='philox', seed=77)
    c = c_oracle.COracleDE(obj, lb, ub, x0=x0, population=NP, stream='philox', seed=77)
    assert np.array_equal(c.pop, o.pop)
    for _ in range(6):
        a = o.step()
        c.step()
        assert np.array_equal(c.accept.astype(bool), a)
        assert c.trial_best_idx == o.history['trial_best_idx'][-1]
    assert np.array_equal(c.pop, o.pop)
    if obj == 'ackley':
        assert np.allclose(c.vals, o.vals, rtol=RTOL, atol=0)
    else:
        assert np.array_equal(c.vals, o.vals)
