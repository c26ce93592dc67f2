"""The oracle (oracle/de_oracle.py) against the golden fixtures recorded from the unmodified
reference (tests/golden/make_golden.py).  CPU only; draws are regenerated from CPython's
`random` exactly as the reference draws them (de_oo.py:16-53)."""
import glob
import hashlib
import os

import numpy as np
import pytest

import de_oracle as orc
import helpers
from conftest import GOLDEN_DIR

CASES = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN_DIR, '*.npz')))


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()[:16]


def load(name):
    g = np.load(os.path.join(GOLDEN_DIR, name + '.npz'))
    x0 = g['x0'] if g['x0'].size else None
    return g, x0


def test_fixture_inventory():
    assert len(CASES) >= 24
    assert sum(c.startswith('con_') for c in CASES) >= 5
    assert sum(c.startswith('strat_') for c in CASES) >= 7
    assert 'c1_rosenbrock_d10_np100' in CASES


@pytest.mark.parametrize('name', CASES)
def test_oracle_matches_reference_golden(name):
    g, x0 = load(name)
    NP, gens = int(g['NP']), int(g['gens'])
    o = orc.OracleDE(orc.OBJECTIVES[str(g['obj'])], g['lb'], g['ub'], x0=x0, population=NP,
                     stream='cpython', invert=(str(g['goal']) == 'max'), strategy=helpers.golden_strategy(g),
                     constraints=helpers.golden_constraints(g))
    constrained = o.constraints is not None
    accept = np.unpackbits(g['accept_bits'], axis=1)[:, :NP].astype(bool)
    sign = -1.0 if str(g['goal']) == 'max' else 1.0
    for k in range(gens):
        a = o.step()
        assert np.array_equal(a, accept[k]), 'accept mask differs at generation %d' % (k + 1)
        assert sha(o.vals) == str(g['vals_sha'][k])
        assert sha(o.last_trial_vals) == str(g['trial_vals_sha'][k + 1])
        if constrained and name in helpers.EXACT_CONSTRAINED_CASES:
            assert sha(o.constr_vals) == str(g['cvals_sha'][k])
    assert np.array_equal(np.array(o.history['trial_best_idx']), g['trial_best_idx'])
    assert np.array_equal(np.array(o.history['n_repaired']), g['n_repaired'])
    if constrained:
        assert np.array_equal(np.array(o.history['trial_best_feasible']), g['trial_best_feasible'])
        assert np.allclose(o.history['trial_best_c'], g['trial_best_c'], rtol=1e-12, atol=0)
        assert np.allclose(o.constr_vals, g['final_cvals'], rtol=1e-12, atol=0)
    else:
        assert np.array_equal(sign * np.array(o.history['best_f'][1:]), g['iter_f'][1:])
        assert np.array_equal(o.best_x, g['xf'])
    assert np.array_equal(o.vals, g['final_vals'], equal_nan=True)
    assert sha(o.pop) == str(g['final_pop_sha'])
    assert o.n_evals == NP * (gens + 1)


def test_c1_survey_golden_values():
    """values recorded independently from the reference for config 1."""
    g, _ = load('c1_rosenbrock_d10_np100')
    assert float(g['ff']) == 5.215496384628898
    assert int(g['n_evals']) == 101003 and int(g['istop']) == -7
    assert list(g['trial_best_idx'][:12]) == [0, 0, 96, 63, 13, 39, 30, 34, 0, 61, 61, 58]
    assert sha(g['trial_best_idx'].astype(np.int64)) == 'c0de54ed828d6d29'
    accept = np.unpackbits(g['accept_bits'], axis=1)[:, :100]
    assert int(accept.sum()) == 37146 and sha(accept.astype(np.uint8)) == '30f76733bf078732'
    assert sha(g['iter_f']) == '705678ec8cbf1ec3'
    assert g['iter_f'][1] == 9.0 and g['iter_f'][500] == 5.828978142694421
    assert sha(g['final_vals']) == '85881645b71514fc'


def test_philox_known_answers():
    """Random123 kat_vectors for philox4x32-10."""
    z = np.uint32

    def kat(c, k):
        return [int(w) for w in orc.philox4x32_10(z(c[0]), z(c[1]), z(c[2]), z(c[3]), k[0], k[1])]
    assert kat([0] * 4, [0, 0]) == [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]
    assert kat([0xffffffff] * 4, [0xffffffff] * 2) == [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]
    assert kat([0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], [0xa4093822, 0x299f31d0]) == \
        [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]


def test_pairwise_sum_restatement_matches_numpy():
    rng = np.random.default_rng(7)
    for n in list(range(1, 300)) + [999, 1000, 1001, 4097, 10000]:
        a = rng.standard_normal(n) * 10.0 ** rng.uniform(-3, 3, n)
        assert orc.pairwise_sum(a) == a.sum()
        assert np.stack([a, a[::-1]]).copy().sum(1)[0] == a.sum()      # axis=1 rows use the same loop
        assert sum(l for _, l in orc.pairwise_leaves(n)) == n


def test_philox_stream_is_partition_independent():
    s = orc.PhiloxStream(150880)
    full = s.generation_draws(3, 64, 10)
    part = s.generation_draws(3, 64, 10, rows=np.arange(16, 48))
    for a, b in zip(full, part):
        assert np.array_equal(a[16:48], b)
    assert full[3].min() >= 0.0 and full[3].max() < 1.0
    assert full[0].min() >= 0 and full[0].max() < 64
