"""The arithmetic behind two galileo kernels, restated in numpy and checked on the CPU (the kernels themselves are checked
on the GPU by tests/test_gpu_galileo.py):

* ga_scan_* (csrc/ga_kernels.cuh): the reference's running fitness sums are sequentially rounded, s[k] = RN(s[k-1] + f[k])
  (galileo_oo.py:268, :340-343).  While s stays inside one binade, s = S*u with an integer S in [2^52, 2^53) and u = ulp(s),
  and RN(s + f) = (S + rint(f/u))*u unless f/u is a tie -- an integer prefix sum, which is associative.  `chain_by_binade`
  below is the same chunked procedure (predicted binade from approximate carries, HARD chunks walked one by one) and must
  equal numpy's sequential cumsum bit for bit.
* ga_merge_rank_kernel: the place of every element in the reference's sort-then-reverse order (:472-473) from binary
  searches against the already ordered current generation and the sorted new entries.
"""
import math

import numpy as np
import pytest

CHUNK = 1024
LO, HI = (1 << 52) + 1, (1 << 53) - 1


def exponent(x):
    return math.frexp(abs(x))[1] - 1            # |x| in [2^E, 2^(E+1))


def chain_by_binade(f):
    n = len(f)
    nch = (n + CHUNK - 1) // CHUNK
    out = np.empty(n)
    tot = np.array([f[c * CHUNK:(c + 1) * CHUNK].sum() for c in range(nch)])          # ga_scan_totals_kernel (any order)
    with np.errstate(all='ignore'):
        approx_in = np.concatenate([[0.0], np.cumsum(tot)[:-1]])                       # ga_scan_predict_kernel
    summaries = []
    for c in range(nch):                                                               # ga_scan_quantize_kernel
        a, x = approx_in[c], f[c * CHUNK:(c + 1) * CHUNK]
        if a == 0.0 or not np.isfinite(a) or not -900 <= exponent(a) <= 900:
            summaries.append(None)
            continue
        E = exponent(a)
        u = math.ldexp(1.0, E - 52)
        with np.errstate(all='ignore'):
            t = x * math.ldexp(1.0, 52 - E)
            exact = (t * u == x) & (np.abs(t) < 2.0 ** 51)
            tie = np.abs(t - np.trunc(t)) == 0.5
        if not exact.all() or tie.any():
            summaries.append(None)
            continue
        pre = np.cumsum(np.rint(t).astype(np.int64))
        summaries.append((E, u, pre, int(pre.min()), int(pre.max())))
    s, hard = 0.0, 0
    for c in range(nch):                                                               # ga_scan_carry_kernel
        x, sm = f[c * CHUNK:(c + 1) * CHUNK], summaries[c]
        easy = False
        if sm is not None and s != 0.0 and np.isfinite(s) and exponent(s) == sm[0]:
            E, u, pre, qmin, qmax = sm
            S = int(s / u)
            lo, hi = S + qmin, S + qmax
            easy = (lo >= LO and hi <= HI) if S > 0 else (-hi >= LO and -lo <= HI)
        if easy:
            out[c * CHUNK:c * CHUNK + len(x)] = (S + pre).astype(np.float64) * u         # ga_scan_finish_kernel
            s = float(out[c * CHUNK + len(x) - 1])
        else:
            hard += 1
            for k in range(len(x)):
                s = s + float(x[k])
                out[c * CHUNK + k] = s
    return out, hard, nch


def distributions(n):
    rng = np.random.default_rng(0)
    return {
        'positive': rng.random(n) * 100, 'negative': -(rng.random(n) * 4000 + 100), 'mixed_signs': rng.standard_normal(n),
        'integers': rng.integers(-5, 50, n).astype(np.float64), 'quarters': rng.integers(0, 50, n) * 0.5 + 0.25,
        'tiny': rng.random(n) * 1e-300, 'huge': rng.random(n) * 1e300, 'growing': np.exp(rng.random(n) * 40),
        'cancel': np.concatenate([rng.random(n // 3) * 1e6, -rng.random(n // 3) * 1e6, rng.random(n - 2 * (n // 3))]),
        'halves_of_ulp': np.concatenate([[2.0 ** 60], np.full(n - 1, 2.0 ** 7)]),    # every addend is exactly half an ulp of the sum
    }


@pytest.mark.parametrize('kind', sorted(distributions(8)))
def test_chunked_integer_chain_equals_the_sequential_sum(kind):
    f = distributions(60_000)[kind]
    with np.errstate(all='ignore'):
        want = np.cumsum(f)                       # numpy's cumsum is the sequential loop
    got, hard, nch = chain_by_binade(f)
    assert np.array_equal(got, want), kind
    if kind in ('positive', 'negative', 'growing', 'integers'):
        assert hard < nch // 2, (hard, nch)       # only the binade crossings and ties are walked one by one
    if kind == 'halves_of_ulp':
        assert hard == nch                        # ties everywhere: round-half-even depends on the sum's parity


def test_merge_by_rank_equals_sort_then_reverse():
    rng = np.random.default_rng(1)
    for trial in range(300):
        N, M = int(rng.integers(1, 40)), int(rng.integers(0, 40))
        vals = rng.integers(0, 6, N + M).astype(float) * (1 if trial % 2 else 0.5)       # many ties
        cur = np.sort(vals[:N])[::-1]                     # the current generation: non-increasing cached fitness
        ch, kept = vals[N:], rng.random(M) < 0.7
        members = [(cur[k], k) for k in range(N)] + [(ch[i], N + i) for i in range(M) if kept[i]]
        members.sort(key=lambda t: t[0])                  # list.sort() is stable (galileo_oo.py:472)
        want = members[::-1][:N]                          # .reverse(), truncate (:473-474)
        cent = sorted([(ch[i], N + i) for i in range(M) if kept[i]], key=lambda t: (t[0], t[1]), reverse=True)
        ckeys = np.array([c[0] for c in cent]) if cent else np.zeros(0)
        top = [None] * N
        for k in range(N):                                # a chromosome: larger keys, the LATER members of its tie run, entries >= it
            K = cur[k]
            r = int(np.sum(cur > K)) + (int(np.sum(cur >= K)) - 1 - k) + int(np.sum(ckeys >= K))
            if r < N:
                assert top[r] is None
                top[r] = (K, k)
        for t, c in enumerate(cent):                      # a sorted entry: the entries before it and the chromosomes with a larger key
            r = t + int(np.sum(cur > c[0]))
            if r < N:
                assert top[r] is None
                top[r] = c
        assert top == want, trial
