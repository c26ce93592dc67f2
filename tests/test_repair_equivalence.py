"""The bound repair as the CUDA kernels do it (csrc/de_device.cuh::repair_gene: one reflection about the violated
bound, the general loops only if that leaves [lb, ub]) restated in Python and compared with the oracle's
restatement of _correct_box_constraints (de_oo.py:346-375) on random and edge inputs, including overshoots of
many box widths (F > 1) where several reflections and the final clamp come into play.  CPU only: this pins the
ALGORITHMIC equivalence; the kernels themselves are compared with the oracle in tests/test_gpu_parity.py."""
import numpy as np

import de_oracle as orc


def slow(v, lo, hi):                       # repair_gene_slow
    d = v - lo
    if d < 0:
        scale, it = 1.0, 0
        while True:
            v = v - (1.0 + scale) * d
            d = v - lo
            scale *= 0.5
            it += 1
            if not (d < 0 and it < orc.MAX_REPAIR_PASSES):
                break
    d = v - hi
    if d > 0:
        scale, it = 1.0, 0
        while True:
            v = v - (1.0 + scale) * d
            d = v - hi
            scale *= 0.5
            it += 1
            if not (d > 0 and it < orc.MAX_REPAIR_PASSES):
                break
    return min(max(v, lo), hi)


def fast(v, lo, hi):                       # repair_gene
    lt, gt = v < lo, v > hi
    b = lo if lt else hi
    refl = v - 2.0 * (v - b)               # fma(-2, v - b, v): 2*(v-b) is exact, so one rounding either way
    x = refl if (lt or gt) else v
    if not (lo <= x <= hi):
        x = slow(v, lo, hi)
    return x, (lt or gt)


def test_single_reflection_fast_path_equals_the_reference_loops():
    rng = np.random.default_rng(3)
    for _ in range(120):
        D = int(rng.integers(1, 30))
        lo = rng.uniform(-10, 0, D) * rng.choice([1, 1e-3, 1e3])
        hi = lo + rng.uniform(1e-3, 10, D) * rng.choice([1, 1e-3, 1e3])
        V = rng.uniform(-1, 2, (40, D)) * (hi - lo) * rng.choice([0.5, 1.5, 4.0, 30.0]) + lo
        V[0], V[1] = lo, hi
        V[2], V[3] = np.nextafter(lo, -np.inf), np.nextafter(hi, np.inf)
        V[4], V[5] = lo - (hi - lo), hi + (hi - lo)
        want, nrep = orc.correct_box_constraints(lo, hi, V)
        got = np.empty_like(V)
        cnt = 0
        for i in range(V.shape[0]):
            for j in range(D):
                got[i, j], viol = fast(float(V[i, j]), float(lo[j]), float(hi[j]))
                cnt += viol
        assert np.array_equal(got, want) and cnt == nrep
