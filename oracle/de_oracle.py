"""CPU oracle for the differential-evolution hot path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A numpy restatement of the reference's algorithm for ``p.solve('de')``
(reference: openopt/solvers/UkrOpt/de_oo.py, OpenOpt 0.43).  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl reference`` leg may import
it, and only as the checker / the timed CPU baseline.  The product path
(``openopt_b200``) never imports anything from ``oracle/``.

Parity pinning: the reference's own tests hold NO golden values for this path
(openopt/tests/rastrigin.py:1-9 asserts nothing), so this oracle is pinned against outputs of the
unmodified reference itself, run in the build container by ``tests/golden/make_golden.py``
(fixtures ``tests/golden/*.npz``; ``tests/test_oracle_golden.py`` replays them).

Every function cites the reference lines it restates.  Element-wise arithmetic is written in the
reference's own operation order (numpy does one IEEE rounding per operation, no FMA).
"""
import math
import random as _pyrandom

import numpy as np

TWO_PI = 2 * math.pi            # the python-float constant the canonical objectives use

# --------------------------------------------------------------------------------------------
# Random streams
# --------------------------------------------------------------------------------------------


class CPythonStream:
    """The reference's live RNG: stdlib ``random`` (MT19937) drawn one element at a time,
    row-major (de_oo.py:16-53 -- the ``except`` branch is what runs because ``from np import
    random`` at :12 always raises).  ``Seed`` = ``random.seed`` (:19,144)."""

    name = 'cpython'

    def __init__(self, seed):
        self.r = _pyrandom.Random()
        self.r.seed(seed)

    def rand(self, *shape):                         # de_oo.py:20-32
        n = int(np.prod(shape))
        rnd = self.r.random
        out = np.array([rnd() for _ in range(n)], dtype=np.float64)
        return out.reshape(shape)

    def randint(self, low, size):                   # de_oo.py:33-53
        n = int(np.prod(size))
        ri = self.r.randint
        out = np.array([ri(0, low - 1) for _ in range(n)], dtype=np.int64)
        return out.reshape(size)

    # draw schedule of one run (de_oo.py:147 and :188,200-201,236,245)
    def init_uniforms(self, NP, D):
        return self.rand(NP, D)

    def generation_draws(self, gen, NP, D, rows=None):
        ib = self.randint(NP, NP)
        r1 = self.randint(NP, (1, NP))[0]
        r2 = self.randint(NP, (1, NP))[0]
        Uf = self.rand(NP, D)
        Uc = self.rand(NP, D)
        return ib, r1, r2, Uf, Uc

    def strategy_draws(self, gen, NP, D, strat):
        """Draws of one generation for any strategy combination, in the reference's order:
        base vector indices (:188, only 'random'), direction indices (:200-201 two (h,NP) blocks, or
        :210 one block of 2h for the directed search), Ft uniforms (:236, only 'random'), crossover
        uniforms (:245)."""
        h = strat.hndvi
        d = Draws()
        if strat.base == 'random':
            d.ib = self.randint(NP, NP)
        if strat.direction == 'random':
            d.r1 = self.randint(NP, (h, NP))
            d.r2 = self.randint(NP, (h, NP))
        else:
            d.r = self.randint(NP, (2 * h,))
        if strat.factor == 'random':
            d.Uf = self.rand(NP, D)
        d.Uc = self.rand(NP, D)
        return d


class Strategy:
    """The solver's strategy parameters (de_oo.py:71-89,157-176)."""

    def __init__(self, base='random', direction='random', factor='random', hndvi=1):
        assert base in ('random', 'best') and direction in ('random', 'best') and factor in ('random', 'constant')
        assert int(hndvi) >= 1
        self.base, self.direction, self.factor, self.hndvi = base, direction, factor, int(hndvi)

    @property
    def is_default(self):
        return (self.base, self.direction, self.factor, self.hndvi) == ('random', 'random', 'random', 1)


class Draws:
    """One generation's draws; fields a strategy does not draw stay None."""
    ib = r1 = r2 = r = Uf = Uc = None


PHILOX_M0 = np.uint64(0xD2511F53)
PHILOX_M1 = np.uint64(0xCD9E8D57)
PHILOX_W0 = 0x9E3779B9
PHILOX_W1 = 0xBB67AE85
_M32 = np.uint64(0xFFFFFFFF)


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """Philox4x32-10 (Salmon et al., SC'11; Random123 reference constants).  Counter words are
    uint32 arrays (broadcastable), key words python ints.  Returns four uint32 arrays.
    No reference counterpart: this is the native-mode generator of the new backend
    (BASELINE.json north_star: 'Philox counter-based random draws')."""
    c0 = np.asarray(c0, dtype=np.uint64)
    c1 = np.asarray(c1, dtype=np.uint64)
    c2 = np.asarray(c2, dtype=np.uint64)
    c3 = np.asarray(c3, dtype=np.uint64)
    c0, c1, c2, c3 = np.broadcast_arrays(c0, c1, c2, c3)
    k0 = int(k0) & 0xFFFFFFFF
    k1 = int(k1) & 0xFFFFFFFF
    for _ in range(10):
        p0 = PHILOX_M0 * c0
        p1 = PHILOX_M1 * c2
        hi0, lo0 = p0 >> np.uint64(32), p0 & _M32
        hi1, lo1 = p1 >> np.uint64(32), p1 & _M32
        c0, c1, c2, c3 = (hi1 ^ c1 ^ np.uint64(k0), lo1, hi0 ^ c3 ^ np.uint64(k1), lo0)
        k0 = (k0 + PHILOX_W0) & 0xFFFFFFFF
        k1 = (k1 + PHILOX_W1) & 0xFFFFFFFF
    return (c0.astype(np.uint32), c1.astype(np.uint32), c2.astype(np.uint32), c3.astype(np.uint32))


def u32_to_unit(w):
    """One 32-bit word -> double in [0,1): exactly w * 2**-32."""
    return w.astype(np.float64) * (1.0 / 4294967296.0)


def mulhi64(w_hi, w_lo, n):
    """floor(((w_hi<<32 | w_lo) * n) / 2**64) for n < 2**32 (unbiased-enough index draw)."""
    n = np.uint64(n)
    return ((w_hi.astype(np.uint64) * n + ((w_lo.astype(np.uint64) * n) >> np.uint64(32))) >> np.uint64(32)).astype(np.int64)


TAG_ELEM = 0     # per-gene draws; generation 0 = initial population
TAG_INDEX = 1    # per-row donor indices
TAG_DIRECT = 2   # the 2*hndvi population-wide indices of the directed search


class PhiloxStream:
    """Native-mode draws, keyed by (seed, generation, global row, global column) so that results
    do not depend on how the population is split over GPUs.  One Philox4x32-10 call serves a PAIR
    of adjacent columns (one 32-bit word per uniform):

    counter = [col >> 1, row & 0xffffffff, (row >> 32) | tag << 24, generation], key = seed lo/hi.
      * gene draw (tag 0): even column: Uf = w0 * 2^-32, Uc = w1 * 2^-32; odd column: Uf = w2 * 2^-32,
        Uc = w3 * 2^-32.  Generation 0: the Uf word seeds the initial population.
      * index draw (tag 1): counter[0] = 0 -> ib = mulhi64(w0,w1,NP), r1 = mulhi64(w2,w3,NP);
                            counter[0] = 1 -> r2 = mulhi64(w0,w1,NP).
    """

    name = 'philox'

    def __init__(self, seed):
        seed = int(seed) & 0xFFFFFFFFFFFFFFFF
        self.k0, self.k1 = seed & 0xFFFFFFFF, seed >> 32

    def _genes(self, gen, rows, D):
        rows = np.asarray(rows, dtype=np.uint64).reshape(-1, 1)
        pairs = np.arange((D + 1) // 2, dtype=np.uint64).reshape(1, -1)
        w = philox4x32_10(pairs, rows & _M32, (rows >> np.uint64(32)) | np.uint64(TAG_ELEM << 24),
                          np.uint64(gen), self.k0, self.k1)
        Uf = np.empty((rows.shape[0], 2 * pairs.shape[1]))
        Uc = np.empty_like(Uf)
        Uf[:, 0::2], Uc[:, 0::2] = u32_to_unit(w[0]), u32_to_unit(w[1])
        Uf[:, 1::2], Uc[:, 1::2] = u32_to_unit(w[2]), u32_to_unit(w[3])
        return np.ascontiguousarray(Uf[:, :D]), np.ascontiguousarray(Uc[:, :D])

    def init_uniforms(self, NP, D, rows=None):
        rows = np.arange(NP) if rows is None else rows
        return self._genes(0, rows, D)[0]

    def generation_draws(self, gen, NP, D, rows=None):
        rows = np.arange(NP, dtype=np.uint64) if rows is None else np.asarray(rows, dtype=np.uint64)
        hi = (rows >> np.uint64(32)) | np.uint64(TAG_INDEX << 24)
        a0, a1, a2, a3 = philox4x32_10(np.uint64(0), rows & _M32, hi, np.uint64(gen), self.k0, self.k1)
        b0, b1, _, _ = philox4x32_10(np.uint64(1), rows & _M32, hi, np.uint64(gen), self.k0, self.k1)
        ib, r1, r2 = mulhi64(a0, a1, NP), mulhi64(a2, a3, NP), mulhi64(b0, b1, NP)
        Uf, Uc = self._genes(gen, rows, D)
        return ib, r1, r2, Uf, Uc

    def strategy_draws(self, gen, NP, D, strat):
        """Any strategy combination.  Per-row index draws beyond the default three come from further
        counters of the same (row, generation): counter[0] = 1 + k (k = 1..hndvi-1) gives
        r1[k] = mulhi64(w0,w1,NP), r2[k] = mulhi64(w2,w3,NP).  The 2h indices of the directed search are
        not per row: counter = [0, m, TAG_DIRECT << 24, generation], r[m] = mulhi64(w0,w1,NP).  Words a
        strategy does not use (ib under 'best', Uf under 'constant') are simply not consumed."""
        h = strat.hndvi
        ib, r1_0, r2_0, Uf, Uc = self.generation_draws(gen, NP, D)
        d = Draws()
        d.Uc = Uc
        if strat.factor == 'random':
            d.Uf = Uf
        if strat.base == 'random':
            d.ib = ib
        if strat.direction == 'random':
            rows = np.arange(NP, dtype=np.uint64)
            hi = (rows >> np.uint64(32)) | np.uint64(TAG_INDEX << 24)
            r1, r2 = [r1_0], [r2_0]
            for k in range(1, h):
                w0, w1, w2, w3 = philox4x32_10(np.uint64(1 + k), rows & _M32, hi, np.uint64(gen), self.k0, self.k1)
                r1.append(mulhi64(w0, w1, NP))
                r2.append(mulhi64(w2, w3, NP))
            d.r1, d.r2 = np.array(r1), np.array(r2)
        else:
            m = np.arange(2 * h, dtype=np.uint64)
            w0, w1, _, _ = philox4x32_10(np.uint64(0), m, np.uint64(TAG_DIRECT << 24), np.uint64(gen), self.k0, self.k1)
            d.r = mulhi64(w0, w1, NP)
        return d


# --------------------------------------------------------------------------------------------
# Canonical objectives: the host numpy form IS the parity definition.
# The reference evaluates the user's f one contiguous 1-D row at a time
# (kernel/nonLinFuncs.py:190-200); numpy's .sum() over the contiguous axis is pairwise, .prod()
# sequential, and axis=1 reductions of a C-contiguous 2-D array use the same inner loop per row.
# --------------------------------------------------------------------------------------------


def rosenbrock(x):
    return (100.0 * (x[..., 1:] - x[..., :-1] ** 2) ** 2 + (1.0 - x[..., :-1]) ** 2).sum(-1)


def rastrigin(x):                                   # openopt/tests/rastrigin.py:4
    return (x * x - 10 * np.cos(TWO_PI * x) + 10).sum(-1)


def ackley(x):
    D = x.shape[-1]
    return (-20.0 * np.exp(-0.2 * np.sqrt((x * x).sum(-1) / D))
            - np.exp(np.cos(TWO_PI * x).sum(-1) / D) + 20.0 + math.e)


def griewank(x):
    D = x.shape[-1]
    return 1.0 + (x * x).sum(-1) / 4000.0 - np.cos(x / np.sqrt(np.arange(1, D + 1))).prod(-1)


def sphere(x):
    return (x * x).sum(-1)


def shifted_sphere(x, shift):                       # examples/glp_3.py:6 ((x-shift)**2).sum()
    return ((x - shift) ** 2).sum(-1)


def glp1(x):
    """examples/glp_1.py:4 / examples/glp_Ab_c.py:6 (D = 4), the reference's own DE example objective.  The reference
    calls it with one row at a time, so every operation runs on numpy SCALARS (libm sin / cos / pow); a 2-D argument
    is therefore looped row by row here instead of being vectorised (numpy's SIMD array loops for sin / cos / pow can
    differ from the scalar ones in the last bit)."""
    x = np.asarray(x, dtype=np.float64)
    if x.ndim > 1:
        return np.array([glp1(row) for row in x], dtype=np.float64)
    return (x[0] - 1.5) ** 2 + np.sin(0.8 * x[1] ** 2 + 15) ** 4 + np.cos(0.8 * x[2] ** 2 + 15) ** 4 + (x[3] - 7.5) ** 4


OBJECTIVES = {'rosenbrock': rosenbrock, 'rastrigin': rastrigin, 'ackley': ackley,
              'griewank': griewank, 'sphere': sphere, 'glp1': glp1}


def eval_rows(f, pop):
    """p.f(pop): the user function sees one contiguous row at a time
    (kernel/nonLinFuncs.py:190-200).  The canonical objectives above accept 2-D input and reduce
    along the contiguous axis, which numpy evaluates with the same per-row loops; arbitrary
    callables are looped."""
    pop = np.ascontiguousarray(pop)
    if getattr(f, '__module__', None) == __name__:
        return np.asarray(f(pop), dtype=np.float64)
    return np.array([f(row) for row in pop], dtype=np.float64)


# --------------------------------------------------------------------------------------------
# The generation
# --------------------------------------------------------------------------------------------

MAX_REPAIR_PASSES = 4096          # termination guard shared with the CUDA kernel


def correct_box_constraints(lb, ub, pop):
    """_correct_box_constraints (de_oo.py:346-375), restated per element.

    Lower loop (:347-355): d = x-lb; while d<0: x = x - (1+scale)*d; d = x-lb; scale /= 2,
    scale = 1, 1/2, ...  (first pass = reflection about lb).  Upper loop (:357-365) likewise with
    d = x-ub > 0.  Final clamp (:368-373).  The reference's ``while`` is over the whole matrix, but
    an element is modified only while it violates, so each element sees scale = 2**-(its own pass
    number) -- identical to a per-element loop.  Returns (pop, number of repaired elements)."""
    pop = pop.copy()
    lb_b = np.broadcast_to(lb, pop.shape)
    ub_b = np.broadcast_to(ub, pop.shape)
    repaired = np.zeros(pop.shape, dtype=bool)
    for bound, sign in ((lb_b, -1), (ub_b, +1)):
        d = pop - bound
        viol = d < 0.0 if sign < 0 else d > 0.0
        scale = 1.0
        passes = 0
        while viol.any() and passes < MAX_REPAIR_PASSES:
            repaired |= viol
            c = 1 + scale
            pop[viol] = pop[viol] - c * d[viol]
            d = pop - bound
            viol = d < 0.0 if sign < 0 else d > 0.0
            scale /= 2
            passes += 1
    low = pop < lb_b
    pop[low] = lb_b[low]
    high = pop > ub_b
    pop[high] = ub_b[high]
    repaired |= low | high
    return pop, int(repaired.sum())


def make_trial(old_pop, ib, r1, r2, Uf, Uc, F, Cr, lb, ub):
    """Mutation + binomial crossover + bound repair (de_oo.py:186-254), default strategies
    ('random' base vector, 'random' direction with hndvi=1, 'random' difference factor).

      beta  = old_pop[ib]                         :188
      delta = old_pop[r2] - old_pop[r1]           :200-204,233
      Ft    = Uf*F ; mutant = beta + Ft*delta     :236-241   (t=U*F; m=t*delta; v=beta+m)
      gene kept from the parent iff ceil(Uc-Cr)==1, i.e. Uc > Cr   :244-250
    The reference blends with 0/1 masks; a select is identical for finite operands except for the
    sign of an exact zero (documented deviation)."""
    mutant = old_pop[ib] + (Uf * F) * (old_pop[r2] - old_pop[r1])
    trial = np.where(Uc > Cr, old_pop, mutant)
    return correct_box_constraints(lb, ub, trial)


def make_trial_strategy(old_pop, vals, constr_vals, last_best_x, d, strat, F, Cr, lb, ub):
    """Mutation + crossover + repair for any strategy combination (de_oo.py:186-254).

      base 'best'        beta = ones((NP,D))*best[2]: every row is the best row of the LAST evaluated
                         (trial, or initial) population, not the global Best               :191-193
      direction 'random' barycenter_k = old_pop[r_k].sum(0) over the (h,NP) index block (numpy adds the h
                         gathered rows in order), both divided by h when h != 1            :200-204,229-231
      direction 'best'   one block of 2h indices for the whole population, sorted by (constr_val, val)
                         -- ties broken by the index, the structured array's remaining field -- the
                         first h are 'best', the rest 'worst'; barycenter = rows.sum(0)/h, and divided
                         by h AGAIN when h != 1 (reference quirk); delta is one (D,) vector
                         broadcast over the rows                                         :209-231
      factor 'constant'  Ft = F                                                          :238-239
    """
    NP, D = old_pop.shape
    h = strat.hndvi
    beta = old_pop[d.ib] if strat.base == 'random' else np.ones((NP, D)) * last_best_x
    if strat.direction == 'random':
        b1 = old_pop[d.r1].sum(0)
        b2 = old_pop[d.r2].sum(0)
    else:
        arr = np.array([(int(j), constr_vals[j], vals[j]) for j in d.r],
                       dtype=[('i', int), ('constr_val', float), ('val', float)])
        arr.sort(order=['constr_val', 'val'])
        best_arr = arr['i'][:h]
        worst_arr = arr['i'][h:2 * h]
        b1 = old_pop[worst_arr].sum(0) / h
        b2 = old_pop[best_arr].sum(0) / h
    if h != 1:
        b1 = b1 / h
        b2 = b2 / h
    delta = b2 - b1
    Ft = d.Uf * F if strat.factor == 'random' else F
    mutant = beta + Ft * delta
    trial = np.where(d.Uc > Cr, old_pop, mutant)
    return correct_box_constraints(lb, ub, trial)


def first_argmin_nan_to_inf(vals):
    """_eval_pop (de_oo.py:297-305): NaN -> +inf, first-occurrence argmin."""
    vals = vals.copy()
    vals[np.isnan(vals)] = np.inf
    return vals, int(vals.argmin())


NAN_PENALTY = 1e10                 # de_oo.py:6


class Constraints:
    """General constraints of the problem as _eval_pop sees them (de_oo.py:306-338): linear A x <= b,
    Aeq x = beq (kernel/residuals.py:33-51, evaluated for the whole population at once as
    (A @ pop.T - b[:,None]).T), nonlinear c(x) <= 0, h(x) = 0 (one row at a time,
    kernel/nonLinFuncs.py:190-200) and the constraint tolerance the solver runs under."""

    def __init__(self, A=None, b=None, Aeq=None, beq=None, c=None, h=None, contol=1e-6):
        as2d = lambda M: None if M is None else np.atleast_2d(np.asarray(M, dtype=np.float64))
        as1d = lambda v: None if v is None else np.asarray(v, dtype=np.float64).ravel()
        self.A, self.b, self.Aeq, self.beq = as2d(A), as1d(b), as2d(Aeq), as1d(beq)
        self.c, self.h, self.contol = c, h, contol

    def constr_vals(self, pop):
        """P.mr(checkBoxBounds=False) + nanPenalty*P.nNaNs() for a multi-array point (de_oo.py:314-316;
        kernel/Point.py:128-167,363-374): the largest of 0, the inequality residuals and the absolute
        equality residuals, NaN entries ignored (nanmax), plus 1e10 per NaN among c and h."""
        NP = pop.shape[0]
        r = np.zeros(NP)
        nnan = np.zeros(NP)
        cv = hv = None
        if self.c is not None:
            cv = np.array([np.atleast_1d(self.c(row)) for row in pop], dtype=np.float64)
            nnan = nnan + np.isnan(cv).sum(1)
        if self.h is not None:
            hv = np.array([np.atleast_1d(self.h(row)) for row in pop], dtype=np.float64)
            nnan = nnan + np.isnan(hv).sum(1)
        ineq = []
        if self.A is not None and self.A.size:
            ineq.append((np.dot(self.A, pop.T) - self.b.reshape(-1, 1)).T)
        if cv is not None:
            ineq.append(cv)
        eq = []
        if self.Aeq is not None and self.Aeq.size:
            eq.append((np.dot(self.Aeq, pop.T) - self.beq.reshape(-1, 1)).T)
        if hv is not None:
            eq.append(hv)
        with np.errstate(invalid='ignore'), __import__('warnings').catch_warnings():
            __import__('warnings').simplefilter('ignore')
            for fv in ineq + [np.abs(e) for e in eq]:
                if fv.size:
                    vm = np.nanmax(fv, 1)
                    r = np.where(r < vm, vm, r)
        return r + NAN_PENALTY * nnan

    def point_mr(self, x):
        """mr() + nanPenalty*nNaNs() of ONE point, as the host does for bestPoint (de_oo.py:338)."""
        return float(self.constr_vals(np.asarray(x, dtype=np.float64)[None, :])[0])


def constrained_best(vals, constr_vals, contol):
    """_eval_pop's best individual under general constraints (de_oo.py:318-338): the lowest f among
    the rows with constr_val < contol (first occurrence), or, if there is none, the row with the
    smallest constr_val.  Returns (index, is_feasible)."""
    ind = constr_vals < contol
    if not np.any(ind):
        return int(np.argmin(constr_vals)), False
    IND = np.where(ind)[0]
    return int(IND[int(np.nanargmin(vals[IND]))]), True


def select_constrained(old_vals, trial_vals, old_c, trial_c):
    """Greedy selection with constraint values (de_oo.py:260-282):
    keep_old = (4*[oc<nc]-2)*[(nc>0)+(oc>0)] + (2*[of<nf]-1) > 0, i.e. if either side violates
    anything the smaller constr_val wins (ties -> the trial), otherwise the smaller f (ties -> trial).
    vals and constr_vals are blended arithmetically like the reference."""
    bool_constr = old_c < trial_c
    bool_v = old_vals < trial_vals
    zero_constr = 1 * ((trial_c > 0) + (old_c > 0))
    keep_old = ((bool_constr * 4 - 2) * zero_constr + (bool_v * 2 - 1)) > 0
    k = keep_old.astype(np.float64)
    with np.errstate(invalid='ignore'):
        new_vals = old_vals * k + trial_vals * (1.0 - k)
        new_c = old_c * k + trial_c * (1.0 - k)
    return keep_old, new_vals, new_c


def select(old_vals, trial_vals):
    """Greedy selection, box-only branch (de_oo.py:260-282).  constr_vals == 0 on both sides, so
    keep_old = old_f < trial_f (ties and NaN-old -> the trial wins).  vals are blended
    arithmetically exactly like the reference (:281) so that inf*0 = NaN enters vals where the
    reference lets it."""
    keep_old = old_vals < trial_vals
    k = keep_old.astype(np.float64)
    with np.errstate(invalid='ignore'):
        new_vals = old_vals * k + trial_vals * (1.0 - k)
    return keep_old, new_vals


class OracleDE:
    """Whole-run restatement of de.__solver__ (de_oo.py:118-290) for box-bounded problems."""

    def __init__(self, f, lb, ub, x0=None, population=None, F=0.8, Cr=0.5, seed=150880,
                 stream='cpython', invert=False, strategy=None, constraints=None):
        self.f = f
        self.strategy = strategy or Strategy()
        self.constraints = constraints          # None: box-bounded (de_oo.py:303-305)
        self.lb = np.asarray(lb, dtype=np.float64)
        self.ub = np.asarray(ub, dtype=np.float64)
        self.D = self.lb.size
        self.NP = 10 * self.D if population is None else int(population)     # :130-133
        self.F, self.Cr = F, Cr
        self.sign = -1.0 if invert else 1.0       # nonLinFuncs.py:297-298 (goal='max')
        self.stream = (CPythonStream(seed) if stream == 'cpython' else
                       PhiloxStream(seed) if stream == 'philox' else stream)
        NP, D = self.NP, self.D
        U0 = self.stream.init_uniforms(NP, D)
        pop = U0 * (self.ub - self.lb) + self.lb                               # :147
        if x0 is None:
            x0 = (self.lb + self.ub) / 2.0                                     # GLP.py:48
        x0 = np.asarray(x0, dtype=np.float64)
        if np.any(np.isfinite(x0)):                                            # :149-150
            pop[0] = x0
        self.pop = pop
        vals = self._eval(pop)
        self.vals, bi = first_argmin_nan_to_inf(vals)                          # :153
        self.constr_vals = np.zeros(NP)            # box-bounded problems: _eval_pop :296
        self.best_feasible = True
        if constraints is not None:
            self.constr_vals = constraints.constr_vals(pop)
            bi, self.best_feasible = constrained_best(self.vals, self.constr_vals, constraints.contol)
        self.best_f, self.best_x = self.vals[bi], pop[bi].copy()               # :155
        self.best_c = float(self.constr_vals[bi])
        self.last_best_x = pop[bi].copy()          # best[2] of the last _eval_pop (base vector 'best', :192)
        self.gen = 0
        self.n_evals = NP
        self.history = dict(trial_best_idx=[bi], trial_best_f=[self.best_f], best_f=[self.best_f],
                            n_accepted=[], n_repaired=[], best_changed=[True],
                            trial_best_c=[self.best_c], trial_best_feasible=[self.best_feasible])
        self.trial_best_x = pop[bi].copy()
        self.last_accept = None
        self.last_trial = None
        self.last_trial_vals = None

    def _eval(self, pop):
        v = eval_rows(self.f, pop)
        return self.sign * v if self.sign < 0 else v

    def step(self, draws=None):
        """One trip of the hot loop (de_oo.py:178-290)."""
        NP, D = self.NP, self.D
        self.gen += 1
        if self.strategy.is_default:
            if draws is None:
                draws = self.stream.generation_draws(self.gen, NP, D)
            ib, r1, r2, Uf, Uc = draws
            trial, n_rep = make_trial(self.pop, ib, r1, r2, Uf, Uc, self.F, self.Cr, self.lb, self.ub)
        else:
            if draws is None:
                draws = self.stream.strategy_draws(self.gen, NP, D, self.strategy)
            trial, n_rep = make_trial_strategy(self.pop, self.vals, self.constr_vals, self.last_best_x, draws,
                                               self.strategy, self.F, self.Cr, self.lb, self.ub)
        tvals, bi = first_argmin_nan_to_inf(self._eval(trial))                 # :256
        feasible = True
        if self.constraints is not None:
            tc = self.constraints.constr_vals(trial)                           # :314-316
            bi, feasible = constrained_best(tvals, tc, self.constraints.contol)   # :318-338
            keep_old, new_vals, new_c = select_constrained(self.vals, tvals, self.constr_vals, tc)
            self.last_trial_constr = tc
        else:
            tc = np.zeros(NP)
            keep_old, new_vals = select(self.vals, tvals)                      # :260-282
            new_c = tc
        self.last_best_x = trial[bi].copy()
        self.trial_best_x = trial[bi].copy()
        self.n_evals += NP
        accept = ~keep_old
        self.pop = np.where(accept[:, None], trial, self.pop)                  # :277
        self.vals = new_vals
        self.constr_vals = new_c
        # Best tracking: `if Old_best.betterThan(Best): Best = Old_best` (:285-286); for a
        # box-bounded problem betterThan is f_self < f_other (kernel/Point.py:309-333), so on a tie
        # the newer trial-best replaces Best.
        # (With general constraints Best is tracked by Point.betterThan on the HOST, which needs the
        # host's own mr() of the point; this field then only mirrors the trial-best.)
        changed = not (self.best_f < tvals[bi]) if self.constraints is None else True
        if changed:
            self.best_f, self.best_x = tvals[bi], trial[bi].copy()
        h = self.history
        h['trial_best_c'].append(float(tc[bi]))
        h['trial_best_feasible'].append(feasible)
        h['trial_best_idx'].append(bi)
        h['trial_best_f'].append(tvals[bi])
        h['best_f'].append(self.best_f)
        h['n_accepted'].append(int(accept.sum()))
        h['n_repaired'].append(n_rep)
        h['best_changed'].append(changed)
        self.last_accept, self.last_trial, self.last_trial_vals = accept, trial, tvals
        return accept

    def run(self, n_gens, record_accept=False):
        masks = []
        for _ in range(n_gens):
            a = self.step()
            if record_accept:
                masks.append(a.astype(np.uint8))
        return np.array(masks, dtype=np.uint8) if record_accept else None


# --------------------------------------------------------------------------------------------
# numpy's pairwise summation, restated (numpy/_core/src/umath/loops_utils.h.src
# DOUBLE_pairwise_sum; numpy is an unpinned third-party dependency of the reference,
# setup.py:91 -- version 2.3.5 in this image).  Used to derive the reduction plan the CUDA kernel
# follows and to cross-check the C oracle; verified against np.sum in tests/test_oracle_golden.py.
# --------------------------------------------------------------------------------------------


def pairwise_sum(a):
    n = len(a)
    if n < 8:
        res = 0.0
        for v in a:
            res = res + v
        return res
    if n <= 128:
        r = [a[j] for j in range(8)]
        i = 8
        while i < n - (n % 8):
            for j in range(8):
                r[j] = r[j] + a[i + j]
            i += 8
        res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]))
        while i < n:
            res = res + a[i]
            i += 1
        return res
    n2 = n // 2
    n2 -= n2 % 8
    return pairwise_sum(a[:n2]) + pairwise_sum(a[n2:])


def pairwise_leaves(n, off=0):
    """Leaf blocks [(offset, length)] of the pairwise tree for a length-n contiguous sum."""
    if n <= 128:
        return [(off, n)]
    n2 = n // 2
    n2 -= n2 % 8
    return pairwise_leaves(n2, off) + pairwise_leaves(n - n2, off + n2)
