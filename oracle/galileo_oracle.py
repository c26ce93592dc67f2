"""CPU restatement of the `galileo` genetic algorithm -- TEST INFRASTRUCTURE ONLY (the oracle of the sibling
population solver).  Only tests/, __graft_entry__.smoke() and the golden generator import it.

Reference: openopt/solvers/Standalone/galileo_oo.py (all line numbers below refer to that file).  The solver
hard-wires one configuration of Donald Goodman's GA library (:35-66): roulette selection, one-point crossover,
per-gene re-initialising mutation, steady-state replacement without duplicates, replacementSize = population.

How the reference is pinned.  galileo_oo.py is Python-2 code: `list.sort()` over `Chromosome.__cmp__` (:188-191),
slicing through `__getslice__` (:169-176) and an UNSEEDED `random.Random()` (:223).  Under this image's Python 3.12
it dies in `replace_SteadyStateNoDuplicates` (:478).  oracle/galileo_ref_shim.py restores the three Python-2
behaviours at run time (nothing under /root/reference is edited): `__lt__` := `__cmp__(...) < 0`, slice objects
routed to `__getslice__`, and `Random()` replaced by `Random(seed)`; tests/golden/make_golden_galileo.py then
records the reference's own runs and asserts that this file reproduces them bit for bit.  What is NOT reproduced:
Python 2's `randint` mapping (`int(random()*n)`; Python 3 uses `_randbelow`) and the int-truncating comparison of
old-style instances -- the goldens are Python-3 runs of the unmodified file.

One generation (`__solver__` loop :73-92), as restated here:
  evaluate (:255-276)   cached fitness = -f(genes); sumFitness accumulated left to right from 0.0; the best is the
                        FIRST chromosome with the strictly largest fitness.
  select (:290-301)     2*ceil(N/2) roulette spins (:326-346): wheel = 0 + (sum - 0)*random(); the first chromosome
                        whose running sum (again left to right) is >= wheel, else the last one.  The selected
                        objects are ALIASES of current chromosomes, not copies.
  crossover (:303-312, :354-369)  per pair: prob = random(); if prob <= crossoverRate: cut = randint(0, D-1) and
                        two NEW chromosomes a[:cut]+b[cut:], b[:cut]+a[cut:]; else the pair stays aliased.
  mutate (:281-288, :425-441)  for the first N entries only: per gene prob = random(); if prob <= mutationRate the
                        gene becomes uniform(lb, ub) = lb + (ub-lb)*random().  An aliased entry is mutated IN PLACE in
                        the current generation and keeps its cached (now stale) fitness.
  replace (:458-474)    each entry in order is appended unless its genes equal those of a chromosome already in the
                        list (an alias always equals itself); `sort()` ascending by cached fitness (stable), reverse,
                        keep the first N: descending fitness, ties broken towards the LATER list position.
  report (:86-89)       fval = -maxFitness and bestFitIndividual.genes of the evaluate step ABOVE (so the population
                        at the start of the iteration; the genes are read after the in-place mutations).
"""
import random

import numpy as np

from de_oracle import _M32, philox4x32_10, u32_to_unit, mulhi64, PhiloxStream

TAG_GA = 3          # counter tag of the GA's per-pick / per-pair draws (de_oracle.py: tags 0..2 belong to de)


class Draws:
    """The draws of one generation: wheel[2P] in [0,1); cut[P] (-1: the pair is not crossed);
    mut[N][D] bool and mut_u[N][D] (the uniform behind a mutated gene)."""


class GACPythonStream:
    """The reference's own generator: stdlib `random.Random(seed)` consumed in the reference's order."""
    name = 'cpython'

    def __init__(self, seed):
        self.r = random.Random(seed)

    def init_uniforms(self, N, D):
        rnd = self.r.random
        return np.array([rnd() for _ in range(N * D)], dtype=np.float64).reshape(N, D)     # :127-133 via :246-252

    def generation_draws(self, gen, N, D, cr, mr):
        rnd, P = self.r, (N + 1) // 2
        d = Draws()
        d.wheel = np.array([rnd.random() for _ in range(2 * P)], dtype=np.float64)           # :295-297, :337
        cut = []
        for _ in range(P):                                                                   # :309-312
            prob = rnd.random()                                                              # :360
            cut.append(rnd.randint(0, D - 1) if prob <= cr else -1)                          # :361-363
        d.cut = np.array(cut, dtype=np.int64)
        d.mut = np.zeros((N, D), dtype=bool)
        d.mut_u = np.zeros((N, D), dtype=np.float64)
        for i in range(N):                                                                   # :287-288
            for j in range(D):                                                               # :430
                if rnd.random() <= mr:                                                       # :431-432
                    d.mut[i, j] = True
                    d.mut_u[i, j] = rnd.random()                                             # :439 (uniform's one draw)
        return d


class GAPhiloxStream:
    """Native-mode draws of the device backend (no reference counterpart), Philox4x32-10 keyed by the seed:
      initial genes  : de's gene stream at generation 0 (PhiloxStream._genes: Uf word of (row, column pair))
      spin m         : counter = [0, m & 0xffffffff, (m >> 32) | TAG_GA << 24, gen]; wheel = 53 bits of (w0, w1)
      pair p         : counter = [1, p & 0xffffffff, (p >> 32) | TAG_GA << 24, gen]; prob = 53 bits of (w0, w1),
                       cut = mulhi64(w2, w3, D)
      gene (i, j)    : de's gene stream at generation gen: prob = Uf word * 2^-32, new value uniform = Uc word * 2^-32
    (53 bits of (a, b) = ((a >> 5) * 2^26 + (b >> 6)) * 2^-53, CPython's own mapping.)"""
    name = 'philox'

    def __init__(self, seed):
        self.genes = PhiloxStream(seed)
        self.k0, self.k1 = self.genes.k0, self.genes.k1

    @staticmethod
    def _u53(a, b):
        return ((a >> np.uint32(5)).astype(np.float64) * 67108864.0 + (b >> np.uint32(6)).astype(np.float64)) * (1.0 / 9007199254740992.0)

    def init_uniforms(self, N, D):
        return self.genes.init_uniforms(N, D)

    def selection_draws(self, gen, N, D, cr):
        """(wheel[2P], cut[P]) alone -- cheap at any population size."""
        P = (N + 1) // 2
        m = np.arange(2 * P, dtype=np.uint64)
        w = philox4x32_10(np.uint64(0), m & _M32, (m >> np.uint64(32)) | np.uint64(TAG_GA << 24), np.uint64(gen), self.k0, self.k1)
        wheel = self._u53(w[0], w[1])
        p = np.arange(P, dtype=np.uint64)
        w = philox4x32_10(np.uint64(1), p & _M32, (p >> np.uint64(32)) | np.uint64(TAG_GA << 24), np.uint64(gen), self.k0, self.k1)
        prob = self._u53(w[0], w[1])
        return wheel, np.where(prob <= cr, mulhi64(w[2], w[3], D), -1).astype(np.int64)

    def generation_draws(self, gen, N, D, cr, mr):
        d = Draws()
        d.wheel, d.cut = self.selection_draws(gen, N, D, cr)
        Uf, Uc = self.genes._genes(gen, np.arange(N), D)
        d.mut = Uf <= mr
        d.mut_u = Uc
        return d


def make_stream(kind, seed):
    return GACPythonStream(seed) if kind == 'cpython' else GAPhiloxStream(seed)


class OracleGalileo:
    """f: objective over the last axis of an array (de_oracle.OBJECTIVES style).  State: `pop` [N][D], `fit` [N]
    (cached fitness, possibly stale after an in-place mutation), in the reference's list order."""

    def __init__(self, f, lb, ub, population=15, crossover_rate=1.0, mutation_rate=0.05, seed=150880,
                 stream='cpython', invert=False):
        self.f = f
        self.lb = np.asarray(lb, dtype=np.float64).copy()
        self.ub = np.asarray(ub, dtype=np.float64).copy()
        self.N, self.D = int(population), self.lb.size
        self.cr, self.mr = float(crossover_rate), float(mutation_rate)
        self.sign = -1.0 if invert else 1.0          # goal='max': p.f is the negated user objective
        self.stream = make_stream(stream, seed) if isinstance(stream, str) else stream
        U = self.stream.init_uniforms(self.N, self.D)
        self.pop = self.lb + (self.ub - self.lb) * U                                          # uniform(a, b) = a + (b-a)*random()
        self.fit = -self._f(self.pop)                                                        # :39 evalFunc = -p.f
        self.n_evals = self.N
        self.gen = 0
        self.history = dict(best_f=[], best_x=[], best_idx=[], n_appended=[], n_aliased=[], sum_fitness=[])

    def _f(self, rows):
        rows = np.atleast_2d(rows)        # the objectives reduce over the last axis with numpy's per-row inner loop (de_oracle.py)
        return self.sign * np.asarray(self.f(rows), dtype=np.float64).reshape(-1)

    @staticmethod
    def _key(row):
        return (row + 0.0).tobytes()               # list equality of floats: -0.0 == 0.0 (NaN genes cannot occur inside finite bounds)

    def step(self):
        N, D, P = self.N, self.D, (self.N + 1) // 2
        pop, fit = self.pop, self.fit
        self.gen += 1
        # ---- evaluate (:255-276)
        prefix = np.empty(N)
        s = 0.0
        for k in range(N):
            s = s + fit[k]
            prefix[k] = s
        best = 0
        for k in range(N):
            if fit[k] > fit[best]:
                best = k
        d = self.stream.generation_draws(self.gen, N, D, self.cr, self.mr)
        # ---- select (:290-301, :326-346)
        picks = np.empty(2 * P, dtype=np.int64)
        for m in range(2 * P):
            wheel = 0.0 + (s - 0.0) * d.wheel[m]
            hit = np.nonzero(prefix >= wheel)[0]
            picks[m] = hit[0] if hit.size else N - 1
        # ---- crossover (:303-312, :354-369): fresh children are built from the not yet mutated rows
        fresh = {}
        for p in range(P):
            a, b, cut = picks[2 * p], picks[2 * p + 1], int(d.cut[p])
            if cut >= 0:
                fresh[2 * p] = np.concatenate([pop[a, :cut], pop[b, cut:]])
                fresh[2 * p + 1] = np.concatenate([pop[b, :cut], pop[a, cut:]])
        # ---- mutate (:281-288, :425-441): entries 0..N-1 in order; aliases change the current generation in place
        for i in range(N):
            tgt = fresh[i] if i in fresh else pop[picks[i]]
            js = np.nonzero(d.mut[i])[0]
            tgt[js] = self.lb[js] + (self.ub[js] - self.lb[js]) * d.mut_u[i, js]
        best_x = pop[best].copy()                     # :86-89 read after the in-place mutations
        # ---- replace (:458-474)
        seen = {self._key(pop[k]) for k in range(N)}
        rows, new_fit = [pop[k] for k in range(N)], list(fit)
        appended = []
        for i in range(2 * P):
            if i not in fresh:
                continue                             # an alias is identical to itself
            key = self._key(fresh[i])
            if key in seen:
                continue
            seen.add(key)
            appended.append(i)
        if appended:
            f_new = -self._f(np.array([fresh[i] for i in appended]))
            self.n_evals += len(appended)
            rows += [fresh[i] for i in appended]
            new_fit += list(f_new)
        new_fit = np.array(new_fit, dtype=np.float64)
        order = sorted(range(len(rows)), key=lambda t: (new_fit[t] + 0.0, t), reverse=True)[:N]   # stable sort + reverse
        self.pop = np.array([rows[t] for t in order], dtype=np.float64).reshape(N, D)
        self.fit = new_fit[order]
        h = self.history
        h['best_f'].append(-fit[best]); h['best_x'].append(best_x); h['best_idx'].append(best)
        h['n_appended'].append(len(appended)); h['n_aliased'].append(2 * P - len(fresh)); h['sum_fitness'].append(s)
        self.last_picks, self.last_cut = picks, d.cut
        return -fit[best], best_x

    def step_fast(self):
        """The same iteration with numpy array operations instead of Python loops over the population (only the in-place
        mutations of aliased entries keep their list-order loop), for checks at populations of 10^5 ... 10^6.
        tests/test_galileo_oracle.py checks it against step() on random problems."""
        N, D, P = self.N, self.D, (self.N + 1) // 2
        pop, fit = self.pop, self.fit
        self.gen += 1
        prefix = np.cumsum(fit)                                    # numpy's cumsum is the sequential loop
        s = float(prefix[-1])
        best = int(np.argmax(fit))                                 # first occurrence of the largest fitness
        d = self.stream.generation_draws(self.gen, N, D, self.cr, self.mr)
        wheel = 0.0 + (s - 0.0) * d.wheel
        picks = np.minimum(np.searchsorted(np.maximum.accumulate(prefix), wheel, side='left'), N - 1).astype(np.int64)
        a, b, cut = picks[0::2], picks[1::2], d.cut
        left = np.arange(D)[None, :] < cut[:, None]                # columns taken from the first parent
        entries = np.empty((2 * P, D))
        entries[0::2] = np.where(left, pop[a], pop[b])
        entries[1::2] = np.where(left, pop[b], pop[a])
        fresh = np.repeat(cut >= 0, 2)
        new = self.lb + (self.ub - self.lb) * d.mut_u
        entries[:N] = np.where(d.mut & fresh[:N, None], new, entries[:N])
        for i in np.nonzero(~fresh[:N])[0]:                        # aliased entries: the current chromosome itself, in list order
            js = np.nonzero(d.mut[i])[0]
            pop[picks[i], js] = new[i, js]
        best_x = pop[best].copy()
        # replacement without duplicates: an entry is dropped if an identical row sits earlier in the list
        live = np.concatenate([np.arange(N), N + np.nonzero(fresh)[0]])
        members = np.concatenate([pop, entries[fresh]])             # the list the entries are appended to
        rows = members + 0.0                                       # compared by value: -0.0 == 0.0
        mult = (np.arange(1, D + 1, dtype=np.uint64) * np.uint64(0x9E3779B97F4A7C15)) | np.uint64(1)
        h = (rows.view(np.uint64) * mult).sum(axis=1, dtype=np.uint64)
        _, first, inv = np.unique(h, return_index=True, return_inverse=True)
        first = first[inv]                                         # earliest list member with the same hash
        cand = np.nonzero((first < np.arange(len(rows))) & (np.arange(len(rows)) >= N))[0]
        same = (rows[cand] == rows[first[cand]]).all(axis=1)
        if not same.all():                                         # a 64-bit collision between different rows: compare them all
            for t in cand[~same]:
                same[np.nonzero(cand == t)[0][0]] = any((rows[u] == rows[t]).all() for u in np.nonzero(h[:t] == h[t])[0])
        keep = np.ones(len(rows), dtype=bool)
        keep[cand[same]] = False
        kept = np.nonzero(keep[N:])[0] + N
        f_new = -self._f(rows[kept]) if len(kept) else np.zeros(0)
        self.n_evals += len(kept)
        all_fit = np.concatenate([fit, f_new])
        all_pos = np.concatenate([np.arange(N), live[kept]])
        src = np.concatenate([np.arange(N), kept])
        order = np.lexsort((all_pos, all_fit + 0.0))[::-1][:N]      # fitness descending, ties towards the later position
        self.pop = members[src[order]]
        self.fit = all_fit[order]
        hst = self.history
        hst['best_f'].append(-fit[best]); hst['best_x'].append(best_x); hst['best_idx'].append(best)
        hst['n_appended'].append(len(kept)); hst['n_aliased'].append(int(2 * P - fresh.sum())); hst['sum_fitness'].append(s)
        self.last_picks, self.last_cut = picks, d.cut
        return -fit[best], best_x

    def run(self, iterations):
        for _ in range(iterations):
            self.step()
        return self
