"""Test helpers: golden loading, draw regeneration, and an oracle-backed stand-in for the device
handle so the HOST logic (driver, stop criteria, history, plugin protocol) can be tested on a
machine without a GPU.  Never used by the product path."""
import glob
import os
from types import SimpleNamespace

import numpy as np

import de_oracle as orc
from conftest import GOLDEN_DIR

from openopt_b200 import objectives

CASES = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN_DIR, '*.npz')))
STRATEGY_CASES = [c for c in CASES if c.startswith('strat_')]      # non-default strategies
CONSTRAINED_CASES = [c for c in CASES if c.startswith('con_')]     # general constraints
DEFAULT_CASES = [c for c in CASES if not c.startswith(('strat_', 'con_'))]
# constraint residuals that are identical in any summation order (<= 2 non-zero coefficients per row)
EXACT_CONSTRAINED_CASES = [c for c in CONSTRAINED_CASES if 'dense' not in c]


def golden_constraint_kwargs(g):
    """Problem-constructor kwargs (A, b, Aeq, beq, c, h) a constrained fixture was recorded with."""
    import json
    from openopt_b200.constraints import DiagQuadratic
    if 'con_json' not in g.files:
        return {}
    con = json.loads(str(g['con_json']))
    kw = {}
    for k, v in con.items():
        kw[k] = DiagQuadratic(*v) if k in ('c', 'h') else np.asarray(v, dtype=np.float64)
    return kw


def golden_constraints(g):
    """oracle Constraints of a fixture (None for box-bounded ones)."""
    kw = golden_constraint_kwargs(g)
    return orc.Constraints(contol=float(g['contol']), **kw) if kw else None


def solver_kwargs(strat):
    """The reference's p.solve('de', ...) parameter names for a strategy (de_oo.py:71-89)."""
    return dict(baseVectorStrategy=strat.base, searchDirectionStrategy=strat.direction,
                differenceFactorStrategy=strat.factor, hndvi=strat.hndvi)
# objectives whose device evaluation is bit-exact with numpy (no transcendental functions)
EXACT_OBJECTIVES = ('rosenbrock', 'sphere')
# objectives only the numpy oracle restates (the C port knows the six built-in ones)
EXPRESSION_OBJECTIVES = ('glp1',)


def load_golden(name):
    g = np.load(os.path.join(GOLDEN_DIR, name + '.npz'))
    x0 = g['x0'] if g['x0'].size else None
    accept = np.unpackbits(g['accept_bits'], axis=1)[:, :int(g['NP'])].astype(np.uint8)
    return g, x0, accept


def golden_strategy(g):
    """The strategy a fixture was recorded with (fixtures older than the strategy cases: default)."""
    if 'strat_base' not in g.files:
        return orc.Strategy()
    return orc.Strategy(str(g['strat_base']), str(g['strat_direction']), str(g['strat_factor']), int(g['strat_hndvi']))


LAST_EXPR_NAME = {}     # device id -> name of the expression objectives handed out by device_objective()


def device_objective(name):
    o = objectives.get(name)
    if not isinstance(objectives.BY_NAME[name], type):        # an expression objective: compiled + registered on first use
        LAST_EXPR_NAME[o.device_id] = name
    return o


def oracle_objective(obj):
    return orc.OBJECTIVES[obj.name]


class OracleBackedDE:
    """Same interface as openopt_b200._lib.DeviceDE, computed by oracle/de_oracle.py."""

    def __init__(self, objective, lb, ub, x0=None, population=None, diff_factor=0.8, cross_rate=0.5,
                 seed=150880, obj_aux=None, rng='philox', invert=False, device=0, max_batch=64, base='random',
                 direction='random', factor='random', hndvi=1, **kw):
        builtin = {v.device_id: k for k, v in objectives.BY_NAME.items() if isinstance(v, type)}
        name = builtin[objective] if objective in builtin else LAST_EXPR_NAME[objective]
        self.o_args = dict(f=orc.OBJECTIVES[name], lb=lb, ub=ub, x0=x0, population=population, F=diff_factor,
                           Cr=cross_rate, seed=seed, stream=rng, invert=invert,
                           strategy=orc.Strategy(base, direction, factor, hndvi))
        self.o = None
        self.best_hist = {}
        self.generation = -1
        self.NP = int(population)
        self.constrained = False

    def set_constraints(self, A=None, b=None, Aeq=None, beq=None, c=None, h=None, contol=1e-6):
        self.o_args['constraints'] = orc.Constraints(A=A, b=b, Aeq=Aeq, beq=beq, c=c, h=h, contol=contol)
        self.constrained = True

    def _rec(self, g):
        h = self.o.history
        self.best_hist[g] = (self.o.trial_best_x if self.constrained else self.o.best_x).copy()
        return SimpleNamespace(trial_best_c=h['trial_best_c'][g], trial_best_feasible=int(h['trial_best_feasible'][g]),
                               trial_best_f=h['trial_best_f'][g], best_f=h['best_f'][g],
                               trial_best_idx=h['trial_best_idx'][g],
                               n_accepted=(h['n_accepted'][g - 1] if g else 0),
                               n_repaired=(h['n_repaired'][g - 1] if g else 0),
                               best_changed=int(h['best_changed'][g]), generation=g)

    def init(self, u0=None):
        self.o = orc.OracleDE(**self.o_args)
        self.generation = 0
        return self._rec(0)

    def step(self, n_gens=1, draws=None):
        out = []
        for _ in range(n_gens):
            self.o.step()
            self.generation += 1
            out.append(self._rec(self.generation))
        return out

    def best_x(self, generation=None):
        return self.best_hist[self.generation if generation is None else generation].copy()

    def counters(self):
        return dict(kernel_ms=0.0, generations=self.generation, trial_evals=0, gpu_launches=0, bytes_algorithmic=0)

    def close(self):
        pass


class OracleBackedGalileo:
    """Same interface as openopt_b200._lib.DeviceGalileo, computed by oracle/galileo_oracle.py."""

    def __init__(self, objective, lb, ub, population=15, crossover_rate=1.0, mutation_rate=0.05, seed=150880,
                 obj_aux=None, rng='philox', invert=False, device=0, max_batch=64):
        import galileo_oracle as gorc
        builtin = {v.device_id: k for k, v in objectives.BY_NAME.items() if isinstance(v, type)}
        name = builtin[objective] if objective in builtin else LAST_EXPR_NAME[objective]
        self.make = lambda: gorc.OracleGalileo(orc.OBJECTIVES[name], lb, ub, population=population,
                                               crossover_rate=crossover_rate, mutation_rate=mutation_rate, seed=seed,
                                               stream=rng, invert=invert)
        self.N, self.D = int(population), np.size(lb)
        self.o = None

    def init(self):
        self.o = self.make()

    def step(self, n_gens=1):
        recs, xs = [], []
        for _ in range(n_gens):
            before = self.o.n_evals
            f, x = self.o.step()
            h = self.o.history
            recs.append(SimpleNamespace(best_f=f, sum_fitness=h['sum_fitness'][-1], best_idx=h['best_idx'][-1],
                                        n_appended=self.o.n_evals - before, n_aliased=h['n_aliased'][-1],
                                        generation=self.o.gen))
            xs.append(x)
        return recs, np.array(xs)

    def gpu_launches(self):
        return 0

    def close(self):
        pass
