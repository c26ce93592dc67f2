"""`de` solver plugin, CUDA backend: same plugin contract, parameters and outer protocol as the
reference's openopt/solvers/UkrOpt/de_oo.py:56-290, with the generation loop on the GPU.

What stays on the host (this file): parameter handling (:127-176), building `Best` Points from
the per-generation records and calling ``p.iterfcn(Best)`` once per generation, stopping on
``p.istop`` (:257,285-290).  What moves to the device (liboode, include/oode.h): population init
(:147-155) and everything inside the loop (:178-282).  Generations are run in batches of up to
``gensPerCall`` per C call; the host then replays ``p.iterfcn`` over the recorded history, so stop
criteria fire on exactly the generation they would in the reference and speculative generations
past the stop are discarded.

The objective must be a registered device objective (openopt_b200.objectives); there is no CPU
fallback -- a plain Python callable is rejected with p.err().
"""
import math

import numpy as np

from ... import _lib
from ...host.registry import baseSolver
from ...host.stopping import SMALL_DELTA_F, SMALL_DELTA_X
from ...objectives import DeviceObjective


class de(baseSolver):
    __name__ = 'de'
    __license__ = "BSD"
    __authors__ = ("B200-native CUDA backend of OpenOpt's de (algorithm: Stepan Hlushak's two-array "
                   "differential evolution as connected to OpenOpt by Dmitrey)")
    __alg__ = ("Two array differential evolution algorithm, Feoktistov V. Differential Evolution. "
               "In Search of Solutions (Springer, 2006)(ISBN 0387368965).")
    iterfcnConnected = True
    __homepage__ = ''
    __isIterPointAlwaysFeasible__ = lambda self, p: p.__isNoMoreThanBoxBounded__()
    __optionalDataThatCanBeHandled__ = ['lb', 'ub', 'A', 'b', 'Aeq', 'beq', 'c', 'h']
    _requiresFiniteBoxBounds = True

    # ---- the reference's parameters (de_oo.py:71-90), same names and defaults ----
    baseVectorStrategy = 'random'
    searchDirectionStrategy = 'random'
    differenceFactorStrategy = 'random'
    population = 'default: 10*nVars will be used'
    differenceFactor = 0.8
    crossoverRate = 0.5
    hndvi = 1
    seed = 150880

    # ---- backend parameters (new) ----
    rng = 'philox'        # 'philox' (device, counter based) | 'cpython' (the reference's own stream)
    gensPerCall = 64      # generations per C call (history is replayed on the host afterwards)
    device = None         # CUDA device ordinal (default: LOCAL_RANK under torch.distributed, else 0)
    shard = 'auto'        # over the ranks of an initialised torch.distributed group: 'auto' = column shards when
                          # the row is wide enough, else row shards; 'cols' / 'rows' force one;
                          # 'none': every rank runs its own full copy

    __info__ = """
        Two array differential evolution, CUDA (sm_100a) backend.

        Finite box-bound constraints lb <= x <= ub are required; the objective must be a registered
        device objective (openopt_b200.objectives).

        Parameters (as in OpenOpt's de):
            population - population number (should be ~ 10*nVars),
            differenceFactor - difference factor (should be in (0,1] range),
            crossoverRate - constant of crossover (should be ~ 0.5, in (0,1] range),
            baseVectorStrategy - 'random' or 'best' (base vector = best row of the last trial population),
            searchDirectionStrategy - 'random' or 'best' (directed search),
            differenceFactorStrategy - 'random' (F' in [0,F) per gene) or 'constant',
            hndvi - half of number of individuals that take part in creation of difference vector,
            seed - random seed.
        Backend parameters:
            rng - 'philox' (default; Philox4x32-10 on the device) or 'cpython' (the reference's
                  random-module stream, regenerated natively: same iterates as OpenOpt's de),
            gensPerCall - generations per library call, device - CUDA device ordinal.
        """

    def __init__(self):
        pass

    def __solver__(self, p):
        if not p.__isFiniteBoxBounded__():
            p.err('this solver requires finite lb, ub: lb <= x <= ub')
        p.kernelIterFuncs.pop(SMALL_DELTA_X, None)
        p.kernelIterFuncs.pop(SMALL_DELTA_F, None)

        for attr, allowed in (('baseVectorStrategy', ('random', 'best')),
                              ('searchDirectionStrategy', ('random', 'best')),
                              ('differenceFactorStrategy', ('random', 'constant'))):
            val = getattr(self, attr)
            if val not in allowed:
                p.err('incorrect %s, should be "%s" or "%s", got %s' % (attr, allowed[0], allowed[1], str(val)))
        if int(self.hndvi) != self.hndvi or not 1 <= self.hndvi <= _lib.MAX_HNDVI:
            p.err('hndvi should be an integer in [1, %d], got %s' % (_lib.MAX_HNDVI, str(self.hndvi)))
        constrained = not p.__isNoMoreThanBoxBounded__()
        con = {}
        if constrained:                           # de_oo.py:306-338: general constraints
            from ...constraints import DeviceConstraint
            for kind in ('c', 'h'):
                if getattr(p.userProvided, kind):
                    fs = getattr(p.user, kind)
                    if len(fs) != 1 or not isinstance(fs[0], DeviceConstraint) or getattr(p.args, kind) != ():
                        p.err('the CUDA de backend needs nonlinear constraints (%s) as ONE registered device constraint '
                              '(see openopt_b200.constraints); arbitrary Python callables cannot run on the GPU' % kind)
                    con[kind] = fs[0]
            if np.size(p.A):
                con['A'], con['b'] = np.asarray(p.A, dtype=np.float64), p.b
            if np.size(p.Aeq):
                con['Aeq'], con['beq'] = np.asarray(p.Aeq, dtype=np.float64), p.beq

        funcs = p.user.f
        if len(funcs) != 1 or not isinstance(funcs[0], DeviceObjective):
            p.err('the CUDA de backend needs a registered device objective (see openopt_b200.objectives); '
                  'arbitrary Python callables cannot run on the GPU and there is no CPU fallback')
        if p.args.f != ():
            p.err('extra objective args are not supported by the CUDA de backend')
        obj = funcs[0]
        D = p.n
        NP = 10 * D if isinstance(self.population, str) else int(self.population)
        batch = max(1, int(self.gensPerCall))

        from ... import distributed as ddist
        world, rank, local_rank = ddist.world()
        if self.shard not in ('auto', 'none', 'cols', 'rows'):
            p.err('incorrect shard, should be "auto", "cols", "rows" or "none", got ' + str(self.shard))
        sharded = world > 1 and self.shard != 'none'
        if sharded and constrained:
            p.err("general constraints are not available on a sharded population yet: use shard='none'")
        device = self.device if self.device is not None else (local_rank if world > 1 else 0)
        try:
            extra = {}
            if sharded:
                kind = self.shard
                if kind == 'auto':
                    # column shards (only per-row leaf sums cross the GPUs) whenever the row has a leaf per
                    # GPU and the objective does not couple neighbouring columns; else row shards
                    try:
                        _lib.shard_layout(D, obj.device_id, world, rank)
                        kind = 'cols'
                    except _lib.LibraryError:
                        kind = 'rows'
                extra = dict(world_size=world, rank=rank, nccl_id=ddist.shared_nccl_id(),
                             shard=_lib.SHARD_COLS if kind == 'cols' else _lib.SHARD_ROWS)
            dev = _lib.DeviceDE(obj.device_id, p.lb, p.ub, x0=p.x0, population=NP,
                                diff_factor=self.differenceFactor, cross_rate=self.crossoverRate,
                                seed=self.seed, obj_aux=obj.aux(D), rng=self.rng, invert=p.invertObjFunc,
                                device=device, max_batch=batch, base=self.baseVectorStrategy,
                                direction=self.searchDirectionStrategy, factor=self.differenceFactorStrategy,
                                hndvi=int(self.hndvi), **extra)
        except _lib.LibraryError as e:
            p.err(str(e))
        # A sharded handle is advanced by collective calls: the ranks agree on every stop (distributed.agree_stop).
        # Stops that only depend on the replicated records are checked once per batch; wall-clock limits and user
        # callbacks can fire on different generations on different ranks, so with those the agreement is per generation.
        self._sharded = sharded
        self._agree_each_gen = sharded and (math.isfinite(p.maxTime) or math.isfinite(p.maxCPUTime) or bool(p.callback))
        try:
            if constrained:
                dev.set_constraints(contol=p.contol, **con)
                self._run_constrained(p, dev, NP, batch, con)
            else:
                self._run(p, dev, NP, batch)
            self.counters = dev.counters()
        except _lib.LibraryError as e:
            p.err(str(e))
        finally:
            dev.close()

    def _run(self, p, dev, NP, batch):
        rec = dev.init()
        p.nEvals['f'] += NP                      # what p.f(pop) would have counted (nonLinFuncs.py:300-302)
        Best = p.point(dev.best_x(0), f=rec.best_f, mr=0, mrName=None, mrInd=0)      # de_oo.py:155

        total = p.maxIter + 10                   # de_oo.py:178
        done = 0
        timed = math.isfinite(p.maxTime) or math.isfinite(p.maxCPUTime)
        while done < total:
            k = min(batch, total - done)
            # do not speculate far past a stop the host can predict (maxIter, maxFunEvals)
            k = min(k, max(1, p.maxIter - 1 - p.iter + 1))
            k = min(k, max(1, -(-(p.maxFunEvals - p.nEvals['f']) // NP)))
            if timed:
                k = min(k, 8)
            if self._agree_each_gen:
                k = 1
            recs = dev.step(k)
            # sharded: Best.x of a generation is assembled collectively, so every snapshot of the batch is fetched
            # before the replay -- no rank may skip a collective because it stopped earlier than another
            xs = {r.generation: dev.best_x(r.generation) for r in recs if r.best_changed} if self._sharded else None
            for rec in recs:
                p.nEvals['f'] += NP
                if rec.best_changed:             # de_oo.py:257,285-286 decided on the device
                    x = xs[rec.generation] if xs is not None else dev.best_x(rec.generation)
                    Best = p.point(x, f=rec.best_f, mr=0, mrName=None, mrInd=0)
                p.iterfcn(Best)                  # de_oo.py:288
                if p.istop and not self._sharded:      # de_oo.py:289-290
                    return
                if p.istop:
                    break
            if self._sharded and self._stop_together(p):
                return
            done += k

    def _stop_together(self, p):
        """Sharded runs: True when the ranks have agreed to stop (all of them, on this generation batch)."""
        from ... import distributed as ddist
        agreed = ddist.agree_stop(p.istop)
        if agreed and not p.istop:
            p.istop = agreed
            p.msg = 'stopped together with another rank of the sharded population (istop %s there)' % str(agreed)
        return bool(agreed)

    def _run_constrained(self, p, dev, NP, batch, con):
        """The same loop with general constraints.  The library takes _eval_pop's decisions for the whole
        population (constr_vals, the constrained choice of the best row :306-338, selection :260-282); the
        reference then evaluates that ONE row again on the host -- f only if nothing was feasible (:322),
        its max-residual always (:338) -- and compares Points with betterThan (:285-286), which is done here
        with the host Point as well."""
        nan_penalty = 1e10                       # de_oo.py:6

        def count_population_evals():            # p.f(pop), P.mr() -> p.c(pop), p.h(pop)  (:297,:316)
            p.nEvals['f'] += NP
            for kind in ('c', 'h'):
                if kind in con:
                    p.nEvals[kind] += NP

        def best_of(rec):
            x = dev.best_x(rec.generation)
            bp = p.point(x, f=rec.trial_best_f) if rec.trial_best_feasible else p.point(x)   # :333 / :322
            best = (bp.f(), bp.mr() + nan_penalty * bp.nNaNs(), bp.x)                          # :338
            return p.point(best[2], f=best[0], mr=best[1], mrName=None, mrInd=0)

        rec = dev.init()
        count_population_evals()
        Best = best_of(rec)                      # de_oo.py:153-155
        total = p.maxIter + 10
        done = 0
        timed = math.isfinite(p.maxTime) or math.isfinite(p.maxCPUTime)
        while done < total:
            k = min(batch, total - done)
            k = min(k, max(1, p.maxIter - 1 - p.iter + 1))
            k = min(k, max(1, -(-(p.maxFunEvals - p.nEvals['f']) // NP)))
            if timed:
                k = min(k, 8)
            for rec in dev.step(k):
                count_population_evals()
                Old_best, Best = Best, best_of(rec)          # :183,:256-257
                if Old_best.betterThan(Best):                # :285-286
                    Best = Old_best
                p.iterfcn(Best)
                if p.istop:
                    return
            done += k
