"""`galileo` solver plugin, CUDA backend: same plugin contract, parameters and outer protocol as the reference's
openopt/solvers/Standalone/galileo_oo.py:6-92 (a genetic algorithm: roulette selection, one-point crossover,
re-initialising mutation, steady-state replacement without duplicates), with the population on the GPU.

What stays on the host (this file): the loop bound ``range(p.maxIter+1)`` (:73), ``p.xf / p.ff`` tracking (:87-88),
``p.iterfcn(x, fval)`` once per iteration and the return on ``p.istop`` (:89-92), replayed over batches of iterations
the library has already run (speculative iterations past the stop are discarded, as in the de plugin).  What moves
to the device (liboode, include/ooga.h): ``Population.prepPopulation`` and everything inside the loop (:75-84).

The objective must be a registered device objective (openopt_b200.objectives); there is no CPU fallback.
"""
import math

import numpy as np

from ... import _lib
from ...host.registry import baseSolver
from ...host.stopping import SMALL_DELTA_F, SMALL_DELTA_X
from ...objectives import DeviceObjective


class galileo(baseSolver):
    __name__ = 'galileo'
    __license__ = "BSD"
    __authors__ = ("B200-native CUDA backend of OpenOpt's galileo (algorithm: Donald Goodman's GA library as "
                   "connected to OpenOpt by Dmitrey)")
    __alg__ = "Genetic Algorithm, same as ga2, the C++ canonical genetic algorithm lib, also by Donald Goodman"
    iterfcnConnected = True
    __homepage__ = ''
    __optionalDataThatCanBeHandled__ = ['lb', 'ub']
    __isIterPointAlwaysFeasible__ = lambda self, p: True
    _requiresFiniteBoxBounds = True

    # ---- the reference's parameters (galileo_oo.py:20-24), same names and defaults ----
    population = 15
    crossoverRate = 1.0       # probability that a selected pair is crossed
    mutationRate = 0.05       # probability that a gene of an entry is re-drawn
    useInteger = False        # integer genes: not available here (see __solver__)

    # ---- backend parameters (new) ----
    seed = 150880             # the reference draws from an unseeded random.Random() (:223)
    rng = 'philox'            # 'philox' (device, counter based) | 'cpython' (random.Random(seed), the reference's generator)
    gensPerCall = 64          # iterations per C call (history is replayed on the host afterwards)
    device = None             # CUDA device ordinal (default: LOCAL_RANK under torch.distributed, else 0); every rank of a
                              # job runs its own full copy (replicas only: the roulette couples every pick to the whole list)

    __info__ = """
        Genetic algorithm (galileo), CUDA (sm_100a) backend.

        requires finite lb, ub: lb <= x <= ub; the objective must be a registered device objective
        (openopt_b200.objectives).

        Parameters (as in OpenOpt's galileo): population, crossoverRate, mutationRate
        (useInteger = True is not available: under Python 3 the reference itself fails on it).
        Backend parameters: seed, rng - 'philox' (default) or 'cpython' (the stream of random.Random(seed):
        same iterates as OpenOpt's galileo with that generator), gensPerCall, device.
        """

    def __init__(self):
        pass

    def __solver__(self, p):
        p.kernelIterFuncs.pop(SMALL_DELTA_X, None)          # galileo_oo.py:29-30
        p.kernelIterFuncs.pop(SMALL_DELTA_F, None)
        p.ff = np.inf                                       # :31

        if self.useInteger:
            # Chromosome.randomInit would call generator.randint(lb[i], ub[i]) with the float entries of p.lb.tolist()
            # (:42-45, :127-133), a TypeError since Python 3.12: nothing to be compatible with
            p.err('the CUDA galileo backend has no useInteger mode (the reference fails on it under Python 3)')
        funcs = p.user.f
        if len(funcs) != 1 or not isinstance(funcs[0], DeviceObjective):
            p.err('the CUDA galileo backend needs a registered device objective (see openopt_b200.objectives); '
                  'arbitrary Python callables cannot run on the GPU and there is no CPU fallback')
        if p.args.f != ():
            p.err('extra objective args are not supported by the CUDA galileo backend')
        if self.rng not in ('philox', 'cpython'):
            p.err('incorrect rng, should be "philox" or "cpython", got ' + str(self.rng))
        obj = funcs[0]
        batch = max(1, int(self.gensPerCall))
        device = self.device
        if device is None:
            from ... import distributed as ddist
            world, _rank, local_rank = ddist.world()
            device = local_rank if world > 1 else 0
        entries = 2 * ((int(self.population) + 1) // 2)      # at most this many evaluations per iteration
        timed = math.isfinite(p.maxTime) or math.isfinite(p.maxCPUTime)
        try:
            dev = _lib.DeviceGalileo(obj.device_id, p.lb, p.ub, population=int(self.population),
                                     crossover_rate=self.crossoverRate, mutation_rate=self.mutationRate, seed=self.seed,
                                     obj_aux=obj.aux(p.n), rng=self.rng, invert=p.invertObjFunc, device=int(device),
                                     max_batch=batch)
        except _lib.LibraryError as e:
            p.err(str(e))
        try:
            dev.init()                                      # P.prepPopulation() :70 and the first evaluate :77
            p.nEvals['f'] += dev.N
            total = p.maxIter + 1                           # :73
            done = 0
            while done < total:
                k = min(batch, total - done)
                k = min(k, max(1, p.maxIter - p.iter))      # do not speculate far past a stop the host can predict
                k = min(k, max(1, -(-(p.maxFunEvals - p.nEvals['f']) // entries)))
                if timed:
                    k = min(k, 8)
                recs, xs = dev.step(k)
                for rec, x in zip(recs, xs):
                    p.nEvals['f'] += rec.n_appended         # the fresh entries replace() had to evaluate (:470-472)
                    fval = np.asarray(rec.best_f, dtype=np.float64)      # asfarray(-P.maxFitness) :86
                    if p.ff > fval:                         # :87-88
                        p.xf, p.ff = np.array(x, dtype=np.float64), fval
                    p.iterfcn(np.array(x, dtype=np.float64), fval)       # :89
                    if p.istop:                             # :91-93
                        self.gpu_launches = dev.gpu_launches()
                        return
                done += k
            self.gpu_launches = dev.gpu_launches()
        except _lib.LibraryError as e:
            p.err(str(e))
        finally:
            dev.close()
