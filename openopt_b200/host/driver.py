"""p.solve(): resolve the solver, route kwargs, prepare, run, finalise, build the result.

Reference: kernel/runProbSolver.py:45-379 (runProbSolver), :382-418 (final text), :430-499
(OpenOptResult), kernel/check.py:3-22.  Converter strings ('nlp:ralg'), FuncDesigner models, plotting
and the GUI manager are outside the scope of this package.
"""
import time

import numpy as np

from . import stopping
from .log import OpenOptException, isSolved
from .registry import getSolverFromStringName
from .. import OPENOPT_API_VERSION as version       # runProbSolver.py:213 prints the OpenOpt release

ConTolMultiplier = 0.8

try:
    import setproctitle
except ImportError:            # optional, exactly as in the reference
    setproctitle = None


class _Bag:
    pass


def check(p):
    """kernel/check.py:3-22: goal allowed?  can the solver handle the optional data present?"""
    if p.allowedGoals is not None and p.goal is not None and p.goal not in p.allowedGoals:
        p.err('goal ' + p.goal + ' is not available for the ' + p.probType + ' class (at least not implemented yet)')
    for fn in p._optionalData:
        attr = getattr(p, fn, None)
        if attr is None or (isinstance(attr, (list, tuple)) and len(attr) == 0) \
                or fn in p.solver.__optionalDataThatCanBeHandled__:
            continue
        if (callable(attr) and getattr(p.userProvided, fn)) or \
                (type(attr) in (set, dict, list, tuple) and len(attr) != 0) or \
                (not callable(attr) and np.asarray(attr).size > 0 and np.any(np.isfinite(attr))):
            p.err('the solver ' + p.solver.__name__ + ' cannot handle ' + "'" + fn + "' data")
    return 0


def runProbSolver(p, solver_str_or_instance=None, *args, **kwargs):
    if len(args) != 0:
        p.err('unexpected args for p.solve()')
    if hasattr(p, 'was_involved'):
        p.err("You can't run same prob instance for twice. \n    Please reassign prob struct.")
    p.was_involved = True

    if solver_str_or_instance is None:
        if hasattr(p, 'solver'):
            solver_str_or_instance = p.solver
        elif 'solver' in kwargs:
            solver_str_or_instance = kwargs['solver']
    if solver_str_or_instance is None:
        p.err('you should provide name of solver')
    if type(solver_str_or_instance) is str:
        if ':' in solver_str_or_instance:
            p.err('problem converters ("%s") are outside the scope of this package' % solver_str_or_instance)
        p.solver = getSolverFromStringName(p, solver_str_or_instance)
    else:
        if not getattr(solver_str_or_instance, 'isInstalled', True):
            p.err('solver %s seems to be uninstalled yet' % solver_str_or_instance.__name__)
        p.solver = solver_str_or_instance
        for key, value in solver_str_or_instance.fieldsForProbInstance.items():
            setattr(p, key, value)
    p.isConverterInvolved = False

    old_err = np.seterr(all='ignore')
    if 'debug' in kwargs:
        p.debug = kwargs['debug']
    # an attribute present on both the problem instance and the solver instance goes to the solver
    for elem in set(p.__dict__).intersection(set(p.solver.__dict__)):
        setattr(p.solver, elem, getattr(p, elem))

    solver = p.solver.__solver__

    # kwargs: solver attribute if the solver has it, else problem attribute, else a warning (:99-107)
    for key, value in kwargs.items():
        if hasattr(p.solver, key):
            setattr(p.solver, key, value)
        elif hasattr(p, key):
            setattr(p, key, value)
        else:
            p.warn('incorrect parameter for prob.solve(): "' + str(key) + '" - will be ignored (this one has been '
                   'not found in neither prob nor ' + p.solver.__name__ + ' solver parameters)')

    iv = p.iterValues = _Bag()
    p.iterCPUTime, p.iterTime = [], []
    iv.x, iv.f, iv.r, iv.rt, iv.ri = [], [], [], [], []
    p.solutions = []
    if p._baseClassName == 'NonLin':
        iv.nNaNs = []
    if p.goal in ('max', 'maximum'):
        p.invertObjFunc = True
    p.advanced = _Bag()
    p.istop = 0
    p.iter = 0
    p.graphics.nPointsPlotted = 0
    p.finalIterFcnFinished = False
    p.msg = ''
    if type(p.callback) not in (list, tuple):
        p.callback = [p.callback]
    if hasattr(p, 'xlabel'):
        p.graphics.xlabel = p.xlabel
    if p.graphics.xlabel == 'nf':
        iv.nf = []

    T, Cpu = time.time(), time.process_time()
    p._Prepare()
    T, Cpu = time.time() - T, time.process_time() - Cpu
    if T > 1 or Cpu > 1:
        p.disp('Initialization: Time = %0.1f CPUTime = %0.1f' % (T, Cpu))

    for fn in ('FunEvals', 'Iter', 'Time', 'CPUTime'):
        lo, hi = getattr(p, 'min' + fn, None), getattr(p, 'max' + fn, None)
        if lo is not None and hi is not None and hi < lo:
            p.warn('min' + fn + ' (' + str(lo) + ') exceeds ' + 'max' + fn + '(' + str(hi) + '), setting latter to former')
            setattr(p, 'max' + fn, lo)
    for fn in ('maxFunEvals', 'maxIter'):
        setattr(p, fn, int(getattr(p, fn)))

    if hasattr(p, 'x0'):
        p.x0 = np.atleast_1d(np.array(p.x0, dtype=np.float64))
    for fn in ('lb', 'ub', 'b', 'beq'):
        if hasattr(p, fn):
            fv = getattr(p, fn)
            if fv is not None:
                if isinstance(fv, map):
                    p.err("Python3 incompatibility with previous versions: you can't use 'map' here, use rendered value instead")
                setattr(p, fn, np.array(fv, dtype=np.float64).flatten())
            else:
                setattr(p, fn, np.array([], dtype=np.float64))

    if p.solver._requiresFiniteBoxBounds:
        ind1, ind2 = np.isinf(p.lb), np.isinf(p.ub)
        if np.isscalar(p.implicitBounds):
            p.implicitBounds = (-p.implicitBounds, p.implicitBounds)
        ib = p.implicitBounds
        p.lb[ind1] = ib[0] if np.asarray(ib[0]).size == 1 else ib[0][ind1]
        p.ub[ind2] = ib[1] if np.asarray(ib[1]).size == 1 else ib[0][ind2]

    p.stopdict = {}
    for s in ('b', 'beq'):
        if hasattr(p, s):
            setattr(p, 'n' + s, len(getattr(p, s)))
    p.isUC = p._isUnconstrained()

    feas = p.solver.__isIterPointAlwaysFeasible__
    if (feas if type(feas) == bool else feas(p)):
        if p.data4TextOutput[-1] == 'log10(maxResidual)':
            p.data4TextOutput = p.data4TextOutput[:-1]
    elif p.useScaledResidualOutput:
        p.data4TextOutput[-1] = 'log10(MaxResidual/ConTol)'
    if p.showFeas and p.data4TextOutput[-1] != 'isFeasible':
        p.data4TextOutput.append('isFeasible')

    if not p.solver.iterfcnConnected:
        p.kernelIterFuncs.pop(stopping.SMALL_DELTA_X, None)
        p.kernelIterFuncs.pop(stopping.SMALL_DELTA_F, None)

    p.primalConTol = p.contol
    if not p.solver.__name__.startswith('interalg'):
        p.contol *= ConTolMultiplier
    p.timeStart = time.time()
    p.cpuTimeStart = time.process_time()
    p.plotOnlyCurrentMinimum = p.__isNoMoreThanBoxBounded__()

    if p.iprint >= 0:
        p.disp('\n' + '-' * 25 + ' OpenOpt %s ' % version + '-' * 25)
        s = 'solver: ' + p.solver.__name__ + '   problem: ' + p.name + '    type: %s' % p.probType
        if p.showGoal:
            s += '   goal: ' + p.goal
        p.disp(s)
    p.extras = {}

    try:
        if check(p):
            p.err('prob check results: ERRORS!')
        p.iterfcn(p.x0)                                   # iteration 0 (runProbSolver.py:259)
        original_title = None
        if setproctitle is not None:
            try:
                original_title = setproctitle.getproctitle()
                if original_title.startswith('OpenOpt-'):
                    original_title = None
                else:
                    setproctitle.setproctitle('OpenOpt-' + p.solver.__name__ + '-' + p.name)
            except Exception:
                original_title = None
        try:
            solver(p)                                     # the hot path (runProbSolver.py:276)
        finally:
            if setproctitle is not None and original_title is not None:
                setproctitle.setproctitle(original_title)
    except isSolved:
        if p.istop == 0:
            p.istop = 1000
    finally:
        np.seterr(**old_err)
    p.contol = p.primalConTol

    # ---- finalisation (runProbSolver.py:301-365) ----
    if hasattr(p, '_bestPoint') and not np.any(np.isnan(p._bestPoint.x)):
        p.iterfcn(p._bestPoint)
    if not hasattr(p, 'xf') and not hasattr(p, 'xk'):
        p.xf = p.xk = np.ones(p.n) * np.nan
    if hasattr(p, 'xf') and (not hasattr(p, 'xk') or np.array_equal(p.xk, p.x0)):
        p.xk = p.xf
    if not hasattr(p, 'xf') or np.all(np.isnan(p.xf)):
        p.xf = p.xk
    p.isFeasible = bool(p.isFeas(p.xf))
    p.isFinished = True
    p.ff = p.fk = p.objFunc(p.xk)                         # host re-evaluation at xf (:326)
    if type(p.ff) == np.ndarray and p.ff.size == 1:
        p.ff = p.fk = np.asarray(p.ff).item()
    if p.invertObjFunc:
        p.fk, p.ff = -p.fk, -p.ff
    if np.asarray(p.ff).size > 1:
        p.ff = p.objFuncMultiple2Single(p.fk)
    if type(p.xf) in (list, tuple) or np.isscalar(p.xf):
        p.xf = np.asarray(p.xf)
    p.xf = p.xf.flatten()
    p.rf = p.getMaxResidual(p.xf)
    if not p.isFeasible and p.istop > 0:
        p.istop = -100 - p.istop / 1000.0
    if p.istop == 0 and p.iter >= p.maxIter:
        p.istop, p.msg = stopping.IS_MAX_ITER_REACHED, 'Max Iter has been reached'
    p.stopcase = stopping.stopcase(p)
    p.xk, p.rk = p.xf, p.rf
    if p.invertObjFunc:
        p.fk = -p.ff
        p.iterfcn(p.xf, -p.ff, p.rf)
    else:
        p.fk = p.ff
        p.iterfcn(p.xf, p.ff, p.rf)
    p.__finalize__()
    if not p.storeIterPoints:
        delattr(p.iterValues, 'x')
    r = OpenOptResult(p)
    p.invertObjFunc = False
    finalTextOutput(p, r)
    return r


def finalTextOutput(p, r):
    if p.iprint < 0:
        return
    p.disp('istop: ' + str(r.istop) + (' (' + p.msg + ')' if len(p.msg) else ''))
    p.disp('Solver:   Time Elapsed = ' + str(r.elapsed['solver_time']) + ' \tCPU Time Elapsed = ' +
           str(r.elapsed['solver_cputime']))
    rMsg = ('max(residuals/requiredTolerances) = %g' % (r.rf / p.contol)) if p.useScaledResidualOutput \
        else ('MaxResidual = %g' % r.rf)
    if not p.isFeasible:
        p.disp('NO FEASIBLE SOLUTION has been obtained (%s, objFunc = %0.8g)' % (rMsg, r.ff))
    else:
        msg = 'objFunValue: ' + (p.finalObjFunTextFormat % r.ff)
        if not p.isUC:
            msg += ' (feasible, %s)' % rMsg
        p.disp(msg)


class OpenOptResult:
    """Result fields of kernel/runProbSolver.py:447-499."""

    def __call__(self, *args):
        raise OpenOptException('Is callable for FuncDesigner models only')

    def __init__(self, p):
        self.rf = np.asarray(p.rf).item()
        self.ff = np.asarray(p.ff).item()
        self.isFDmodel = p.isFDmodel
        self.probType = p.probType
        self.xf = p.xf
        self.elapsed = {'solver_time': round(100.0 * (time.time() - p.timeStart)) / 100.0,
                        'solver_cputime': round(100.0 * (time.process_time() - p.cpuTimeStart)) / 100.0,
                        'plot_time': 0, 'plot_cputime': 0}
        for fn in ('ff', 'istop', 'duals', 'isFeasible', 'msg', 'stopcase', 'iterValues', 'special', 'extras',
                   'solutions'):
            if hasattr(p, fn):
                setattr(self, fn, getattr(p, fn))
        if hasattr(p.solver, 'innerState'):
            self.extras['innerState'] = p.solver.innerState
        self.solverInfo = {fn: getattr(p.solver, '__' + fn + '__')
                           for fn in ('homepage', 'alg', 'authors', 'license', 'info', 'name')}
        self.evals = {key: (val if type(val) == int else round(val * 10) / 10.0) for key, val in p.nEvals.items()}
        self.evals['iter'] = p.iter
