"""Problem classes that drive the de solver: GLP and NLP.

Reference: kernel/baseProblem.py:48-240 (defaults, ctor, solve), :307-316,:623-668 (_prepare, the
non-FuncDesigner branch), :880-1037 (NonLinProblem), kernel/GLP.py:5-48, kernel/NLP.py:6-14,
kernel/residuals.py:22-172, kernel/nonLinFuncs.py:15-317 (the non-derivative path of p.f/p.c/p.h),
kernel/iterPrint.py:20-43.  The FuncDesigner (symbolic) front end, derivative machinery, GUI and
plotting of the reference are outside this package's scope.
"""
import numpy as np
from numpy import inf

from . import stopping
from .iteration import ooIter, iter_print
from .log import (OpenOptException, killThread, oodisp, ooerr, oohint, ooinfo, ooPWarn, oowarn)
from .point import Point


class _Bag:
    """Attribute container (the reference's EmptyClass / autocreate / user)."""


class baseProblem:
    # ---- class-level defaults (kernel/baseProblem.py:52-154) ----
    isObjFunValueASingleNumber = True
    prepared = False
    _baseProblemIsPrepared = False
    name = 'unnamed'
    state = 'init'
    castFrom = ''
    nonStopMsg = ''
    xlabel = 'time'
    plot = False
    show = True
    iter = 0
    cpuTimeElapsed = 0.
    TimeElapsed = 0.
    isFinished = False
    invertObjFunc = False
    nProc = 1
    lastPrintedIter = -1
    iterObjFunTextFormat = '%0.3e'
    finalObjFunTextFormat = '%0.8g'
    debug = 0
    iprint = 10
    maxDistributionSize = 0
    maxIter = 1000
    maxFunEvals = 10000
    maxCPUTime = inf
    maxTime = inf
    maxLineSearch = 500
    xtol = 1e-6
    gtol = 1e-6
    ftol = 1e-6
    contol = 1e-6
    fTol = None
    minIter = 0
    minFunEvals = 0
    minCPUTime = 0.0
    minTime = 0.0
    storeIterPoints = False
    userStop = False
    useSparse = 'auto'
    useAttachedConstraints = False
    x0 = None
    isFDmodel = False
    noise = 0
    showFeas = False
    useScaledResidualOutput = False
    hasLogicalConstraints = False
    A = None
    b = None
    Aeq = None
    beq = None
    scale = None
    goal = None
    showGoal = False
    color = 'b'
    specifier = '-'
    plotOnlyCurrentMinimum = False
    legend = ''
    fixedVars = None
    freeVars = None
    istop = 0
    maxSolutions = 1
    fEnough = -inf
    fOpt = None
    implicitBounds = inf
    convex = 'unknown'
    _linear_objective = False
    allowedGoals = None

    def __init__(self, *args, **kwargs):
        self.err, self.warn, self.info, self.hint = ooerr, oowarn, ooinfo, oohint
        self.pWarn, self.disp = ooPWarn, oodisp
        self.data4TextOutput = ['objFunVal', 'log10(maxResidual)']
        self.nEvals = {}
        expected = getattr(self, 'expectedArgs', None)
        if expected is not None:
            if len(expected) < len(args):
                self.err('Too much arguments for ' + self.probType + ': ' + str(len(args)) +
                         ' are got, at most ' + str(len(expected)) + ' were expected')
            for name, arg in zip(expected, args):
                setattr(self, name, arg)
        self.norm = lambda x, k=2: np.linalg.norm(np.atleast_1d(x), k)
        self.denyingStopFuncs = stopping.denying_stop_funcs()
        self.iterfcn = lambda *a, **kw: ooIter(self, *a, **kw)
        self.graphics = _Bag()
        self.graphics.xlabel = 'time'
        self.graphics.nPointsPlotted = 0
        self.user = _Bag()
        self.F = lambda x: self.objFuncMultiple2Single(self.objFunc(x))
        self.point = lambda *a, **kw: Point(self, *a, **kw)
        self.timeElapsedForPlotting = [0.]
        self.cpuTimeElapsedForPlotting = [0.]
        self.debugmsg = lambda msg: print('OpenOpt debug msg: %s' % msg) if self.debug else None
        self.constraints = []
        self.callback = []
        self.solverParams = _Bag()
        self.userProvided = _Bag()
        self.special = _Bag()
        self.intVars = []
        self.binVars = []
        self.optionalData = []
        if self.allowedGoals is not None:
            if 'min' in self.allowedGoals:
                self.minimize = lambda *a, **kw: _with_goal(self, 'minimum', ('min', 'minimum'), 'minimize', *a, **kw)
            if 'max' in self.allowedGoals:
                self.maximize = lambda *a, **kw: _with_goal(self, 'maximum', ('max', 'maximum'), 'maximize', *a, **kw)
        for key, val in kwargs.items():            # assignScript, kernel/ooMisc.py:201-206
            if key != 'manage':
                setattr(self, key, val)

    def __finalize__(self):
        pass

    def objFunc(self, x):
        return self.f(x)

    def __isFiniteBoxBounded__(self):
        return bool(np.all(np.isfinite(self.ub)) and np.all(np.isfinite(self.lb)))

    def __isNoMoreThanBoxBounded__(self):
        return self.b.size == 0 and self.beq.size == 0 and \
            (self._baseClassName == 'Matrix' or (not self.userProvided.c and not self.userProvided.h))

    def solve(self, *args, **kwargs):
        from .driver import runProbSolver
        return runProbSolver(self, *args, **kwargs)

    def _solve(self, *args, **kwargs):
        self.debug = True
        return self.solve(*args, **kwargs)

    def objFuncMultiple2Single(self, f):
        if np.asarray(f, dtype=np.float64).size != 1:
            self.err('unexpected f size. The function should be redefined in OO child class, inform OO developers')
        return f

    iterPrint = iter_print

    # ---- kernel/baseProblem.py:307-316, 623-668 (non-FuncDesigner branch) ----
    def _prepare(self):
        if self._baseProblemIsPrepared:
            return
        if self.fixedVars is not None or self.freeVars is not None:
            self.err('fixedVars and freeVars are valid for optimization of FuncDesigner models only')
        if self.x0 is None:
            for fn in ('lb', 'ub'):
                if hasattr(self, fn):
                    fv = np.asarray(getattr(self, fn))
                    if np.any(np.isfinite(fv)):
                        self.x0 = np.zeros(fv.size)
                        break
        self.x0 = np.ravel(self.x0)
        if not hasattr(self, 'n'):
            self.n = self.x0.size
        if not hasattr(self, 'lb'):
            self.lb = -inf * np.ones(self.n)
        if not hasattr(self, 'ub'):
            self.ub = inf * np.ones(self.n)
        for fn in ('A', 'Aeq'):
            fv = getattr(self, fn)
            if fv is not None:
                afv = np.asarray(fv, dtype=np.float64) if type(fv) in (list, tuple) else fv
                if afv.ndim > 1:
                    if afv.shape[1] != self.n:
                        self.err('incorrect ' + fn + ' size')
                elif afv.shape != () and afv.shape[0] == self.n:
                    afv = afv.reshape(1, self.n)
                setattr(self, fn, afv)
            else:
                setattr(self, fn, np.zeros((0, self.n)))
        self._baseProblemIsPrepared = True

    # ---- residuals (kernel/residuals.py:22-172) ----
    def _get_nonLinInEq_residuals(self, x):
        return self.c(x) if self.userProvided.c else np.array([])

    def _get_nonLinEq_residuals(self, x):
        return self.h(x) if self.userProvided.h else np.array([])

    def _get_AX_Less_B_residuals(self, x):
        if self.A is not None and np.size(self.A) > 0:
            return np.dot(x, np.asarray(self.A).T) - self.b if np.ndim(x) > 1 else np.dot(self.A, x) - self.b
        return np.array([])

    def _get_AeqX_eq_Beq_residuals(self, x):
        if self.Aeq is not None and np.size(self.Aeq) > 0:
            return np.dot(x, np.asarray(self.Aeq).T) - self.beq if np.ndim(x) > 1 else np.dot(self.Aeq, x) - self.beq
        return np.array([])

    def getMaxResidual(self, x, retAll=False):
        """(max residual, constraint family, index) -- kernel/residuals.py:77-115."""
        fields = (('c', self._get_nonLinInEq_residuals(x) if self._baseClassName == 'NonLin' else 0, False),
                  ('lin_ineq', self._get_AX_Less_B_residuals(x), False),
                  ('lb', self.lb - x, False), ('ub', x - self.ub, False),
                  ('h', self._get_nonLinEq_residuals(x) if self._baseClassName == 'NonLin' else 0, True),
                  ('lin_eq', self._get_AeqX_eq_Beq_residuals(x), True))
        r, fname, ind = 0, None, None
        for field, fv, is_eq in fields:
            fv = np.asarray(fv).flatten()
            if fv.size > 0:
                if is_eq:
                    fv = np.abs(fv)
                ind_max = np.argmax(fv)
                val_max = fv[ind_max]
                if r < val_max:
                    r, ind, fname = val_max, ind_max, field
        return (r, fname, ind) if retAll else r

    def isFeas(self, x):
        if np.any(np.isnan(self._get_nonLinEq_residuals(x))) or np.any(np.isnan(self._get_nonLinInEq_residuals(x))):
            return False
        finite_x = np.all(np.isfinite(x))
        contol_ok = self.getMaxResidual(x) <= self.contol
        return bool(contol_ok and finite_x and np.all(np.isfinite(self.objFunc(x))))


def _with_goal(p, goal, same, fname, *args, **kwargs):
    if 'goal' in kwargs:
        if kwargs['goal'] in same:
            p.warn("you shouldn't pass 'goal' to the function '%s'" % fname)
        else:
            p.err('ambiguous goal has been requested: function "%s", goal: %s' % (fname, kwargs['goal']))
    p.goal = goal
    from .driver import runProbSolver
    return runProbSolver(p, *args, **kwargs)


class NonLinProblem(baseProblem):
    _baseClassName = 'NonLin'
    diffInt = 1.5e-8
    c = None
    h = None
    maxViolation = 1e-2
    JacobianApproximationStencil = 1

    def __init__(self, *args, **kwargs):
        baseProblem.__init__(self, *args, **kwargs)
        if not hasattr(self, 'args'):
            self.args = _Bag()
            self.args.f = self.args.c = self.args.h = ()
        self.functype = {}
        self.kernelIterFuncs = stopping.default_criteria('NonLin')

    def _makeCorrectArgs(self):
        a = self.args
        if not (hasattr(a, 'f') and hasattr(a, 'c') and hasattr(a, 'h')):
            tmp, self.args = a, _Bag()
            self.args.f = self.args.c = self.args.h = tmp
        for j in ('f', 'c', 'h'):
            v = getattr(self.args, j)
            if type(v) != tuple:
                setattr(self.args, j, (v,))

    def _Prepare(self):
        """kernel/baseProblem.py:960-1029: wrap the user's f/c/h, zero the evaluation counters."""
        baseProblem._prepare(self)
        if np.asarray(self.implicitBounds).size == 1:
            self.implicitBounds = sorted([-self.implicitBounds, self.implicitBounds])
        if self.prepared:
            return
        self._makeCorrectArgs()
        for s in ('f', 'df', 'd2f', 'c', 'dc', 'd2c', 'h', 'dh', 'd2h'):
            self.nEvals[s] = 0
            attr = getattr(self, s, None)
            provided = attr is not None and (not isinstance(attr, (list, tuple)) or len(attr) != 0)
            setattr(self.userProvided, s, provided)
            if provided:
                setattr(self.user, s, attr if type(attr) in (list, tuple) else (attr,))
            if len(s) == 1:
                setattr(self, s, lambda x, IND=None, kind=s: self.wrapped_func(x, IND, kind))
            elif provided:
                self.err('derivatives (%s) are outside the scope of this package: the de solver is derivative-free' % s)
        self.diffInt = np.ravel(self.diffInt)
        if self.scale is not None:
            self.scale = np.ravel(self.scale)
        for s in ('c', 'h', 'f'):
            if not getattr(self.userProvided, s):
                setattr(self, 'n' + s, 0)
            else:
                self._set_funcs_number(s)
        self.prepared = True

    def _set_funcs_number(self, kind):
        """kernel/ooMisc.py:208-255: one direct call of the user function at x0 fixes n<kind>
        (not counted in nEvals)."""
        fv = getattr(self.user, kind)
        args = getattr(self.args, kind)
        if len(fv) == 1:
            self.functype[kind] = 'single func'
            tmp = fv[0](*(self.x0,) + args)
            n = sum(np.asarray(e).size for e in tmp) if isinstance(tmp, (list, tuple)) else np.asarray(tmp).size
        else:
            n = sum(np.asarray(func(*(self.x0,) + args)).size for func in fv)
            self.functype[kind] = 'some funcs R^nvars -> R'
        setattr(self, 'n' + kind, n)

    def wrapped_func(p, x, IND, kind):
        """p.f / p.c / p.h (kernel/nonLinFuncs.py:15-317, non-FuncDesigner, non-derivative path):
        shape check, row-by-row evaluation of 2-D input (:190-200), sign flip for goal='max'
        (:297-298), evaluation counting (:300-304)."""
        if not getattr(p.userProvided, kind):
            return np.array([])
        if p.istop == stopping.USER_DEMAND_EXIT:
            if p.solver.useStopByException:
                raise killThread
            return np.nan
        if IND is not None:
            p.err('evaluating a subset of functions (IND) is outside the scope of this package')
        funcs = getattr(p.user, kind)
        args = getattr(p.args, kind)
        x = np.atleast_1d(x)
        if x.shape[0] != p.n and (x.ndim < 2 or x.shape[1] != p.n):
            p.err('x with incorrect shape passed to non-linear function')
        n_vectors = 1 if (x.ndim <= 1 or x.shape[0] == 1) else x.shape[0]

        def call_all(row):
            tmp = [fun(*(row,) + args) for fun in funcs]
            if len(tmp) == 1:
                return np.hstack(tmp[0]) if isinstance(tmp[0], (list, tuple)) else tmp[0]
            return np.hstack(tmp)

        if n_vectors == 1:
            r = call_all(x)
        else:
            r = np.hstack([call_all(x[i]) for i in range(n_vectors)])
        if kind == 'f' and p.isObjFunValueASingleNumber and not np.isscalar(r):
            if np.prod(np.shape(r)) > 1 and n_vectors == 1:
                p.err('implicit summation in objective is no longer available to prevent possibly hidden bugs')
        if kind != 'f' and n_vectors != 1:
            r = np.asarray(r).reshape(n_vectors, int(np.size(r) / n_vectors))
        if n_vectors == 1:
            r = np.asarray(r, dtype=np.float64).flatten() if not np.isscalar(r) else np.atleast_1d(r)
        if p.invertObjFunc and kind == 'f':
            r = -r
        p.nEvals[kind] += n_vectors
        return r

    def _isUnconstrained(self):
        return self.b.size == 0 and self.beq.size == 0 and not self.userProvided.c and not self.userProvided.h \
            and (len(self.lb) == 0 or np.all(np.isinf(self.lb))) and (len(self.ub) == 0 or np.all(np.isinf(self.ub)))


class GLP(NonLinProblem):
    """Global problem: f(x) -> min|max, lb <= x <= ub (kernel/GLP.py:5-48)."""
    probType = 'GLP'
    _optionalData = ['lb', 'ub', 'c', 'A', 'b']
    expectedArgs = ['f', 'x0']
    allowedGoals = ['minimum', 'min', 'maximum', 'max']
    goal = 'minimum'
    showGoal = False
    plotOnlyCurrentMinimum = True
    _currentBestPoint = None
    _nonSuccessCounter = 0
    maxNonSuccess = 15

    def __init__(self, *args, **kwargs):
        NonLinProblem.__init__(self, *args, **kwargs)

        def maxNonSuccess(p):
            # kernel/GLP.py:23-36: stop after maxNonSuccess iterations without a better point
            new = p.point(p.xk)
            if self._currentBestPoint is None:
                self._currentBestPoint = new
                return False
            if new.betterThan(self._currentBestPoint):
                self._currentBestPoint = new
                self._nonSuccessCounter = 0
                return False
            self._nonSuccessCounter += 1
            if self._nonSuccessCounter > self.maxNonSuccess:
                return (True, 'Non-Success Number > maxNonSuccess = ' + str(self.maxNonSuccess))
            return False

        self.kernelIterFuncs[stopping.MAX_NON_SUCCESS] = maxNonSuccess
        if 'lb' in kwargs:
            self.n = len(kwargs['lb'])
        elif 'ub' in kwargs:
            self.n = len(kwargs['ub'])
        if hasattr(self, 'n'):
            if not hasattr(self, 'lb'):
                self.lb = -inf * np.ones(self.n)
            if not hasattr(self, 'ub'):
                self.ub = inf * np.ones(self.n)
            if 'x0' not in kwargs:
                self.x0 = (np.asarray(self.lb) + np.asarray(self.ub)) / 2.0


class NLP(NonLinProblem):
    """Nonlinear problem (kernel/NLP.py:6-14); reaches the de solver when box-bounded."""
    _optionalData = ['A', 'Aeq', 'b', 'beq', 'lb', 'ub', 'c', 'h']
    expectedArgs = ['f', 'x0']
    probType = 'NLP'
    allowedGoals = ['minimum', 'min', 'maximum', 'max']
    showGoal = True

    def __init__(self, *args, **kwargs):
        self.goal = 'minimum'
        NonLinProblem.__init__(self, *args, **kwargs)


__all__ = ['GLP', 'NLP', 'OpenOptException']
