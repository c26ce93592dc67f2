"""istop codes and the kernel stop criteria evaluated once per iteration.

Reference: kernel/setDefaultIterFuncs.py:5-28 (codes), :40-57 (default set for NonLin problems),
:59-118 (the criteria), :122-132 (denying functions).  Each criterion takes the problem and returns
False or (True, message).
"""
import numpy as np

SMALL_DF = 2
SMALL_DELTA_X = 3
SMALL_DELTA_F = 4
FVAL_IS_ENOUGH = 10
MAX_NON_SUCCESS = 11
USER_DEMAND_STOP = 80
BUTTON_ENOUGH_HAS_BEEN_PRESSED = 88
SOLVED_WITH_UNIMPLEMENTED_OR_UNKNOWN_REASON = 1000
UNDEFINED = 0
IS_NAN_IN_X = -4
IS_LINE_SEARCH_FAILED = -5
IS_MAX_ITER_REACHED = -7
IS_MAX_CPU_TIME_REACHED = -8
IS_MAX_TIME_REACHED = -9
IS_MAX_FUN_EVALS_REACHED = -10
IS_ALL_VARS_FIXED = -11
FAILED_TO_OBTAIN_MOVE_DIRECTION = -13
USER_DEMAND_EXIT = -99
FAILED_WITH_UNIMPLEMENTED_OR_UNKNOWN_REASON = -1000

LIMIT_CODES = (IS_MAX_ITER_REACHED, IS_MAX_CPU_TIME_REACHED, IS_MAX_TIME_REACHED, IS_MAX_FUN_EVALS_REACHED)


def stopcase(arg):
    istop = arg.istop if hasattr(arg, 'istop') else arg
    if istop > 0:
        return 1
    if istop in (0,) + LIMIT_CODES:
        return 0
    return -1


def small_df(p):
    df = getattr(p, '_df', None)
    if df is None:
        return False
    if p.norm(df) >= p.gtol or not np.all(np.isfinite(df)) or not p.isFeas(p.iterValues.x[-1]):
        return False
    return (True, '|| gradient F(X[k]) || < gtol')


def small_deltaX(p):
    if p.iter == 0:
        return False
    dx = p.iterValues.x[-1] - p.iterValues.x[-2]
    if p.scale is not None:
        dx = dx * p.scale
    return False if p.norm(dx) >= p.xtol else (True, '|| X[k] - X[k-1] || < xtol')


def small_deltaF(p):
    if p.iter == 0:
        return False
    f1, f0 = p.iterValues.f[-1], p.iterValues.f[-2]
    if np.isnan(f1) or np.isnan(f0) or p.norm(f1 - f0) >= p.ftol:
        return False
    return (True, '|| F[k] - F[k-1] || < ftol')


def isEnough(p):
    if p.fEnough > np.asarray(p.Fk).item() and p.isFeas(p.xk):
        return (True, 'fEnough has been reached')
    return False


def isNanInX(p):
    return (True, 'NaN in X[k] coords has been obtained') if np.any(np.isnan(p.xk)) else False


def isMaxIterReached(p):
    # iteration numbering starts from zero, and iteration 0 is spent on x0
    return (True, 'Max Iter has been reached') if p.iter >= p.maxIter - 1 else False


def isMaxCPUTimeReached(p):
    if p.iterCPUTime[-1] < p.maxCPUTime + p.cpuTimeElapsedForPlotting[-1]:
        return False
    return (True, 'max CPU time limit has been reached')


def isMaxTimeReached(p):
    if p.currtime - p.timeStart < p.maxTime + p.timeElapsedForPlotting[-1]:
        return False
    return (True, 'max time limit has been reached')


def isMaxFunEvalsReached(p):
    return (True, 'max objfunc evals limit has been reached') if p.nEvals['f'] >= p.maxFunEvals else False


def default_criteria(class_name):
    """Ordered like the reference's dict (insertion order = evaluation order, ooIter.py:68)."""
    d = {SMALL_DF: small_df, SMALL_DELTA_X: small_deltaX, SMALL_DELTA_F: small_deltaF,
         FVAL_IS_ENOUGH: isEnough, IS_NAN_IN_X: isNanInX, IS_MAX_ITER_REACHED: isMaxIterReached,
         IS_MAX_CPU_TIME_REACHED: isMaxCPUTimeReached, IS_MAX_TIME_REACHED: isMaxTimeReached}
    if class_name == 'NonLin':
        d[IS_MAX_FUN_EVALS_REACHED] = isMaxFunEvalsReached
    return d


def _min_iter(p):
    return p.iter >= p.minIter


def _min_time(p):
    return p.currtime - p.timeStart > p.minTime + p.timeElapsedForPlotting[-1]


def _min_cputime(p):
    return p.iterCPUTime[-1] >= p.minCPUTime + p.cpuTimeElapsedForPlotting[-1]


def _min_funevals(p):
    return p.minFunEvals == 0 or ('f' in p.nEvals and p.nEvals['f'] >= p.minFunEvals)


def denying_stop_funcs():
    return {_min_iter: 'min iter is not reached yet', _min_time: 'min time is not reached yet',
            _min_cputime: 'min cputime is not reached yet',
            _min_funevals: 'min objective function evaluations nuber is not reached yet'}
