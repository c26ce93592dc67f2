"""The per-iteration callback p.iterfcn and the text table.

Reference: kernel/ooIter.py:22-122 (ooIter), kernel/baseSolver.py:35-111 (__decodeIterFcnArgs__,
here `decode_iterfcn_args`, also reachable as solver.__decodeIterFcnArgs__), kernel/iterPrint.py:20-43.
Called once per generation by the de plugin with a Point (de_oo.py:288); the driver calls it with
(x) at iteration 0 and (x, f, r) at the end (kernel/runProbSolver.py:259,354-360).
"""
import time

import numpy as np

from . import stopping
from .log import isSolved
from .point import Point


def decode_iterfcn_args(p, *args, **kwargs):
    """Set p.xk, p.fk, p.rk (+ rtk, rik, nNaNs) and append to p.iterValues."""
    f_given = True
    if len(args) > 0 and isinstance(args[0], Point):
        if len(args) != 1:
            p.err('incorrect iterfcn args, if you see this contact OO developers')
        point = args[0]
        p.xk, p.fk = point.x, point.f()
        p.rk, p.rtk, p.rik = point.mr(True)
        p.nNaNs = point.nNaNs()
        if p.solver._requiresBestPointDetection and (p.iter == 0 or point.betterThan(p._bestPoint)):
            p._bestPoint = point
    else:
        if len(args) > 0:
            p.xk = args[0]
        elif 'xk' in kwargs:
            p.xk = kwargs['xk']
        elif not hasattr(p, 'xk'):
            p.err('iterfcn must get x value, if you see it inform oo developers')
        if p._baseClassName == 'NonLin':
            C, H = p.c(p.xk), p.h(p.xk)
            p.nNaNs = int(np.isnan(C).sum() + np.isnan(H).sum())
        if p.solver._requiresBestPointDetection:
            cur = p.point(p.xk)
            if p.iter == 0 or cur.betterThan(p._bestPoint):
                p._bestPoint = cur
        if len(args) > 1:
            p.fk = args[1]
        elif 'fk' in kwargs:
            p.fk = kwargs['fk']
        else:
            f_given = False
        if len(args) > 2:
            p.rk = args[2]
        elif 'rk' in kwargs:
            p.rk = kwargs['rk']
        else:
            p.rk, p.rtk, p.rik = p.getMaxResidual(p.xk, True)

    iv = p.iterValues
    iv.r.append(np.ravel(p.rk)[0])
    p.rk, p.rtk, p.rik = p.getMaxResidual(p.xk, True)
    iv.rt.append(p.rtk)
    iv.ri.append(p.rik)
    if p._baseClassName == 'NonLin':
        iv.nNaNs.append(p.nNaNs)
    iv.x.append(np.copy(p.xk))
    if not p.storeIterPoints and len(iv.x) > 2:
        iv.x.pop(0)
    if not f_given:
        p.Fk = p.F(p.xk)
        p.fk = np.copy(p.Fk)
    elif np.asarray(p.fk).size > 1:
        p.Fk = p.objFuncMultiple2Single(np.asarray(p.fk))
    else:
        p.Fk = p.fk
    v = np.ravel(p.Fk)[0]
    iv.f.append(-v if p.invertObjFunc else v)
    if not np.isscalar(p.fk) and p.fk.size == 1:
        p.fk = np.asarray(p.fk).item()


def ooIter(p, *args, **kwargs):
    if p.finalIterFcnFinished:
        return
    if not hasattr(p, 'timeStart'):
        return
    p.currtime = time.time()
    if not p.iter:
        p.lastDrawTime = p.currtime
        p.lastDrawIter = 0

    if not p.isFinished or len(p.iterValues.f) == 0:
        p.solver.__decodeIterFcnArgs__(p, *args, **kwargs)
        same_as_last = hasattr(p, 'xk_prev') and np.array_equal(p.xk, p.xk_prev)
        p.xk_prev = p.xk.copy()
        if p.graphics.xlabel == 'nf':
            p.iterValues.nf.append(p.nEvals['f'])
        p.iterCPUTime.append(time.process_time() - p.cpuTimeStart)
        p.iterTime.append(p.currtime - p.timeStart)

        population_based = p.probType in ('GLP', 'MILP') or p.solver.__name__ in ('de', 'pswarm', 'interalg')
        if not population_based and ((p.iter == 1 and np.array_equal(p.xk, p.x0)) or same_as_last):
            # a repeated point is dropped from the history for trajectory solvers (ooIter.py:56-65)
            for lst in list(vars(p.iterValues).values()) + [p.iterTime, p.iterCPUTime]:
                if type(lst) == list:
                    lst.pop(-1)
            if not (p.isFinished and same_as_last):
                return

        if not p.userStop and (not same_as_last or population_based):
            for key, fun in p.kernelIterFuncs.items():
                r = fun(p)
                if r is not False:
                    p.stopdict[key] = True
                    if p.istop == 0 or key not in stopping.LIMIT_CODES:
                        p.istop = key
                        p.msg = r[1] if type(r) == tuple else 'unkown, if you see the message inform openopt developers'
            if stopping.IS_NAN_IN_X in p.stopdict:
                pass
            elif stopping.SMALL_DELTA_X in p.stopdict and np.array_equal(p.iterValues.x[-1], p.iterValues.x[-2]):
                pass
            else:
                p.nonStopMsg = ''
                for fun, why in p.denyingStopFuncs.items():
                    if not fun(p):
                        p.istop, p.stopdict, p.msg, p.nonStopMsg = 0, {}, '', why
                        break
                for fun in p.callback:
                    r = fun(p)
                    if r is None:
                        p.err('user-defined callback function returned None, that is forbidden, '
                              'see /doc/userCallback.py for allowed return values')
                    if r not in [0, False]:
                        if r in [True, 1]:
                            p.istop = stopping.USER_DEMAND_STOP
                        elif np.isscalar(r):
                            p.istop, p.msg = r, 'user-defined'
                        else:
                            p.istop, p.msg = r[0], r[1]
                        p.stopdict[p.istop] = True
                        p.userStop = True
    if not p.solver.properTextOutput:
        p.iterPrint()
    if p.plot:
        # Graphics are outside this package's scope: behave like the reference on a machine without pylab
        # (kernel/oographics.py:46-52, reached from ooIter.py:105-108): warn once and turn them off.
        p.pWarn('to use OpenOpt graphics you need pylab (Python module) installed. Turning graphics off...')
        p.plot = 0
    p.iter += 1
    if p.isFinished:
        p.finalIterFcnFinished = True
    if p.istop and p.istop != 1000 and not p.solver.iterfcnConnected and not p.isFinished and p.solver.useStopByException:
        raise isSolved


_DELIM = '   '


def _col_objFunVal(p):
    return p.iterObjFunTextFormat % (-p.fk if p.invertObjFunc else p.fk)


_COLUMNS = {
    'objFunVal': _col_objFunVal,
    'log10(maxResidual)': lambda p: '%0.2f' % np.log10(p.rk + 1e-100),
    'log10(MaxResidual/ConTol)': lambda p: '%0.2f' % np.log10(max((p.rk / p.contol, 1e-100))),
    'isFeasible': lambda p: '+' if p.isFeas(p.xk) else '-',
}


def iter_print(p):
    """kernel/iterPrint.py:24-43: header at iteration 0, then every iprint-th row and the last."""
    if p.lastPrintedIter == p.iter:
        return
    if p.iter == 0 and p.iprint >= 0:
        print(' iter' + _DELIM + ''.join(fn + _DELIM for fn in p.data4TextOutput))
    elif p.iprint < 0 or (((p.iprint > 0 and p.iter % p.iprint != 0) or p.iprint == 0)
                          and not (p.isFinished or p.iter == 0)):
        return
    s = str(p.iter).rjust(5) + '  '
    for col in p.data4TextOutput:
        s += _COLUMNS[col](p).rjust(len(col)) + ' '
    print(s)
    p.lastPrintedIter = p.iter
