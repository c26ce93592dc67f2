"""Solver plugin base class and registry.

Reference: kernel/baseSolver.py:5-33 (the plugin contract), kernel/nonOptMisc.py:89-98 (solvers are
discovered by walking the solvers/ tree for ``<name>_oo.py`` files holding ``class <name>``),
:101-130 (instantiation by name), :132-160 (oosolver).
"""
import importlib
import os

from .iteration import decode_iterfcn_args
from .log import OpenOptException

_UNDEF = 'Undefined. If you are a user and got the message, inform developers please.'


class baseSolver:
    __name__ = _UNDEF
    __license__ = _UNDEF
    __authors__ = _UNDEF
    __alg__ = 'Undefined'
    __solver__ = _UNDEF
    __homepage__ = 'Undefined. Use web search'
    __info__ = 'None'
    _requiresBestPointDetection = False
    _requiresFiniteBoxBounds = False
    useStopByException = True
    __optionalDataThatCanBeHandled__ = []
    __isIterPointAlwaysFeasible__ = lambda self, p: p.isUC
    iterfcnConnected = False
    funcForIterFcnConnection = 'df'
    _canHandleScipySparse = False
    properTextOutput = False
    useLinePoints = False
    __expectedArgs__ = ['xk', 'fk', 'rk']

    def __init__(self):
        pass

    def __decodeIterFcnArgs__(self, p, *args, **kwargs):
        decode_iterfcn_args(p, *args, **kwargs)


_PKG = __name__.rsplit('.', 2)[0]                      # 'openopt_b200'
_SOLVERS_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'solvers')


def _walk():
    paths = {}
    for root, _dirs, files in os.walk(_SOLVERS_DIR):
        if '__pycache__' in root.split(os.sep):
            continue
        rel = os.path.relpath(root, _SOLVERS_DIR)
        mods = [] if rel == '.' else rel.split(os.sep)
        for fn in files:
            if fn.endswith('_oo.py'):
                paths[fn[:-6]] = '.'.join([_PKG, 'solvers'] + mods + [fn[:-3]])
    return paths


solverPaths = _walk()


def solver_import(name):
    return getattr(importlib.import_module(solverPaths[name]), name)


def getSolverFromStringName(p, solver_str):
    if solver_str not in solverPaths:
        p.err('incorrect solver is called, maybe the solver "%s" is misspelled or requires special '
              'installation and is not installed (this package provides: %s)' % (solver_str, sorted(solverPaths)))
    r = solver_import(solver_str)()
    if not hasattr(r, 'fieldsForProbInstance'):
        r.fieldsForProbInstance = {}
    return r


def oosolver(solverName, *args, **kwargs):
    """Pre-bind solver parameters: oosolver('de', population=100) (kernel/nonOptMisc.py:132-160)."""
    if args != ():
        raise OpenOptException("Error: oosolver() doesn't consume any *args, use **kwargs only")
    if isinstance(solverName, baseSolver):
        return solverName
    if ':' in solverName:
        solverName = solverName.split(':')[1]
    try:
        inst = solver_import(solverName)()
        inst.fieldsForProbInstance = {}
        for key, value in kwargs.items():
            if hasattr(inst, key):
                setattr(inst, key, value)
            else:
                inst.fieldsForProbInstance[key] = value
        inst.isInstalled = True
    except (ImportError, KeyError):
        inst = baseSolver()
        inst.__name__ = solverName
        inst.fieldsForProbInstance = {}
        inst.isInstalled = False
    return inst
