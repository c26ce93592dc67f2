"""ctypes wrapper of oracle/_build/libde_oracle.so (the C restatement, de_oracle.c).
Test infrastructure / CPU baseline only -- see the header of de_oracle.c."""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, '_build', 'libde_oracle.so')
OBJ_IDS = {'sphere': 0, 'rosenbrock': 1, 'rastrigin': 2, 'ackley': 3, 'griewank': 4, 'shifted_sphere': 5}
_dp = C.POINTER(C.c_double)
_lib = None


def load():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB):
            subprocess.check_call(['make', '-s', '-C', HERE])
        lib = C.CDLL(LIB)
        lib.orc_create.restype = C.c_void_p
        lib.orc_create.argtypes = [C.c_int, C.c_int, C.c_int64, _dp, _dp, _dp, C.c_double, C.c_double, C.c_uint64,
                                   C.c_int, C.c_int, _dp]
        for name, res in (('orc_pop', _dp), ('orc_vals', _dp), ('orc_tvals', _dp), ('orc_best_x', _dp),
                          ('orc_accept', C.POINTER(C.c_uint8)), ('orc_best_f', C.c_double),
                          ('orc_trial_best_f', C.c_double), ('orc_trial_best_idx', C.c_int64),
                          ('orc_n_accepted', C.c_int64), ('orc_n_repaired', C.c_int64), ('orc_best_changed', C.c_int)):
            getattr(lib, name).restype = res
            getattr(lib, name).argtypes = [C.c_void_p]
        lib.orc_cvals.restype, lib.orc_cvals.argtypes = _dp, [C.c_void_p]
        lib.orc_trial_best_feasible.restype, lib.orc_trial_best_feasible.argtypes = C.c_int, [C.c_void_p]
        lib.orc_trial_best_c.restype, lib.orc_trial_best_c.argtypes = C.c_double, [C.c_void_p]
        lib.orc_set_strategy.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int]
        lib.orc_set_constraints.argtypes = [C.c_void_p, C.POINTER(C.c_int), C.POINTER(_dp), C.POINTER(_dp), C.c_double]
        lib.orc_init.argtypes = [C.c_void_p]
        lib.orc_step.argtypes = [C.c_void_p, C.c_int]
        lib.orc_destroy.argtypes = [C.c_void_p]
        lib.orc_eval.argtypes = [C.c_void_p, _dp, C.c_int64, _dp]
        lib.orc_set_threads.argtypes = [C.c_int]
        lib.orc_set_threads.restype = None
        _lib = lib
    return _lib


def _d(a):
    return a.ctypes.data_as(_dp)


def set_threads(n=None):
    """OpenMP threads of the C port (default: every core this process may run on), whatever OMP_NUM_THREADS
    says -- torchrun exports OMP_NUM_THREADS=1.  Returns the count now in effect."""
    lib = load()
    if n is None:
        n = len(os.sched_getaffinity(0)) if hasattr(os, 'sched_getaffinity') else (os.cpu_count() or 1)
    lib.orc_set_threads(int(n))
    return lib.orc_threads()


class COracleDE:
    def __init__(self, obj, lb, ub, x0=None, population=None, F=0.8, Cr=0.5, seed=150880, stream='philox',
                 invert=False, aux=None, strategy=None, constraints=None):
        """strategy: an object with base / direction / factor / hndvi (de_oracle.Strategy); constraints: an object
        with A, b, Aeq, beq, c, h (c, h: DiagQuadratic-like with .Q, .r) and contol (de_oracle.Constraints)."""
        self.lib = load()
        self.lb = np.ascontiguousarray(lb, dtype=np.float64)
        self.ub = np.ascontiguousarray(ub, dtype=np.float64)
        self.D = self.lb.size
        self.NP = int(10 * self.D if population is None else population)
        x0 = (self.lb + self.ub) / 2.0 if x0 is None else np.ascontiguousarray(x0, dtype=np.float64)
        if obj == 'griewank' and aux is None:
            aux = np.sqrt(np.arange(1, self.D + 1))
        aux = None if aux is None else np.ascontiguousarray(aux, dtype=np.float64)
        self._keep = (x0, aux)
        self.h = self.lib.orc_create(OBJ_IDS[obj], self.D, self.NP, _d(self.lb), _d(self.ub), _d(x0), F, Cr, seed,
                                     1 if stream == 'cpython' else 0, int(invert), _d(aux) if aux is not None else None)
        if strategy is not None:
            self.lib.orc_set_strategy(self.h, int(strategy.base == 'best'), int(strategy.direction == 'best'),
                                      int(strategy.factor == 'constant'), int(strategy.hndvi))
        if constraints is not None:
            mats = [(constraints.A, constraints.b), (constraints.Aeq, constraints.beq),
                    (getattr(constraints.c, 'Q', None), getattr(constraints.c, 'r', None)),
                    (getattr(constraints.h, 'Q', None), getattr(constraints.h, 'r', None))]
            m = (C.c_int * 4)()
            M, R = (_dp * 4)(), (_dp * 4)()
            keep = []
            for k, (Mk, rk) in enumerate(mats):
                if Mk is None or np.size(Mk) == 0:
                    continue
                Mk = np.ascontiguousarray(np.atleast_2d(Mk), dtype=np.float64)
                rk = np.ascontiguousarray(rk, dtype=np.float64).ravel()
                keep += [Mk, rk]
                m[k], M[k], R[k] = Mk.shape[0], _d(Mk), _d(rk)
            self.lib.orc_set_constraints(self.h, m, M, R, float(constraints.contol))
        self.lib.orc_init(self.h)

    def step(self, n=1):
        self.lib.orc_step(self.h, n)

    def _arr(self, name, shape, dtype=np.float64):
        return np.ctypeslib.as_array(getattr(self.lib, name)(self.h), shape=shape).astype(dtype, copy=True)

    pop = property(lambda s: s._arr('orc_pop', (s.NP, s.D)))
    vals = property(lambda s: s._arr('orc_vals', (s.NP,)))
    tvals = property(lambda s: s._arr('orc_tvals', (s.NP,)))
    accept = property(lambda s: s._arr('orc_accept', (s.NP,), np.uint8))
    best_x = property(lambda s: s._arr('orc_best_x', (s.D,)))
    best_f = property(lambda s: s.lib.orc_best_f(s.h))
    trial_best_idx = property(lambda s: s.lib.orc_trial_best_idx(s.h))
    n_accepted = property(lambda s: s.lib.orc_n_accepted(s.h))
    n_repaired = property(lambda s: s.lib.orc_n_repaired(s.h))
    cvals = property(lambda s: s._arr('orc_cvals', (s.NP,)))
    trial_best_feasible = property(lambda s: bool(s.lib.orc_trial_best_feasible(s.h)))
    trial_best_c = property(lambda s: s.lib.orc_trial_best_c(s.h))

    def eval(self, X):
        X = np.ascontiguousarray(np.atleast_2d(X), dtype=np.float64)
        out = np.empty(X.shape[0])
        self.lib.orc_eval(self.h, _d(X), X.shape[0], _d(out))
        return out

    @property
    def threads(self):
        return self.lib.orc_threads()

    def close(self):
        if self.h:
            self.lib.orc_destroy(self.h)
            self.h = None

    __del__ = close
