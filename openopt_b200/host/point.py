"""Lazy point object: caches f and residuals of one x (or a 2-D array of rows).

Reference: kernel/Point.py:15-40 (ctor, f), :62-118 (residual accessors), :128-167 (mr),
:280-340 (betterThan), :363-374 (nNaNs).  Only what the de path touches is provided; the
line-search / gradient helpers of the reference (:376-676) belong to other solvers.
"""
import numpy as np

_EMPTY = np.array(())


class Point:
    __expectedArgs__ = ['x', 'f', 'mr']

    def __init__(self, p, x, *args, **kwargs):
        self.p = p
        self.x = np.copy(x)
        self.isMultiArray = self.x.ndim > 1
        for name, val in zip(self.__expectedArgs__, args):
            setattr(self, '_' + name, val)
        for name, val in kwargs.items():
            setattr(self, '_' + name, val)

    def f(self):
        if not hasattr(self, '_f'):
            p = self.p
            if p._baseClassName == 'NonLin' and p.probType != 'IP':
                self._f = (p.f if p.isObjFunValueASingleNumber else p.F)(self.x)
            else:
                self._f = p.objFunc(self.x)
        return np.copy(self._f)

    # ---- residual accessors (positive = violated) ----
    def lb(self):
        if not hasattr(self, '_lb'):
            self._lb = self.p.lb - self.x
        return np.copy(self._lb)

    def ub(self):
        if not hasattr(self, '_ub'):
            self._ub = self.x - self.p.ub
        return np.copy(self._ub)

    def lin_ineq(self):
        if not hasattr(self, '_lin_ineq'):
            self._lin_ineq = self.p._get_AX_Less_B_residuals(self.x)
        return np.copy(self._lin_ineq)

    def lin_eq(self):
        if not hasattr(self, '_lin_eq'):
            self._lin_eq = self.p._get_AeqX_eq_Beq_residuals(self.x)
        return np.copy(self._lin_eq)

    def c(self):
        if not self.p.userProvided.c:
            return _EMPTY.copy()
        if not hasattr(self, '_c'):
            self._c = self.p.c(self.x)
            self._nNaNs_C = np.isnan(self._c).sum(self._c.ndim - 1)
        return np.copy(self._c)

    def h(self):
        if not self.p.userProvided.h:
            return _EMPTY.copy()
        if not hasattr(self, '_h'):
            self._h = self.p.h(self.x)
            self._nNaNs_H = np.isnan(self._h).sum(self._h.ndim - 1)
        return np.copy(self._h)

    def nNaNs(self):
        p = self.p
        if p._baseClassName != 'NonLin':
            return 0
        n = 0
        if p.userProvided.c:
            self.c()
            n = n + self._nNaNs_C
        if p.userProvided.h:
            self.h()
            n = n + self._nNaNs_H
        return n

    def mr(self, retAll=False, checkBoxBounds=True):
        """Max residual (and, for a single point, which constraint gives it)."""
        if not hasattr(self, '_mr'):
            r, fname, ind = 0, None, 0
            if self.isMultiArray:
                r = np.zeros(self.x.shape[0])
            ineqs = ['lin_ineq', 'lb', 'ub'] if checkBoxBounds else ['lin_ineq']
            eqs = ['lin_eq']
            if self.p._baseClassName == 'NonLin':
                ineqs.append('c')
                eqs.append('h')
            for fields, is_eq in ((ineqs, False), (eqs, True)):
                for field in fields:
                    fv = np.array(getattr(self, field)())
                    if fv.size == 0:
                        continue
                    if is_eq:
                        fv = np.abs(fv)
                    val_max = np.nanmax(fv, fv.ndim - 1)
                    # The reference raises r first and only then tests `r < val_max` to record which
                    # constraint it was (Point.py:146-151), so name/index are never recorded: a point
                    # always reports (mr, None, 0).  Kept as is -- it is what callers observe.
                    r = np.where(r < val_max, val_max, r)
            if self.isMultiArray:
                return r
            self._mr, self._mrName, self._mrInd = r, fname, ind
        if retAll:
            return (np.asarray(self._mr).item(), self._mrName, np.asarray(self._mrInd).item())
        return np.asarray(self._mr).item()

    def betterThan(self, other):
        """True iff self is strictly better than `other` (kernel/Point.py:280-340)."""
        p = self.p
        if p.isUC:
            return self.f() < other.f()
        mr, omr = self.mr(), other.mr()
        n_self, n_other = self.nNaNs(), other.nNaNs()
        if n_other > n_self:
            return True
        if n_other < n_self:
            return False
        if n_self == 0:
            if mr > p.contol and mr > omr:
                return False
            if omr > p.contol and omr > mr:
                return True
        else:
            return mr < omr
        f_self, f_other = self.f(), other.f()
        if not np.isnan(f_other):
            return bool(f_self < f_other) if not np.isnan(f_self) else False
        return bool(mr < omr) if np.isnan(f_self) else True
